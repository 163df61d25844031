// TMEM read throughput with wide loads (tcgen05.ld.32x32b.x32 = 32 columns x 32 lanes x 4 B = 4 KB per warp)
// + the OR-reduction the tensor-core epilogue needs.  One CTA per SM, W warps (each on its lane quarter).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int REDUCE>
__global__ void __launch_bounds__(1024, 1) bench(uint32_t* out, long long* cycles, int iters) {
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_tmem)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t trow = s_tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int c0 = 0; c0 < 256; c0 += 32) {
            uint32_t v[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
                "tcgen05.wait::ld.sync.aligned;"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(trow + (uint32_t)c0 + (uint32_t)((it & 1) * 256)) : "memory");
            if (REDUCE) {
                uint32_t o0 = v[0] | v[1] | v[2], o1 = v[3] | v[4] | v[5], o2 = v[6] | v[7] | v[8], o3 = v[9] | v[10] | v[11];
#pragma unroll
                for (int k = 12; k < 32; k += 4) { o0 |= v[k]; o1 |= v[k + 1]; o2 |= v[k + 2]; o3 |= v[k + 3]; }
                acc |= o0 | o1 | o2 | o3;
            } else acc |= v[0] ^ v[31];
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + tid] = acc;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(512));
}

int main() {
    uint32_t* d_out; long long* d_cyc;
    cudaMalloc(&d_out, 148 * 1024 * 4); cudaMalloc(&d_cyc, 148 * 8);
    const int iters = 2000;
    for (int warps : {1, 4, 8, 16, 32})
        for (int red = 0; red < 2; ++red) {
            if (red) bench<1><<<148, warps * 32>>>(d_out, d_cyc, iters); else bench<0><<<148, warps * 32>>>(d_out, d_cyc, iters);
            cudaError_t e = cudaDeviceSynchronize();
            long long cyc[148];
            cudaMemcpy(cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
            const double loads = (double)iters * 8 * warps;
            printf("warps=%2d reduce=%d err=%s: %.1f clk per x32 load per SM (4 KB) -> %.0f B/clk/SM; %.1f clk per 128x256 tile\n", warps, red,
                   cudaGetErrorString(e), cyc[0] / loads, 4096.0 * loads / cyc[0], cyc[0] / (double)iters * 4 / warps * (warps < 4 ? 4.0 / warps * warps / 4 : 1));
        }
    return 0;
}
