// Micro-benchmark: TMEM (tcgen05.ld.32x32b.x1 with a data-dependent, warp-uniform column) as a
// look-up table port, against the equivalent shared-memory look-up.  One CTA per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t tmem_ld(uint32_t taddr) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
    return v;
}
__device__ __forceinline__ void tmem_st(uint32_t taddr, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" :: "r"(taddr), "r"(v));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 15; x *= 0x2C1B3C6Du; x ^= x >> 12; x *= 0x297A2D39u; x ^= x >> 15; return x; }

template <int MODE>   // 0 = TMEM loads, 1 = LDS loads, 2 = both interleaved
__global__ void __launch_bounds__(1024, 1) bench(uint32_t* out, long long* cycles, int iters, int* errs) {
    extern __shared__ __align__(16) uint32_t tab[];     // [512][32] words = 64 KB
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&s_tmem)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = s_tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    // fill: value(col, lane) = mix(col * 32 + lane); every quarter gets the same table (warps 0..3 fill)
    for (int i = tid; i < 512 * 32; i += blockDim.x) tab[i] = mix((uint32_t)i);
    if (warp < 4) {
        for (int c = 0; c < 512; ++c) tmem_st(tbase + c, mix((uint32_t)(c * 32 + lane)));
        tmem_wait_st();
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    // correctness
    int bad = 0;
    for (int c = warp; c < 512; c += (blockDim.x >> 5)) {
        const uint32_t v = tmem_ld(tbase + c);
        tmem_wait_ld();
        if (v != mix((uint32_t)(c * 32 + lane))) ++bad;
    }
    if (bad) atomicAdd(errs, bad);
    __syncthreads();
    uint32_t acc = 0, x = 0x1234567u + warp * 77u;      // warp-uniform address stream
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(tab) + lane * 4;
    uint32_t col[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { x = x * 1664525u + 1013904223u; col[k] = (x >> 9) & 511u; }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {       // fixed (warp-specific) addresses; "memory" clobbers keep the loads in the loop
            if (MODE == 0 || (MODE == 2 && (k & 1)))
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1]; // %2" : "=r"(v[k]) : "r"(tbase | col[k]), "r"(it) : "memory");
            else asm volatile("ld.volatile.shared.u32 %0, [%1]; // %2" : "=r"(v[k]) : "r"(sbase + col[k] * 128u), "r"(it) : "memory");
        }
        if (MODE != 1) tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 8; ++k) acc ^= v[k];
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + tid] = acc;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(s_tmem), "n"(512));
}

int main() {
    uint32_t* d_out; long long* d_cyc; int* d_err;
    cudaMalloc(&d_out, 148 * 1024 * 4); cudaMalloc(&d_cyc, 148 * 8); cudaMalloc(&d_err, 4);
    cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const int iters = 4096;
    for (int threads = 128; threads <= 1024; threads *= 2) {
        for (int mode = 0; mode < 3; ++mode) {
            cudaMemset(d_err, 0, 4);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            if (mode == 0) bench<0><<<148, threads, 65536>>>(d_out, d_cyc, iters, d_err);
            if (mode == 1) bench<1><<<148, threads, 65536>>>(d_out, d_cyc, iters, d_err);
            if (mode == 2) bench<2><<<148, threads, 65536>>>(d_out, d_cyc, iters, d_err);
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long cyc[148]; int err;
            cudaMemcpy(cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
            cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost);
            const double loads = (double)iters * 8 * (threads / 32);
            printf("threads=%4d mode=%d (%s) err=%s bad=%d ms=%.3f cycles(sm0)=%lld  -> %.3f clk per warp-load per SM\n", threads, mode,
                   mode == 0 ? "tmem" : (mode == 1 ? "lds " : "both"), cudaGetErrorString(e), err, ms, cyc[0], (double)cyc[0] / loads);
        }
    }
    return 0;
}
