// Derisking experiment for a tensor-core first stage (round-2 candidate): tcgen05.mma kind::i8 with
// the A operand described as an OVERLAPPING (Toeplitz) view of a linear one-hot byte stream:
//   row r, k-byte x  ->  O[16 r + x]      (no-swizzle K-major descriptor with LBO = 16 B, SBO = 128 B)
// so that row r is the one-hot encoding of the 8 bases starting at position i0 + p + 4 r, without
// materialising an im2col matrix.  Checks the result against the CPU and times back-to-back MMAs.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);               // start address, bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;      // leading byte offset, bits [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;      // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                // descriptor version (sm_100)
    return d;                                              // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}

constexpr int kM = 128, kN = 256, kK = 32;
constexpr uint32_t kIdesc = (2u << 4)      /* c_format S32 */
                          | (0u << 7)      /* a_format UINT8 */
                          | (0u << 10)     /* b_format UINT8 */
                          | (0u << 15) | (0u << 16)   /* both K-major */
                          | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);

__global__ void __launch_bounds__(128, 1) kernel(const uint8_t* __restrict__ bases /* >= 4*128+16 */, const uint8_t* __restrict__ Bmat /* [256][32] */,
                                                 int32_t* __restrict__ out /* [4 phases][128][256] */, long long* cycles, int reps, int* status) {
    __shared__ __align__(128) uint8_t onehot[4][4 * (4 * 128 + 64) + 64];   // 4 phase copies of the one-hot stream
    __shared__ __align__(128) uint8_t bsm[kN * kK];                         // canonical K-major, no swizzle
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // one-hot streams: copy p holds bases starting at index p
    for (int p = 0; p < 4; ++p)
        for (int i = tid; i < 4 * 128 + 64; i += 128) {
            const uint32_t b = bases[i + p];
            *reinterpret_cast<uint32_t*>(&onehot[p][4 * i]) = 1u << (8 * b);
        }
    // B: element (n, x) at (n/8)*256 + (x/16)*128 + (n%8)*16 + x%16
    for (int i = tid; i < kN * kK; i += 128) {
        const int n = i / kK, x = i % kK;
        bsm[(n / 8) * 256 + (x / 16) * 128 + (n % 8) * 16 + (x % 16)] = Bmat[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");          // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem;
    uint32_t parity = 0;
    int ok = 1;
    for (int p = 0; p < 4 && ok; ++p) {
        const long long t0 = clock64();
        if (tid == 0) {
            const uint64_t adesc = make_desc(smem_u32(&onehot[p][0]), 16, 128);     // Toeplitz view
            const uint64_t bdesc = make_desc(smem_u32(bsm), 128, 256);
            for (int r = 0; r < reps; ++r) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                    ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(0u), "r"(0u));
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)));
        }
        // bounded wait (never hang the GPU on a bad descriptor)
        uint32_t done = 0;
        for (long long spin = 0; spin < 20000000 && !done; ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&mbar)), "r"(parity));
        const long long t1 = clock64();
        if (!done) { ok = 0; if (tid == 0) *status = 100 + p; }
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;");
        if (tid == 0) cycles[p] = t1 - t0;
        if (ok) {
            const uint32_t trow = tmem + ((uint32_t)(32 * warp) << 16);
            for (int c = 0; c < kN; ++c) {
                uint32_t v;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(trow + c));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                out[((size_t)p * kM + 32 * warp + lane) * kN + c] = (int32_t)v;
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
}

int main() {
    const int nb = 4 * 128 + 64 + 8;
    std::vector<uint8_t> bases(nb), B(kN * kK);
    srand(7);
    for (auto& b : bases) b = rand() & 3;
    for (auto& v : B) v = rand() % 100;
    uint8_t *d_bases, *d_B; int32_t* d_out; long long* d_cyc; int* d_status;
    cudaMalloc(&d_bases, nb); cudaMalloc(&d_B, kN * kK); cudaMalloc(&d_out, 4 * kM * kN * 4); cudaMalloc(&d_cyc, 64); cudaMalloc(&d_status, 4);
    cudaMemcpy(d_bases, bases.data(), nb, cudaMemcpyHostToDevice);
    cudaMemcpy(d_B, B.data(), kN * kK, cudaMemcpyHostToDevice);
    for (int reps : {1, 64, 512}) {
        cudaMemset(d_status, 0, 4); cudaMemset(d_out, 0xFF, 4 * kM * kN * 4);
        kernel<<<1, 128>>>(d_bases, d_B, d_out, d_cyc, reps, d_status);
        cudaError_t e = cudaDeviceSynchronize();
        int st; long long cyc[4]; std::vector<int32_t> out(4 * kM * kN);
        cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost); cudaMemcpy(cyc, d_cyc, 32, cudaMemcpyDeviceToHost);
        cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost);
        long long bad = 0, first = -1;
        for (int p = 0; p < 4; ++p) for (int r = 0; r < kM; ++r) for (int n = 0; n < kN; ++n) {
            int ref = 0;
            for (int k = 0; k < 8; ++k) ref += B[n * kK + 4 * k + bases[p + 4 * r + k]];
            if (out[((size_t)p * kM + r) * kN + n] != ref) { if (first < 0) first = ((long long)p * kM + r) * kN + n; ++bad; }
        }
        printf("reps=%d err=%s status=%d mismatches=%lld (first %lld: got %d) cycles/phase=%lld -> %.1f clk per MMA (M128 N256 K32)\n", reps, cudaGetErrorString(e), st, bad, first,
               first >= 0 ? out[first] : 0, cyc[1], (double)cyc[1] / reps);
        if (e != cudaSuccess) break;
    }
    return 0;
}
