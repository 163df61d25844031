// l2gather.cu — how fast can every sequence position gather three random 32-byte sectors from
// L2-resident tables (index = the 8-mer at j, j+8, j+16) and AND them?  Upper bound for a
// "k-mer alive-mask" first stage (see experiments/exp_masktable.py for its selectivity).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o l2gather l2gather.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint4 ldg16(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
struct u8x { uint32_t v[8]; };
__device__ __forceinline__ u8x ldg32(const void* p) {
    u8x r;
#if USE_V8
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
#else
    const uint4 a = ldg16(p), b = ldg16((const char*)p + 16);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
#endif
    return r;
}

// NT tables of 65536 entries x WORDS uint32
template <int NT, int WORDS, int UNROLL>
__global__ void __launch_bounds__(512) gather(const uint32_t* __restrict__ seq2, long long npos, const uint32_t* __restrict__ tab,
                                              unsigned long long* out) {
    const long long stride = (long long)gridDim.x * blockDim.x * UNROLL;
    unsigned long long found = 0;
    for (long long base = ((long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) * UNROLL + (threadIdx.x & 31); base < npos; base += stride) {
        uint32_t acc[UNROLL][WORDS];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long long j = base + 32 * u;
            const uint32_t* w = seq2 + (j >> 4);
            const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
            const int sh = 2 * (int)(j & 15);
            const uint32_t a = __funnelshift_r(w0, w1, sh), b = __funnelshift_r(w1, w2, sh);
            uint32_t idx[3] = {a & 0xFFFFu, a >> 16, b & 0xFFFFu};
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const uint32_t* p = tab + ((size_t)t * 65536 + idx[t]) * WORDS;
                if (WORDS == 8) {
                    const u8x r = ldg32(p);
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[u][k] = t == 0 ? r.v[k] : (acc[u][k] & r.v[k]);
                } else if (WORDS == 4) {
                    const uint4 r = ldg16(p);
                    if (t == 0) { acc[u][0] = r.x; acc[u][1] = r.y; acc[u][2] = r.z; acc[u][3] = r.w; }
                    else { acc[u][0] &= r.x; acc[u][1] &= r.y; acc[u][2] &= r.z; acc[u][3] &= r.w; }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            uint32_t any = 0;
#pragma unroll
            for (int k = 0; k < WORDS; ++k) any |= acc[u][k];
            if (any) {
#pragma unroll
                for (int k = 0; k < WORDS; ++k) found += __popc(acc[u][k]);
            }
        }
    }
    if (found) atomicAdd(out, found);
}

int main(int argc, char** argv) {
    const long long npos = 248956422LL;
    const double density = argc > 1 ? atof(argv[1]) : 0.08;
    const size_t nwords = (size_t)(npos / 16) + 64;
    std::vector<uint32_t> h(nwords);
    uint64_t s = 0x9E3779B97F4A7C15ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    for (auto& x : h) x = (uint32_t)rnd();
    uint32_t* d_seq; CK(cudaMalloc(&d_seq, nwords * 4)); CK(cudaMemcpy(d_seq, h.data(), nwords * 4, cudaMemcpyHostToDevice));
    const size_t tw = (size_t)3 * 65536 * 8;
    std::vector<uint32_t> ht(tw);
    for (auto& x : ht) { uint32_t v = 0; for (int b = 0; b < 32; ++b) if ((rnd() & 0xFFFF) < density * 65536) v |= 1u << b; x = v; }
    uint32_t* d_tab; CK(cudaMalloc(&d_tab, tw * 4)); CK(cudaMemcpy(d_tab, ht.data(), tw * 4, cudaMemcpyHostToDevice));
    unsigned long long* d_out; CK(cudaMalloc(&d_out, 8));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto run = [&](const char* name, auto kern, int nt, int words, int ctas_per_sm, int threads) {
        float best = 1e9;
        unsigned long long found = 0;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaMemset(d_out, 0, 8));
            cudaEventRecord(e0);
            kern<<<sms * ctas_per_sm, threads>>>(d_seq, npos, d_tab, d_out);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
            CK(cudaMemcpy(&found, d_out, 8, cudaMemcpyDeviceToHost));
        }
        const double bytes = (double)npos * nt * words * 4;
        printf("%-28s %7.3f ms  %6.2f TB/s of table sectors  (%.3f survivors/pos)\n", name, best, bytes / best * 1e-9, (double)found / npos);
    };
    run("3 tables x 32 B, unroll 2", gather<3, 8, 2>, 3, 8, 4, 512);
    run("3 tables x 32 B, unroll 4", gather<3, 8, 4>, 3, 8, 2, 512);
    run("2 tables x 32 B, unroll 4", gather<2, 8, 4>, 2, 8, 2, 512);
    run("1 table  x 32 B, unroll 4", gather<1, 8, 4>, 1, 8, 2, 512);
    run("3 tables x 16 B, unroll 4", gather<3, 4, 4>, 3, 4, 4, 512);
    run("3 tables x 16 B, unroll 8", gather<3, 4, 8>, 3, 4, 2, 512);
    run("1 table  x 16 B, unroll 8", gather<1, 4, 8>, 1, 4, 2, 512);
    return 0;
}
