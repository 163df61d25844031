// l2gather_L.cu — gather rate vs table size: two mask tables indexed by L-mers (L = 8, 9, 10: 2, 8, 32 MiB each),
// 32-byte sectors, 256-bit loads; second lookup unconditional or only when the first mask is non-zero.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o l2gather_L l2gather_L.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
struct u8x { uint32_t v[8]; };
__device__ __forceinline__ u8x ldg32(const void* p) {
    u8x r;
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
    return r;
}
template <int L, bool COND>
__global__ void __launch_bounds__(1024, 1) gather(const uint32_t* __restrict__ seq2, long long npos, const uint32_t* __restrict__ tab, unsigned long long* out) {
    constexpr uint32_t MASK = (1u << (2 * L)) - 1u;
    constexpr size_t ENT = (size_t)1 << (2 * L);
    const long long stride = (long long)gridDim.x * blockDim.x * 2;
    unsigned long long found = 0;
    for (long long base = ((long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) * 2 + (threadIdx.x & 31); base < npos; base += stride) {
        u8x m[2]; unsigned long long x[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long j = base + 32 * u;
            const uint32_t* w = seq2 + (j >> 4);
            const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
            const int sh = 2 * (int)(j & 15);
            x[u] = (unsigned long long)__funnelshift_r(w0, w1, sh) | ((unsigned long long)__funnelshift_r(w1, w2, sh) << 32);
            m[u] = ldg32(tab + ((size_t)((uint32_t)x[u] & MASK)) * 8);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            uint32_t any = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) any |= m[u].v[k];
            if (!COND || any) {
                const u8x t = ldg32(tab + (ENT + (size_t)((uint32_t)(x[u] >> (2 * L)) & MASK)) * 8);
#pragma unroll
                for (int k = 0; k < 8; ++k) m[u].v[k] &= t.v[k];
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            uint32_t any = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) any |= m[u].v[k];
            if (any) {
#pragma unroll
                for (int k = 0; k < 8; ++k) found += __popc(m[u].v[k]);
            }
        }
    }
    if (found) atomicAdd(out, found);
}
int main(int argc, char** argv) {
    const long long npos = 248956422LL;
    const size_t nwords = (size_t)(npos / 16) + 64;
    std::vector<uint32_t> h(nwords);
    uint64_t s = 0x9E3779B97F4A7C15ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    for (auto& x : h) x = (uint32_t)rnd();
    uint32_t* d_seq; CK(cudaMalloc(&d_seq, nwords * 4)); CK(cudaMemcpy(d_seq, h.data(), nwords * 4, cudaMemcpyHostToDevice));
    unsigned long long* d_out; CK(cudaMalloc(&d_out, 8));
    uint32_t* d_flush; CK(cudaMalloc(&d_flush, 256u << 20));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto run = [&](const char* name, auto kern, int L, double density) {
        const size_t tw = (size_t)2 * ((size_t)1 << (2 * L)) * 8;
        std::vector<uint32_t> ht(tw);
        for (auto& x : ht) { uint32_t v = 0; for (int b = 0; b < 32; ++b) if ((rnd() & 0xFFFF) < density * 65536) v |= 1u << b; x = v; }
        uint32_t* d_tab; CK(cudaMalloc(&d_tab, tw * 4)); CK(cudaMemcpy(d_tab, ht.data(), tw * 4, cudaMemcpyHostToDevice));
        float best = 1e9, cold = 0; unsigned long long found = 0;
        for (int rep = 0; rep < 4; ++rep) {
            if (rep == 0) CK(cudaMemset(d_flush, 1, 256u << 20));
            CK(cudaMemset(d_out, 0, 8));
            cudaEventRecord(e0);
            kern<<<sms, 1024>>>(d_seq, npos, d_tab, d_out);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep == 0) cold = ms;
            if (ms < best) best = ms;
            CK(cudaMemcpy(&found, d_out, 8, cudaMemcpyDeviceToHost));
        }
        printf("%-34s tables %5.0f MiB  cold %7.3f ms  best %7.3f ms  (%.3f survivors/pos)\n", name, tw * 4 / 1048576.0, cold, best, (double)found / npos);
        cudaFree(d_tab);
    };
    run("L=8  always 2 lookups", gather<8, false>, 8, 0.03);
    run("L=9  always 2 lookups", gather<9, false>, 9, 0.03);
    run("L=10 always 2 lookups", gather<10, false>, 10, 0.03);
    run("L=8  2nd if non-zero (p=99.9%)", gather<8, true>, 8, 0.03);
    run("L=10 2nd if non-zero (p~40%)", gather<10, true>, 10, 0.002);
    run("L=10 2nd if non-zero (p~12%)", gather<10, true>, 10, 0.0005);
    run("L=9  2nd if non-zero (p~40%)", gather<9, true>, 9, 0.002);
    return 0;
}
