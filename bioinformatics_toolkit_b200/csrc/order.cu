// order.cu — puts the hits of a scan into the reference's order, (segment, motif, strand, position)
// (findTFBS yields ascending positions, Motif/Search.hs:47-58; scanMotif runs motif by motif, strand by
// strand over one region after the other, Data/Bed/Utils.hs:108-118), without a library sort:
//
//   count   one counter per bucket = key >> shift (the 64-bit hit key is [segment | motif | strand | position])
//   scan    exclusive prefix sum of the counters -> where every bucket starts
//   scatter every hit into its bucket (order inside a bucket arbitrary)
//   sort    one CTA per bucket: bitonic sort of (key - bucket base, local index) in shared memory, then
//           the compact columns are written in final order: position (u32), score (f64) and either the
//           run table of (motif, strand) for a single-segment scan or 16-bit motif|strand + segment
//
// The result does not depend on the arrival order of the hits (keys are unique), so it is deterministic.
// shift is chosen from the hit count (about kOrderTarget hits per bucket); a bucket that still exceeds the
// shared-memory capacity raises a flag and the caller repeats with a smaller shift — positions are unique
// per (segment, motif, strand), so a bucket spanning <= kOrderCap positions always fits.
#include "internal.h"

using namespace pwmscan;

namespace {

constexpr int kSortThreads = 128;
constexpr int kScanTile = 4096;                 // counters per CTA of the multi-CTA scan

__global__ void bucket_count_kernel(const unsigned long long* __restrict__ keys, unsigned long long n, int shift,
                                    unsigned int* __restrict__ cnt) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x)
        atomicAdd(&cnt[keys[i] >> shift], 1u);
}

// Block-wide exclusive scan of one value per thread (1024 threads max); returns the exclusive prefix, *total = block sum.
__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long* total) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    unsigned long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = lane < nwarp ? s_warp[lane] : 0ull, wx = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, wx, o); if (lane >= o) wx += y; }
        s_warp[lane] = wx - w;
        if (lane == 31) s_total = wx;
    }
    __syncthreads();
    const unsigned long long r = s_warp[warp] + x - v;
    *total = s_total;
    __syncthreads();
    return r;
}

// nb <= 1024 * per: the whole scan in one CTA.  start[nb] = n; cursor = copy of start for the scatter.
__global__ void __launch_bounds__(1024) bucket_scan_single_kernel(const unsigned int* __restrict__ cnt, unsigned long long nb,
                                                                  unsigned long long* __restrict__ start, unsigned long long* __restrict__ cursor) {
    const unsigned long long per = (nb + blockDim.x - 1) / blockDim.x;
    const unsigned long long b0 = (unsigned long long)threadIdx.x * per, b1 = b0 + per < nb ? b0 + per : nb;
    unsigned long long s = 0;
    for (unsigned long long b = b0; b < b1; ++b) s += cnt[b];
    unsigned long long total;
    unsigned long long run = block_exclusive_scan(s, &total);
    for (unsigned long long b = b0; b < b1; ++b) { start[b] = run; cursor[b] = run; run += cnt[b]; }
    if (threadIdx.x == 0) start[nb] = total;
}

// multi-CTA scan, three launches: tile sums, scan of the tile sums, per-tile scan
__global__ void __launch_bounds__(256) bucket_tile_sum_kernel(const unsigned int* __restrict__ cnt, unsigned long long nb,
                                                              unsigned long long* __restrict__ tile_sum) {
    const unsigned long long t0 = (unsigned long long)blockIdx.x * kScanTile;
    unsigned long long s = 0;
    for (int i = threadIdx.x; i < kScanTile; i += 256) if (t0 + i < nb) s += cnt[t0 + i];
    unsigned long long total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) bucket_tile_scan_kernel(unsigned long long* __restrict__ tile_sum, unsigned long long ntiles) {
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (unsigned long long t0 = 0; t0 < ntiles; t0 += 1024) {
        const unsigned long long t = t0 + threadIdx.x;
        const unsigned long long v = t < ntiles ? tile_sum[t] : 0ull;
        unsigned long long total;
        const unsigned long long ex = block_exclusive_scan(v, &total);
        if (t < ntiles) tile_sum[t] = s_carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) bucket_tile_apply_kernel(const unsigned int* __restrict__ cnt, unsigned long long nb, unsigned long long n,
                                                                const unsigned long long* __restrict__ tile_sum,
                                                                unsigned long long* __restrict__ start, unsigned long long* __restrict__ cursor) {
    constexpr int per = kScanTile / 256;
    const unsigned long long b0 = (unsigned long long)blockIdx.x * kScanTile + (unsigned long long)threadIdx.x * per;
    unsigned int c[per];
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < per; ++i) { c[i] = b0 + i < nb ? cnt[b0 + i] : 0u; s += c[i]; }
    unsigned long long total;
    unsigned long long run = tile_sum[blockIdx.x] + block_exclusive_scan(s, &total);
#pragma unroll
    for (int i = 0; i < per; ++i) if (b0 + i < nb) { start[b0 + i] = run; cursor[b0 + i] = run; run += c[i]; }
    if (blockIdx.x == 0 && threadIdx.x == 0) start[nb] = n;
}

__global__ void bucket_scatter_kernel(const unsigned long long* __restrict__ keys, const double* __restrict__ sc, unsigned long long n, int shift,
                                      unsigned long long* __restrict__ cursor, unsigned long long* __restrict__ tkeys, double* __restrict__ tsc) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long k = keys[i];
        const unsigned long long slot = atomicAdd(&cursor[k >> shift], 1ull);
        tkeys[slot] = k;
        tsc[slot] = sc[i];
    }
}

// One CTA per bucket (grid-stride over the buckets).  flags[0] = largest bucket that did not fit.
__global__ void __launch_bounds__(kSortThreads) bucket_sort_kernel(const unsigned long long* __restrict__ start, unsigned long long nb, int shift,
                                                                   const unsigned long long* __restrict__ tkeys, const double* __restrict__ tsc,
                                                                   KeyLayout kl, int multi_seg, unsigned long long* __restrict__ skeys,
                                                                   double* __restrict__ score, uint32_t* __restrict__ pos32,
                                                                   int32_t* __restrict__ seg32, uint16_t* __restrict__ ms16,
                                                                   unsigned long long* __restrict__ flags) {
    __shared__ unsigned long long v[kOrderCap];
    const int tid = threadIdx.x;
    const unsigned long long pos_mask = (1ull << kl.pb) - 1ull, ms_mask = (1ull << (kl.mb + 1)) - 1ull;
    for (unsigned long long b = blockIdx.x; b < nb; b += gridDim.x) {
        const unsigned long long s = start[b], e = start[b + 1];
        const unsigned long long cnt = e - s;
        if (cnt == 0) continue;
        if (cnt > (unsigned long long)kOrderCap) { if (tid == 0) atomicMax(&flags[0], cnt); continue; }
        const unsigned long long base = shift < 64 ? (b << shift) : 0ull;
        int n2 = 1;
        while (n2 < (int)cnt) n2 <<= 1;
        for (int i = tid; i < n2; i += kSortThreads)
            v[i] = i < (int)cnt ? (((tkeys[s + i] - base) << kOrderIdxBits) | (unsigned long long)i) : ~0ull;
        __syncthreads();
        for (int k = 2; k <= n2; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < n2; i += kSortThreads) {
                    const int p = i ^ j;
                    if (p > i) {
                        const unsigned long long a = v[i], c = v[p];
                        if ((a > c) == ((i & k) == 0)) { v[i] = c; v[p] = a; }
                    }
                }
                __syncthreads();
            }
        for (int i = tid; i < (int)cnt; i += kSortThreads) {
            const unsigned long long x = v[i];
            const unsigned long long key = base + (x >> kOrderIdxBits);
            const unsigned long long o = s + (unsigned long long)i;
            skeys[o] = key;
            pos32[o] = (uint32_t)(key & pos_mask);
            score[o] = tsc[s + (x & ((1ull << kOrderIdxBits) - 1ull))];
            if (multi_seg) {
                ms16[o] = (uint16_t)((key >> kl.pb) & ms_mask);
                seg32[o] = (int32_t)(key >> (kl.pb + 1 + kl.mb));
            }
        }
        __syncthreads();
    }
}

// Single-segment scans: run_off[c] = first hit of combo c = 2 * motif + strand (c = 0 .. 2M): lower bound of c << pb.
__global__ void run_offsets_kernel(const unsigned long long* __restrict__ skeys, unsigned long long n, int pb, int nrun,
                                   long long* __restrict__ run_off) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > nrun) return;
    const unsigned long long target = (unsigned long long)c << pb;
    unsigned long long lo = 0, hi = n;
    while (lo < hi) { const unsigned long long mid = (lo + hi) >> 1; if (skeys[mid] < target) lo = mid + 1; else hi = mid; }
    run_off[c] = (long long)lo;
}

}  // namespace

namespace pwmscan {

// Bucket shift for n hits whose keys have key_bits significant bits.
int order_pick_shift(unsigned long long n, int key_bits) {
    int lb = 0;                                             // log2 of the bucket count aimed at
    while (lb < 26 && ((unsigned long long)kOrderTarget << lb) < n) ++lb;
    int shift = key_bits > lb ? key_bits - lb : 0;
    if (shift > 64 - kOrderIdxBits) shift = 64 - kOrderIdxBits;  // (key - base) << idx bits must fit 64 bits
    return shift;
}

size_t order_bucket_count(int key_bits, int shift) { return key_bits > shift ? ((size_t)1 << (key_bits - shift)) : 1; }

// Stream-ordered.  d_work: cnt u32[nb] + start u64[nb+1] + cursor u64[nb] + tiles u64[ceil(nb / 4096)] (order_work_bytes);
// d_tkeys/d_tsc: n entries of scratch; outputs as described on top.  d_flags[0] must be 0 on entry (the caller
// clears it with the other counters) and holds the size of the largest overflowing bucket afterwards.
size_t order_work_bytes(size_t nb) { return nb * 4 + 64 + (nb + 1) * 8 + nb * 8 + ((nb + kScanTile - 1) / kScanTile + 1) * 8; }

int order_hits(pwmscan_ctx* ctx, cudaStream_t st, const unsigned long long* d_keys, const double* d_sc, unsigned long long n, KeyLayout kl,
               int key_bits, int shift, int nseg, int M, unsigned char* d_work, unsigned long long* d_tkeys, double* d_tsc,
               unsigned long long* d_skeys, double* d_score, uint32_t* d_pos32, int32_t* d_seg32, uint16_t* d_ms16, long long* d_run_off,
               unsigned long long* d_flags, int* launches) {
    const size_t nb = order_bucket_count(key_bits, shift);
    unsigned int* cnt = (unsigned int*)d_work;
    unsigned long long* start = (unsigned long long*)(d_work + ((nb * 4 + 63) & ~(size_t)63));
    unsigned long long* cursor = start + nb + 1;
    unsigned long long* tiles = cursor + nb;
    const int sm = ctx->sm_count;
    PWM_CUDA(cudaMemsetAsync(cnt, 0, nb * 4, st));
    const unsigned grid_n = (unsigned)std::min<unsigned long long>((n + 255) / 256, (unsigned long long)sm * 16);
    bucket_count_kernel<<<grid_n, 256, 0, st>>>(d_keys, n, shift, cnt);
    if (nb <= 32768) {
        bucket_scan_single_kernel<<<1, 1024, 0, st>>>(cnt, nb, start, cursor);
        *launches += 1;
    } else {
        const unsigned long long ntiles = (nb + kScanTile - 1) / kScanTile;
        bucket_tile_sum_kernel<<<(unsigned)ntiles, 256, 0, st>>>(cnt, nb, tiles);
        bucket_tile_scan_kernel<<<1, 1024, 0, st>>>(tiles, ntiles);
        bucket_tile_apply_kernel<<<(unsigned)ntiles, 256, 0, st>>>(cnt, nb, n, tiles, start, cursor);
        *launches += 3;
    }
    bucket_scatter_kernel<<<grid_n, 256, 0, st>>>(d_keys, d_sc, n, shift, cursor, d_tkeys, d_tsc);
    const unsigned grid_b = (unsigned)std::min<size_t>(nb, (size_t)sm * 16);
    bucket_sort_kernel<<<grid_b, kSortThreads, 0, st>>>(start, nb, shift, d_tkeys, d_tsc, kl, nseg > 1 ? 1 : 0, d_skeys, d_score, d_pos32, d_seg32,
                                                       d_ms16, d_flags);
    *launches += 3;
    if (nseg == 1) {
        run_offsets_kernel<<<(2 * M + 1 + 255) / 256, 256, 0, st>>>(d_skeys, n, kl.pb, 2 * M, d_run_off);
        *launches += 1;
    }
    PWM_CUDA(cudaGetLastError());
    return PWMSCAN_OK;
}

}  // namespace pwmscan
