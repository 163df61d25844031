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
constexpr int kScanTile = 2048;                 // counters per CTA of the scan

__global__ void bucket_count_kernel(const unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ n_ptr, unsigned long long cap,
                                    int shift, unsigned int* __restrict__ cnt) {
    const unsigned long long n = *n_ptr < cap ? *n_ptr : cap;
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

// Exclusive scan of the bucket counters in three launches (tile sums, scan of the tile sums, per-tile scan); every global
// access is coalesced: a tile's kScanTile counters go through shared memory (padded: one thread then owns 8 consecutive ones).
__device__ __forceinline__ int pad_idx(int i) { return i + (i >> 5); }

__global__ void __launch_bounds__(256) bucket_tile_sum_kernel(const unsigned int* __restrict__ cnt, unsigned long long nb,
                                                              unsigned long long* __restrict__ tile_sum) {
    const unsigned long long t0 = (unsigned long long)blockIdx.x * kScanTile;
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < kScanTile / 256; ++i) { const unsigned long long b = t0 + (unsigned long long)i * 256 + threadIdx.x; if (b < nb) s += cnt[b]; }
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

__global__ void __launch_bounds__(256) bucket_tile_apply_kernel(const unsigned int* __restrict__ cnt, unsigned long long nb,
                                                                const unsigned long long* __restrict__ tile_sum,
                                                                unsigned long long* __restrict__ start, unsigned long long* __restrict__ cursor) {
    constexpr int per = kScanTile / 256;
    __shared__ unsigned int s_c[kScanTile + kScanTile / 32];
    __shared__ unsigned long long s_o[kScanTile + kScanTile / 32];
    const unsigned long long t0 = (unsigned long long)blockIdx.x * kScanTile;
#pragma unroll
    for (int i = 0; i < per; ++i) { const int e = i * 256 + threadIdx.x; s_c[pad_idx(e)] = t0 + e < nb ? cnt[t0 + e] : 0u; }
    __syncthreads();
    unsigned int c[per];
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < per; ++i) { c[i] = s_c[pad_idx(threadIdx.x * per + i)]; s += c[i]; }
    unsigned long long total;
    unsigned long long run = tile_sum[blockIdx.x] + block_exclusive_scan(s, &total);
#pragma unroll
    for (int i = 0; i < per; ++i) { s_o[pad_idx(threadIdx.x * per + i)] = run; run += c[i]; }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < per; ++i) {
        const int e = i * 256 + threadIdx.x;
        if (t0 + e < nb) { const unsigned long long v = s_o[pad_idx(e)]; start[t0 + e] = v; cursor[t0 + e] = v; }
    }
    if (t0 + kScanTile >= nb && threadIdx.x == 0) start[nb] = tile_sum[blockIdx.x] + total;      // the last tile: total = number of hits
}

__global__ void bucket_scatter_kernel(const unsigned long long* __restrict__ keys, const double* __restrict__ sc, const unsigned long long* __restrict__ n_ptr,
                                      int shift, unsigned long long* __restrict__ cursor, unsigned long long* __restrict__ tkeys, double* __restrict__ tsc) {
    const unsigned long long n = *n_ptr;
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

// Single-segment scans: run_off[c] = first hit of combo c = 2 * motif + strand (c = 0 .. 2M).  A bucket never spans two
// combos when shift <= pb: the offset is the bucket's start; otherwise (few hits) the lower bound of c << pb in the sorted keys.
__global__ void run_offsets_kernel(const unsigned long long* __restrict__ skeys, const unsigned long long* __restrict__ start, unsigned long long nb,
                                   int shift, int pb, int nrun, long long* __restrict__ run_off) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > nrun) return;
    const unsigned long long target = (unsigned long long)c << pb;
    if (shift <= pb) {
        const unsigned long long b = target >> shift;
        run_off[c] = (long long)start[b < nb ? b : nb];
        return;
    }
    unsigned long long lo = 0, hi = start[nb];
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

size_t order_work_bytes(size_t nb) { return ((nb * 4 + 63) & ~(size_t)63) + (nb + 1) * 8 + nb * 8 + ((nb + kScanTile - 1) / kScanTile + 1) * 8; }

OrderWork order_work_view(unsigned char* d_work, size_t nb) {
    OrderWork w;
    w.nb = nb;
    w.cnt = (unsigned int*)d_work;
    w.start = (unsigned long long*)(d_work + ((nb * 4 + 63) & ~(size_t)63));
    w.cursor = w.start + nb + 1;
    w.tiles = w.cursor + nb;
    return w;
}

// Counts the stored hits per bucket (only needed when the counting was not fused into the kernels that produce the hits:
// the retry with a smaller shift).  w.cnt must be zero.  Stream-ordered.
int order_count(pwmscan_ctx* ctx, cudaStream_t st, const unsigned long long* d_keys, const unsigned long long* d_nhits, unsigned long long cap,
                int shift, const OrderWork& w, int* launches) {
    bucket_count_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(d_keys, d_nhits, cap, shift, w.cnt);
    PWM_CUDA(cudaGetLastError());
    *launches += 1;
    return PWMSCAN_OK;
}

// Everything after the counting: scan, scatter, per-bucket sort into the output columns, run table.  n_est only sizes the
// grids (the kernels read the real count from the scan's total).  d_flags[0] must be 0 on entry and holds the size of the
// largest bucket that did not fit afterwards.  Stream-ordered, no host synchronisation.
int order_finish(pwmscan_ctx* ctx, cudaStream_t st, const unsigned long long* d_keys, const double* d_sc, unsigned long long n_est, KeyLayout kl,
                 int shift, int nseg, int M, const OrderWork& w, unsigned long long* d_tkeys, double* d_tsc,
                 unsigned long long* d_skeys, double* d_score, uint32_t* d_pos32, int32_t* d_seg32, uint16_t* d_ms16, long long* d_run_off,
                 unsigned long long* d_flags, int* launches) {
    const int sm = ctx->sm_count;
    const unsigned long long ntiles = (w.nb + kScanTile - 1) / kScanTile;
    bucket_tile_sum_kernel<<<(unsigned)ntiles, 256, 0, st>>>(w.cnt, w.nb, w.tiles);
    bucket_tile_scan_kernel<<<1, 1024, 0, st>>>(w.tiles, ntiles);
    bucket_tile_apply_kernel<<<(unsigned)ntiles, 256, 0, st>>>(w.cnt, w.nb, w.tiles, w.start, w.cursor);
    const unsigned grid_n = (unsigned)std::max<unsigned long long>(1, std::min<unsigned long long>((n_est + 255) / 256, (unsigned long long)sm * 16));
    bucket_scatter_kernel<<<grid_n, 256, 0, st>>>(d_keys, d_sc, w.start + w.nb, shift, w.cursor, d_tkeys, d_tsc);
    const unsigned grid_b = (unsigned)std::min<size_t>(w.nb, (size_t)sm * 16);
    bucket_sort_kernel<<<grid_b, kSortThreads, 0, st>>>(w.start, w.nb, shift, d_tkeys, d_tsc, kl, nseg > 1 ? 1 : 0, d_skeys, d_score, d_pos32, d_seg32,
                                                       d_ms16, d_flags);
    *launches += 5;
    if (nseg == 1) {
        run_offsets_kernel<<<(2 * M + 1 + 255) / 256, 256, 0, st>>>(d_skeys, w.start, w.nb, shift, kl.pb, 2 * M, d_run_off);
        *launches += 1;
    }
    PWM_CUDA(cudaGetLastError());
    return PWMSCAN_OK;
}

}  // namespace pwmscan
