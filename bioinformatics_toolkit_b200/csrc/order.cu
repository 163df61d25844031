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

constexpr int kSortThreads = 256;               // CTA of the shared-memory sort (big buckets)
constexpr int kSpanCap = 64;                    // hits per span of whole buckets in the main sorting kernel (2 per lane)
constexpr int kWarpCap = 256;                   // hits one warp orders in registers at most (8 per lane: the medium kernel)
constexpr int kWarpIdxBits = 8;
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

// (also zeroes the two queue counters of the sorting kernels: a cudaMemsetAsync in the middle of the chain is a copy-engine
//  operation and waits behind whatever hit-list copies of other items that engine is busy with)
__global__ void __launch_bounds__(1024) bucket_tile_scan_kernel(unsigned long long* __restrict__ tile_sum, unsigned long long ntiles,
                                                                unsigned long long* __restrict__ zero_a, unsigned long long* __restrict__ zero_b) {
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) { s_carry = 0; *zero_a = 0ull; *zero_b = 0ull; }
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

// Writes the element that ended up at sorted position `o` of a bucket: v = (key - base) << idx_bits | index among the
// bucket's scattered elements.
struct SortOut {
    const double* tsc; unsigned long long* skeys; double* score; uint32_t* pos32; int32_t* seg32; uint16_t* ms16;
    KeyLayout kl; int multi_seg;
};
__device__ __forceinline__ void sort_emit(const SortOut& so, unsigned long long s, unsigned long long o, unsigned long long base, unsigned long long v, int idx_bits) {
    const unsigned long long key = base + (v >> idx_bits);
    so.skeys[o] = key;
    so.pos32[o] = (uint32_t)(key & ((1ull << so.kl.pb) - 1ull));
    so.score[o] = so.tsc[s + (v & ((1ull << idx_bits) - 1ull))];
    if (so.multi_seg) {
        so.ms16[o] = (uint16_t)((key >> so.kl.pb) & ((1ull << (so.kl.mb + 1)) - 1ull));
        so.seg32[o] = (int32_t)(key >> (so.kl.pb + 1 + so.kl.mb));
    }
}

// Bitonic sort of 32 * E values held E per lane (value index = lane * E + r), ascending, entirely in registers: partners
// closer than E live in the same lane, the others are one shuffle away.  No shared memory, no barrier.
template <int E>
__device__ __forceinline__ void warp_bitonic(unsigned long long (&v)[E]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 2; k <= 32 * E; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= E) {
                const int lj = j / E;                                   // partner lane = lane ^ lj, same r
                const bool lower = (lane & lj) == 0;
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, v[r], lj);
                    const bool up = ((lane * E + r) & k) == 0;
                    const bool keep_min = (up == lower);
                    const unsigned long long mn = v[r] < o ? v[r] : o, mx = v[r] < o ? o : v[r];
                    v[r] = keep_min ? mn : mx;
                }
            } else {
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    if ((r & j) == 0) {
                        const bool up = ((lane * E + r) & k) == 0;
                        const unsigned long long a = v[r], c = v[r | j];
                        if ((a > c) == up) { v[r] = c; v[r | j] = a; }
                    }
                }
            }
        }
    }
}

// Orders the span [s, s + cnt) of the scattered hits — whole consecutive buckets, so ascending key order IS the final
// order — in registers.  base = smallest key the span can hold; (key - base) < 2^(shift + 5) leaves room for the index.
template <int E>
__device__ __forceinline__ void warp_sort_span(unsigned long long s, int cnt, unsigned long long base, const unsigned long long* __restrict__ tkeys,
                                               const SortOut& so) {
    const int lane = threadIdx.x & 31;
    unsigned long long v[E];
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const int i = lane * E + r;
        v[r] = i < cnt ? (((tkeys[s + i] - base) << kWarpIdxBits) | (unsigned long long)i) : ~0ull;
    }
    warp_bitonic<E>(v);
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const int i = lane * E + r;
        if (i < cnt) sort_emit(so, s, s + (unsigned long long)i, base, v[r], kWarpIdxBits);
    }
}

// One WARP per group of 32 consecutive buckets (grid-stride over the groups): the lanes fetch the 33 bucket boundaries with
// two coalesced loads, then the warp walks the group in spans of whole buckets holding at most kSpanCap hits, each ordered
// in registers (empty and tiny buckets cost nothing extra; short bitonic networks: ~6 instructions per hit).  A single
// bucket above kSpanCap goes to the list of bucket_sort_medium_kernel (<= kWarpCap hits) or bucket_sort_big_kernel:
// list[0] = number of listed buckets, list[1..] their indices.
__device__ __forceinline__ void list_push(unsigned long long* list, unsigned long long cap, unsigned long long b, unsigned long long cnt, unsigned long long* flags) {
    const unsigned long long k = atomicAdd(&list[0], 1ull);
    if (k < cap) list[1 + k] = b; else atomicMax(&flags[0], cnt);      // (cannot happen: the lists have room for every such bucket)
}

__global__ void __launch_bounds__(256) bucket_sort_kernel(const unsigned long long* __restrict__ start, unsigned long long nb, int shift,
                                                          const unsigned long long* __restrict__ tkeys, SortOut so,
                                                          unsigned long long* __restrict__ med, unsigned long long med_cap,
                                                          unsigned long long* __restrict__ big, unsigned long long big_cap, unsigned long long* __restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const unsigned long long nwarp = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    const unsigned long long ngroup = (nb + 31) >> 5;
    for (unsigned long long g = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < ngroup; g += nwarp) {
        const unsigned long long b0 = g << 5;
        const unsigned long long bl = b0 + lane < nb ? b0 + lane : nb;
        const unsigned long long st = start[bl], en = start[bl < nb ? bl + 1 : nb];      // lanes past the last bucket: empty
        if (__shfl_sync(0xFFFFFFFFu, en, 31) == __shfl_sync(0xFFFFFFFFu, st, 0)) continue;   // nothing in this group
        int l0 = 0;
        while (l0 < 32) {
            const unsigned long long s0 = __shfl_sync(0xFFFFFFFFu, st, l0);
            const unsigned fit = __ballot_sync(0xFFFFFFFFu, lane >= l0 && en - s0 <= (unsigned long long)kSpanCap);
            if (fit == 0u) {                                          // bucket l0 alone is too large for this kernel
                if (lane == l0) {
                    if (en - st <= (unsigned long long)kWarpCap) list_push(med, med_cap, b0 + l0, en - st, flags);
                    else list_push(big, big_cap, b0 + l0, en - st, flags);
                }
                ++l0;
                continue;
            }
            const int l1 = 32 - __clz((int)fit);                     // boundaries are non-decreasing: the fitting lanes are l0 .. l1-1
            const int cnt = (int)(__shfl_sync(0xFFFFFFFFu, en, l1 - 1) - s0);
            const unsigned long long base = shift < 64 ? ((b0 + (unsigned long long)l0) << shift) : 0ull;
            if (cnt > 32) warp_sort_span<2>(s0, cnt, base, tkeys, so);
            else if (cnt > 0) warp_sort_span<1>(s0, cnt, base, tkeys, so);
            l0 = l1;
        }
    }
}

// One warp per listed bucket of kSpanCap < hits <= kWarpCap: the same register sort, 4 or 8 hits per lane.
__global__ void __launch_bounds__(256) bucket_sort_medium_kernel(const unsigned long long* __restrict__ start, int shift,
                                                                 const unsigned long long* __restrict__ tkeys, SortOut so,
                                                                 const unsigned long long* __restrict__ med, unsigned long long med_cap) {
    const unsigned long long nwarp = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    const unsigned long long nmed = med[0] < med_cap ? med[0] : med_cap;
    for (unsigned long long q = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < nmed; q += nwarp) {
        const unsigned long long b = med[1 + q];
        const unsigned long long s = start[b];
        const int cnt = (int)(start[b + 1] - s);
        const unsigned long long base = shift < 64 ? (b << shift) : 0ull;
        if (cnt > 128) warp_sort_span<8>(s, cnt, base, tkeys, so);
        else warp_sort_span<4>(s, cnt, base, tkeys, so);
    }
}

// One CTA per listed bucket: bitonic sort in shared memory, up to kOrderCap hits.  flags[0] = largest bucket that did not fit.
__global__ void __launch_bounds__(kSortThreads) bucket_sort_big_kernel(const unsigned long long* __restrict__ start, int shift,
                                                                       const unsigned long long* __restrict__ tkeys, SortOut so,
                                                                       const unsigned long long* __restrict__ big, unsigned long long big_cap,
                                                                       unsigned long long* __restrict__ flags) {
    __shared__ unsigned long long v[kOrderCap];
    const int tid = threadIdx.x;
    const unsigned long long nbig = big[0] < big_cap ? big[0] : big_cap;
    for (unsigned long long q = blockIdx.x; q < nbig; q += gridDim.x) {
        const unsigned long long b = big[1 + q];
        const unsigned long long s = start[b], e = start[b + 1];
        const unsigned long long cnt = e - s;
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
        for (int i = tid; i < (int)cnt; i += kSortThreads) sort_emit(so, s, s + (unsigned long long)i, base, v[i], kOrderIdxBits);
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
    if (shift > 64 - kOrderIdxBits - 5) shift = 64 - kOrderIdxBits - 5;   // (key - base of a 32-bucket span) << idx bits must fit 64 bits
    return shift;
}

size_t order_bucket_count(int key_bits, int shift) { return key_bits > shift ? ((size_t)1 << (key_bits - shift)) : 1; }

// cnt u32[nb] | start u64[nb+1] | cursor u64[nb] | tile sums | list of big buckets (1 + big_cap entries; a big bucket holds
// more than kWarpCap hits, so hit_cap / kWarpCap + 1 entries always suffice)
static size_t big_cap_of(size_t nb, size_t hit_cap) { return std::min(nb, hit_cap / kWarpCap + 1); }
static size_t med_cap_of(size_t nb, size_t hit_cap) { return std::min(nb, hit_cap / kSpanCap + 1); }
size_t order_work_bytes(size_t nb, size_t hit_cap) {
    return ((nb * 4 + 63) & ~(size_t)63) + (nb + 1) * 8 + nb * 8 + ((nb + kScanTile - 1) / kScanTile + 1) * 8 + (big_cap_of(nb, hit_cap) + 1) * 8 +
           (med_cap_of(nb, hit_cap) + 1) * 8;
}

OrderWork order_work_view(unsigned char* d_work, size_t nb, size_t hit_cap) {
    OrderWork w;
    w.nb = nb;
    w.big_cap = big_cap_of(nb, hit_cap);
    w.cnt = (unsigned int*)d_work;
    w.start = (unsigned long long*)(d_work + ((nb * 4 + 63) & ~(size_t)63));
    w.cursor = w.start + nb + 1;
    w.tiles = w.cursor + nb;
    w.big = w.tiles + (nb + kScanTile - 1) / kScanTile + 1;
    w.med = w.big + w.big_cap + 1;
    w.med_cap = med_cap_of(nb, hit_cap);
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
    bucket_tile_scan_kernel<<<1, 1024, 0, st>>>(w.tiles, ntiles, w.big, w.med);
    bucket_tile_apply_kernel<<<(unsigned)ntiles, 256, 0, st>>>(w.cnt, w.nb, w.tiles, w.start, w.cursor);
    const unsigned grid_n = (unsigned)std::max<unsigned long long>(1, std::min<unsigned long long>((n_est + 255) / 256, (unsigned long long)sm * 16));
    bucket_scatter_kernel<<<grid_n, 256, 0, st>>>(d_keys, d_sc, w.start + w.nb, shift, w.cursor, d_tkeys, d_tsc);
    SortOut so{d_tsc, d_skeys, d_score, d_pos32, d_seg32, d_ms16, kl, nseg > 1 ? 1 : 0};
    const unsigned grid_b = (unsigned)std::max<size_t>(1, std::min<size_t>((w.nb + 255) / 256, (size_t)sm * 8));      // 8 warps = 8 groups of 32 buckets per CTA
    bucket_sort_kernel<<<grid_b, 256, 0, st>>>(w.start, w.nb, shift, d_tkeys, so, w.med, w.med_cap, w.big, w.big_cap, d_flags);
    bucket_sort_medium_kernel<<<sm * 4, 256, 0, st>>>(w.start, shift, d_tkeys, so, w.med, w.med_cap);
    bucket_sort_big_kernel<<<sm * 6, kSortThreads, 0, st>>>(w.start, shift, d_tkeys, so, w.big, w.big_cap, d_flags);
    *launches += 2;
    *launches += 5;
    if (nseg == 1) {
        run_offsets_kernel<<<(2 * M + 1 + 255) / 256, 256, 0, st>>>(d_skeys, w.start, w.nb, shift, kl.pb, 2 * M, d_run_off);
        *launches += 1;
    }
    PWM_CUDA(cudaGetLastError());
    return PWMSCAN_OK;
}

}  // namespace pwmscan
