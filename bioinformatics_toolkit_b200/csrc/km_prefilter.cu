// km_prefilter.cu — k-mer mask first stage of the thresholded scan (plan_km.h has the maths): decides, for every
// sequence position and every (motif, strand), whether the reference's lookAheadSearch' (Motif/Search.hs:169-198)
// could still report the window — never a false negative; survivors are replayed exactly (rescore_kernel).
//
//   alive(u) = T_0[k-mer at u] & T_1[k-mer at u + L_0] & ... & T_{nt-1}[k-mer at u + L_0 + .. + L_{nt-2}]    (256-bit masks)
//
// (table t indexed by an 8-, 9- or 10-mer: 2, 8 or 32 MiB.)  The tables of the running block stay
// in L2; every stored position gathers one 32-byte sector per table with a single 256-bit load
// (ld.global.nc.v8.u32 — sm_100), the next table only while the AND is still non-zero.  The
// measured bound is the gather rate of the SM's L1 (about one sector per clock and SM, the same
// for 4 MiB and 64 MiB of tables: experiments/masktab/l2gather.cu, l2gather_L.cu), not bytes —
// hence 256 combos per sector and a selective first table.
//
// One persistent CTA per SM (32 warps) pulls work items (block, 8192 positions).  A surviving bit
// (position, combo) goes into the warp's shared-memory queue; whenever 32 are queued the warp runs
// the middle stage on them, one per lane: the whole-window deficit in 16-bit fixed point from the
// block's q16 table in shared memory (bases from the item's staged 2-bit words), early exit once
// the budget kKmQ is exceeded.  What passes (about the hit rate) is appended to the candidate
// list and replayed exactly in fp64 by rescore_kernel.
#include <cstdio>
#include <cstdlib>

#include "internal.h"
#include "plan_km.h"

using namespace pwmscan;

namespace {

constexpr int kKmThreads = 1024;
constexpr int kKmIterPos = 2048;                 // item sizes are multiples of this (2 positions per thread)
constexpr int kKmMaxIters = 16;                  // iterations per work item: 16 (32 768 positions), fewer for short inputs
constexpr int kKmMinItemPos = kKmIterPos;        // launches cover multiples of kAnchorRound, a multiple of every item size
constexpr int kKmHaloWords = 4;                  // 64 stored bases in front of the item (window starts reach back n-1 <= 63)
constexpr int kKmSeqWords = kKmHaloWords + kKmMaxIters * kKmIterPos / 16 + 5;   // windows end <= 7 + 64 bases after the item's last position
constexpr int kKmQueue = 64;
constexpr int kKmHitQueue = 96;                  // < 32 waiting + up to 32 new per ballot, rounded up

struct Mask { uint32_t v[8]; };

// One 256-bit load = one mask.  The tables are the only data worth keeping in L2 (the genome and the hit
// lists stream through once), so the loads carry an evict_last cache policy.
__device__ __forceinline__ uint64_t l2_keep_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ Mask ldg_mask(const uint8_t* p, uint64_t pol) {
    Mask r;
    asm volatile("ld.global.nc.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ uint32_t mask_any(const Mask& m) {
    return (m.v[0] | m.v[1] | m.v[2]) | (m.v[3] | m.v[4] | m.v[5]) | (m.v[6] | m.v[7]);
}

// Table build: bit j of T[x] = (sum_{p<L} wdef[j][p][base p of x] <= dsafe[j]).  One L-mer x per thread.
__global__ void __launch_bounds__(256) km_build_kernel(const double* __restrict__ wdef, const double* __restrict__ dsafe, int L,
                                                       uint32_t* __restrict__ table) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_w = reinterpret_cast<double*>(smem_raw);               // [256][kKmMaxL][4]
    double* s_d = s_w + kKmCombos * kKmMaxL * 4;                     // [256]
    for (int i = threadIdx.x; i < kKmCombos * kKmMaxL * 4; i += blockDim.x) s_w[i] = wdef[i];
    for (int i = threadIdx.x; i < kKmCombos; i += blockDim.x) s_d[i] = dsafe[i];
    __syncthreads();
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    int off[kKmMaxL];
#pragma unroll
    for (int p = 0; p < kKmMaxL; ++p) off[p] = 4 * p + (int)((x >> (2 * p)) & 3u);     // p >= L: base 0 of an all-zero row
    uint32_t out[kKmWords];
#pragma unroll 1
    for (int w = 0; w < kKmWords; ++w) {
        uint32_t bits = 0;
#pragma unroll 2
        for (int j = 0; j < 32; ++j) {
            const double* wd = s_w + (w * 32 + j) * (kKmMaxL * 4);
            double s = wd[off[0]];
#pragma unroll
            for (int p = 1; p < kKmMaxL; ++p) if (p < L) s = __dadd_rn(s, wd[off[p]]);
            bits |= (s <= s_d[w * 32 + j]) ? (1u << j) : 0u;
        }
        out[w] = bits;
    }
    uint4* dst = reinterpret_cast<uint4*>(table + (size_t)x * kKmWords);
    dst[0] = make_uint4(out[0], out[1], out[2], out[3]);
    dst[1] = make_uint4(out[4], out[5], out[6], out[7]);
}

// bit x of the bitmap = (mask of k-mer x is non-zero); one warp per 32 consecutive k-mers
__global__ void __launch_bounds__(256) km_bitmap_kernel(const uint32_t* __restrict__ table, uint32_t* __restrict__ bitmap) {
    const size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint4* e = reinterpret_cast<const uint4*>(table + x * kKmWords);
    const uint4 a = e[0], b = e[1];
    const bool nz = (a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w) != 0u;
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, nz);
    if ((threadIdx.x & 31) == 0) bitmap[x >> 5] = bal;
}

struct KmShared {
    uint32_t seq[kKmSeqWords];
    uint32_t queue[kKmThreads / 32][kKmQueue];
    uint16_t hitq[kKmThreads / 32][kKmHitQueue]; // bitmap path: positions whose first mask is non-zero, waiting for their gathers
    uint32_t cn[kKmCombos];                      // c (low 16 bits, signed) | n << 16
    long long item;
};

// Middle stage on up to 32 queued (position, combo) entries of one warp, one per lane.
__device__ __forceinline__ void km_drain(const KmShared& sh, const uint16_t* __restrict__ s_q16, int max_n, uint32_t entry, bool active,
                                         long long U0, long long len, unsigned long long combo_base,
                                         unsigned long long* __restrict__ cand, unsigned long long cand_cap, unsigned long long* counters) {
    const int lane = threadIdx.x & 31;
    const uint32_t combo = entry & 255u;
    const int p = (int)(entry >> 8);
    const uint32_t cn = sh.cn[combo];
    const int c = (int)(short)(cn & 0xFFFFu), n = (int)(cn >> 16);
    const long long i = U0 + p - c - kPadFront;                     // window start in sequence coordinates
    bool ok = active && i >= 0 && i + n <= len;
    if (ok) {
        const int rel = p - c + 16 * kKmHaloWords;                  // window start relative to sh.seq[0]
        const int wi = rel >> 4, s2 = 2 * (rel & 15);
        const uint32_t w0 = sh.seq[wi], w1 = sh.seq[wi + 1], w2 = sh.seq[wi + 2];
        const uint32_t x0 = __funnelshift_r(w0, w1, s2), x1 = __funnelshift_r(w1, w2, s2);
        uint32_t x2 = 0, x3 = 0;
        if (n > 32) {
            const uint32_t w3 = sh.seq[wi + 3], w4 = sh.seq[wi + 4];
            x2 = __funnelshift_r(w2, w3, s2); x3 = __funnelshift_r(w3, w4, s2);
        }
        const uint2* row = reinterpret_cast<const uint2*>(s_q16) + (size_t)combo * max_n;
        uint32_t sum = 0;
        for (int k = 0; k < n; ++k) {                               // rows in order of decreasing expected deficit
            const uint2 r = row[k];
            const uint32_t col = ((r.x >> 12) & 15u) | ((r.x >> 24) & 0x30u);
            const uint32_t xs = col < 32u ? (col < 16u ? x0 : x1) : (col < 48u ? x2 : x3);
            const uint32_t b = (xs >> (2u * (col & 15u))) & 3u;
            const uint32_t f = (b & 2u) ? r.y : r.x;
            sum += (f >> (16u * (b & 1u))) & 0xFFFu;
            if (sum > (uint32_t)kKmQ) break;
        }
        ok = sum <= (uint32_t)kKmQ;
    }
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, ok);
    if (bal) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&counters[0], (unsigned long long)__popc(bal));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (ok) {
            const unsigned long long slot = base + (unsigned)__popc(bal & ((1u << lane) - 1u));
            if (slot < cand_cap) cand[slot] = ((combo_base + combo) << 40) | (unsigned long long)(U0 + p);
        }
    }
}

// Expands the set bits of the 32 lanes' masks (position p per lane) into the warp's queue, draining it through the
// middle stage whenever 32 entries are queued.  All 32 lanes must call it.
__device__ __forceinline__ void km_expand(Mask& m, uint32_t p, uint32_t* q, int& qcnt, const KmShared& sh, const uint16_t* __restrict__ s_q16, int max_n,
                                          long long U0, long long len, unsigned long long combo_base,
                                          unsigned long long* __restrict__ cand, unsigned long long cand_cap, unsigned long long* counters) {
    const int lane = threadIdx.x & 31;
    uint32_t any = mask_any(m);
    unsigned bal = __ballot_sync(0xFFFFFFFFu, any != 0u);
    while (bal) {
        if (any) {
            uint32_t bit;
            if (m.v[0]) { bit = __ffs((int)m.v[0]) - 1; m.v[0] &= m.v[0] - 1; }
            else if (m.v[1]) { bit = 32 + __ffs((int)m.v[1]) - 1; m.v[1] &= m.v[1] - 1; }
            else if (m.v[2]) { bit = 64 + __ffs((int)m.v[2]) - 1; m.v[2] &= m.v[2] - 1; }
            else if (m.v[3]) { bit = 96 + __ffs((int)m.v[3]) - 1; m.v[3] &= m.v[3] - 1; }
            else if (m.v[4]) { bit = 128 + __ffs((int)m.v[4]) - 1; m.v[4] &= m.v[4] - 1; }
            else if (m.v[5]) { bit = 160 + __ffs((int)m.v[5]) - 1; m.v[5] &= m.v[5] - 1; }
            else if (m.v[6]) { bit = 192 + __ffs((int)m.v[6]) - 1; m.v[6] &= m.v[6] - 1; }
            else { bit = 224 + __ffs((int)m.v[7]) - 1; m.v[7] &= m.v[7] - 1; }
            q[qcnt + __popc(bal & ((1u << lane) - 1u))] = (p << 8) | bit;
            any = mask_any(m);
        }
        qcnt += __popc(bal);
        __syncwarp();
        if (qcnt >= 32) {
            const uint32_t e = q[lane];
            const uint32_t rest = (lane + 32 < qcnt) ? q[lane + 32] : 0u;
            __syncwarp();
            if (lane + 32 < qcnt) q[lane] = rest;
            qcnt -= 32;
            km_drain(sh, s_q16, max_n, e, true, U0, len, combo_base, cand, cand_cap, counters);
            __syncwarp();
        }
        bal = __ballot_sync(0xFFFFFFFFu, any != 0u);
    }
}

// One work item of a block WITHOUT a first-table bitmap: per iteration a thread takes two adjacent positions (the
// same three staged 2-bit words serve both), gathers the first mask, the next ones only while the AND is non-zero,
// expands surviving bits into the warp's queue and drains it 32 entries at a time.
// (Software-pipelined variants — gathers of the next stage / the next batch of bitmap hits issued before the
// current masks are expanded — were measured 1.5-2x SLOWER: a third live mask, or two masks live across the middle
// stage, no longer fit 64 registers and the spills go through the very L1 path that bounds the kernel.)
template <int NT>
__device__ __forceinline__ void km_item(KmShared& sh, const uint16_t* __restrict__ s_q16, const DevKmBlock& db, const uint8_t* __restrict__ tab,
                                        long long U0, int iters, long long len, unsigned long long combo_base,
                                        unsigned long long* __restrict__ cand, unsigned long long cand_cap, unsigned long long* counters) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* q = sh.queue[warp];
    int qcnt = 0;
    const uint64_t pol = l2_keep_policy();
    const uint8_t* tp[NT];
    int tsh[NT];
    uint32_t tmk[NT];
#pragma unroll
    for (int s = 0; s < NT; ++s) { tp[s] = tab + db.table_off[s]; tsh[s] = db.shift[s]; tmk[s] = db.kmask[s]; }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        Mask m[2];
        unsigned long long x[2];
        {
            const int p = it * kKmIterPos + warp * 64 + 2 * lane;
            const int wi = (p >> 4) + kKmHaloWords, s2 = 2 * (p & 15);
            const uint32_t w0 = sh.seq[wi], w1 = sh.seq[wi + 1], w2 = sh.seq[wi + 2];
            x[0] = (unsigned long long)__funnelshift_r(w0, w1, s2) | ((unsigned long long)__funnelshift_r(w1, w2, s2) << 32);
            x[1] = (unsigned long long)__funnelshift_r(w0, w1, s2 + 2) | ((unsigned long long)__funnelshift_r(w1, w2, s2 + 2) << 32);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) m[j] = ldg_mask(tp[0] + (size_t)((uint32_t)(x[j] >> tsh[0]) & tmk[0]) * 32, pol);
#pragma unroll
        for (int s = 1; s < NT; ++s) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (mask_any(m[j])) {
                    const Mask t = ldg_mask(tp[s] + (size_t)((uint32_t)(x[j] >> tsh[s]) & tmk[s]) * 32, pol);
#pragma unroll
                    for (int w = 0; w < 8; ++w) m[j].v[w] &= t.v[w];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
            km_expand(m[j], (uint32_t)(it * kKmIterPos + warp * 64 + 2 * lane + j), q, qcnt, sh, s_q16, db.max_n, U0, len, combo_base, cand, cand_cap, counters);
    }
    if (qcnt > 0) {
        const uint32_t e = lane < qcnt ? q[lane] : 0u;
        km_drain(sh, s_q16, db.max_n, e, lane < qcnt, U0, len, combo_base, cand, cand_cap, counters);
    }
}

// Gathers, ANDs and expands the masks of up to 32 queued positions (one per lane; lanes >= count idle).
template <int NT>
__device__ __forceinline__ void km_hits(KmShared& sh, const uint16_t* __restrict__ s_q16, const DevKmBlock& db, const uint8_t* const (&tp)[NT],
                                        const int (&tsh)[NT], const uint32_t (&tmk)[NT], uint64_t pol, uint32_t p, bool active,
                                        uint32_t* q, int& qcnt, long long U0, long long len, unsigned long long combo_base,
                                        unsigned long long* __restrict__ cand, unsigned long long cand_cap, unsigned long long* counters) {
    Mask m;
#pragma unroll
    for (int w = 0; w < 8; ++w) m.v[w] = 0u;
    if (active) {
        const int wi = (int)(p >> 4) + kKmHaloWords, s2 = 2 * (int)(p & 15u);
        const uint32_t w0 = sh.seq[wi], w1 = sh.seq[wi + 1], w2 = sh.seq[wi + 2];
        const unsigned long long x = (unsigned long long)__funnelshift_r(w0, w1, s2) | ((unsigned long long)__funnelshift_r(w1, w2, s2) << 32);
        m = ldg_mask(tp[0] + (size_t)((uint32_t)(x >> tsh[0]) & tmk[0]) * 32, pol);
        if (NT > 1) {                                       // the first mask is known to be non-zero: no need to wait for it
            const Mask t = ldg_mask(tp[1] + (size_t)((uint32_t)(x >> tsh[1]) & tmk[1]) * 32, pol);
#pragma unroll
            for (int w = 0; w < 8; ++w) m.v[w] &= t.v[w];
        }
#pragma unroll
        for (int s = 2; s < NT; ++s) {
            if (mask_any(m)) {
                const Mask t = ldg_mask(tp[s] + (size_t)((uint32_t)(x >> tsh[s]) & tmk[s]) * 32, pol);
#pragma unroll
                for (int w = 0; w < 8; ++w) m.v[w] &= t.v[w];
            }
        }
    }
    km_expand(m, p, q, qcnt, sh, s_q16, db.max_n, U0, len, combo_base, cand, cand_cap, counters);
}

// Work item of a block with a first-table bitmap, in two phases.  Scan: a thread tests the bitmap for 8 consecutive
// positions (four staged words, two funnel shifts and one shared-memory bit per position) and the positions whose
// bit is set are compacted into the warp's hit queue.  Hits: whenever 32 are queued, one per lane gathers its masks
// (km_hits).  Sparse blocks — most positions of a many-motif scan — so cost ~0.2 instructions per position.
template <int NT>
__device__ __forceinline__ void km_item_bitmap(KmShared& sh, const uint16_t* __restrict__ s_q16, const uint32_t* __restrict__ s_bm, const DevKmBlock& db,
                                               const uint8_t* __restrict__ tab, long long U0, int iters, long long len, unsigned long long combo_base,
                                               unsigned long long* __restrict__ cand, unsigned long long cand_cap, unsigned long long* counters) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* q = sh.queue[warp];
    uint16_t* hq = sh.hitq[warp];
    int qcnt = 0, hcnt = 0;
    const uint64_t pol = l2_keep_policy();
    const uint8_t* tp[NT];
    int tsh[NT];
    uint32_t tmk[NT];
#pragma unroll
    for (int s = 0; s < NT; ++s) { tp[s] = tab + db.table_off[s]; tsh[s] = db.shift[s]; tmk[s] = db.kmask[s]; }
    const int per_warp = iters * (kKmIterPos / (kKmThreads / 32));      // 64 * iters positions, contiguous
    const int kpt = per_warp >= 256 ? 8 : per_warp / 32;                // positions per thread and scan step (2, 4 or 8)
    const int nscan = per_warp / (32 * kpt);
    const int sh0 = tsh[0];
    const uint32_t mk0 = tmk[0];
#pragma unroll 1
    for (int si = 0; si < nscan; ++si) {
        const int pb = (si * (kKmThreads / 32) + warp) * 32 * kpt + lane * kpt;    // scan steps interleaved over the warps; (pb & 15) + kpt <= 16
        const int wi = (pb >> 4) + kKmHaloWords;
        const uint32_t w0 = sh.seq[wi], w1 = sh.seq[wi + 1], w2 = sh.seq[wi + 2], w3 = sh.seq[wi + 3];
        const int s0 = 2 * (pb & 15);
        uint32_t hm = 0;                                                // bit k: position pb + k has a non-zero first mask
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (k < kpt) {
                // bases pb+k .. pb+k+47: enough for the first table's k-mer at any offset (shift + 2L <= 64 bits, plan_km.h)
                const int s2 = s0 + 2 * k;
                const uint32_t lo = __funnelshift_r(w0, w1, s2), mid = __funnelshift_r(w1, w2, s2), hi = __funnelshift_r(w2, w3, s2);
                const uint32_t i0 = (sh0 < 32 ? __funnelshift_r(lo, mid, sh0) : __funnelshift_r(mid, hi, sh0 - 32)) & mk0;
                hm |= ((s_bm[i0 >> 5] >> (i0 & 31u)) & 1u) << k;
            }
        }
        // compaction: per round every thread that still has a hit queues its lowest one
        unsigned bal = __ballot_sync(0xFFFFFFFFu, hm != 0u);
        while (bal) {
            if (hm) {
                hq[hcnt + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)(pb + __ffs((int)hm) - 1);
                hm &= hm - 1u;
            }
            hcnt += __popc(bal);
            __syncwarp();
            if (hcnt >= 32) {
                const uint32_t p = hq[lane];
                const uint32_t rest = (lane + 32 < hcnt) ? hq[lane + 32] : 0u;
                __syncwarp();
                if (lane + 32 < hcnt) hq[lane] = (uint16_t)rest;
                hcnt -= 32;
                km_hits<NT>(sh, s_q16, db, tp, tsh, tmk, pol, p, true, q, qcnt, U0, len, combo_base, cand, cand_cap, counters);
                __syncwarp();
            }
            bal = __ballot_sync(0xFFFFFFFFu, hm != 0u);
        }
    }
    if (hcnt > 0) {
        const uint32_t p = lane < hcnt ? hq[lane] : 0u;
        km_hits<NT>(sh, s_q16, db, tp, tsh, tmk, pol, p, lane < hcnt, q, qcnt, U0, len, combo_base, cand, cand_cap, counters);
    }
    if (qcnt > 0) {
        const uint32_t e = lane < qcnt ? q[lane] : 0u;
        km_drain(sh, s_q16, db.max_n, e, lane < qcnt, U0, len, combo_base, cand, cand_cap, counters);
    }
}

__global__ void __launch_bounds__(kKmThreads, 1)
km_prefilter_kernel(const uint32_t* __restrict__ seq2, long long anchor0, long long n_anchor, long long total_items, int iters,
                    const DevKmBlock* __restrict__ blocks, const uint8_t* __restrict__ tables, const uint16_t* __restrict__ q16,
                    const DevCombo* __restrict__ combos, long long len, unsigned long long* __restrict__ cand, unsigned long long cand_cap,
                    unsigned long long* counters, unsigned long long* work_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ KmShared sh;
    uint16_t* s_q16 = reinterpret_cast<uint16_t*>(smem_raw);
    const int tid = threadIdx.x;
    const int item_pos = iters * kKmIterPos;
    const long long items_per_block = n_anchor / item_pos;
    int cur_blk = -1;
    DevKmBlock db = blocks[0];
    const uint32_t* s_bm = nullptr;
    for (;;) {
        if (tid == 0) sh.item = (long long)atomicAdd(work_counter, 1ull);
        __syncthreads();
        const long long item = sh.item;
        if (item >= total_items) break;
        const int blk = (int)(item / items_per_block);                 // block-major: tables change rarely
        const long long U0 = anchor0 + (item - (long long)blk * items_per_block) * item_pos;
        if (blk != cur_blk) {
            db = blocks[blk];
            const uint4* src = reinterpret_cast<const uint4*>(q16 + db.q16_off);
            uint4* dst = reinterpret_cast<uint4*>(s_q16);
            const int n16 = kKmCombos * db.max_n * 8 / 16;
            for (int i = tid; i < n16; i += kKmThreads) dst[i] = __ldg(src + i);
            s_bm = reinterpret_cast<const uint32_t*>(smem_raw + (size_t)n16 * 16);       // bitmap right behind the rows
            if (db.bitmap_words > 0) {
                const uint4* bsrc = reinterpret_cast<const uint4*>(tables + db.bitmap_off);
                uint4* bdst = reinterpret_cast<uint4*>(smem_raw + (size_t)n16 * 16);
                for (int i = tid; i < db.bitmap_words / 4; i += kKmThreads) bdst[i] = __ldg(bsrc + i);
            }
            if (tid < kKmCombos) {
                const DevCombo dc = combos[(size_t)blk * kKmCombos + tid];
                sh.cn[tid] = ((uint32_t)dc.c & 0xFFFFu) | ((uint32_t)dc.n << 16);
            }
            cur_blk = blk;
        }
        {   // stage the item's 2-bit words (64 bases of history in front; the allocation has one front word)
            const long long w0 = (U0 >> 4) - kKmHaloWords;
            const int nw = kKmHaloWords + item_pos / 16 + 5;
            for (int i = tid; i < nw; i += kKmThreads) sh.seq[i] = (w0 + i >= -1) ? __ldg(seq2 + w0 + i) : 0u;
        }
        __syncthreads();
        const uint8_t* tab = tables;
        const unsigned long long combo_base = (unsigned long long)blk * kKmCombos;
#define KM_RUN(NT_) \
        if (db.bitmap_words > 0) km_item_bitmap<NT_>(sh, s_q16, s_bm, db, tab, U0, iters, len, combo_base, cand, cand_cap, counters); \
        else km_item<NT_>(sh, s_q16, db, tab, U0, iters, len, combo_base, cand, cand_cap, counters);
        switch (db.nt) {
            case 1: KM_RUN(1) break;
            case 2: KM_RUN(2) break;
            case 3: KM_RUN(3) break;
            default: KM_RUN(4) break;
        }
#undef KM_RUN
        __syncthreads();   // everyone is done with sh.seq / s_q16 / sh.item
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Second generation of the kernel (PWMSCAN_KM_KERNEL=2): the same first stage, table layout, middle stage and candidate
// format, but organised around WARPS instead of CTAs:
//   * a work unit is (block, kKm2UnitPos positions) and belongs to one warp, which draws it from the global counter
//     (the next one is requested while the current one is processed), stages the unit's 2-bit words in its own slice
//     of shared memory and runs it to the end — no CTA-wide barrier per item, no waiting for the slowest warp;
//   * the CTA only synchronises when the block in shared memory (middle-stage rows, bitmap) has to change: a warp that
//     has drawn a unit of a later block waits at the barrier until every warp has, the smallest wanted block is loaded;
//   * bitmap scan: 16 positions per thread from one 64-bit window (two shifts per position instead of five);
//   * blocks whose single table covers every column of every motif (short motifs) skip the middle stage: a surviving
//     bit already is the whole-window test.
constexpr int kKm2UnitPos = 2048;
constexpr int kKm2SmallUnit = 512;               // the last units of every block (and the unit size of short inputs)
constexpr int kKm2SeqWords = kKmHaloWords + kKm2UnitPos / 16 + 5;

struct Km2Shared {
    uint32_t seq[kKmThreads / 32][kKm2SeqWords];
    uint32_t queue[kKmThreads / 32][kKmQueue];
    uint16_t hitq[kKmThreads / 32][kKmHitQueue];
    uint32_t cn[kKmCombos];
    int want;
};

struct Km2Ctx {                     // what the per-warp helpers need (all warp-uniform)
    const uint32_t* sseq;           // the unit's staged words (halo in front)
    const uint32_t* cn;
    const uint16_t* q16;
    uint32_t* q;
    int max_n, direct;
    long long U0, len;
    unsigned long long combo_base;
    unsigned long long* cand;
    unsigned long long cand_cap;
    unsigned long long* counters;
};

__device__ __forceinline__ void km2_drain(const Km2Ctx& c, uint32_t entry, bool active) {
    const int lane = threadIdx.x & 31;
    const uint32_t combo = entry & 255u;
    const int p = (int)(entry >> 8);
    const uint32_t cn = c.cn[combo];
    const int cc = (int)(short)(cn & 0xFFFFu), n = (int)(cn >> 16);
    const long long i = c.U0 + p - cc - kPadFront;                   // window start in sequence coordinates
    bool ok = active && i >= 0 && i + n <= c.len;
    if (ok && !c.direct) {
        const int rel = p - cc + 16 * kKmHaloWords;                  // window start relative to sseq[0]
        const int wi = rel >> 4, s2 = 2 * (rel & 15);
        const uint32_t w0 = c.sseq[wi], w1 = c.sseq[wi + 1], w2 = c.sseq[wi + 2];
        const uint32_t x0 = __funnelshift_r(w0, w1, s2), x1 = __funnelshift_r(w1, w2, s2);
        uint32_t x2 = 0, x3 = 0;
        if (n > 32) {
            const uint32_t w3 = c.sseq[wi + 3], w4 = c.sseq[wi + 4];
            x2 = __funnelshift_r(w2, w3, s2); x3 = __funnelshift_r(w3, w4, s2);
        }
        const uint2* row = reinterpret_cast<const uint2*>(c.q16) + (size_t)combo * c.max_n;
        uint32_t sum = 0;
        for (int k = 0; k < n; ++k) {                               // rows in order of decreasing expected deficit
            const uint2 r = row[k];
            const uint32_t col = ((r.x >> 12) & 15u) | ((r.x >> 24) & 0x30u);
            const uint32_t xs = col < 32u ? (col < 16u ? x0 : x1) : (col < 48u ? x2 : x3);
            const uint32_t b = (xs >> (2u * (col & 15u))) & 3u;
            const uint32_t f = (b & 2u) ? r.y : r.x;
            sum += (f >> (16u * (b & 1u))) & 0xFFFu;
            if (sum > (uint32_t)kKmQ) break;
        }
        ok = sum <= (uint32_t)kKmQ;
    }
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, ok);
    if (bal) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&c.counters[0], (unsigned long long)__popc(bal));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (ok) {
            const unsigned long long slot = base + (unsigned)__popc(bal & ((1u << lane) - 1u));
            if (slot < c.cand_cap) c.cand[slot] = ((c.combo_base + combo) << 40) | (unsigned long long)(c.U0 + p);
        }
    }
}

// as km_expand, on the warp's own queue
__device__ __forceinline__ void km2_expand(const Km2Ctx& c, Mask& m, uint32_t p, int& qcnt) {
    const int lane = threadIdx.x & 31;
    uint32_t any = mask_any(m);
    unsigned bal = __ballot_sync(0xFFFFFFFFu, any != 0u);
    while (bal) {
        if (any) {
            uint32_t bit;
            if (m.v[0]) { bit = __ffs((int)m.v[0]) - 1; m.v[0] &= m.v[0] - 1; }
            else if (m.v[1]) { bit = 32 + __ffs((int)m.v[1]) - 1; m.v[1] &= m.v[1] - 1; }
            else if (m.v[2]) { bit = 64 + __ffs((int)m.v[2]) - 1; m.v[2] &= m.v[2] - 1; }
            else if (m.v[3]) { bit = 96 + __ffs((int)m.v[3]) - 1; m.v[3] &= m.v[3] - 1; }
            else if (m.v[4]) { bit = 128 + __ffs((int)m.v[4]) - 1; m.v[4] &= m.v[4] - 1; }
            else if (m.v[5]) { bit = 160 + __ffs((int)m.v[5]) - 1; m.v[5] &= m.v[5] - 1; }
            else if (m.v[6]) { bit = 192 + __ffs((int)m.v[6]) - 1; m.v[6] &= m.v[6] - 1; }
            else { bit = 224 + __ffs((int)m.v[7]) - 1; m.v[7] &= m.v[7] - 1; }
            c.q[qcnt + __popc(bal & ((1u << lane) - 1u))] = (p << 8) | bit;
            any = mask_any(m);
        }
        qcnt += __popc(bal);
        __syncwarp();
        if (qcnt >= 32) {
            const uint32_t e = c.q[lane];
            const uint32_t rest = (lane + 32 < qcnt) ? c.q[lane + 32] : 0u;
            __syncwarp();
            if (lane + 32 < qcnt) c.q[lane] = rest;
            qcnt -= 32;
            km2_drain(c, e, true);
            __syncwarp();
        }
        bal = __ballot_sync(0xFFFFFFFFu, any != 0u);
    }
}

template <int NT>
__device__ __forceinline__ void km2_hits(const Km2Ctx& c, const uint8_t* const (&tp)[NT], const int (&tsh)[NT], const uint32_t (&tmk)[NT], uint64_t pol,
                                         uint32_t p, bool active, int& qcnt) {
    Mask m;
#pragma unroll
    for (int w = 0; w < 8; ++w) m.v[w] = 0u;
    if (active) {
        const int wi = (int)(p >> 4) + kKmHaloWords, s2 = 2 * (int)(p & 15u);
        const uint32_t w0 = c.sseq[wi], w1 = c.sseq[wi + 1], w2 = c.sseq[wi + 2];
        const unsigned long long x = (unsigned long long)__funnelshift_r(w0, w1, s2) | ((unsigned long long)__funnelshift_r(w1, w2, s2) << 32);
        m = ldg_mask(tp[0] + (size_t)((uint32_t)(x >> tsh[0]) & tmk[0]) * 32, pol);
        if (NT > 1) {                                       // the first mask is known to be non-zero: no need to wait for it
            const Mask t = ldg_mask(tp[1] + (size_t)((uint32_t)(x >> tsh[1]) & tmk[1]) * 32, pol);
#pragma unroll
            for (int w = 0; w < 8; ++w) m.v[w] &= t.v[w];
        }
#pragma unroll
        for (int s = 2; s < NT; ++s) {
            if (mask_any(m)) {
                const Mask t = ldg_mask(tp[s] + (size_t)((uint32_t)(x >> tsh[s]) & tmk[s]) * 32, pol);
#pragma unroll
                for (int w = 0; w < 8; ++w) m.v[w] &= t.v[w];
            }
        }
    }
    km2_expand(c, m, p, qcnt);
}

// One unit of a block with a first-table bitmap.  Scan: a thread tests 16 consecutive positions; the k-mers of all of them
// come out of ONE 64-bit window (the 32 bases from the first k-mer's first base), shifted by 2 bits per position.
template <int NT>
__device__ __forceinline__ void km2_unit_bitmap(const Km2Ctx& c, const uint32_t* __restrict__ s_bm, const DevKmBlock& db, const uint8_t* __restrict__ tab,
                                                uint16_t* hq, int unit_pos) {
    const int lane = threadIdx.x & 31;
    int qcnt = 0, hcnt = 0;
    const uint64_t pol = l2_keep_policy();
    const uint8_t* tp[NT];
    int tsh[NT];
    uint32_t tmk[NT];
#pragma unroll
    for (int s = 0; s < NT; ++s) { tp[s] = tab + db.table_off[s]; tsh[s] = db.shift[s]; tmk[s] = db.kmask[s]; }
    const int off0 = tsh[0] >> 1;                                       // bases between a position and its first table's k-mer
    const uint32_t mk0 = tmk[0];
#pragma unroll 1
    for (int pb0 = 0; pb0 < unit_pos; pb0 += 512) {
        const int pb = pb0 + 16 * lane;
        const int B = pb + off0 + 16 * kKmHaloWords;                    // first base of position pb's k-mer, relative to sseq[0]
        const int wi = B >> 4, sB = 2 * (B & 15);
        const uint32_t w0 = c.sseq[wi], w1 = c.sseq[wi + 1], w2 = c.sseq[wi + 2];
        const uint32_t xlo = __funnelshift_r(w0, w1, sB), xhi = __funnelshift_r(w1, w2, sB);
        uint32_t hm = 0;                                                // bit 15 - k: position pb + k has a non-zero first mask
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const uint32_t i0 = __funnelshift_r(xlo, xhi, 2 * k) & mk0;
            const uint32_t word = s_bm[i0 >> 5];
            hm = __funnelshift_l(word << (~i0 & 31u), hm, 1);           // hm = hm << 1 | bit i0 of word
        }
        // compaction: per round every thread that still has a hit queues one
        unsigned bal = __ballot_sync(0xFFFFFFFFu, hm != 0u);
        while (bal) {
            if (hm) {
                const int b = __ffs((int)hm) - 1;
                hq[hcnt + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)(pb + 15 - b);
                hm &= hm - 1u;
            }
            hcnt += __popc(bal);
            __syncwarp();
            if (hcnt >= 32) {
                const uint32_t p = hq[lane];
                const uint32_t rest = (lane + 32 < hcnt) ? hq[lane + 32] : 0u;
                __syncwarp();
                if (lane + 32 < hcnt) hq[lane] = (uint16_t)rest;
                hcnt -= 32;
                km2_hits<NT>(c, tp, tsh, tmk, pol, p, true, qcnt);
                __syncwarp();
            }
            bal = __ballot_sync(0xFFFFFFFFu, hm != 0u);
        }
    }
    if (hcnt > 0) {
        const uint32_t p = lane < hcnt ? hq[lane] : 0u;
        km2_hits<NT>(c, tp, tsh, tmk, pol, p, lane < hcnt, qcnt);
    }
    if (qcnt > 0) {
        const uint32_t e = lane < qcnt ? c.q[lane] : 0u;
        km2_drain(c, e, lane < qcnt);
    }
}

// One unit of a block without a bitmap: two adjacent positions per thread and iteration (km_item's loop).
template <int NT>
__device__ __forceinline__ void km2_unit_plain(const Km2Ctx& c, const DevKmBlock& db, const uint8_t* __restrict__ tab, int unit_pos) {
    const int lane = threadIdx.x & 31;
    int qcnt = 0;
    const uint64_t pol = l2_keep_policy();
    const uint8_t* tp[NT];
    int tsh[NT];
    uint32_t tmk[NT];
#pragma unroll
    for (int s = 0; s < NT; ++s) { tp[s] = tab + db.table_off[s]; tsh[s] = db.shift[s]; tmk[s] = db.kmask[s]; }
#pragma unroll 1
    for (int pb0 = 0; pb0 < unit_pos; pb0 += 64) {
        Mask m[2];
        unsigned long long x[2];
        const int p = pb0 + 2 * lane;
        {
            const int wi = (p >> 4) + kKmHaloWords, s2 = 2 * (p & 15);
            const uint32_t w0 = c.sseq[wi], w1 = c.sseq[wi + 1], w2 = c.sseq[wi + 2];
            x[0] = (unsigned long long)__funnelshift_r(w0, w1, s2) | ((unsigned long long)__funnelshift_r(w1, w2, s2) << 32);
            x[1] = (unsigned long long)__funnelshift_r(w0, w1, s2 + 2) | ((unsigned long long)__funnelshift_r(w1, w2, s2 + 2) << 32);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) m[j] = ldg_mask(tp[0] + (size_t)((uint32_t)(x[j] >> tsh[0]) & tmk[0]) * 32, pol);
#pragma unroll
        for (int s = 1; s < NT; ++s) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (mask_any(m[j])) {
                    const Mask t = ldg_mask(tp[s] + (size_t)((uint32_t)(x[j] >> tsh[s]) & tmk[s]) * 32, pol);
#pragma unroll
                    for (int w = 0; w < 8; ++w) m[j].v[w] &= t.v[w];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) km2_expand(c, m[j], (uint32_t)(p + j), qcnt);
    }
    if (qcnt > 0) {
        const uint32_t e = lane < qcnt ? c.q[lane] : 0u;
        km2_drain(c, e, lane < qcnt);
    }
}

__global__ void __launch_bounds__(kKmThreads, 1)
km_prefilter_kernel_v2(const uint32_t* __restrict__ seq2, long long anchor0, long long n_big, long long n_small, int unit_pos, int nblocks,
                       const DevKmBlock* __restrict__ blocks, const uint8_t* __restrict__ tables, const uint16_t* __restrict__ q16,
                       const DevCombo* __restrict__ combos, long long len, unsigned long long* __restrict__ cand, unsigned long long cand_cap,
                       unsigned long long* counters, unsigned long long* work_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ Km2Shared sh;
    uint16_t* s_q16 = reinterpret_cast<uint16_t*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Units of a block: n_big of unit_pos positions, then n_small of kKm2SmallUnit covering the block's last stretch — every
    // change of block is a CTA-wide barrier at which a warp idles for up to one unit, so the last units are short ones.
    const long long units_per_block = n_big + n_small, total_units = units_per_block * nblocks;
    if (tid == 0) sh.want = 0x7FFFFFFF;
    __syncthreads();
    int cur_blk = -1;
    DevKmBlock db = blocks[0];
    const uint32_t* s_bm = nullptr;
    // the warp's next unit is always requested one unit ahead
    long long next_unit = 0;
    if (lane == 0) next_unit = (long long)atomicAdd(work_counter, 1ull);
    for (;;) {
        const long long unit = __shfl_sync(0xFFFFFFFFu, next_unit, 0);
        if (lane == 0 && unit < total_units) next_unit = (long long)atomicAdd(work_counter, 1ull);     // in flight while this unit runs
        const int blk = unit < total_units ? (int)(unit / units_per_block) : nblocks;                    // nblocks = nothing left
        while (blk != cur_blk) {
            // the block in shared memory has to change: wait until no warp needs the current one any more, load the
            // smallest block any warp wants (every warp is here with a block >= that)
            __syncthreads();
            if (lane == 0) atomicMin(&sh.want, blk);
            __syncthreads();
            const int nb = sh.want;
            __syncthreads();
            if (tid == 0) sh.want = 0x7FFFFFFF;
            if (nb >= nblocks) return;                                   // every warp has run out of units
            db = blocks[nb];
            const uint4* src = reinterpret_cast<const uint4*>(q16 + db.q16_off);
            uint4* dst = reinterpret_cast<uint4*>(s_q16);
            const int n16 = kKmCombos * db.max_n * 8 / 16;
            for (int i = tid; i < n16; i += kKmThreads) dst[i] = __ldg(src + i);
            s_bm = reinterpret_cast<const uint32_t*>(smem_raw + (size_t)n16 * 16);       // bitmap right behind the rows
            if (db.bitmap_words > 0) {
                const uint4* bsrc = reinterpret_cast<const uint4*>(tables + db.bitmap_off);
                uint4* bdst = reinterpret_cast<uint4*>(smem_raw + (size_t)n16 * 16);
                for (int i = tid; i < db.bitmap_words / 4; i += kKmThreads) bdst[i] = __ldg(bsrc + i);
            }
            if (tid < kKmCombos) {
                const DevCombo dc = combos[(size_t)nb * kKmCombos + tid];
                sh.cn[tid] = ((uint32_t)dc.c & 0xFFFFu) | ((uint32_t)dc.n << 16);
            }
            cur_blk = nb;
            __syncthreads();
        }
        const long long ub = unit - (long long)blk * units_per_block;
        const int ulen = ub < n_big ? unit_pos : kKm2SmallUnit;
        const long long U0 = anchor0 + (ub < n_big ? ub * unit_pos : n_big * unit_pos + (ub - n_big) * kKm2SmallUnit);
        uint32_t* sseq = sh.seq[warp];
        {   // stage the unit's 2-bit words (64 bases of history in front; the allocation has one front word)
            const long long w0 = (U0 >> 4) - kKmHaloWords;
            const int nw = kKmHaloWords + ulen / 16 + 5;
            for (int i = lane; i < nw; i += 32) sseq[i] = (w0 + i >= -1) ? __ldg(seq2 + w0 + i) : 0u;
        }
        __syncwarp();
        Km2Ctx c;
        c.sseq = sseq; c.cn = sh.cn; c.q16 = s_q16; c.q = sh.queue[warp]; c.max_n = db.max_n; c.direct = db.direct;
        c.U0 = U0; c.len = len; c.combo_base = (unsigned long long)blk * kKmCombos; c.cand = cand; c.cand_cap = cand_cap; c.counters = counters;
#define KM2_RUN(NT_) \
        if (db.bitmap_words > 0) km2_unit_bitmap<NT_>(c, s_bm, db, tables, sh.hitq[warp], ulen); \
        else km2_unit_plain<NT_>(c, db, tables, ulen);
        switch (db.nt) {
            case 1: KM2_RUN(1) break;
            case 2: KM2_RUN(2) break;
            case 3: KM2_RUN(3) break;
            default: KM2_RUN(4) break;
        }
#undef KM2_RUN
        __syncwarp();
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Third generation (PWMSCAN_KM_KERNEL=3; measured and NOT the default): the warp-organised unit code of the second, but not
// persistent.  A CTA is one (block, range of positions) item of the grid: it loads the block's middle-stage rows / bitmap
// once, its 32 warps draw the units of the range from a counter in shared memory (no global atomic, no block-change
// barrier) and the CTA retires, so that the exact replay / ordering kernels of the previous item (high-priority stream) get
// SMs within microseconds and the next item's CTAs fill the tail of this one (launches of different lanes not chained).
// Measured on a config-4 chunk: alone 2.5 / 1.75 / 1.68 ms with ranges of 1 Mi / 512 Ki / 256 Ki positions (wave
// quantisation: 448 CTAs on 148 SMs; warps idle at the end of every CTA) against 1.36 ms for the persistent kernel, and a
// whole genome 125 / 100 ms against 86 ms: the hardware interleaves the CTAs of the 2-3 launches in flight instead of
// running them first-in first-out, so the mask tables of several motif blocks compete for L2.
__global__ void __launch_bounds__(kKmThreads, 1)
km_prefilter_kernel_v3(const uint32_t* __restrict__ seq2, long long anchor0, long long n_anchor, long long range_pos, int unit_pos, int ranges_per_block,
                       const DevKmBlock* __restrict__ blocks, const uint8_t* __restrict__ tables, const uint16_t* __restrict__ q16,
                       const DevCombo* __restrict__ combos, long long len, unsigned long long* __restrict__ cand, unsigned long long cand_cap,
                       unsigned long long* counters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ Km2Shared sh;
    uint16_t* s_q16 = reinterpret_cast<uint16_t*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int blk = (int)(blockIdx.x / (unsigned)ranges_per_block);
    const long long R0 = (long long)(blockIdx.x - (unsigned)blk * (unsigned)ranges_per_block) * range_pos;
    const DevKmBlock db = blocks[blk];
    const int n16 = kKmCombos * db.max_n * 8 / 16;
    {
        const uint4* src = reinterpret_cast<const uint4*>(q16 + db.q16_off);
        uint4* dst = reinterpret_cast<uint4*>(s_q16);
        for (int i = tid; i < n16; i += kKmThreads) dst[i] = __ldg(src + i);
        if (db.bitmap_words > 0) {
            const uint4* bsrc = reinterpret_cast<const uint4*>(tables + db.bitmap_off);
            uint4* bdst = reinterpret_cast<uint4*>(smem_raw + (size_t)n16 * 16);       // bitmap right behind the rows
            for (int i = tid; i < db.bitmap_words / 4; i += kKmThreads) bdst[i] = __ldg(bsrc + i);
        }
        if (tid < kKmCombos) {
            const DevCombo dc = combos[(size_t)blk * kKmCombos + tid];
            sh.cn[tid] = ((uint32_t)dc.c & 0xFFFFu) | ((uint32_t)dc.n << 16);
        }
        if (tid == 0) sh.want = 0;                                             // next unit of this CTA's range
    }
    __syncthreads();
    const uint32_t* s_bm = reinterpret_cast<const uint32_t*>(smem_raw + (size_t)n16 * 16);
    const long long left = n_anchor - R0;
    const int nunits = (int)((left < range_pos ? left : range_pos) / unit_pos);
    uint32_t* sseq = sh.seq[warp];
    Km2Ctx c;
    c.sseq = sseq; c.cn = sh.cn; c.q16 = s_q16; c.q = sh.queue[warp]; c.max_n = db.max_n; c.direct = db.direct;
    c.len = len; c.combo_base = (unsigned long long)blk * kKmCombos; c.cand = cand; c.cand_cap = cand_cap; c.counters = counters;
    for (;;) {
        int u = 0;
        if (lane == 0) u = atomicAdd(&sh.want, 1);
        u = __shfl_sync(0xFFFFFFFFu, u, 0);
        if (u >= nunits) break;
        const long long U0 = anchor0 + R0 + (long long)u * unit_pos;
        {   // stage the unit's 2-bit words (64 bases of history in front; the allocation has one front word)
            const long long w0 = (U0 >> 4) - kKmHaloWords;
            const int nw = kKmHaloWords + unit_pos / 16 + 5;
            for (int i = lane; i < nw; i += 32) sseq[i] = (w0 + i >= -1) ? __ldg(seq2 + w0 + i) : 0u;
        }
        __syncwarp();
        c.U0 = U0;
#define KM3_RUN(NT_) \
        if (db.bitmap_words > 0) km2_unit_bitmap<NT_>(c, s_bm, db, tables, sh.hitq[warp], unit_pos); \
        else km2_unit_plain<NT_>(c, db, tables, unit_pos);
        switch (db.nt) {
            case 1: KM3_RUN(1) break;
            case 2: KM3_RUN(2) break;
            case 3: KM3_RUN(3) break;
            default: KM3_RUN(4) break;
        }
#undef KM3_RUN
        __syncwarp();
    }
}

}  // namespace

namespace pwmscan {

// Bitmap of one mask table: bit x = mask of k-mer x is non-zero (stream-ordered).
int km_build_bitmap(pwmscan_ctx* ctx, const uint8_t* d_table, int L, uint8_t* d_bitmap) {
    km_bitmap_kernel<<<(unsigned)(((size_t)1 << (2 * L)) / 256), 256, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(d_table), reinterpret_cast<uint32_t*>(d_bitmap));
    PWM_CUDA(cudaGetLastError());
    return PWMSCAN_OK;
}

// Builds one mask table (4^L entries) on the device (stream-ordered).  d_wdef: [256][kKmMaxL][4], d_dsafe: [256].
int km_build_table(pwmscan_ctx* ctx, const double* d_wdef, const double* d_dsafe, int L, uint8_t* d_table) {
    const int smem = (kKmCombos * kKmMaxL * 4 + kKmCombos) * 8;
    if (!ctx->km_build_attr_set) {
        PWM_CUDA(cudaFuncSetAttribute(km_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        ctx->km_build_attr_set = true;
    }
    km_build_kernel<<<(unsigned)(((size_t)1 << (2 * L)) / 256), 256, smem, ctx->stream>>>(d_wdef, d_dsafe, L, reinterpret_cast<uint32_t*>(d_table));
    PWM_CUDA(cudaGetLastError());
    return PWMSCAN_OK;
}

static int km_kernel_choice() {
    static const int which = std::getenv("PWMSCAN_KM_KERNEL") ? std::atoi(std::getenv("PWMSCAN_KM_KERNEL")) : 2;    // 2 = persistent, warp-organised (default); 1 = persistent, CTA-organised; 3 = not persistent
    return which;
}
bool km_prefilter_is_persistent() { return km_kernel_choice() != 3; }

int launch_km_prefilter(pwmscan_ctx* ctx, ScanLane* lane, cudaStream_t stream, const uint32_t* d_seq2, long long a0, long long a1, long long len, int nblocks, size_t smem_bytes,
                        const DevKmBlock* d_blocks, const uint8_t* d_tables, const uint16_t* d_q16, const DevCombo* d_combos,
                        unsigned long long* d_cand, unsigned long long cand_cap, unsigned long long* work_counter) {
    if (nblocks == 0 || a1 <= a0) return PWMSCAN_OK;
    if (!ctx->km_attr_set) {
        PWM_CUDA(cudaFuncSetAttribute(km_prefilter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kKmSmemBudget));
        PWM_CUDA(cudaFuncSetAttribute(km_prefilter_kernel_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kKmSmemBudget));
        PWM_CUDA(cudaFuncSetAttribute(km_prefilter_kernel_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kKmSmemBudget));
        ctx->km_attr_set = true;
    }
    const int which = km_kernel_choice();
    if (which == 3) {
        // CTA = (block, range of positions); ranges of 1 Mi positions, smaller for short inputs so that every SM gets a few CTAs;
        // warp-owned units of 2048 positions, at least two per warp
        static const long long range_max = std::getenv("PWMSCAN_KM_RANGE") ? std::max(16384LL, std::atoll(std::getenv("PWMSCAN_KM_RANGE"))) : (1LL << 20);   // experiments
        long long range_pos = range_max;
        while (range_pos > 16384 && (long long)nblocks * ((a1 - a0 + range_pos - 1) / range_pos) < 2LL * ctx->sm_count) range_pos >>= 1;
        int unit_pos = kKm2UnitPos;
        while (unit_pos > 512 && range_pos / unit_pos < 2 * (kKmThreads / 32)) unit_pos >>= 1;
        const long long ranges = (a1 - a0 + range_pos - 1) / range_pos;
        if (ranges * nblocks > 0x7FFFFFFFLL) return set_err(ctx, PWMSCAN_ERANGE, "pwmscan_scan: sequence x motif blocks exceed the first stage's grid");
        km_prefilter_kernel_v3<<<(unsigned)(ranges * nblocks), kKmThreads, smem_bytes, stream>>>(d_seq2, a0, a1 - a0, range_pos, unit_pos, (int)ranges, d_blocks, d_tables, d_q16,
                                                                                                 d_combos, len, d_cand, cand_cap, lane->d_counters);
        PWM_CUDA(cudaGetLastError());
        return PWMSCAN_OK;
    }
    if (which == 2) {
        // warp-owned units of 2048 positions (512 for short inputs, so that every warp of every SM gets some)
        // (measured on config-4 chunks of 16 / 32 Mbp: units of 2048 beat 1024 and 512 by 6 % / 33 % and 12 % / 40 % — at the end of
        //  every unit the warp flushes its partly filled hit and survivor queues, one gather latency each)
        int unit_pos = kKm2UnitPos;
        while (unit_pos > 512 && (a1 - a0) / unit_pos * nblocks < 4LL * ctx->sm_count * (kKmThreads / 32)) unit_pos >>= 1;
        // the last eighth of a block's positions (at most two short units per warp of the device) goes out in short units
        long long tail_pos = 0;
        if (unit_pos > kKm2SmallUnit) {
            static const bool taper = std::getenv("PWMSCAN_KM_NO_TAPER") == nullptr;      // experiments
            tail_pos = std::min<long long>((a1 - a0) / 8, 2LL * ctx->sm_count * (kKmThreads / 32) * kKm2SmallUnit) / unit_pos * unit_pos;
            if (!taper) tail_pos = 0;
        }
        const long long n_big = (a1 - a0 - tail_pos) / unit_pos, n_small = tail_pos / kKm2SmallUnit;
        const long long units = (n_big + n_small) * nblocks;
        const int grid = (int)std::min<long long>(ctx->sm_count, (units + kKmThreads / 32 - 1) / (kKmThreads / 32));
        km_prefilter_kernel_v2<<<grid, kKmThreads, smem_bytes, stream>>>(d_seq2, a0, n_big, n_small, unit_pos, nblocks, d_blocks, d_tables, d_q16, d_combos,
                                                                               len, d_cand, cand_cap, lane->d_counters, work_counter);
        PWM_CUDA(cudaGetLastError());
        return PWMSCAN_OK;
    }
    int iters = kKmMaxIters;                                  // short inputs: smaller items so that every SM gets several
    while (iters > 1 && (a1 - a0) / (iters * kKmIterPos) * nblocks < 8LL * ctx->sm_count) iters >>= 1;
    const long long items = (a1 - a0) / (iters * kKmIterPos) * nblocks;
    const int grid = (int)std::min<long long>(ctx->sm_count, items);
    km_prefilter_kernel<<<grid, kKmThreads, smem_bytes, stream>>>(d_seq2, a0, a1 - a0, items, iters, d_blocks, d_tables, d_q16, d_combos, len,
                                                                                    d_cand, cand_cap, lane->d_counters, work_counter);
    PWM_CUDA(cudaGetLastError());
    return PWMSCAN_OK;
}

}  // namespace pwmscan
