// scan_core.cuh — per-lane body of the prefilter kernel and the exact fp64 replay, written as
// __host__ __device__ code so that the CPU unit test (tests/emul) can run the very same index
// arithmetic lane by lane.  On the device `tab` is the shared-memory table base; on the host it is
// a plain array with the same layout [slot][kmer][lane] of 32-bit words (4 combos, one byte each).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PWM_HD __host__ __device__ __forceinline__
#else
#define PWM_HD inline
#endif

namespace pwmscan {

PWM_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, int s) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    return s == 0 ? lo : ((lo >> s) | (hi << (32 - s)));
#endif
}

// Table handle (template parameter Tab): on the device a 32-bit address in the shared window (the
// per-lane base is one register and a table read is `LDS R, [R + imm]`), on the host a pointer.
PWM_HD uint32_t lds32(const unsigned char* p) { return *reinterpret_cast<const uint32_t*>(p); }
#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
#endif
PWM_HD uint32_t byte_of(uint32_t w, int q) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(w, 0u, 0x4440u + (uint32_t)q);
#else
    return (w >> (8 * q)) & 0xFFu;
#endif
}

// byte `b` (0..15) of the 16-byte window Y[0..3]
PWM_HD uint32_t window_byte(uint32_t y0, uint32_t y1, uint32_t y2, uint32_t y3, int b) {
    const uint32_t w = b < 8 ? (b < 4 ? y0 : y1) : (b < 12 ? y2 : y3);
    return (w >> (8 * (b & 3))) & 0xFFu;
}

// Survivor of the first stage: add the remaining slots one by one (dead byte lanes are parked at
// 0x80 so nothing ever carries into a neighbour), then report the byte lanes still alive.
// Returns the alive mask (bit 8q+7 set <=> combo 4*lane+q is a candidate).
// tab_lane = table base + 4*lane.
template <class Tab>
PWM_HD uint32_t second_stage(uint32_t acc, uint32_t y0, uint32_t y1, uint32_t y2, uint32_t y3, int byte0 /* 4+j-nl */,
                             int nl, int s1, int nslot, Tab tab_lane) {
    for (int s = nl - 1; s >= 0; --s) {                    // slots left of the first stage
        const uint32_t m = acc & 0x80808080u;
        if (m == 0x80808080u) return 0u;
        acc &= ~(m - (m >> 7));
        const uint32_t kmer = window_byte(y0, y1, y2, y3, byte0 + s);
        acc += lds32(tab_lane + ((uint32_t)s * 32768u + kmer * 128u));
    }
    for (int s = nl + s1; s < nslot; ++s) {                // slots right of it
        const uint32_t m = acc & 0x80808080u;
        if (m == 0x80808080u) return 0u;
        acc &= ~(m - (m >> 7));
        const uint32_t kmer = window_byte(y0, y1, y2, y3, byte0 + s);
        acc += lds32(tab_lane + ((uint32_t)s * 32768u + kmer * 128u));
    }
    return ~acc & 0x80808080u;
}

// 32 consecutive anchors u = U .. U+31 (U a multiple of 16) for one lane.
// X[q] = packed word (U/16 - 1 + q), q = 0..4 (16 bases per word, base t at bits 2t..2t+1).
// emit(q, u) is called for every candidate byte lane q at anchor u.
//
// The first stage is straight-line code: for each of the 4 phases r (anchors U+4j+r, j = 0..7) the
// 16-byte window Y is the sequence shifted by r bases, so the 4-mer of slot t at anchor j is the
// plain byte 4+j+t of Y; each table address is computed once (byte extract + multiply-add) and
// shared by S1 anchors.  The 8 accumulators of a phase are tested together (one branch per 8
// anchors); only phases in which some byte lane is still below 128 enter the per-anchor handler.
template <int S1, class Tab, class Emit>
PWM_HD void process32(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t x4, int64_t U, int lane,
                      Tab tab, int nl, int nslot, Emit&& emit) {
    const Tab tab_lane = tab + (uint32_t)lane * 4u;
    const Tab tab1 = tab_lane + (uint32_t)nl * 32768u;          // first-stage slot 0, this lane's bank
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        // Y covers the bases U-16+r .. U+47+r; byte b holds the 4 bases from U-16+r+4b
        const uint32_t y0 = funnel_r(x0, x1, 2 * r), y1 = funnel_r(x1, x2, 2 * r);
        const uint32_t y2 = funnel_r(x2, x3, 2 * r), y3 = funnel_r(x3, x4, 2 * r);
        Tab idx[8 + S1 - 1];                                    // bytes 4 .. 4+7+S1-1
#pragma unroll
        for (int b = 0; b < 8 + S1 - 1; ++b) {
            const int bb = 4 + b;
            const uint32_t w = bb < 8 ? y1 : (bb < 12 ? y2 : y3);
            idx[b] = tab1 + byte_of(w, bb & 3) * 128u;
        }
        uint32_t acc[8];
        uint32_t any = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t a = lds32(idx[j]) + lds32(idx[j + 1] + 32768u);
            if (S1 >= 3) a += lds32(idx[j + 2] + 65536u);
            if (S1 >= 4) a += lds32(idx[j + 3] + 98304u);
            acc[j] = a;
            any |= ~a;
        }
        if ((any & 0x80808080u) != 0u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if ((~acc[j] & 0x80808080u) != 0u) {
                    uint32_t alive = second_stage(acc[j], y0, y1, y2, y3, 4 + j - nl, nl, S1, nslot, tab_lane);
                    while (alive) {
#if defined(__CUDA_ARCH__)
                        const int bit = __ffs((int)alive) - 1;
#else
                        const int bit = __builtin_ctz(alive);
#endif
                        alive &= alive - 1;
                        emit(bit >> 3, U + 4 * j + r);
                    }
                }
            }
        }
    }
}

#if defined(__CUDACC__)
// ---- Blackwell tensor memory (TMEM) as a second look-up port ---------------------------------------
// tcgen05.ld.32x32b.x1 hands thread l of a warp the 32-bit cell (lane 32*(warp%4)+l, column c) of
// TMEM for a warp-uniform column c — exactly the access pattern of the first stage (all lanes of a
// warp look at the same anchor, lane l wants row `kmer`, bank/column l of the table).  Measured on
// B200 (experiments/tmem/tmem_bench.cu): LDTM 0.77 clk per warp-load per SM, LDS 1.00, interleaved 0.6,
// i.e. the two ports run side by side.  The first two first-stage slots live in TMEM (512 columns =
// 2 x 256 k-mers, every 32-lane quarter holds a copy), the others stay in shared memory.
__device__ __forceinline__ uint32_t tmem_ld32(uint32_t taddr) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
    return v;
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" :: "r"(taddr), "r"(v));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Same contract as process32, first-stage slots 0 and 1 read from TMEM (tq = TMEM address of this
// warp's lane quarter, column 0; its low byte is 0 so one PRMT splices the k-mer byte in).
template <int S1, class Emit>
__device__ __forceinline__ void process32_tmem(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t x4, int64_t U,
                                               int lane, uint32_t tab, uint32_t tq, int nl, int nslot, Emit&& emit) {
    const uint32_t tab_lane = tab + (uint32_t)lane * 4u;
    const uint32_t tab1 = tab_lane + (uint32_t)nl * 32768u;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint32_t y0 = funnel_r(x0, x1, 2 * r), y1 = funnel_r(x1, x2, 2 * r);
        const uint32_t y2 = funnel_r(x2, x3, 2 * r), y3 = funnel_r(x3, x4, 2 * r);
        uint32_t tix[9];                                          // TMEM addresses of bytes 4..12 (slots 0 and 1)
#pragma unroll
        for (int b = 0; b < 9; ++b) {
            const int bb = 4 + b;
            const uint32_t w = bb < 8 ? y1 : (bb < 12 ? y2 : y3);
            tix[b] = __byte_perm(w, tq, 0x7650u + (uint32_t)(bb & 3));
        }
        uint32_t six[S1 >= 3 ? 8 + S1 - 3 : 1];                   // shared-memory addresses of bytes 6.. (slots 2, 3)
        if (S1 >= 3) {
#pragma unroll
            for (int b = 0; b < 8 + S1 - 3; ++b) {
                const int bb = 6 + b;
                const uint32_t w = bb < 8 ? y1 : (bb < 12 ? y2 : y3);
                six[b] = tab1 + byte_of(w, bb & 3) * 128u;
            }
        }
        uint32_t acc[8];
        uint32_t any = 0u;
#pragma unroll
        for (int h = 0; h < 2; ++h) {                             // two half-phases of 4 anchors: 8 TMEM loads in flight
            uint32_t t0[4], t1[4], sm[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) t0[q] = tmem_ld32(tix[4 * h + q]);
#pragma unroll
            for (int q = 0; q < 4; ++q) t1[q] = tmem_ld32(tix[4 * h + q + 1] + 256u);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                sm[q] = 0u;
                if (S1 >= 3) sm[q] = lds32(six[4 * h + q] + 65536u);
                if (S1 >= 4) sm[q] += lds32(six[4 * h + q + 1] + 98304u);
            }
            // t0/t1 are valid only after the wait; tying them to it keeps the compiler from moving their uses above it
            asm volatile("tcgen05.wait::ld.sync.aligned;"
                         : "+r"(t0[0]), "+r"(t0[1]), "+r"(t0[2]), "+r"(t0[3]), "+r"(t1[0]), "+r"(t1[1]), "+r"(t1[2]), "+r"(t1[3])
                         :: "memory");
#pragma unroll
            for (int q = 0; q < 4; ++q) { const uint32_t a = t0[q] + t1[q] + sm[q]; acc[4 * h + q] = a; any |= ~a; }
        }
        if ((any & 0x80808080u) != 0u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if ((~acc[j] & 0x80808080u) != 0u) {
                    uint32_t alive = second_stage(acc[j], y0, y1, y2, y3, 4 + j - nl, nl, S1, nslot, tab_lane);
                    while (alive) {
                        const int bit = __ffs((int)alive) - 1;
                        alive &= alive - 1;
                        emit(bit >> 3, U + 4 * j + r);
                    }
                }
            }
        }
    }
}
#endif

// 2-bit code stored for a base that is not A/C/G/T (or for padding): a position hash, so that long
// N runs look like random sequence to the prefilter instead of a poly-A tract.  The invalid-base
// bit plane is what makes such windows non-hits (exact replay).
PWM_HD uint32_t pad_hash(int64_t u) {
    uint32_t x = (uint32_t)u * 0x9E3779B1u + (uint32_t)((uint64_t)u >> 32) * 0x85EBCA77u;
    x ^= x >> 15; x *= 0x2C1B3C6Du; x ^= x >> 12;
    return x & 3u;
}

// 2-bit base code / invalid flag of stored base index u
PWM_HD uint32_t packed_base(const uint32_t* seq2, int64_t u) { return (seq2[u >> 4] >> (2 * (int)(u & 15))) & 3u; }
PWM_HD uint32_t invalid_bit(const uint32_t* mask, int64_t u) { return (mask[u >> 5] >> (int)(u & 31)) & 1u; }

#if defined(__CUDA_ARCH__)
#define PWM_DADD(a, b) __dadd_rn((a), (b))
#define PWM_NEG_INF __longlong_as_double((long long)0xFFF0000000000000ull)
#else
#define PWM_DADD(a, b) ((a) + (b))
#define PWM_NEG_INF (-__builtin_huge_val())
#endif

// lookAheadSearch' replayed exactly (Motif/Search.hs:176-189): acc_k = acc_{k-1} + s[k][base]
// (-inf for a non-ACGT base), window rejected as soon as acc_k < th[k] (th[k] = thres - sigma[k]).
// tab16: n x 16 doubles, first four codes A,C,G,T.
PWM_HD bool replay_skip(const uint32_t* seq2, const uint32_t* mask, int64_t u0, int n, const double* tab16,
                        const double* th, double* score) {
    double acc = 0.0;
    for (int k = 0; k < n; ++k) {
        const int64_t u = u0 + k;
        const double sc = invalid_bit(mask, u) ? PWM_NEG_INF : tab16[16 * k + packed_base(seq2, u)];
        acc = PWM_DADD(acc, sc);
        if (acc < th[k]) return false;
    }
    *score = acc;
    return true;
}

}  // namespace pwmscan
