// plan_mma.h — host-side planning of the tensor-core first stage (mma_prefilter.cu).
//
// Same idea as the shared-memory prefilter (plan.h): a window can be a hit only if its total
// deficit sum_k d_k(b_k) stays below Dsafe.  Here the first stage evaluates the quantised deficits
// of K1 = 8, 16, 24 or 32 consecutive motif columns (starting at the combo's anchor column c) for
// up to 256 combos (128 motifs x 2 strands) at once as an int8 GEMM on the 5th-gen tensor cores:
//
//     D[position, combo] = sum_{k < K1} sum_b onehot[position + k][b] * W[combo][4k + b]
//
// W[combo][4k+b] = min(127, floor(scale * d_{c+k}(b))) with scale = 126 / Dsafe, and the bias -127
// folded into the four entries of column k = 0 (exactly one of them is selected by the one-hot), so
// that "still possible" <=> D < 0 (the sign bit).  The A operand is never materialised: the UMMA
// shared-memory descriptor walks a linear one-hot byte stream with overlapping rows (row r = the
// 32 bytes from 16 r: LBO = 16 B, SBO = 128 B), i.e. a Toeplitz view (experiments/mma/i8_toeplitz.cu).
// Survivors go straight to the exact fp64 replay (rescore_kernel).
#pragma once
#include "plan.h"

namespace pwmscan {

constexpr int kMmaMotifsPerBlock = 128;
constexpr int kMmaCombos = 256;
constexpr int kMmaQ = 126;
constexpr int kMmaKill = 127;
constexpr int kMmaMaxSteps = 4;                 // K1 = 8 * steps columns

struct MmaCombo {
    int motif = -1, strand = 0, n = 0, c = 0;
    bool enabled = false;
    double sigma1 = 0;                          // P(first stage survives) under uniform iid bases
};

struct MmaBlockPlan {
    int steps = 1;                              // MMAs (K = 32 bytes = 8 columns each) per phase tile
    int ncombo = 32;                            // N of the MMA: combos padded to a multiple of 32
    MmaCombo combo[kMmaCombos];
    std::vector<int8_t> btiles;                 // steps x (ncombo x 32 bytes), canonical K-major no-swizzle layout
    double cand_rate = 0;                       // expected candidates per sequence position
    double cost = 0;
};

namespace detail {
// canonical no-swizzle K-major layout of a [N][32-byte] operand: core matrices of 8 rows x 16 bytes,
// K-adjacent core matrices 128 B apart (LBO), 8-row groups 256 B apart (SBO)
inline size_t btile_offset(int n, int x) { return (size_t)(n / 8) * 256 + (size_t)(x / 16) * 128 + (size_t)(n % 8) * 16 + (size_t)(x % 16); }
}  // namespace detail

// Fills W for `steps` and returns sum of sigma1 over combos (candidates per position).
inline double mma_fill_block(const std::vector<ComboInput>& in, const std::vector<ComboDef>& defs, int steps, MmaBlockPlan& bp) {
    const int K1 = 8 * steps;
    int used = 0;
    for (size_t ci = 0; ci < in.size(); ++ci) if (defs[ci].present) used = (int)ci + 1;
    bp.steps = steps;
    bp.ncombo = std::max(32, (used + 31) / 32 * 32);      // whole 32-column epilogue chunks; padding rows have zero weights
    bp.btiles.assign((size_t)steps * bp.ncombo * 32, 0);
    double rate = 0;
    for (int ci = 0; ci < kMmaCombos; ++ci) {
        MmaCombo& mc = bp.combo[ci];
        mc = MmaCombo();
        if (ci >= (int)in.size() || !defs[ci].present) continue;
        const ComboInput& ip = in[ci];
        const ComboDef& cd = defs[ci];
        mc.motif = ip.motif; mc.strand = ip.strand; mc.n = ip.n;
        if (!cd.possible) continue;                       // no bias => accumulator >= 0 => never a candidate
        mc.enabled = true;
        const int n = ip.n;
        int best_c = 0; double best_w = -1;
        for (int c = 0; c <= std::max(0, n - K1); ++c) {
            double t = 0;
            for (int k = c; k < std::min(n, c + K1); ++k) t += cd.w[k];
            if (t > best_w) { best_w = t; best_c = c; }
        }
        mc.c = best_c;
        const double scale = cd.dsafe > 0 ? (double)kMmaQ / cd.dsafe : 1e300;
        // exact survival probability: saturating convolution of the per-column value distributions
        double dist[kMmaKill + 1];
        std::fill(dist, dist + kMmaKill + 1, 0.0);
        dist[0] = 1.0;
        for (int k = 0; k < K1; ++k) {
            const int col = best_c + k;
            int q[4] = {0, 0, 0, 0};
            if (col < n)
                for (int b = 0; b < 4; ++b) {
                    const double v = std::floor(scale * cd.dk[col][b] * (1.0 - 1e-9));
                    q[b] = v >= (double)kMmaKill ? kMmaKill : (v < 0 ? 0 : (int)v);
                }
            for (int b = 0; b < 4; ++b) {
                const int w = q[b] - (k == 0 ? kMmaKill : 0);             // bias -127 folded into column 0
                bp.btiles[(size_t)(k / 8) * bp.ncombo * 32 + detail::btile_offset(ci, 4 * (k % 8) + b)] = (int8_t)w;
            }
            if (col < n) {
                double nd[kMmaKill + 1];
                std::fill(nd, nd + kMmaKill + 1, 0.0);
                for (int s = 0; s <= kMmaKill; ++s) if (dist[s] != 0)
                    for (int b = 0; b < 4; ++b) nd[std::min(kMmaKill, s + q[b])] += 0.25 * dist[s];
                std::copy(nd, nd + kMmaKill + 1, dist);
            }
        }
        double s1 = 0;
        for (int s = 0; s <= kMmaQ; ++s) s1 += dist[s];
        mc.sigma1 = s1;
        rate += s1;
    }
    bp.cand_rate = rate;
    // clocks per phase tile (128 positions): tensor pipe vs epilogue scan, plus the survivors' handling
    // (measured on B200: a tcgen05.commit completes about every 500 clk at best, so a phase tile costs
    //  max(tensor time, ~500 clk) plus the epilogue that does not overlap; records cost ~600 clk each downstream)
    const double nfrac = bp.ncombo / 256.0;
    bp.cost = std::max(128.0 * steps * nfrac, 500.0) + 600.0 * rate;
    return rate;
}

inline void mma_build_block(const std::vector<ComboInput>& in, MmaBlockPlan& best, int force_steps = 0) {
    std::vector<ComboDef> defs(kMmaCombos);
    for (size_t i = 0; i < in.size() && i < (size_t)kMmaCombos; ++i)
        if (in[i].motif >= 0) combo_deficits(in[i], defs[i]);
    bool have = false;
    for (int steps = 1; steps <= kMmaMaxSteps; ++steps) {
        if (force_steps && steps != force_steps) continue;
        MmaBlockPlan bp;
        mma_fill_block(in, defs, steps, bp);
        if (!have || bp.cost < best.cost) { best = std::move(bp); have = true; }
    }
}

struct MmaHostPlan {
    std::vector<double> th;
    std::vector<MmaBlockPlan> blocks;
    std::vector<std::pair<int, int>> generic;
    double cand_rate = 0;
    int max_n = 0;
};

inline void build_mma_host_plan(const std::vector<MotifHost>& mh, const std::vector<int64_t>& col_off, int64_t total_cols2,
                                const double* thres, int strands, int skip, MmaHostPlan& hp, int force_steps = 0) {
    const int M = (int)mh.size();
    hp.th.assign((size_t)total_cols2, 0.0);
    for (int m = 0; m < M; ++m)
        for (int s = 0; s < 2; ++s) {
            for (int k = 0; k < mh[m].n; ++k) hp.th[(size_t)col_off[2 * m + s] + k] = thres[m] - mh[m].sigma[s][k];
            hp.max_n = std::max(hp.max_n, mh[m].n);
        }
    std::vector<int> pre;
    for (int m = 0; m < M; ++m) {
        bool ok = skip != 0;
        for (int s = 0; s < strands && ok; ++s) {
            if (!std::isfinite(hp.th[(size_t)col_off[2 * m + s] + mh[m].n - 1])) ok = false;
            for (int k = 0; k < mh[m].n && ok; ++k)
                for (int b = 0; b < 4; ++b) if (!std::isfinite(mh[m].tab16[s][16 * k + b])) ok = false;
        }
        if (ok) pre.push_back(m);
        else for (int s = 0; s < strands; ++s) hp.generic.emplace_back(m, s);
    }
    std::stable_sort(pre.begin(), pre.end(), [&](int a, int b) { return mh[a].n < mh[b].n; });
    const int per_block = strands == 2 ? kMmaMotifsPerBlock : kMmaCombos;      // forward-only: 256 motifs per block
    const int nblocks = (int)((pre.size() + per_block - 1) / per_block);
    hp.blocks.resize(nblocks);
    for (int b = 0; b < nblocks; ++b) {
        std::vector<ComboInput> in(kMmaCombos);
        for (int ci = 0; ci < kMmaCombos; ++ci) {
            in[ci].motif = -1; in[ci].strand = 0; in[ci].n = 0; in[ci].tab16 = nullptr; in[ci].th_last = 0;
            const size_t mi = (size_t)b * per_block + (strands == 2 ? ci >> 1 : ci);
            const int s = strands == 2 ? (ci & 1) : 0;
            if (mi >= pre.size()) continue;
            const int m = pre[mi];
            in[ci].motif = m; in[ci].strand = s; in[ci].n = mh[m].n; in[ci].tab16 = mh[m].tab16[s].data();
            in[ci].th_last = hp.th[(size_t)col_off[2 * m + s] + mh[m].n - 1];
        }
        mma_build_block(in, hp.blocks[b], force_steps);
        hp.cand_rate += hp.blocks[b].cand_rate;
    }
}

}  // namespace pwmscan
