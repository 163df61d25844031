// plan.h — host-side planning for the thresholded scan: log-odds tables (host libm), look-ahead
// bounds, and the quantised 4-mer deficit tables the prefilter kernel keeps in shared memory.
// Pure C++ (no CUDA), so the CPU unit tests can compile it too.
//
// Reference semantics restated here (bioinformatics-toolkit/src/Bio/...):
//   score table   s[k][b] = log(p'/bg_b), p' = p==0 ? 0.0001 : p      Motif/Search.hs:185-196
//   IUPAC table   Motif/Search.hs:141-157 == Motif.hs:169-185
//   sigma         Motif/Search.hs:111-124 (optimalScoresSuffix)
//   rcPWM         Motif.hs:77-78
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#include <algorithm>
#include <numeric>
#include <thread>
#include <atomic>

namespace pwmscan {

constexpr int kMaxCols = 64;
constexpr int kLanes = 32;
constexpr int kCombosPerLane = 4;                       // 2 motifs x 2 strands, one byte each
constexpr int kMotifsPerBlock = 64;
constexpr int kCombosPerBlock = kLanes * kCombosPerLane;  // 128
constexpr int kMaxSlots = 7;                            // 7 x 32 KB of shared memory
constexpr int kSlotBytes = 256 * kLanes * 4;            // 32768
constexpr int kPadFront = 32;                           // stored base index u = position + kPadFront
constexpr int kQ2 = 120;                                // quantisation levels, 2-slot first stage
constexpr int kQ3 = 60;                                 // quantisation levels, 3-slot first stage
constexpr int kQ4 = 41;                                 // quantisation levels, 4-slot first stage
// (bias 127-Q plus S1 saturated entries Q+1 must stay below 256: Q <= (128 - S1) / (S1 - 1))
constexpr int kIupacCodes = 16;                         // ACGTNVHDBMKWSYR + invalid(15)

inline double add_some(double x) { return x == 0 ? 0.0001 : x; }

// code of an (upper-case) IUPAC letter: A0 C1 G2 T3 N4 V5 H6 D7 B8 M9 K10 W11 S12 Y13 R14, other 15
inline int iupac_code(unsigned char ch) {
    switch (ch) {
        case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; case 'N': return 4;
        case 'V': return 5; case 'H': return 6; case 'D': return 7; case 'B': return 8; case 'M': return 9;
        case 'K': return 10; case 'W': return 11; case 'S': return 12; case 'Y': return 13; case 'R': return 14;
        default: return 15;
    }
}

// One column of the IUPAC score table (16 doubles; code 15 = NaN marker for `error`).
inline void iupac_column(const double* row, const double* bg, double* out) {
    const double a = bg[0], c = bg[1], g = bg[2], t = bg[3];
    const double mA = add_some(row[0]), mC = add_some(row[1]), mG = add_some(row[2]), mT = add_some(row[3]);
    out[0] = std::log(mA / a); out[1] = std::log(mC / c); out[2] = std::log(mG / g); out[3] = std::log(mT / t);
    out[4] = 0.0;
    out[5] = std::log((mA + mC + mG) / (a + c + g));
    out[6] = std::log((mA + mC + mT) / (a + c + t));
    out[7] = std::log((mA + mG + mT) / (a + g + t));
    out[8] = std::log((mC + mG + mT) / (c + g + t));
    out[9] = std::log((mA + mC) / (a + c));
    out[10] = std::log((mG + mT) / (g + t));
    out[11] = std::log((mA + mT) / (a + t));
    out[12] = std::log((mC + mG) / (c + g));
    out[13] = std::log((mC + mT) / (c + t));
    out[14] = std::log((mA + mG) / (a + g));
    out[15] = std::numeric_limits<double>::quiet_NaN();
}

// U.maximumBy (comparing snd): last of equal maxima (vector >= 0.12.2; unpinned by the reference).
inline int argmax_last(const double* row) {
    int bi = 0; double bs = row[0];
    for (int i = 1; i < 4; ++i) if (!(row[i] < bs)) { bs = row[i]; bi = i; }
    return bi;
}

inline void optimal_scores_suffix(const double* bg, int n, const double* mat, double* out) {
    std::vector<double> s(n + 1);
    s[0] = 0.0;
    for (int i = 0; i < n; ++i) { const int j = argmax_last(mat + 4 * i); s[i + 1] = s[i] + std::log(mat[4 * i + j] / bg[j]); }
    for (int k = 0; k < n; ++k) out[k] = s[n] - s[k + 1];
}

struct MotifHost {
    int n = 0;
    std::vector<double> mat[2];     // [0] the PWM, [1] rcPWM                      (n x 4)
    std::vector<double> tab16[2];   // IUPAC log-odds table                          (n x 16)
    std::vector<double> sigma[2];   // optimalScoresSuffix (or caller override)      (n)
};

inline void motif_prepare(MotifHost& mh, int n, const double* mat, const double* bg) {
    mh.n = n;
    mh.mat[0].assign(mat, mat + 4 * n);
    mh.mat[1].resize(4 * n);
    for (int i = 0; i < 4 * n; ++i) mh.mat[1][i] = mat[4 * n - 1 - i];
    for (int s = 0; s < 2; ++s) {
        mh.tab16[s].resize((size_t)n * 16);
        for (int k = 0; k < n; ++k) iupac_column(mh.mat[s].data() + 4 * k, bg, mh.tab16[s].data() + 16 * k);
        mh.sigma[s].resize(n);
        optimal_scores_suffix(bg, n, mh.mat[s].data(), mh.sigma[s].data());
    }
}

// ------------------------------------------------------------------------------------------
// Prefilter plan.
//
// A "combo" is (motif, strand).  A block holds up to 64 motifs = 128 combos: lane l of a warp owns
// combos 4l..4l+3 = (motif 2l, +), (motif 2l, -), (motif 2l+1, +), (motif 2l+1, -), one byte each
// of a 32-bit shared-memory word.  Every combo has an anchor column c: the window starting at
// sequence position i is examined at anchor a = i + c.  Slot s (s = 0..nslot-1, rho = s - nl) is a
// 256-entry table indexed by the 4 bases at a+4rho .. a+4rho+3 and holds the quantised sum of the
// deficits of motif columns c+4rho .. c+4rho+3 (columns outside [0,n) contribute 0).  Slots
// nl .. nl+s1-1 are the first stage (the kernel reads them for every anchor); the other slots are
// read only for anchors that survive.
//
// Conservativeness: with d_k(b) = max_b' s_k(b') - s_k(b) >= 0, a window can be a
// hit only if sum_k d_k(b_k) <= Dsafe = (sum_k colmax_k - th_last) + margin; entries are
// floor(scale * slot_deficit * (1 - 1e-9)) with scale = Q / Dsafe, saturated at Q+1, so for a hit
// the byte accumulator (bias 127-Q) never reaches 128.
struct ComboPlan {
    int motif = -1;       // index into the caller's motif list (-1 = empty byte lane)
    int strand = 0;
    int n = 0;
    int c = 0;            // anchor column
    bool enabled = false; // false: no window can be a hit (or empty lane)
    double dsafe = 0, scale = 0;
    double sigma1 = 0;    // P(first stage survives) under uniform iid bases
    double sigma_all = 0; // P(candidate)
};

struct BlockPlan {
    int s1 = 2, nl = 0, nr = 0, nslot = 2, Q = kQ2;
    // Few motifs: only lanes 0..groups-1 carry motifs; their table columns are replicated rep =
    // 32/groups times and every replica scans its own stretch of the sequence (position-parallel).
    int rep = 1, groups = kLanes;
    ComboPlan combo[kCombosPerBlock];
    std::vector<uint32_t> table;   // nslot * 256 * 32 words: [slot][kmer][lane]
    double p_warp = 0;             // P(some lane of a warp survives the first stage)
    double cost = 0;
};

struct ComboInput {
    int motif, strand, n;
    const double* tab16;   // n x 16 (first four codes = A,C,G,T)
    double th_last;        // threshold of the final comparison (thres - sigma[n-1])
};

namespace detail {

inline void slot_hist(const int* tab, int Q, std::vector<double>& h) {
    h.assign(Q + 2, 0.0);
    for (int k = 0; k < 256; ++k) h[tab[k]] += 1.0 / 256.0;
}
inline void conv_sat(const std::vector<double>& a, const std::vector<double>& b, int Q, std::vector<double>& out) {
    std::vector<double> r(Q + 2, 0.0);
    for (int i = 0; i <= Q + 1; ++i) if (a[i] != 0)
        for (int j = 0; j <= Q + 1; ++j) if (b[j] != 0) r[std::min(Q + 1, i + j)] += a[i] * b[j];
    out.swap(r);
}

}  // namespace detail

struct ComboDef {          // per-combo deficits, shared by placement and table building
    bool present = false, possible = false;
    int n = 0;
    double dsafe = 0;
    double dk[kMaxCols][4];
    double w[kMaxCols];    // expected clamped deficit of a column (x4)
};

inline void combo_deficits(const ComboInput& ip, ComboDef& cd) {
    cd.present = true; cd.n = ip.n;
    double A = 0, smax = 0;
    for (int k = 0; k < ip.n; ++k) {
        const double* s = ip.tab16 + 16 * k;
        double mx = s[0], amax = std::fabs(s[0]);
        for (int b = 1; b < 4; ++b) { mx = std::max(mx, s[b]); amax = std::max(amax, std::fabs(s[b])); }
        A += amax; smax += mx;
        for (int b = 0; b < 4; ++b) cd.dk[k][b] = mx - s[b];
    }
    const double D = smax - ip.th_last;
    cd.dsafe = D + 1e-9 * (1.0 + A);
    // (non-finite cutoffs never get here: the caller routes them to the exact kernel)
    cd.possible = (cd.dsafe >= 0) && std::isfinite(cd.dsafe);
    for (int k = 0; k < ip.n; ++k) {
        double t = 0;
        for (int b = 0; b < 4; ++b) t += std::min(cd.dk[k][b], cd.possible ? cd.dsafe : 0.0);
        cd.w[k] = t;
    }
}

// Anchor columns for a given (s1, nl, nr): maximise the expected clamped deficit seen by the
// first stage.  Returns false if some combo cannot be placed; *total = sum of the maxima.
inline bool place_block(const std::vector<ComboDef>& defs, int s1, int nl, int nr, int* c_out, double* total) {
    const int nslot = nl + s1 + nr;
    double tot = 0;
    for (size_t ci = 0; ci < defs.size(); ++ci) {
        const ComboDef& cd = defs[ci];
        c_out[ci] = 0;
        if (!cd.present) continue;
        const int n = cd.n;
        if (n > 4 * nslot) return false;
        const int c_lo = n - 4 * (s1 + nr), c_hi = 4 * nl;
        if (c_lo > c_hi) return false;
        int best_c = c_lo; double best_w = -1;
        for (int c = c_lo; c <= c_hi; ++c) {
            double t = 0;
            for (int k = std::max(0, c); k < std::min(n, c + 4 * s1); ++k) t += cd.w[k];
            if (t > best_w) { best_w = t; best_c = c; }
        }
        c_out[ci] = best_c;
        tot += best_w;
    }
    *total = tot;
    return true;
}

// Quantised tables + survival estimates for a placed block.
inline void build_block_tables(const std::vector<ComboInput>& in, const std::vector<ComboDef>& defs, int s1, int nl,
                               int nr, const int* c_in, BlockPlan& bp, bool want_sigma_all) {
    const int nslot = nl + s1 + nr;
    const int Q = (s1 == 2) ? kQ2 : (s1 == 3 ? kQ3 : kQ4);
    bp.s1 = s1; bp.nl = nl; bp.nr = nr; bp.nslot = nslot; bp.Q = Q;
    bp.table.assign((size_t)nslot * 256 * kLanes, 0u);
    std::vector<int> qtab((size_t)nslot * 256);
    std::vector<double> h, acc;
    double log_nosurv = 0.0;
    for (int ci = 0; ci < kCombosPerBlock; ++ci) {
        ComboPlan& cp = bp.combo[ci];
        cp = ComboPlan();
        const int lane = ci >> 2, byte = ci & 3;
        auto put = [&](int slot, int kmer, uint32_t v) {
            bp.table[((size_t)slot * 256 + kmer) * kLanes + lane] |= (v & 0xFFu) << (8 * byte);
        };
        if (ci >= (int)in.size() || !defs[ci].present) {
            for (int k = 0; k < 256; ++k) put(nl, k, 0x80u);      // dead from the start
            continue;
        }
        const ComboInput& ip = in[ci];
        const ComboDef& cd = defs[ci];
        const int n = ip.n;
        cp.motif = ip.motif; cp.strand = ip.strand; cp.n = n; cp.dsafe = cd.dsafe; cp.c = c_in[ci];
        if (!cd.possible) {                                        // no window can reach the cutoff
            for (int k = 0; k < 256; ++k) put(nl, k, 0x80u);
            continue;
        }
        cp.enabled = true;
        const double scale = (cd.dsafe > 0) ? (double)Q / cd.dsafe : 1e300;
        cp.scale = scale;
        for (int s = 0; s < nslot; ++s) {
            const int col0 = cp.c + 4 * (s - nl);
            double e4[4][4];                                       // per base position t, per base b
            for (int t = 0; t < 4; ++t) {
                const int col = col0 + t;
                for (int b = 0; b < 4; ++b) e4[t][b] = (col >= 0 && col < n) ? cd.dk[col][b] : 0.0;
            }
            for (int kmer = 0; kmer < 256; ++kmer) {
                const double e = ((e4[0][kmer & 3] + e4[1][(kmer >> 2) & 3]) + e4[2][(kmer >> 4) & 3]) + e4[3][(kmer >> 6) & 3];
                const double q = std::floor(scale * e * (1.0 - 1e-9));
                const int qi = (q >= (double)(Q + 1)) ? Q + 1 : (q < 0 ? 0 : (int)q);
                qtab[(size_t)s * 256 + kmer] = qi;
                put(s, kmer, (uint32_t)(qi + (s == nl ? 127 - Q : 0)));
            }
        }
        // survival probabilities under uniform iid bases
        detail::slot_hist(&qtab[(size_t)nl * 256], Q, acc);
        for (int s = nl + 1; s < nl + s1; ++s) { detail::slot_hist(&qtab[(size_t)s * 256], Q, h); detail::conv_sat(acc, h, Q, acc); }
        double s1p = 0; for (int i = 0; i <= Q; ++i) s1p += acc[i];
        cp.sigma1 = s1p;
        if (want_sigma_all) {
            for (int s = 0; s < nslot; ++s) if (s < nl || s >= nl + s1) { detail::slot_hist(&qtab[(size_t)s * 256], Q, h); detail::conv_sat(acc, h, Q, acc); }
            double sa = 0; for (int i = 0; i <= Q; ++i) sa += acc[i];
            cp.sigma_all = sa;
        } else cp.sigma_all = s1p;
        log_nosurv += std::log1p(-std::min(s1p, 0.999999));
    }
    // replicate the used lanes' columns
    int used = 0;
    for (int ci = 0; ci < kCombosPerBlock; ++ci) if (bp.combo[ci].motif >= 0) used = ci / kCombosPerLane + 1;
    int groups = 1;
    while (groups < used) groups <<= 1;
    bp.groups = groups; bp.rep = kLanes / groups;
    if (bp.rep > 1)
        for (size_t row = 0; row < (size_t)nslot * 256; ++row)
            for (int lane = groups; lane < kLanes; ++lane) bp.table[row * kLanes + lane] = bp.table[row * kLanes + (lane % groups)];
    // a warp's 32 lanes see the same anchor: it takes the slow path if any combo survives
    bp.p_warp = 1.0 - std::exp(log_nosurv * bp.rep);
    // cost model in shared-memory-load slots per anchor: s1 loads, the per-phase handler (8 anchors
    // share one branch) and the second stage of the survivors
    const double p_phase = 1.0 - std::pow(1.0 - bp.p_warp, 8.0);
    bp.cost = (double)s1 + 0.5 + p_phase * 5.0 + bp.p_warp * 5.0;
}

// Chooses (s1, nl, nr) for a block.  Returns false if the block cannot use the prefilter
// (a motif longer than 28 columns).
inline bool build_block(const std::vector<ComboInput>& in, BlockPlan& best, int force_s1 = 0) {
    int nmax = 0;
    std::vector<ComboDef> defs(kCombosPerBlock);
    for (size_t i = 0; i < in.size() && i < (size_t)kCombosPerBlock; ++i)
        if (in[i].motif >= 0) { nmax = std::max(nmax, in[i].n); combo_deficits(in[i], defs[i]); }
    if (nmax > 4 * kMaxSlots) return false;
    bool have = false;
    int cbuf[kCombosPerBlock], cbest[kCombosPerBlock], cchosen[kCombosPerBlock];
    for (int s1 = 2; s1 <= 4; ++s1) {
        if (force_s1 && s1 != force_s1) continue;
        const int nslot = std::max(s1, std::min(kMaxSlots, (nmax + 3) / 4 + 1));
        int best_nl = -1; double best_tot = -1;
        for (int nl = 0; nl <= std::min(4, nslot - s1); ++nl) {
            const int nr = nslot - s1 - nl;
            if (s1 + nr > 5) continue;
            double tot;
            if (!place_block(defs, s1, nl, nr, cbuf, &tot)) continue;
            if (tot > best_tot) { best_tot = tot; best_nl = nl; std::memcpy(cbest, cbuf, sizeof(cbuf)); }
        }
        if (best_nl < 0) continue;
        BlockPlan bp;
        build_block_tables(in, defs, s1, best_nl, nslot - s1 - best_nl, cbest, bp, false);
        if (!have || bp.cost < best.cost) { best = std::move(bp); have = true; std::memcpy(cchosen, cbest, sizeof(cbest)); }
        if (have && best.cost < (double)(s1 + 1) + 0.5) break;    // a larger first stage cannot be cheaper than its own loads
    }
    if (have) {                                                    // candidate probabilities only for the chosen configuration
        const int s1 = best.s1, nl = best.nl, nr = best.nr;
        build_block_tables(in, defs, s1, nl, nr, cchosen, best, true);
    }
    return have;
}

// ------------------------------------------------------------------------------------------
// Whole-plan assembly (shared by pwmscan_plan_create and the CPU emulation test).
struct HostPlan {
    std::vector<double> th;                          // th[k] = thres - sigma[k], laid out by col_off[2m+s]
    std::vector<BlockPlan> blocks;                   // prefilter blocks (motifs sorted by length)
    std::vector<std::pair<int, int>> generic;        // (motif, strand) handled by the exact kernel only
    int max_n = 0, max_nslot = 0;
    double cand_rate = 0;                            // expected candidates per sequence position
};

inline bool build_host_plan(const std::vector<MotifHost>& mh, const std::vector<int64_t>& col_off, int64_t total_cols2,
                            const double* thres, int strands, int skip, HostPlan& hp, int force_s1 = 0) {
    const int M = (int)mh.size();
    hp.th.assign((size_t)total_cols2, 0.0);
    for (int m = 0; m < M; ++m)
        for (int s = 0; s < 2; ++s) {
            for (int k = 0; k < mh[m].n; ++k) hp.th[(size_t)col_off[2 * m + s] + k] = thres[m] - mh[m].sigma[s][k];   // Search.hs:139,182
            hp.max_n = std::max(hp.max_n, mh[m].n);
        }
    std::vector<int> pre;
    for (int m = 0; m < M; ++m) {
        bool ok = skip && mh[m].n <= 4 * kMaxSlots;
        for (int s = 0; s < strands && ok; ++s) {
            if (!std::isfinite(hp.th[(size_t)col_off[2 * m + s] + mh[m].n - 1])) ok = false;
            for (int k = 0; k < mh[m].n && ok; ++k)
                for (int b = 0; b < 4; ++b) if (!std::isfinite(mh[m].tab16[s][16 * k + b])) ok = false;
        }
        if (ok) pre.push_back(m);
        else for (int s = 0; s < strands; ++s) hp.generic.emplace_back(m, s);
    }
    std::stable_sort(pre.begin(), pre.end(), [&](int a, int b) { return mh[a].n < mh[b].n; });
    const int nblocks = (int)((pre.size() + kMotifsPerBlock - 1) / kMotifsPerBlock);
    hp.blocks.resize(nblocks);
    std::vector<char> okb(nblocks, 1);
    auto build_one = [&](int b) {
        std::vector<ComboInput> in(kCombosPerBlock);
        for (int ci = 0; ci < kCombosPerBlock; ++ci) {
            in[ci].motif = -1; in[ci].strand = 0; in[ci].n = 0; in[ci].tab16 = nullptr; in[ci].th_last = 0;
            const size_t mi = (size_t)b * kMotifsPerBlock + (ci >> 1);
            const int s = ci & 1;
            if (mi >= pre.size() || s >= strands) continue;
            const int m = pre[mi];
            in[ci].motif = m; in[ci].strand = s; in[ci].n = mh[m].n; in[ci].tab16 = mh[m].tab16[s].data();
            in[ci].th_last = hp.th[(size_t)col_off[2 * m + s] + mh[m].n - 1];
        }
        if (!build_block(in, hp.blocks[b], force_s1)) okb[b] = 0;
    };
    {   // blocks are independent: build them on a few host threads
        const int nth = std::max(1, std::min<int>(nblocks, (int)std::min<unsigned>(8u, std::thread::hardware_concurrency())));
        std::atomic<int> next(0);
        auto worker = [&]() { for (;;) { const int b = next.fetch_add(1); if (b >= nblocks) return; build_one(b); } };
        if (nth <= 1) worker();
        else { std::vector<std::thread> th; for (int t = 0; t < nth; ++t) th.emplace_back(worker); for (auto& t : th) t.join(); }
    }
    for (int b = 0; b < nblocks; ++b) {
        if (!okb[b]) return false;
        hp.max_nslot = std::max(hp.max_nslot, hp.blocks[b].nslot);
        for (int ci = 0; ci < kCombosPerBlock; ++ci)
            if (hp.blocks[b].combo[ci].enabled) hp.cand_rate += hp.blocks[b].combo[ci].sigma_all;
    }
    return true;
}

}  // namespace pwmscan
