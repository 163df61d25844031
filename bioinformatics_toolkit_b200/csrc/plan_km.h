// plan_km.h — host-side planning of the k-mer mask first stage (km_prefilter.cu).
//
// Necessary condition used (same deficits as plan.h): a window can be a hit only if its total
// deficit sum_k d_k(b_k) stays <= Dsafe, hence only if the deficit of EVERY sub-window stays
// <= Dsafe.  A block holds up to 256 combos (128 motifs x 2 strands) = the 256 bits of a 32-byte
// mask.  Table t (t < nt <= 4) has one mask per L_t-mer, L_t = 8, 9 or 10 (4^L x 32 B = 2, 8 or 32 MiB,
// L2-resident): bit j of T_t[x] says "the L_t bases x, read as motif columns c_j + off_t ..
// c_j + off_t + L_t - 1 of combo j, cost at most Dsafe_j" (columns outside [0, n) cost nothing;
// off_t = L_0 + .. + L_{t-1}).  For the stored sequence position u the kernel gathers
// T_t[L_t-mer at u + off_t], ANDs the masks, and a surviving bit j is the window that starts at
// u - c_j.  Every combo picks its own offset c_j in [-(L_0 - 1), n-1] (which windows the tables see);
// the product of the windows' pass rates under iid uniform bases is what c_j minimises.  The block
// picks the layout (L_t) that minimises gathered sectors + 15 x first-stage survivors per position.
//
// Survivors (~0.1 per position for 200 combos at p = 1e-5) go through a middle stage in the same
// kernel: the whole-window deficit in 16-bit fixed point from shared memory (q16), and what is left
// (about the hit rate) to the exact fp64 replay (rescore_kernel).  The selectivity experiment behind
// this design is experiments/exp_masktable.py, the L2 gather rate experiments/masktab/l2gather.cu.
#pragma once
#include <cstdlib>

#include "plan.h"

namespace pwmscan {

constexpr int kKmCombos = 256;                  // bits of one mask entry
constexpr int kKmMotifsPerBlock = 128;
constexpr int kKmWords = kKmCombos / 32;        // 8 x uint32 = one 32-byte sector
constexpr int kKmMaxTables = 4;
constexpr int kKmMinL = 8, kKmMaxL = 10;        // window lengths: 8-, 9- or 10-mers (2, 8, 32 MiB per table)
constexpr size_t kKmMaxTableBytes = (size_t)72 << 20;   // per block, so that the tables of the running block stay in the 126 MB L2
constexpr size_t kKmSmemBudget = (size_t)194 << 10;     // dynamic shared memory of the kernel: middle-stage rows + first-table bitmap
                                                        // (+ ~32 KB static: staged sequence words and queues of the 32 warps <= 227 KB)
constexpr double kKmBitmapClk = 0.4;            // clocks per position and SM of the shared-memory bitmap look-up: its wavefronts (~0.11, bank
                                                // conflicts included) share L1TEX with the gathers, plus index/test instructions — fitted on C2
// The bitmap path compacts its hits through a per-warp queue, one hit per lane and round; once more than about a fifth of the
// positions are hits the rounds stop amortising and its cost grows faster than linearly (measured on config-4 chunks, clocks per
// position and SM: P(hit) 0.05 -> 0.54, 0.20 -> 1.20, 0.34 -> 2.87, against 1.9 - 2.0 for the plain path on the same blocks).
constexpr double kKmBitmapDenseFrom = 0.2, kKmBitmapDenseClk = 50.0;
constexpr int kKmQ = 4000;                      // fixed-point budget of the middle stage; entries saturate at 4095 (12 bits)

inline size_t km_table_bytes(int L) { return ((size_t)1 << (2 * L)) * kKmWords * 4; }

struct KmCombo {
    int motif = -1, strand = 0, n = 0, c = 0;
    bool enabled = false;
    double sigma1 = 0;                          // P(all windows pass): first-stage survivors per position
    double sigma_mid = 0;                       // P(middle stage passes): candidates per position
    double pass[kKmMaxTables] = {1, 1, 1, 1};   // per-table pass rates
};

struct KmBlockPlan {
    int nt = 1;
    int max_n = 1;
    int L[kKmMaxTables] = {8, 8, 8, 8};         // window length of table t
    int off[kKmMaxTables] = {0, 8, 16, 24};     // table t sees motif columns c + off[t] .. c + off[t] + L[t] - 1
    int order[kKmMaxTables] = {0, 1, 2, 3};     // gather order: most selective table first
    // One bit per k-mer of table order[0], "the mask is non-zero", kept in shared memory (4^L / 8 bytes): a position
    // whose bit is clear gathers nothing at all.  Used whenever it fits next to the middle-stage rows.
    bool bitmap = false;
    KmCombo combo[kKmCombos];
    std::vector<double> wdef;                   // [nt][256][kKmMaxL][4] deficit of base b at window position p (0 outside the motif)
    std::vector<double> dsafe;                  // [256]; < 0: bit never set
    // Middle-stage rows, [256][max_n] x 8 bytes, visited in order of decreasing expected deficit (early exit):
    // four 16-bit fields (bases A, C, G, T), bits 0..11 the quantised deficit of the row's motif column;
    // bits 12..15 of field 0 / field 1 carry the low 4 / high 2 bits of that column's index.
    std::vector<uint16_t> q16;
    double surv_rate = 0;                       // expected first-stage survivors per position
    double cand_rate = 0;                       // expected candidates (middle-stage survivors) per position
    double req_rate = 0;                        // expected table sectors gathered per position
    double cost = 0;
    size_t table_bytes() const { size_t t = 0; for (int i = 0; i < nt; ++i) t += km_table_bytes(L[i]); return t; }
    size_t q16_bytes() const { return (size_t)kKmCombos * max_n * 8; }
    size_t bitmap_bytes() const { return bitmap ? ((size_t)1 << (2 * L[order[0]])) / 8 : 0; }
};

namespace detail {

// Pass rates P(window deficit <= dsafe) of the windows [a, a+L) ∩ [0, n), L = 8, 9, 10, for every start
// a in [-9, n-1], under iid uniform bases: 128-bin convolution of the per-column deficits (rounded to
// the nearest bin: a planning estimate, not a bound).  out[L-8][a+9].
inline void km_window_rates(const ComboDef& cd, std::vector<float> out[3]) {
    constexpr int NB = 128;
    const int n = cd.n;
    for (int l = 0; l < 3; ++l) out[l].assign(n + 9, 1.0f);
    if (!(cd.dsafe > 0)) {                       // only zero-deficit windows pass
        for (int l = 0; l < 3; ++l)
            for (int a = -9; a < n; ++a) {
                double p = 1;
                for (int k = std::max(0, a); k < std::min(n, a + 8 + l); ++k) { int z = 0; for (int b = 0; b < 4; ++b) z += cd.dk[k][b] <= 0; p *= z / 4.0; }
                out[l][a + 9] = (float)p;
            }
        return;
    }
    std::vector<int> q((size_t)n * 4);
    const double sc = NB / cd.dsafe;
    for (int k = 0; k < n; ++k)
        for (int b = 0; b < 4; ++b) { const double v = std::floor(cd.dk[k][b] * sc + 0.5); q[4 * k + b] = v > NB ? NB + 1 : (int)v; }
    double h[NB + 1], nx[NB + 1];
    for (int a = -9; a < n; ++a) {
        std::fill(h, h + NB + 1, 0.0);
        h[0] = 1.0;
        int top = 0;                              // highest occupied bin
        for (int j = 0; j < kKmMaxL; ++j) {
            const int k = a + j;
            if (k >= 0 && k < n) {
                std::fill(nx, nx + NB + 1, 0.0);
                int ntop = 0;
                for (int s = 0; s <= top; ++s) if (h[s] != 0)
                    for (int b = 0; b < 4; ++b) { const int t = s + q[4 * k + b]; if (t <= NB) { nx[t] += 0.25 * h[s]; ntop = std::max(ntop, t); } }
                std::copy(nx, nx + NB + 1, h);
                top = ntop;
            }
            if (j + 1 >= kKmMinL) {
                double t = 0;
                for (int s = 0; s <= top; ++s) t += h[s];
                out[j + 1 - kKmMinL][a + 9] = (float)t;
            }
        }
    }
}

// P(sum over all columns of the clamped deficits <= limit), 1024-bin convolution (rounded down, so a slight over-estimate)
inline double km_total_pass(const ComboDef& cd, double limit) {
    const int NB = 1024;
    if (!(limit > 0)) return 0.0;
    std::vector<double> h(NB + 1, 0.0), nx(NB + 1);
    h[0] = 1.0;
    const double sc = NB / limit;
    for (int k = 0; k < cd.n; ++k) {
        std::fill(nx.begin(), nx.end(), 0.0);
        int q[4];
        for (int b = 0; b < 4; ++b) { const double v = std::floor(cd.dk[k][b] * sc); q[b] = v >= NB ? NB + 1 : (int)v; }
        for (int s = 0; s <= NB; ++s) if (h[s] != 0)
            for (int b = 0; b < 4; ++b) if (s + q[b] <= NB) nx[s + q[b]] += 0.25 * h[s];
        h.swap(nx);
    }
    double t = 0;
    for (int s = 0; s <= NB; ++s) t += h[s];
    return t;
}

struct KmConfig { int nt; int L[kKmMaxTables]; };
// table layouts the planner chooses from (sum of L <= 32: the kernel indexes them from one 64-bit window of 32 bases)
inline const std::vector<KmConfig>& km_configs() {
    static const std::vector<KmConfig> c = {
        {1, {8, 0, 0, 0}},  {2, {8, 8, 0, 0}},  {3, {8, 8, 8, 0}},  {4, {8, 8, 8, 8}},
        {1, {9, 0, 0, 0}},  {2, {9, 9, 0, 0}},  {3, {9, 9, 9, 0}},
        {1, {10, 0, 0, 0}}, {2, {10, 10, 0, 0}}, {3, {10, 10, 8, 0}}, {3, {10, 10, 9, 0}},
    };
    return c;
}

}  // namespace detail

// Builds the plan of one block.  `in`: 256 combos (motif < 0 = unused bit).  force_nt / force_L restrict the layouts tried.
inline void km_build_block(const std::vector<ComboInput>& in, KmBlockPlan& bp, int force_nt = 0, int force_L = 0) {
    std::vector<ComboDef> defs(kKmCombos);
    int max_n = 1;
    for (int ci = 0; ci < kKmCombos; ++ci)
        if (in[ci].motif >= 0) { combo_deficits(in[ci], defs[ci]); max_n = std::max(max_n, in[ci].n); }
    std::vector<std::vector<float>> rates((size_t)kKmCombos * 3);
    for (int ci = 0; ci < kKmCombos; ++ci)
        if (defs[ci].present && defs[ci].possible) detail::km_window_rates(defs[ci], &rates[(size_t)ci * 3]);
    auto pass_at = [&](int ci, int L, int a) -> double {
        return (a <= -L || a >= defs[ci].n) ? 1.0 : (double)rates[(size_t)ci * 3 + (L - kKmMinL)][a + 9];
    };
    bool have = false;
    KmBlockPlan best;
    for (const detail::KmConfig& cf : detail::km_configs()) {
        if (force_nt && cf.nt != force_nt) continue;
        if (force_L && cf.L[0] != force_L) continue;
        KmBlockPlan cur;
        cur.nt = cf.nt; cur.max_n = max_n;
        int o = 0;
        for (int t = 0; t < cf.nt; ++t) { cur.L[t] = cf.L[t]; cur.off[t] = o; o += cf.L[t]; }
        if (cur.table_bytes() > kKmMaxTableBytes) continue;
        if (!force_nt && cf.nt > 1 && cur.off[cf.nt - 1] >= max_n + cf.L[0] - 1) continue;     // the last table could never see a column
        // Every combo picks its offset c.  What it costs the block: its chance of keeping the AND non-zero after
        // the first gathered table (a second sector, weight alpha = d sectors / d lambda = exp(-lambda)) and its
        // first-stage survivors (weight kSurvClk).  Solved by a few rounds per choice of the first table.
        constexpr double kSurvClk = 15.0;
        double req = 0;
        bool have_cur = false;
        KmBlockPlan trial = cur;
        auto bitmap_fits = [&](int t) { return (size_t)kKmCombos * max_n * 8 + (((size_t)1 << (2 * cur.L[t])) / 8) <= kKmSmemBudget; };
        for (int first = 0; first < cur.nt; ++first) {
            double alpha = 1.0;
            // with the bitmap a non-zero first mask costs the first AND the second sector (both are then gathered at once)
            const double w_first = bitmap_fits(first) ? 2.0 : 1.0;      // (an estimate: the bitmap may still be declined below)
            for (int round = 0; round < 3; ++round) {
                double tab_sum[kKmMaxTables] = {0, 0, 0, 0};
                trial.surv_rate = 0;
                for (int ci = 0; ci < kKmCombos; ++ci) {
                    KmCombo& kc = trial.combo[ci];
                    kc = KmCombo();
                    const ComboDef& cd = defs[ci];
                    if (!cd.present) continue;
                    kc.motif = in[ci].motif; kc.strand = in[ci].strand; kc.n = in[ci].n;
                    if (!cd.possible) continue;
                    kc.enabled = true;
                    double bestv = 1e300, bestp = 1; int bestc = 0;
                    for (int c = -(cur.L[0] - 1); c < cd.n; ++c) {
                        double p = 1;
                        for (int t = 0; t < cur.nt; ++t) p *= pass_at(ci, cur.L[t], c + cur.off[t]);
                        const double v = (cur.nt > 1 ? w_first * alpha * pass_at(ci, cur.L[first], c + cur.off[first]) : 0.0) + kSurvClk * p;
                        if (v < bestv) { bestv = v; bestp = p; bestc = c; }
                    }
                    kc.c = bestc; kc.sigma1 = bestp;
                    for (int t = 0; t < cur.nt; ++t) { kc.pass[t] = pass_at(ci, cur.L[t], bestc + cur.off[t]); tab_sum[t] += kc.pass[t]; }
                    trial.surv_rate += bestp;
                }
                for (int t = 0; t < kKmMaxTables; ++t) trial.order[t] = t;
                std::stable_sort(trial.order, trial.order + cur.nt, [&](int a, int b) { return tab_sum[a] < tab_sum[b]; });
                // sectors gathered per position.  Without the bitmap: table order[0] always, the next one only while
                // the AND is still non-zero.  With it: nothing unless the first mask is non-zero, then the first two
                // tables at once, the third only while the AND is still non-zero.
                double pnz[kKmMaxTables] = {1, 1, 1, 1};                      // P(AND of the first s masks non-zero), s = 1..nt-1
                for (int s = 1; s < cur.nt || s == 1; ++s) {
                    double lam = 0;                                           // expected surviving bits after the first s tables
                    for (int ci = 0; ci < kKmCombos; ++ci) if (trial.combo[ci].enabled) {
                        double p = 1;
                        for (int k = 0; k < s; ++k) p *= trial.combo[ci].pass[trial.order[k]];
                        lam += p;
                    }
                    pnz[s] = 1.0 - std::exp(-lam);
                }
                double r_bm = pnz[1] * (cur.nt > 1 ? 2.0 : 1.0), r_plain = 1;
                for (int s = 2; s < cur.nt; ++s) r_bm += pnz[s];
                for (int s = 1; s < cur.nt; ++s) r_plain += pnz[s];
                static const int force_bm = std::getenv("PWMSCAN_FORCE_BITMAP") ? std::atoi(std::getenv("PWMSCAN_FORCE_BITMAP")) : -1;   // experiments
                const double dense = std::max(0.0, pnz[1] - kKmBitmapDenseFrom);
                const double bm_clk = kKmBitmapClk + kKmBitmapDenseClk * dense * dense;
                trial.bitmap = bitmap_fits(trial.order[0]) && (force_bm < 0 ? r_bm + bm_clk < r_plain : force_bm != 0);
                const double r = trial.bitmap ? r_bm : r_plain;
                const double cost = r + (trial.bitmap ? bm_clk : 0.0) + kSurvClk * trial.surv_rate;
                if (!have_cur || cost < cur.cost) { cur = trial; cur.cost = cost; req = r; have_cur = true; }
                alpha = std::exp(-tab_sum[first]);
                if (cur.nt == 1) break;
            }
            if (cur.nt == 1) break;
        }
        cur.req_rate = req;
        // (clocks per position and SM, measured on B200 / C2: about one gathered sector per clock, ~15 clocks per
        //  first-stage survivor for bit extraction, queueing and the middle stage — fitted to forced layouts)
        if (!have || cur.cost < best.cost) { best = std::move(cur); have = true; }
    }
    bp = std::move(best);
    const int nt = bp.nt;
    bp.wdef.assign((size_t)nt * kKmCombos * kKmMaxL * 4, 0.0);
    bp.dsafe.assign(kKmCombos, -1.0);
    bp.q16.assign((size_t)kKmCombos * max_n * 4, 0);
    bp.cand_rate = 0;
    for (int ci = 0; ci < kKmCombos; ++ci) {
        KmCombo& kc = bp.combo[ci];
        if (!kc.enabled) continue;
        const ComboDef& cd = defs[ci];
        bp.dsafe[ci] = cd.dsafe;
        for (int t = 0; t < nt; ++t)
            for (int p = 0; p < bp.L[t]; ++p) {
                const int col = kc.c + bp.off[t] + p;
                if (col < 0 || col >= cd.n) continue;
                for (int b = 0; b < 4; ++b) bp.wdef[(((size_t)t * kKmCombos + ci) * kKmMaxL + p) * 4 + b] = cd.dk[col][b];
            }
        const double scale = cd.dsafe > 0 ? (double)kKmQ / cd.dsafe : 1e300;
        int perm[kMaxCols];
        for (int k = 0; k < cd.n; ++k) perm[k] = k;
        std::stable_sort(perm, perm + cd.n, [&](int a, int b) { return cd.w[a] > cd.w[b]; });
        for (int r = 0; r < cd.n; ++r) {
            const int col = perm[r];
            uint16_t* row = &bp.q16[((size_t)ci * max_n + r) * 4];
            for (int b = 0; b < 4; ++b) {
                const double v = std::floor(scale * cd.dk[col][b] * (1.0 - 1e-9));
                row[b] = (uint16_t)(v < 4095.0 ? (v < 0 ? 0.0 : v) : 4095.0);
            }
            row[0] |= (uint16_t)((col & 15) << 12);
            row[1] |= (uint16_t)((col >> 4) << 12);
        }
        kc.sigma_mid = detail::km_total_pass(cd, cd.dsafe * (1.0 + (double)(cd.n + 1) / kKmQ));
        bp.cand_rate += kc.sigma_mid;
    }
}

struct KmHostPlan {
    std::vector<double> th;
    std::vector<KmBlockPlan> blocks;
    std::vector<std::pair<int, int>> generic;
    // skip == 0 (IUPAC-aware lookAheadSearch): the combos of the blocks once more — their windows that hold a letter other
    // than A/C/G/T are not the first stage's business (it sees 2-bit codes) and get an exact pass of their own
    std::vector<std::pair<int, int>> ambig;
    double cand_rate = 0, surv_rate = 0, req_rate = 0;
    int max_n = 0;
};

inline void build_km_host_plan(const std::vector<MotifHost>& mh, const std::vector<int64_t>& col_off, int64_t total_cols2,
                               const double* thres, int strands, int skip, KmHostPlan& hp, int force_nt = 0, int force_L = 0) {
    const int M = (int)mh.size();
    hp.th.assign((size_t)total_cols2, 0.0);
    for (int m = 0; m < M; ++m)
        for (int s = 0; s < 2; ++s) {
            for (int k = 0; k < mh[m].n; ++k) hp.th[(size_t)col_off[2 * m + s] + k] = thres[m] - mh[m].sigma[s][k];   // Search.hs:139,182
            hp.max_n = std::max(hp.max_n, mh[m].n);
        }
    std::vector<int> pre;
    for (int m = 0; m < M; ++m) {
        bool ok = true;
        for (int s = 0; s < strands && ok; ++s) {
            if (!std::isfinite(hp.th[(size_t)col_off[2 * m + s] + mh[m].n - 1])) ok = false;
            for (int k = 0; k < mh[m].n && ok; ++k)
                for (int b = 0; b < 4; ++b) if (!std::isfinite(mh[m].tab16[s][16 * k + b])) ok = false;
        }
        if (ok) {
            pre.push_back(m);
            if (!skip) for (int s = 0; s < strands; ++s) hp.ambig.emplace_back(m, s);
        } else {
            for (int s = 0; s < strands; ++s) hp.generic.emplace_back(m, s);
        }
    }
    // Longest motifs first: the windows of long motifs pass most often, so their blocks are the expensive ones, and the cost of a
    // block grows less than linearly with its combos (1 - exp(-sum of pass rates)) — the expensive motifs belong together in
    // full blocks, and the partial last block falls to the shortest motifs, whose blocks cost the bitmap scan and little else.
    std::stable_sort(pre.begin(), pre.end(), [&](int a, int b) { return mh[a].n > mh[b].n; });
    const int per_block = strands == 2 ? kKmMotifsPerBlock : kKmCombos;        // forward-only: 256 motifs per block
    const int nblocks = (int)((pre.size() + per_block - 1) / per_block);
    hp.blocks.resize(nblocks);
    auto build_one = [&](int b) {
        std::vector<ComboInput> in(kKmCombos);
        for (int ci = 0; ci < kKmCombos; ++ci) {
            in[ci].motif = -1; in[ci].strand = 0; in[ci].n = 0; in[ci].tab16 = nullptr; in[ci].th_last = 0;
            const size_t mi = (size_t)b * per_block + (strands == 2 ? ci >> 1 : ci);
            const int s = strands == 2 ? (ci & 1) : 0;
            if (mi >= pre.size()) continue;
            const int m = pre[mi];
            in[ci].motif = m; in[ci].strand = s; in[ci].n = mh[m].n; in[ci].tab16 = mh[m].tab16[s].data();
            in[ci].th_last = hp.th[(size_t)col_off[2 * m + s] + mh[m].n - 1];
        }
        km_build_block(in, hp.blocks[b], force_nt, force_L);
    };
    {
        const int nth = std::max(1, std::min<int>(nblocks, (int)std::min<unsigned>(8u, std::thread::hardware_concurrency())));
        std::atomic<int> next(0);
        auto worker = [&]() { for (;;) { const int b = next.fetch_add(1); if (b >= nblocks) return; build_one(b); } };
        if (nth <= 1) worker();
        else { std::vector<std::thread> th; for (int t = 0; t < nth; ++t) th.emplace_back(worker); for (auto& t : th) t.join(); }
    }
    for (int b = 0; b < nblocks; ++b) {
        hp.cand_rate += hp.blocks[b].cand_rate;
        hp.surv_rate += hp.blocks[b].surv_rate;
        hp.req_rate += hp.blocks[b].req_rate;
    }
}

}  // namespace pwmscan
