// emul_scan.cpp — CPU emulation of the device scan pipeline, TEST ONLY.
//
// Runs the very same planning code (csrc/plan.h) and the very same per-lane kernel body
// (csrc/scan_core.cuh: process32 / second_stage / replay_skip) the CUDA kernels run, with the 32
// lanes of a warp executed one after another on the host.  It lets the CPU test-suite check the
// quantised prefilter for false negatives against the oracle without a GPU.  It is never loaded
// by the product package.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../bioinformatics_toolkit_b200/csrc/plan.h"
#include "../../bioinformatics_toolkit_b200/csrc/scan_core.cuh"

using namespace pwmscan;

extern "C" __attribute__((visibility("default")))
int64_t emul_scan(int M, const int* ncol, const double* mats, const double* bg, const double* thres, int strands,
                  const unsigned char* ascii, int64_t len, int force_s1,
                  int32_t* out_motif, unsigned char* out_strand, int64_t* out_pos, double* out_sc, int64_t cap,
                  int64_t* stats /* [0] candidates [1] first-stage survivors(lane level) [2] nblocks [3] generic combos */) {
    std::vector<MotifHost> mh(M);
    std::vector<int64_t> col_off(2 * (size_t)M);
    int64_t cols = 0;
    const double* p = mats;
    for (int m = 0; m < M; ++m) {
        motif_prepare(mh[m], ncol[m], p, bg);
        p += 4 * ncol[m];
        for (int s = 0; s < 2; ++s) { col_off[2 * m + s] = cols; cols += ncol[m]; }
    }
    HostPlan hp;
    if (!build_host_plan(mh, col_off, cols, thres, strands, 1, hp, force_s1)) return -1;
    // pack (restates pack_kernel)
    const int64_t kChunk = 16384 * 32;    // kAnchorRound of the library
    const int64_t n_anchor = (len + kPadFront + 64 + kChunk - 1) / kChunk * kChunk;
    std::vector<uint32_t> seq2_alloc((size_t)(n_anchor / 16 + 9), 0u), mask((size_t)(n_anchor / 32 + 4), 0u);
    uint32_t* seq2 = seq2_alloc.data() + 1;
    for (int64_t u = 0; u < (int64_t)mask.size() * 32; ++u) {
        const int64_t pos = u - kPadFront;
        const int code = (pos >= 0 && pos < len) ? iupac_code(ascii[pos]) : 15;
        uint32_t c2;
        if (code < 4) c2 = (uint32_t)code; else { c2 = pad_hash(u); mask[u >> 5] |= 1u << (u & 31); }
        seq2[u >> 4] |= c2 << (2 * (u & 15));
    }
    struct Hit { int32_t motif; unsigned char strand; int64_t pos; double sc; };
    std::vector<Hit> hits;
    int64_t ncand = 0;
    for (size_t b = 0; b < hp.blocks.size(); ++b) {
        const BlockPlan& bp = hp.blocks[b];
        const unsigned char* tab = reinterpret_cast<const unsigned char*>(bp.table.data());
        // like the kernel: a warp owns a stretch of 512 * rep anchors; lane replica r scans the r-th 512 of it
        const int64_t stretch = 512 * (int64_t)bp.rep;
        for (int64_t W0 = 0; W0 < n_anchor; W0 += stretch) {
            for (int lane = 0; lane < 32; ++lane) {
                const int replica = lane / bp.groups;
                for (int it = 0; it < 16; ++it) {
                    const int64_t U = W0 + (int64_t)replica * 512 + 32 * it;
                    const uint32_t* w = seq2 + (U >> 4);
                    auto emit = [&](int q, int64_t u) {
                        ++ncand;
                        const ComboPlan& cp = bp.combo[(lane % bp.groups) * 4 + q];
                        if (cp.motif < 0) return;
                        const int64_t i = u - kPadFront - cp.c;
                        if (i < 0 || i + cp.n > len) return;
                        double sc;
                        if (replay_skip(seq2, mask.data(), i + kPadFront, cp.n, mh[cp.motif].tab16[cp.strand].data(),
                                        hp.th.data() + col_off[2 * cp.motif + cp.strand], &sc))
                            hits.push_back({cp.motif, (unsigned char)cp.strand, i, sc});
                    };
                    if (bp.s1 == 2)      process32<2>(w[-1], w[0], w[1], w[2], w[3], U, lane, tab, bp.nl, bp.nslot, emit);
                    else if (bp.s1 == 3) process32<3>(w[-1], w[0], w[1], w[2], w[3], U, lane, tab, bp.nl, bp.nslot, emit);
                    else                 process32<4>(w[-1], w[0], w[1], w[2], w[3], U, lane, tab, bp.nl, bp.nslot, emit);
                }
            }
        }
    }
    // combos the prefilter cannot take: exact replay of every window
    for (auto& g : hp.generic) {
        const int m = g.first, s = g.second, n = mh[m].n;
        for (int64_t i = 0; i + n <= len; ++i) {
            double sc;
            if (replay_skip(seq2, mask.data(), i + kPadFront, n, mh[m].tab16[s].data(), hp.th.data() + col_off[2 * m + s], &sc))
                hits.push_back({m, (unsigned char)s, i, sc});
        }
    }
    std::sort(hits.begin(), hits.end(), [](const Hit& a, const Hit& b) {
        if (a.motif != b.motif) return a.motif < b.motif;
        if (a.strand != b.strand) return a.strand < b.strand;
        return a.pos < b.pos;
    });
    for (size_t k = 0; k < hits.size() && (int64_t)k < cap; ++k) {
        out_motif[k] = hits[k].motif; out_strand[k] = hits[k].strand; out_pos[k] = hits[k].pos; out_sc[k] = hits[k].sc;
    }
    if (stats) {
        stats[0] = ncand; stats[1] = 0; stats[2] = (int64_t)hp.blocks.size(); stats[3] = (int64_t)hp.generic.size();
        double pw = 0; for (auto& bp : hp.blocks) pw += bp.p_warp;
        stats[1] = (int64_t)(1e6 * (hp.blocks.empty() ? 0 : pw / hp.blocks.size()));
        stats[4] = hp.blocks.empty() ? 0 : hp.blocks[0].s1 * 100 + hp.blocks[0].nl * 10 + hp.blocks[0].nr;
    }
    return (int64_t)hits.size();
}

// ---- alignment: the host build of csrc/align_core.h (the function the CUDA kernel runs) -----------
#include "../../bioinformatics_toolkit_b200/csrc/align_core.h"

extern "C" __attribute__((visibility("default")))
void emul_align(int n1, const double* m1, int n2, const double* m2, int penal_kind, double gap, int comb_kind,
                double* d_out, int* same_dir, int* pos, int* ambiguous) {
    AlignCfg cfg;
    cfg.penal_kind = penal_kind; cfg.comb_kind = comb_kind; cfg.gap = gap; cfg.ln2 = std::log(2.0);
    std::vector<double> m2r(4 * n2), l1(4 * n1), l2(4 * n2), l2r(4 * n2), o1(64), o2(64);
    for (int i = 0; i < 4 * n2; ++i) m2r[i] = m2[4 * n2 - 1 - i];
    for (int i = 0; i < 4 * n1; ++i) l1[i] = std::log(m1[i]) / cfg.ln2;
    for (int i = 0; i < 4 * n2; ++i) { l2[i] = std::log(m2[i]) / cfg.ln2; l2r[i] = std::log(m2r[i]) / cfg.ln2; }
    aln_optimal_sc(cfg, n1, n2, o1.data());
    aln_optimal_sc(cfg, n2, n1, o2.data());
    *ambiguous = 0;
    align_pair(cfg, m1, l1.data(), n1, m2, l2.data(), m2r.data(), l2r.data(), n2, o1.data(), o2.data(), d_out, same_dir, pos, ambiguous);
}

// ---- tensor-core first stage: the host plan (csrc/plan_mma.h) evaluated with plain integer loops -------
#include "../../bioinformatics_toolkit_b200/csrc/plan_mma.h"

extern "C" __attribute__((visibility("default")))
int64_t emul_mma_scan(int M, const int* ncol, const double* mats, const double* bg, const double* thres, int strands,
                      const unsigned char* ascii, int64_t len, int force_steps,
                      int32_t* out_motif, unsigned char* out_strand, int64_t* out_pos, double* out_sc, int64_t cap, int64_t* stats) {
    std::vector<MotifHost> mh(M);
    std::vector<int64_t> col_off(2 * (size_t)M);
    int64_t cols = 0;
    const double* p = mats;
    for (int m = 0; m < M; ++m) {
        motif_prepare(mh[m], ncol[m], p, bg);
        p += 4 * ncol[m];
        for (int s = 0; s < 2; ++s) { col_off[2 * m + s] = cols; cols += ncol[m]; }
    }
    MmaHostPlan hp;
    build_mma_host_plan(mh, col_off, cols, thres, strands, 1, hp, force_steps);
    const int64_t kRound = 16384 * 32;
    const int64_t n_anchor = (len + kPadFront + 64 + kRound - 1) / kRound * kRound;
    std::vector<uint32_t> seq2_alloc((size_t)(n_anchor / 16 + 9), 0u), mask((size_t)(n_anchor / 32 + 4), 0u);
    uint32_t* seq2 = seq2_alloc.data() + 1;
    for (int64_t u = 0; u < (int64_t)mask.size() * 32; ++u) {
        const int64_t pos = u - kPadFront;
        const int code = (pos >= 0 && pos < len) ? iupac_code(ascii[pos]) : 15;
        uint32_t c2;
        if (code < 4) c2 = (uint32_t)code; else { c2 = pad_hash(u); mask[u >> 5] |= 1u << (u & 31); }
        seq2[u >> 4] |= c2 << (2 * (u & 15));
    }
    struct Hit { int32_t motif; unsigned char strand; int64_t pos; double sc; };
    std::vector<Hit> hits;
    int64_t ncand = 0;
    const int64_t scan_to = std::min<int64_t>(n_anchor, len + kPadFront + 64);
    for (size_t b = 0; b < hp.blocks.size(); ++b) {
        const MmaBlockPlan& bp = hp.blocks[b];
        for (int64_t u = 0; u < scan_to; ++u)
            for (int ci = 0; ci < bp.ncombo; ++ci) {
                const MmaCombo& mc = bp.combo[ci];
                if (mc.motif < 0) continue;
                int32_t acc = 0;                              // what the int8 GEMM accumulates for (anchor u, combo ci)
                for (int k = 0; k < 8 * bp.steps; ++k) {
                    const uint32_t base = packed_base(seq2, u + k);
                    acc += bp.btiles[(size_t)(k / 8) * bp.ncombo * 32 + detail::btile_offset(ci, 4 * (k % 8) + (int)base)];
                }
                if (acc >= 0) continue;
                ++ncand;
                const int64_t i = u - kPadFront - mc.c;
                if (i < 0 || i + mc.n > len) continue;
                double sc;
                if (replay_skip(seq2, mask.data(), i + kPadFront, mc.n, mh[mc.motif].tab16[mc.strand].data(),
                                hp.th.data() + col_off[2 * mc.motif + mc.strand], &sc))
                    hits.push_back({mc.motif, (unsigned char)mc.strand, i, sc});
            }
    }
    for (auto& g : hp.generic) {
        const int m = g.first, s = g.second, n = mh[m].n;
        for (int64_t i = 0; i + n <= len; ++i) {
            double sc;
            if (replay_skip(seq2, mask.data(), i + kPadFront, n, mh[m].tab16[s].data(), hp.th.data() + col_off[2 * m + s], &sc))
                hits.push_back({m, (unsigned char)s, i, sc});
        }
    }
    std::sort(hits.begin(), hits.end(), [](const Hit& a, const Hit& b) {
        if (a.motif != b.motif) return a.motif < b.motif;
        if (a.strand != b.strand) return a.strand < b.strand;
        return a.pos < b.pos;
    });
    for (size_t k = 0; k < hits.size() && (int64_t)k < cap; ++k) {
        out_motif[k] = hits[k].motif; out_strand[k] = hits[k].strand; out_pos[k] = hits[k].pos; out_sc[k] = hits[k].sc;
    }
    if (stats) {
        stats[0] = ncand; stats[2] = (int64_t)hp.blocks.size(); stats[3] = (int64_t)hp.generic.size();
        stats[1] = (int64_t)(1e6 * hp.cand_rate);
        stats[4] = hp.blocks.empty() ? 0 : hp.blocks[0].steps;
    }
    return (int64_t)hits.size();
}

// ---- k-mer mask first stage: the host plan (csrc/plan_km.h) evaluated with plain loops ---------------------
// The device gathers precomputed masks; here every mask bit is recomputed on the fly with the same fp64 sum
// (km_build_kernel's order) and the middle stage decodes the same packed q16 rows as km_drain.
#include "../../bioinformatics_toolkit_b200/csrc/plan_km.h"

extern "C" __attribute__((visibility("default")))
int64_t emul_km_scan(int M, const int* ncol, const double* mats, const double* bg, const double* thres, int strands,
                     const unsigned char* ascii, int64_t len, int force_nt, int force_L,
                     int32_t* out_motif, unsigned char* out_strand, int64_t* out_pos, double* out_sc, int64_t cap, int64_t* stats) {
    std::vector<MotifHost> mh(M);
    std::vector<int64_t> col_off(2 * (size_t)M);
    int64_t cols = 0;
    const double* p = mats;
    for (int m = 0; m < M; ++m) {
        motif_prepare(mh[m], ncol[m], p, bg);
        p += 4 * ncol[m];
        for (int s = 0; s < 2; ++s) { col_off[2 * m + s] = cols; cols += ncol[m]; }
    }
    KmHostPlan hp;
    build_km_host_plan(mh, col_off, cols, thres, strands, 1, hp, force_nt, force_L);
    const int64_t kRound = 16384 * 32;
    const int64_t n_anchor = (len + kPadFront + 64 + kRound - 1) / kRound * kRound;
    std::vector<uint32_t> seq2_alloc((size_t)(n_anchor / 16 + 9), 0u), mask((size_t)(n_anchor / 32 + 4), 0u);
    uint32_t* seq2 = seq2_alloc.data() + 1;
    for (int64_t u = 0; u < (int64_t)mask.size() * 32; ++u) {
        const int64_t pos = u - kPadFront;
        const int code = (pos >= 0 && pos < len) ? iupac_code(ascii[pos]) : 15;
        uint32_t c2;
        if (code < 4) c2 = (uint32_t)code; else { c2 = pad_hash(u); mask[u >> 5] |= 1u << (u & 31); }
        seq2[u >> 4] |= c2 << (2 * (u & 15));
    }
    struct Hit { int32_t motif; unsigned char strand; int64_t pos; double sc; };
    std::vector<Hit> hits;
    int64_t nsurv = 0, ncand = 0;
    const int64_t scan_to = std::min<int64_t>(n_anchor, len + kPadFront + 64);
    for (size_t b = 0; b < hp.blocks.size(); ++b) {
        const KmBlockPlan& bp = hp.blocks[b];
        for (int ci = 0; ci < kKmCombos; ++ci) {
            const KmCombo& kc = bp.combo[ci];
            if (kc.motif < 0 || bp.dsafe[ci] < 0) continue;
            for (int64_t u = 0; u < scan_to; ++u) {
                bool alive = true;
                for (int t = 0; t < bp.nt && alive; ++t) {
                    const double* wd = &bp.wdef[(((size_t)t * kKmCombos + ci) * kKmMaxL) * 4];
                    double s = wd[packed_base(seq2, u + bp.off[t])];
                    for (int q = 1; q < bp.L[t]; ++q) s = s + wd[4 * q + packed_base(seq2, u + bp.off[t] + q)];
                    alive = s <= bp.dsafe[ci];
                }
                if (!alive) continue;
                ++nsurv;
                const int64_t i = u - kPadFront - kc.c;
                if (i < 0 || i + kc.n > len) continue;
                uint32_t sum = 0;                              // middle stage (km_drain)
                for (int r = 0; r < kc.n; ++r) {
                    const uint16_t* row = &bp.q16[((size_t)ci * bp.max_n + r) * 4];
                    const int col = (row[0] >> 12) | ((row[1] >> 12) << 4);
                    sum += row[packed_base(seq2, i + kPadFront + col)] & 0xFFFu;
                    if (sum > (uint32_t)kKmQ) break;
                }
                if (sum > (uint32_t)kKmQ) continue;
                ++ncand;
                double sc;
                if (replay_skip(seq2, mask.data(), i + kPadFront, kc.n, mh[kc.motif].tab16[kc.strand].data(),
                                hp.th.data() + col_off[2 * kc.motif + kc.strand], &sc))
                    hits.push_back({kc.motif, (unsigned char)kc.strand, i, sc});
            }
        }
    }
    for (auto& g : hp.generic) {
        const int m = g.first, s = g.second, n = mh[m].n;
        for (int64_t i = 0; i + n <= len; ++i) {
            double sc;
            if (replay_skip(seq2, mask.data(), i + kPadFront, n, mh[m].tab16[s].data(), hp.th.data() + col_off[2 * m + s], &sc))
                hits.push_back({m, (unsigned char)s, i, sc});
        }
    }
    std::sort(hits.begin(), hits.end(), [](const Hit& a, const Hit& b) {
        if (a.motif != b.motif) return a.motif < b.motif;
        if (a.strand != b.strand) return a.strand < b.strand;
        return a.pos < b.pos;
    });
    for (size_t k = 0; k < hits.size() && (int64_t)k < cap; ++k) {
        out_motif[k] = hits[k].motif; out_strand[k] = hits[k].strand; out_pos[k] = hits[k].pos; out_sc[k] = hits[k].sc;
    }
    if (stats) {
        stats[0] = ncand; stats[1] = (int64_t)(1e6 * hp.cand_rate); stats[2] = (int64_t)hp.blocks.size(); stats[3] = (int64_t)hp.generic.size();
        stats[4] = hp.blocks.empty() ? 0 : hp.blocks[0].nt; stats[5] = hp.blocks.empty() ? 0 : hp.blocks[0].L[0];
        stats[6] = nsurv; stats[7] = (int64_t)(1e6 * hp.surv_rate);
    }
    return (int64_t)hits.size();
}

// Planner facts per block of the k-mer mask plan: out[8*b + ..] = nt, L of the first gathered table, bitmap used,
// 1e6 * expected sectors per position, 1e6 * expected first-stage survivors, 1e6 * expected candidates, max_n, table MiB.
extern "C" __attribute__((visibility("default")))
int emul_km_plan_info(int M, const int* ncol, const double* mats, const double* bg, const double* thres, int strands,
                      int64_t* out, int max_blocks) {
    std::vector<MotifHost> mh(M);
    std::vector<int64_t> col_off(2 * (size_t)M);
    int64_t cols = 0;
    const double* p = mats;
    for (int m = 0; m < M; ++m) {
        motif_prepare(mh[m], ncol[m], p, bg);
        p += 4 * ncol[m];
        for (int s = 0; s < 2; ++s) { col_off[2 * m + s] = cols; cols += ncol[m]; }
    }
    KmHostPlan hp;
    build_km_host_plan(mh, col_off, cols, thres, strands, 1, hp);
    const int nb = (int)hp.blocks.size();
    for (int b = 0; b < nb && b < max_blocks; ++b) {
        const KmBlockPlan& bp = hp.blocks[b];
        int64_t* o = out + 8 * b;
        o[0] = bp.nt; o[1] = bp.L[bp.order[0]]; o[2] = bp.bitmap ? 1 : 0;
        o[3] = (int64_t)(1e6 * bp.req_rate); o[4] = (int64_t)(1e6 * bp.surv_rate); o[5] = (int64_t)(1e6 * bp.cand_rate);
        o[6] = bp.max_n; o[7] = (int64_t)(bp.table_bytes() >> 20);
        // invariants of the layout the kernel relies on
        int sumL = 0;
        for (int t = 0; t < bp.nt; ++t) sumL += bp.L[t];
        if (sumL > 32 || bp.table_bytes() > kKmMaxTableBytes || bp.q16_bytes() + bp.bitmap_bytes() > kKmSmemBudget) return -1;
    }
    return nb;
}

// ---------------------------------------------------------------------------------------------------------------------
// scoreCDF through csrc/cdf_core.h — the gather form cdf_kernel runs (run boundaries from the inverted index function with
// its certified fast test, one output bin = its four runs merged by (x, base)), four adjacent bins per "thread" as in
// the kernel, executed serially.  Returns the number of compressed CDF points.
#include "../../bioinformatics_toolkit_b200/csrc/cdf_core.h"
#include <cmath>

extern "C" __attribute__((visibility("default")))
int64_t emul_score_cdf(const double* bg, int n, const double* mat, double* out_x, double* out_p, int64_t cap) {
    constexpr int kMaxBin = 200000;
    const double m_epsilon = 2.220446049250313e-16;
    std::vector<double> prev(1, 1.0), nw;
    CdfScFn f{1, 0.0, 0.0};
    int xlo = 0, xhi = 0;
    const double lbg[4] = {std::log(bg[0]), std::log(bg[1]), std::log(bg[2]), std::log(bg[3])};
    for (int col = 0; col < n; ++col) {
        CdfExact ex;
        ex.f = f;
        for (int b = 0; b < 4; ++b) { const double p = mat[4 * col + b]; ex.xb[b] = (p == 0 ? std::log(0.001) : std::log(p)) - lbg[b]; }
        double mn = ex.xb[0], mx = ex.xb[0];
        for (int b = 0; b < 4; ++b) { if (ex.xb[b] < mn) mn = ex.xb[b]; if (ex.xb[b] > mx) mx = ex.xb[b]; }
        const double lo = cdf_scfn(f, xlo) + mn, hi = cdf_scfn(f, xhi) + mx;
        if (!(lo < hi)) continue;
        const double nbd = std::ceil((hi - lo) / 1e-5);
        struct { int nbin; double lo, step; } c;
        c.nbin = nbd > (double)kMaxBin ? kMaxBin : (int)nbd;
        c.lo = lo; c.step = (hi - lo) / (double)c.nbin;
        ex.lo = c.lo; ex.step = c.step;
        CdfFast q;
        cdf_fast_setup(q, ex, bg, c.nbin, xlo, xhi, (int)prev.size());
        nw.assign((size_t)c.nbin, 0.0);
        int nlo = 0x7FFFFFFF, nhi = -1;
        const int last = (int)prev.size() - 1;
        for (int j = 0; j < c.nbin; j += 4) {
            int B[4][5];
            for (int k = 0; k < 4; ++k) for (int g = 0; g < 5; ++g) B[k][g] = cdf_boundary(q, &ex, k, j + g);
            for (int g = 0; g < 4 && j + g < c.nbin; ++g) {
                const double acc = cdf_gather_bin(prev.data(), q, last, B[0][g], B[0][g + 1], B[1][g], B[1][g + 1], B[2][g], B[2][g + 1], B[3][g], B[3][g + 1]);
                nw[(size_t)j + g] = acc;
                if (acc != 0.0) { nlo = std::min(nlo, j + g); nhi = std::max(nhi, j + g); }
            }
        }
        prev.swap(nw);
        xlo = nlo; xhi = nhi;
        f.is_const = 0; f.step = c.step; f.lo = c.lo;
    }
    const int64_t len = (int64_t)prev.size();
    std::vector<double> P((size_t)len);
    double run = 0.0;
    for (int64_t i = 0; i < len; ++i) { run = i == 0 ? prev[0] : run + prev[(size_t)i]; P[(size_t)i] = run; }
    int64_t cnt = 0;
    for (int64_t i = 0; i < len; ++i) {
        bool keep = i == 0;
        if (!keep) {
            keep = P[(size_t)i] - P[(size_t)i - 1] > m_epsilon;
            if (!keep && i + 1 < len) keep = P[(size_t)i + 1] - P[(size_t)i] > m_epsilon;
        }
        if (keep) { if (cnt < cap) { out_x[cnt] = cdf_scfn(f, (int)i); out_p[cnt] = P[(size_t)i]; } ++cnt; }
    }
    return cnt;
}
