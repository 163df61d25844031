// pwm_oracle.cpp — CPU restatement of the bioinformatics-toolkit PWM motif-scan path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
// library.  The CUDA library (libpwmscan.so) never links or calls it.
//
// PARITY STATUS: "parity unpinned" for absolute values.  The reference is 100% Haskell and no
// GHC toolchain exists in this environment, so the reference itself could not be run here.  The
// reference's own tests for this path (bioinformatics-toolkit/tests/Tests/Motif.hs:58-97) hold
// only five self-consistency invariants over tests/data/motifs.fasta (no golden values); this
// oracle is checked against those five invariants (tests/test_oracle_invariants.py).
//
// Every function cites the reference lines it restates (paths relative to
// /root/reference/bioinformatics-toolkit/src/Bio/).  All arithmetic is IEEE-754 double, compiled
// with -ffp-contract=off (GHC never fuses a*b+c), libm log/exp/pow (GHC's log/exp/(**) on Double).
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <thread>
#include <atomic>
#include <limits>

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace {

constexpr double kPseudoScan = 0.0001;   // Motif/Search.hs:164,196 ; Motif.hs:192
constexpr double kPseudoCdf  = 0.001;    // Motif.hs:324
constexpr double kInf = std::numeric_limits<double>::infinity();

inline double add_some(double x) { return x == 0 ? kPseudoScan : x; }   // Search.hs:162-163

// Score of one column for one sequence byte, IUPAC aware.
// Motif/Search.hs:141-157 (lookAheadSearch) == Motif.hs:169-185 (scoreHelp).
// ok=false on an invalid nucleotide (the reference calls `error`).
inline double col_score_iupac(const double* row, const double* bg, unsigned char ch, bool* ok) {
    const double a = bg[0], c = bg[1], g = bg[2], t = bg[3];
    const double mA = add_some(row[0]), mC = add_some(row[1]), mG = add_some(row[2]), mT = add_some(row[3]);
    switch (ch) {
        case 'A': return std::log(mA / a);
        case 'C': return std::log(mC / c);
        case 'G': return std::log(mG / g);
        case 'T': return std::log(mT / t);
        case 'N': return 0.0;
        case 'V': return std::log((mA + mC + mG) / (a + c + g));
        case 'H': return std::log((mA + mC + mT) / (a + c + t));
        case 'D': return std::log((mA + mG + mT) / (a + g + t));
        case 'B': return std::log((mC + mG + mT) / (c + g + t));
        case 'M': return std::log((mA + mC) / (a + c));
        case 'K': return std::log((mG + mT) / (g + t));
        case 'W': return std::log((mA + mT) / (a + t));
        case 'S': return std::log((mC + mG) / (c + g));
        case 'Y': return std::log((mC + mT) / (c + t));
        case 'R': return std::log((mA + mG) / (a + g));
        default: *ok = false; return 0.0;
    }
}

// Motif/Search.hs:184-189 (lookAheadSearch'): anything but ACGT scores -1/0.
inline double col_score_skip(const double* row, const double* bg, unsigned char ch) {
    switch (ch) {
        case 'A': return std::log(add_some(row[0]) / bg[0]);
        case 'C': return std::log(add_some(row[1]) / bg[1]);
        case 'G': return std::log(add_some(row[2]) / bg[2]);
        case 'T': return std::log(add_some(row[3]) / bg[3]);
        default:  return -kInf;
    }
}

// Index of the row maximum as `U.maximumBy (comparing snd)` picks it (Search.hs:116,
// Motif.hs:154).  vector >= 0.12.2 returns the LAST of equal maxima (Data.List semantics);
// the package version is not pinned by the reference (unpinned, irrelevant for a uniform bg).
inline int argmax_last(const double* row) {
    int bi = 0; double bs = row[0];
    for (int i = 1; i < 4; ++i) if (!(row[i] < bs)) { bs = row[i]; bi = i; }
    return bi;
}

// lookAheadSearch / lookAheadSearch' (Search.hs:126-198).  Returns d (depth) and acc.
// A hit is d == n-1.  `hoisted` tables are [n][256] pre-evaluated scores (NaN marks `error`).
struct LookAhead { int d; double acc; bool ok; };

template <bool SKIP, bool HOISTED>
inline LookAhead look_ahead(const double* bg, int n, const double* mat, const double* table,
                            const double* sigma, const unsigned char* dna, int64_t start, double thres) {
    double acc = 0.0, th_d = -1.0;    // loop (0, -1) 0        Search.hs:133,176
    int d = 0;
    for (;;) {
        if (acc < th_d) return {d - 2, acc, true};            // Search.hs:136,179
        if (d >= n)     return {d - 1, acc, true};            // Search.hs:137-138
        double sc;
        const unsigned char ch = dna[start + d];
        if (HOISTED) {
            sc = table[(size_t)d * 256 + ch];
            if (!SKIP && sc != sc) return {d, acc, false};
        } else if (SKIP) {
            sc = col_score_skip(mat + 4 * d, bg, ch);
        } else {
            bool ok = true;
            sc = col_score_iupac(mat + 4 * d, bg, ch, &ok);
            if (!ok) return {d, acc, false};
        }
        acc = acc + sc;                                        // Search.hs:139,182
        th_d = thres - sigma[d];
        ++d;
    }
}

void build_table(const double* bg, int n, const double* mat, bool skip, std::vector<double>& tab) {
    tab.assign((size_t)n * 256, 0.0);
    for (int d = 0; d < n; ++d)
        for (int ch = 0; ch < 256; ++ch) {
            if (skip) tab[(size_t)d * 256 + ch] = col_score_skip(mat + 4 * d, bg, (unsigned char)ch);
            else {
                bool ok = true;
                double v = col_score_iupac(mat + 4 * d, bg, (unsigned char)ch, &ok);
                tab[(size_t)d * 256 + ch] = ok ? v : std::numeric_limits<double>::quiet_NaN();
            }
        }
}

void optimal_scores_suffix(const double* bg, int n, const double* mat, double* out) {
    // Search.hs:111-124: sigma' = scanl f 0 rows ; out = tail (map (last sigma' -) sigma')
    std::vector<double> s(n + 1);
    s[0] = 0.0;
    for (int i = 0; i < n; ++i) {
        const int j = argmax_last(mat + 4 * i);
        s[i + 1] = s[i] + std::log(mat[4 * i + j] / bg[j]);
    }
    for (int k = 0; k < n; ++k) out[k] = s[n] - s[k + 1];
}

// scoreHelp, Motif.hs:164-193.
inline double score_help(const double* bg, int n, const double* mat, const unsigned char* dna, bool* ok) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) acc = acc + col_score_iupac(mat + 4 * i, bg, dna[i], ok);
    return acc;
}

// binarySearchByBounds (Utils/Functions.hs:142-152).  `k = u + l `shiftR` 1` parses as
// u + (l >> 1); with l == 0 at entry the loop walks k = u downwards until it meets the last
// element whose key is < the query and returns that index + 1 (0 if there is none).  For the
// ascending keys of a CDF this is lower_bound; we compute the same value by bisection.
template <class F>
int64_t lower_bound_idx(int64_t n, F key_lt) {   // key_lt(i) == (key[i] < query)
    int64_t lo = 0, hi = n;
    while (lo < hi) { const int64_t m = lo + ((hi - lo) >> 1); if (key_lt(m)) lo = m + 1; else hi = m; }
    return lo;
}

}  // namespace

// ---------------------------------------------------------------- PWM helpers

// rcPWM, Motif.hs:77-78: reverse of the row-major flattening.
ORC_API void orc_rc_pwm(int n, const double* mat, double* out) {
    for (int i = 0; i < 4 * n; ++i) out[i] = mat[4 * n - 1 - i];
}

ORC_API void orc_optimal_scores_suffix(const double* bg, int n, const double* mat, double* out) {
    optimal_scores_suffix(bg, n, mat, out);
}

// optimalScore, Motif.hs:152-162 (foldl' (+) 0 over rows).
ORC_API double orc_optimal_score(const double* bg, int n, const double* mat) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) { const int j = argmax_last(mat + 4 * i); acc = acc + std::log(mat[4 * i + j] / bg[j]); }
    return acc;
}

// The fp64 score table the scan uses (Q1): rows d, 4 entries log(p'/bg).
ORC_API void orc_score_table(const double* bg, int n, const double* mat, double* out /*[n*4]*/) {
    static const char B[4] = {'A', 'C', 'G', 'T'};
    for (int d = 0; d < n; ++d) for (int b = 0; b < 4; ++b) out[4 * d + b] = col_score_skip(mat + 4 * d, bg, B[b]);
}

// ---------------------------------------------------------------- scanning

// findTFBSWith, Search.hs:38-58.  sigma may be NULL (=> findTFBS, Search.hs:33-35).
// mode: 0 = faithful (libm log per visited column, as the reference), 1 = hoisted table.
// Returns the number of hits (may exceed cap; only the first cap are stored), -1 on an
// invalid nucleotide with skip == 0.
ORC_API int64_t orc_find_tfbs(const double* bg, int n, const double* mat, const double* sigma_in,
                              const unsigned char* dna, int64_t len, double thres, int skip, int mode,
                              int64_t* out_pos, double* out_sc, int64_t cap) {
    std::vector<double> sg(n), tab;
    if (sigma_in) std::copy(sigma_in, sigma_in + n, sg.begin()); else optimal_scores_suffix(bg, n, mat, sg.data());
    if (mode == 1) build_table(bg, n, mat, skip != 0, tab);
    int64_t cnt = 0;
    for (int64_t i = 0; i < len - n + 1; ++i) {      // Search.hs:49
        LookAhead r;
        if (skip) r = mode ? look_ahead<true, true>(bg, n, mat, tab.data(), sg.data(), dna, i, thres)
                           : look_ahead<true, false>(bg, n, mat, nullptr, sg.data(), dna, i, thres);
        else      r = mode ? look_ahead<false, true>(bg, n, mat, tab.data(), sg.data(), dna, i, thres)
                           : look_ahead<false, false>(bg, n, mat, nullptr, sg.data(), dna, i, thres);
        if (!r.ok) return -1;
        if (r.d == n - 1) {                          // Search.hs:51-52
            if (cnt < cap) { out_pos[cnt] = i; out_sc[cnt] = r.acc; }
            ++cnt;
        }
    }
    return cnt;
}

// scores / scores', Motif.hs:127-145.  Returns number of windows, -1 on invalid nucleotide.
ORC_API int64_t orc_scores(const double* bg, int n, const double* mat, const unsigned char* dna, int64_t len, double* out) {
    int64_t m = len - n + 1;
    if (m < 0) m = 0;
    for (int64_t i = 0; i < m; ++i) {
        bool ok = true;
        out[i] = score_help(bg, n, mat, dna + i, &ok);
        if (!ok) return -1;
    }
    return m;
}

// findTFBSSlow, Search.hs:87-95: scores' filtered by (>= thres).
ORC_API int64_t orc_find_tfbs_slow(const double* bg, int n, const double* mat, const unsigned char* dna, int64_t len,
                                   double thres, int64_t* out_pos, double* out_sc, int64_t cap) {
    int64_t cnt = 0;
    for (int64_t i = 0; i < len - n + 1; ++i) {
        bool ok = true;
        const double v = score_help(bg, n, mat, dna + i, &ok);
        if (!ok) return -1;
        if (v >= thres) { if (cnt < cap) { out_pos[cnt] = i; out_sc[cnt] = v; } ++cnt; }
    }
    return cnt;
}

// maxMatchSc, Search.hs:98-109 (IUPAC-aware lookAheadSearch with the running max as threshold).
ORC_API double orc_max_match_sc(const double* bg, int n, const double* mat, const unsigned char* dna, int64_t len, int* ok_out) {
    std::vector<double> sg(n);
    optimal_scores_suffix(bg, n, mat, sg.data());
    double mx = -kInf;
    *ok_out = 1;
    for (int64_t i = 0; i < len - n + 1; ++i) {
        LookAhead r = look_ahead<false, false>(bg, n, mat, nullptr, sg.data(), dna, i, mx);
        if (!r.ok) { *ok_out = 0; return mx; }
        if (r.d == n - 1) mx = r.acc;
    }
    return mx;
}

// ---------------------------------------------------------------- CDF

// scoreCDF, Motif.hs:282-326.  Returns the number of compressed points (may exceed cap).
ORC_API int64_t orc_score_cdf(const double* bg, int n, const double* mat, double* out_x, double* out_p, int64_t cap) {
    const double precision = 1e-5;                         // Motif.hs:322
    const double m_epsilon = 2.220446049250313e-16;        // math-functions m_epsilon (Motif.hs:49)
    std::vector<double> prev(1, 1.0), nw;
    bool const_fn = true; double step_p = 0.0, lo_p = 0.0;   // scFn: const 0, then (x+0.5)*step+lo
    auto scfn = [&](int64_t x) -> double { return const_fn ? 0.0 : ((double)x + 0.5) * step_p + lo_p; };
    const double lbg[4] = {std::log(bg[0]), std::log(bg[1]), std::log(bg[2]), std::log(bg[3])};
    for (int i = 0; i < n; ++i) {
        double xat[4];
        for (int b = 0; b < 4; ++b) {                      // Motif.hs:311-314,324-325
            const double p = mat[4 * i + b];
            xat[b] = (p == 0 ? std::log(kPseudoCdf) : std::log(p)) - lbg[b];
        }
        int64_t f = 0, l = (int64_t)prev.size() - 1;
        while (prev[f] == 0) ++f;                          // Motif.hs:307
        while (prev[l] == 0) --l;                          // Motif.hs:308-309
        const double lo1 = scfn(f), hi1 = scfn(l);
        double mn = xat[0], mx = xat[0];                   // Statistics.Function.minMax (Motif.hs:310)
        for (int b = 1; b < 4; ++b) { if (xat[b] < mn) mn = xat[b]; if (xat[b] > mx) mx = xat[b]; }
        const double lo = lo1 + mn, hi = hi1 + mx;         // Motif.hs:303-304
        if (!(lo < hi)) continue;                          // Motif.hs:287,299
        double nb_d = std::ceil((hi - lo) / precision);    // Motif.hs:305
        int64_t nbin = nb_d > 200000.0 ? 200000 : (int64_t)nb_d;
        const double step = (hi - lo) / (double)nbin;      // Motif.hs:306
        nw.assign((size_t)nbin, 0.0);
        const int64_t plen = (int64_t)prev.size();
        for (int64_t x = 0; x < plen; ++x) {               // Motif.hs:290-296
            const double prob = prev[x];
            if (prob != 0) {
                const double sc = scfn(x);
                for (int b = 0; b < 4; ++b) {
                    int64_t j = (int64_t)(((sc + xat[b]) - lo) / step);   // truncate, Motif.hs:301
                    if (j >= nbin) j = nbin - 1;                          // Motif.hs:302
                    nw[j] = bg[b] * prob + nw[j];
                }
            }
        }
        prev.swap(nw);
        const_fn = false; step_p = step; lo_p = lo;        // Motif.hs:298
    }
    // toCDF (Motif.hs:315): scanl1 (+) then compressCDF (316-321).
    const int64_t len = (int64_t)prev.size();
    std::vector<double> P(len);
    double run = 0.0;
    for (int64_t i = 0; i < len; ++i) { run = (i == 0) ? prev[0] : run + prev[i]; P[i] = run; }
    int64_t cnt = 0;
    for (int64_t i = 0; i < len; ++i) {
        bool keep;
        if (i == 0) keep = true;
        else {
            keep = (P[i] - P[i - 1] > m_epsilon);
            // The reference reads P[i+1] with unsafeIndex; for i == len-1 that is out of bounds
            // (undefined).  We treat the missing right neighbour as "no difference" (SURVEY Q5).
            if (!keep && i + 1 < len) keep = (P[i + 1] - P[i] > m_epsilon);
        }
        if (keep) { if (cnt < cap) { out_x[cnt] = scfn(i); out_p[cnt] = P[i]; } ++cnt; }
    }
    return cnt;
}

// cdf, Motif.hs:236-249.
ORC_API double orc_cdf(const double* xs, const double* ps, int64_t n, double x) {
    const int64_t i = lower_bound_idx(n, [&](int64_t k) { return xs[k] < x; });
    if (i >= n) return ps[n - 1];
    if (i == 0) return x == xs[0] ? ps[0] : 0.0;
    const double a = xs[i - 1], a1 = ps[i - 1], b = xs[i], b1 = ps[i];
    const double al = (b - x) / (b - a), be = (x - a) / (b - a);
    return al * a1 + be * b1;
}

// cdf', Motif.hs:252-270.  status: 0 ok, 1 = "p must be in [0,1]", 2 = "cdf': impossible!".
ORC_API double orc_cdf_inv(const double* xs, const double* ps, int64_t n, double p, int* status) {
    *status = 0;
    if (p > 1 || p < 0) { *status = 1; return 0.0; }
    const int64_t i = lower_bound_idx(n, [&](int64_t k) { return ps[k] < p; });
    if (i >= n) return kInf;
    if (i == 0) { if (p == ps[0]) return xs[0]; *status = 2; return 0.0; }
    const double a = xs[i - 1], a1 = ps[i - 1], b = xs[i], b1 = ps[i];
    const double al = (b1 - p) / (b1 - a1), be = (p - a1) / (b1 - a1);
    return al * a + be * b;
}

// truncateCDF, Motif.hs:273-274: keep points with P >= x.
ORC_API int64_t orc_truncate_cdf(const double* xs, const double* ps, int64_t n, double x, double* ox, double* op) {
    int64_t c = 0;
    for (int64_t i = 0; i < n; ++i) if (ps[i] >= x) { ox[c] = xs[i]; op[c] = ps[i]; ++c; }
    return c;
}

// pValueToScore, Motif.hs:213-214.
ORC_API double orc_pvalue_to_score(double p, const double* bg, int n, const double* mat, int* status) {
    std::vector<double> xs(200001), ps(200001);
    const int64_t c = orc_score_cdf(bg, n, mat, xs.data(), ps.data(), 200001);
    return orc_cdf_inv(xs.data(), ps.data(), c, 1 - p, status);
}

// pValueToScoreExact, Motif.hs:218-230 (exhaustive 4^n; n <= 13 here).
ORC_API double orc_pvalue_to_score_exact(double p, const double* bg, int n, const double* mat) {
    const int64_t tot = (int64_t)1 << (2 * n);
    std::vector<std::pair<double, double>> v((size_t)tot);
    std::vector<unsigned char> kmer(n);
    static const char B[4] = {'A', 'C', 'G', 'T'};
    for (int64_t k = 0; k < tot; ++k) {
        double pb = 1.0;
        for (int i = 0; i < n; ++i) {                      // replicateM l "ACGT": first letter most significant
            const int b = (int)((k >> (2 * (n - 1 - i))) & 3);
            kmer[i] = (unsigned char)B[b];
            pb = pb * bg[b];                               // pBkgd, Motif.hs:197-208
        }
        bool ok = true;
        v[(size_t)k] = {score_help(bg, n, mat, kmer.data(), &ok), pb};
    }
    std::stable_sort(v.begin(), v.end(), [](const auto& a, const auto& b) { return a.first > b.first; });
    double acc = 0.0; int64_t i = 0;
    for (;;) {                                             // go, Motif.hs:228-229
        if (acc > p) return v[(size_t)(i - 1)].first;
        if (i >= tot) return v[(size_t)(tot - 1)].first;
        acc = acc + v[(size_t)i].second; ++i;
    }
}

// toAffinity, Data/Bed/Utils.hs:120-123.
ORC_API int64_t orc_to_affinity(double x1) {
    const double x = -(std::log(std::max(1e-20, x1)) / std::log(10.0));   // negate $ logBase 10 $ max 1e-20 x'
    const double sc = 1 / (1 + std::exp(-(x - 5)));
    return (int64_t)std::nearbyint(sc * 1000);                            // Haskell round = half-even
}

// ---------------------------------------------------------------- alignment (Motif/Alignment.hs)

namespace {

inline double log_base2(double v) { return std::log(v) / std::log(2.0); }   // logBase 2 v

// kld, Utils/Functions.hs:114-119 over 4-vectors.
inline double kld4(const double* xs, const double* ys) {
    double acc = 0.0;
    for (int i = 0; i < 4; ++i) {
        const double x = xs[i];
        if (x == 0) continue;
        acc = acc + x * (log_base2(x) - log_base2(ys[i]));
    }
    return acc;
}
// jsd, Utils/Functions.hs:124-126.
inline double jsd4(const double* xs, const double* ys) {
    double zs[4];
    for (int i = 0; i < 4; ++i) zs[i] = (xs[i] + ys[i]) / 2;
    return 0.5 * kld4(xs, zs) + 0.5 * kld4(ys, zs);
}

// Statistics.Sample.mean = (Numeric.Sum.sum kbn xs) / n  (statistics >= 0.14; unpinned).
inline double kbn_mean(const double* v, int n) {
    double s = 0.0, c = 0.0;
    for (int i = 0; i < n; ++i) {
        const double x = v[i], s1 = s + x;
        if (std::fabs(s) >= std::fabs(x)) c = c + ((s - s1) + x); else c = c + ((x - s1) + s);
        s = s1;
    }
    return (s + c) / (double)n;
}

// penalty kinds: 0 lin, 1 quad, 2 cub, 3 exp (Alignment.hs:47-64).
inline double penal(int kind, double g, int64_t nGap, int64_t nMatch) {
    double k;
    switch (kind) {
        case 0: k = (double)nGap; break;
        case 1: k = (double)(nGap * nGap); break;
        case 2: k = (double)(nGap * nGap * nGap); break;
        default: {                                          // fromIntegral (2^nGap - 1 :: Int): wraps mod 2^64
            uint64_t p = nGap >= 64 ? 0ull : (1ull << nGap);
            k = (double)(int64_t)(p - 1ull);
        }
    }
    return k * g / (double)nMatch;
}

// combiners: 0 l1, 1 l2, 2 l3, 3 lInf (Alignment.hs:66-80).
inline double combine(int kind, const double* sc, int n) {
    double t[64];
    switch (kind) {
        case 0: return kbn_mean(sc, n);
        case 1: for (int i = 0; i < n; ++i) t[i] = std::pow(sc[i], 2.0); return std::sqrt(kbn_mean(t, n));
        case 2: for (int i = 0; i < n; ++i) t[i] = std::pow(sc[i], 3.0); return std::pow(kbn_mean(t, n), 1.0 / 3.0);
        default: { double m = sc[0]; for (int i = 1; i < n; ++i) if (sc[i] > m) m = sc[i]; return m; }  // U.maximum
    }
}

// optimalSc x y, Alignment.hs:117-124.
void optimal_sc(int pk, double g, int x, int y, std::vector<double>& out) {
    out.clear();
    for (int i = 0;; ++i) {
        const int nM = std::min(x, y - i);
        if (nM == 0) break;
        const int nG = i + std::abs(x - (y - i));
        out.push_back(penal(pk, g, nG, nM));
    }
    for (int i = (int)out.size() - 2; i >= 0; --i) out[i] = std::min(out[i], out[i + 1]);   // scanr1 min
}

// loop opti (1/0,-1) a b 0, Alignment.hs:103-112.  a has na rows, b has nb rows.
void align_loop(int pk, double g, int ck, const std::vector<double>& opti, const double* a, int na,
                const double* b, int nb, int n1, int n2, double* dmin, int* imin) {
    double mn = kInf; int mi = -1;
    for (int i = 0; i < nb; ++i) {                          // b@(_:xs): stops when b is exhausted
        if (opti[i] >= mn) break;                           // Alignment.hs:104
        const int nMatch = std::min(na, nb - i);
        double sc[64];
        for (int k = 0; k < nMatch; ++k) sc[k] = jsd4(a + 4 * k, b + 4 * (i + k));   // zipWith fn a b
        const int nGap = n1 + n2 - 2 * nMatch;
        const double d = combine(ck, sc, nMatch) + penal(pk, g, nGap, nMatch);
        if (d < mn) { mn = d; mi = i; }
    }
    *dmin = mn; *imin = mi;
}

}  // namespace

// alignmentBy jsd (penalty kind,gap) (combine kind), Alignment.hs:83-132.  n1,n2 <= 64.
ORC_API void orc_alignment(int n1, const double* m1, int n2, const double* m2, int penal_kind, double gap,
                           int comb_kind, double* d_out, int* same_dir, int* pos) {
    std::vector<double> m2rc(4 * n2), opti1, opti2;
    orc_rc_pwm(n2, m2, m2rc.data());
    optimal_sc(penal_kind, gap, n1, n2, opti1);
    optimal_sc(penal_kind, gap, n2, n1, opti2);
    auto one = [&](const double* s2, double* d, int* p) {
        double d1, d2; int i1, i2;
        align_loop(penal_kind, gap, comb_kind, opti2, s2, n2, m1, n1, n1, n2, &d1, &i1);   // loop opti2 .. s2 s1
        align_loop(penal_kind, gap, comb_kind, opti1, m1, n1, s2, n2, n1, n2, &d2, &i2);   // loop opti1 .. s1 s2
        if (d1 < d2) { *d = d1; *p = i1; } else { *d = d2; *p = -i2; }                       // Alignment.hs:92-93
    };
    double df, dr; int pf, pr;
    one(m2, &df, &pf);
    one(m2rc.data(), &dr, &pr);
    if (df <= dr) { *d_out = df; *same_dir = 1; *pos = pf; } else { *d_out = dr; *same_dir = 0; *pos = pr; }
}

// All pairs i<j in the order of iterativeMerge's initialisation (Merge.hs:161-165) /
// MergeMotifs `comb` (app/MergeMotifs.hs:211-213); packed upper triangle, row-major.
ORC_API void orc_alignment_all_pairs(int M, const int* ncol, const double* mats, const int64_t* mat_off,
                                     int penal_kind, double gap, int comb_kind, int nthreads,
                                     double* d_out, unsigned char* same_dir, int* pos) {
    std::vector<int64_t> row_start(M + 1, 0);
    for (int i = 0; i < M; ++i) row_start[i + 1] = row_start[i] + (M - 1 - i);
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= M) return;
            for (int j = i + 1; j < M; ++j) {
                const int64_t k = row_start[i] + (j - i - 1);
                int sd, p; double d;
                orc_alignment(ncol[i], mats + mat_off[i], ncol[j], mats + mat_off[j], penal_kind, gap, comb_kind, &d, &sd, &p);
                d_out[k] = d; same_dir[k] = (unsigned char)sd; pos[k] = p;
            }
        }
    };
    if (nthreads <= 1) { work(); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(work);
    for (auto& t : th) t.join();
}

// ---------------------------------------------------------------- multi-threaded scan driver
// CPU baseline: (motif x 500 000-bp chunk) work items — the decomposition of the reference's own
// abandoned findTFBS' (Search.hs:60-84) — both strands (rcPWM over the forward sequence, same
// cutoff: Data/Bed/Utils.hs:91-98,114-118).  mode 0 = faithful (log per visited column), 1 = hoisted.
// Hits are counted and (optionally) stored unordered as (motif, strand, pos, score).
ORC_API int64_t orc_scan_mt(const double* bg, int M, const int* ncol, const double* mats, const int64_t* mat_off,
                            const double* thres, const unsigned char* dna, int64_t len, int mode, int nthreads,
                            int64_t chunk, int32_t* out_motif, unsigned char* out_strand, int64_t* out_pos,
                            double* out_sc, int64_t cap) {
    struct Prep { std::vector<double> mat[2], sigma[2], tab[2]; int n; };
    std::vector<Prep> prep(M);
    for (int m = 0; m < M; ++m) {
        Prep& p = prep[m]; p.n = ncol[m];
        p.mat[0].assign(mats + mat_off[m], mats + mat_off[m] + 4 * p.n);
        p.mat[1].resize(4 * p.n); orc_rc_pwm(p.n, p.mat[0].data(), p.mat[1].data());
        for (int s = 0; s < 2; ++s) {
            p.sigma[s].resize(p.n); optimal_scores_suffix(bg, p.n, p.mat[s].data(), p.sigma[s].data());
            if (mode == 1) build_table(bg, p.n, p.mat[s].data(), true, p.tab[s]);
        }
    }
    if (chunk <= 0) chunk = 500000;
    const int64_t nchunk = (len + chunk - 1) / chunk;
    const int64_t nitems = nchunk * M;
    std::atomic<int64_t> next(0), total(0);
    auto work = [&]() {
        for (;;) {
            const int64_t it = next.fetch_add(1);
            if (it >= nitems) return;
            const int m = (int)(it / nchunk); const int64_t c = it % nchunk;
            const Prep& p = prep[m];
            const int64_t s0 = c * chunk, s1 = std::min(len - p.n + 1, s0 + chunk);
            for (int s = 0; s < 2; ++s)
                for (int64_t i = s0; i < s1; ++i) {
                    LookAhead r = mode ? look_ahead<true, true>(bg, p.n, p.mat[s].data(), p.tab[s].data(), p.sigma[s].data(), dna, i, thres[m])
                                       : look_ahead<true, false>(bg, p.n, p.mat[s].data(), nullptr, p.sigma[s].data(), dna, i, thres[m]);
                    if (r.d == p.n - 1) {
                        const int64_t k = total.fetch_add(1);
                        if (k < cap) { out_motif[k] = m; out_strand[k] = (unsigned char)s; out_pos[k] = i; out_sc[k] = r.acc; }
                    }
                }
        }
    };
    if (nthreads <= 1) work();
    else { std::vector<std::thread> th; for (int t = 0; t < nthreads; ++t) th.emplace_back(work); for (auto& t : th) t.join(); }
    return total.load();
}

ORC_API int orc_abi_version() { return 1; }
