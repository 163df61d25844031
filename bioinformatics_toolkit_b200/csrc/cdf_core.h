// cdf_core.h — one column of scoreCDF's dynamic programme (Motif.hs:282-310) in GATHER form; the same source is compiled
// into cdf_kernel (cdf.cu) and, for the CPU tests, into tests/emul (g++ -ffp-contract=off; nvcc -fmad=false).
//
// The reference scatters, for x ascending over the non-zero bins of `prev` and b = A,C,G,T,
//     new[getIdx (scFn x + x_b)] = bg_b * prev[x] + new[..]            getIdx v = min (nBin-1) (truncate ((v - lo) / step))
// fp64 addition is not associative, so an output bin must add its contributors in exactly that order.  Every operation of
// getIdx . scFn is monotone (non-decreasing) in x under round-to-nearest, so the x that land in bin j for base b form ONE
// contiguous run [B_b(j), B_b(j+1)).  The run boundaries are found by inverting the index function (an estimate from the
// real-valued inverse, verified with the exact index function on both sides; a binary search if the estimate is off),
// and one thread per output bin merges its four runs by (x, b).  Nothing but `prev` is read and `new` written: no
// compaction of the non-zero bins, no index planes, no atomics.  A zero bin inside [xlo, xhi] contributes
// bg_b * 0 + acc = acc (all terms are >= +0), i.e. exactly what skipping it does (Motif.hs:291 `when (prob /= 0)`).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CDF_HD __host__ __device__ __forceinline__
#else
#define CDF_HD inline
#endif

namespace pwmscan {

struct CdfScFn { int is_const; double step, lo; };       // scFn: const 0 before the first column, then (x + 0.5) * step + lo

struct CdfCol {
    CdfScFn f;              // score of bin x of `prev`
    double lo, step;        // of the new binning (Motif.hs:303-306)
    double rstep;           // 1 / step (fast path of the index function, never decides on its own)
    double est_a;           // real-valued inverse: first x of bin j for base b ~ j * est_a + est_c(b)
    int nbin;
    int xlo, xhi;           // first / last non-zero bin of prev
};
struct CdfBase { double xb, est_c; };     // log' p_b - log bg_b of this column; intercept of the inverse

CDF_HD double cdf_scfn(const CdfScFn& f, int x) { return f.is_const ? 0.0 : ((double)x + 0.5) * f.step + f.lo; }

CDF_HD void cdf_col_setup(CdfCol& c) {
    c.rstep = 1.0 / c.step;
    c.est_a = c.f.is_const ? 0.0 : c.step / c.f.step;
}
CDF_HD CdfBase cdf_base_setup(const CdfCol& c, double xb) {
    CdfBase cb;
    cb.xb = xb;
    cb.est_c = c.f.is_const ? 0.0 : (c.lo - xb - c.f.lo) / c.f.step - 0.5;
    return cb;
}

// truncate ((scFn x + x_b - lo) / step), not clamped to nBin - 1 (callers only compare it with j <= nBin - 1, for which
// the clamp makes no difference); values below 0 are reported as 0 (cannot happen for x >= xlo up to rounding towards 0).
// Fast path: v * (1/step) differs from the correctly rounded quotient by < 4 ulp, so when it is further than 1e-11
// (relative) from an integer both truncate to the same integer; otherwise the IEEE division decides.
CDF_HD long long cdf_idx(const CdfCol& c, const CdfBase& cb, int x) {
    const double v = (cdf_scfn(c.f, x) + cb.xb) - c.lo;
    const double q = v * c.rstep;
    if (q >= 1.0 && q < 1e15) {
        const long long t = (long long)q;
        const double fr = q - (double)t, tol = 1e-11 * q;
        if (fr > tol && fr < 1.0 - tol) return t;
    }
    const double d = v / c.step;
    if (!(d >= 1.0)) return 0;                                      // also NaN
    if (d >= 4e18) return (long long)4e18;
    return (long long)d;
}

// B_b(j): the first x in [xlo, xhi + 1] with cdf_idx(x) >= j   (j in [0, nbin]; B(0) = xlo, B(nbin) = xhi + 1: the top clamp)
CDF_HD int cdf_boundary(const CdfCol& c, const CdfBase& cb, int j) {
    if (j <= 0) return c.xlo;
    if (j >= c.nbin) return c.xhi + 1;
    const double e = (double)j * c.est_a + cb.est_c;
    int x = e <= (double)c.xlo ? c.xlo : (e >= (double)c.xhi + 1.0 ? c.xhi + 1 : (int)e);
    if ((double)x < e) ++x;                                         // ceil
    if (x > c.xhi + 1) x = c.xhi + 1;
    const bool below_ok = x == c.xlo || cdf_idx(c, cb, x - 1) < (long long)j;
    const bool here_ok = x == c.xhi + 1 || cdf_idx(c, cb, x) >= (long long)j;
    if (below_ok && here_ok) return x;
    int lo = c.xlo, hi = c.xhi + 1;                                 // estimate off (degenerate steps): plain binary search
    if (below_ok) lo = x + 1; else hi = x - 1;
    while (lo < hi) { const int mid = lo + ((hi - lo) >> 1); if (cdf_idx(c, cb, mid) >= (long long)j) hi = mid; else lo = mid + 1; }
    return lo;
}

// Consecutive boundaries of one base by walking: the state is x = B_b(j) and ix = cdf_idx(x) (huge when x = xhi + 1).
// B_b(j + 1) >= B_b(j), and because a new bin is at least as wide as an old one it is x or x + 1 almost always: about one
// evaluation of the index function per boundary instead of the two of a stand-alone cdf_boundary.
struct CdfWalk { int x; long long ix; };
constexpr long long kCdfIdxInf = 0x7FFFFFFFFFFFFFFFLL;

CDF_HD void cdf_walk_begin(const CdfCol& c, const CdfBase& cb, int j, CdfWalk& w) {
    w.x = cdf_boundary(c, cb, j);
    w.ix = w.x <= c.xhi ? cdf_idx(c, cb, w.x) : kCdfIdxInf;
}

CDF_HD void cdf_walk_next(const CdfCol& c, const CdfBase& cb, int j /* = previous j + 1 */, CdfWalk& w) {
    if (j >= c.nbin) { w.x = c.xhi + 1; w.ix = kCdfIdxInf; return; }
    for (int s = 0; s < 4; ++s) {
        if (w.ix >= (long long)j) return;                            // idx(x - 1) < previous j < j: x is the boundary of j too
        ++w.x;
        w.ix = w.x <= c.xhi ? cdf_idx(c, cb, w.x) : kCdfIdxInf;
    }
    if (w.ix >= (long long)j) return;
    cdf_walk_begin(c, cb, j, w);                                     // far away (a new bin narrower than an old one): search
}

// new[j] from the four runs [s[b], e[b]) of prev, merged by (x ascending, b = A,C,G,T)
CDF_HD double cdf_gather_bin(const double* __restrict__ prev, const double bg[4], int s0, int e0, int s1, int e1, int s2, int e2, int s3, int e3) {
    double acc = 0.0;
    for (;;) {
        int x = 0x7FFFFFFF, bs = -1;
        if (s0 < e0) { x = s0; bs = 0; }
        if (s1 < e1 && s1 < x) { x = s1; bs = 1; }
        if (s2 < e2 && s2 < x) { x = s2; bs = 2; }
        if (s3 < e3 && s3 < x) { x = s3; bs = 3; }
        if (bs < 0) break;
        acc = bg[bs] * prev[x] + acc;                                // Motif.hs:293-296: (bg_b * prob +) applied to new[idx]
        if (bs == 0) ++s0; else if (bs == 1) ++s1; else if (bs == 2) ++s2; else ++s3;
    }
    return acc;
}

}  // namespace pwmscan
