// cdf_core.h — one column of scoreCDF's dynamic programme (Motif.hs:282-310) in GATHER form; the same source is compiled
// into cdf_kernel (cdf.cu) and, for the CPU tests, into tests/emul (g++ -ffp-contract=off; nvcc -fmad=false).
//
// The reference scatters, for x ascending over the non-zero bins of `prev` and b = A,C,G,T,
//     new[getIdx (scFn x + x_b)] = bg_b * prev[x] + new[..]            getIdx v = min (nBin-1) (truncate ((v - lo) / step))
// fp64 addition is not associative, so an output bin must add its contributors in exactly that order.  Every operation of
// getIdx . scFn is monotone (non-decreasing) in x under round-to-nearest, so the x that land in bin j for base b form ONE
// contiguous run [B_b(j), B_b(j+1)).  The run boundaries are found by inverting the index function, and one thread per
// output bin merges its four runs by (x, b).  Nothing but `prev` is read and `new` written: no compaction of the non-zero
// bins, no index planes, no atomics.  A zero bin inside [xlo, xhi] contributes bg_b * 0 + acc = acc (all terms are
// >= +0), i.e. exactly what skipping it does (Motif.hs:291 `when (prob /= 0)`).
//
// Deciding "idx(x) >= j" cheaply and exactly.  The reference's d = fl(fl(fl(fl(fl(x + .5) * sp) + lop) + x_b) - lo) / step)
// differs from the real number r = ((x + .5) sp + lop + x_b - lo) / step by at most 6 u S / step (u = 2^-53, S = a bound
// on the magnitude of every intermediate score), and so does the two-operation approximation q(x) = x * A + C_b with
// A = sp / step, C_b = (.5 sp + lop + x_b - lo) / step evaluated once per column (<= 11 u S / step).  With
// tol = 2^-40 S / step (several hundred times both bounds together): q >= j + tol implies d >= j, q < j - tol implies
// d < j, and only in between (about 2 tol of every unit of q: ~1e-7 of the evaluations) the IEEE expression itself is
// evaluated.  The fast path never decides anything the exact expression would decide differently.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CDF_HD __host__ __device__ __forceinline__
#define CDF_NOINLINE __host__ __device__ __noinline__
#else
#define CDF_HD inline
#define CDF_NOINLINE inline __attribute__((noinline))
#endif
#define CDF_PRAGMA_UNROLL1 _Pragma("unroll 1")

namespace pwmscan {

struct CdfScFn { int is_const; double step, lo; };       // scFn: const 0 before the first column, then (x + 0.5) * step + lo

// What the reference's own expression needs (kept in shared memory on the device: only the rare exact test reads it).
struct CdfExact {
    CdfScFn f;              // score of bin x of `prev`
    double xb[4];           // log' p_b - log bg_b of this column
    double lo, step;        // of the new binning (Motif.hs:303-306)
};
// What the fast test and the inversion need (registers).
struct CdfFast {
    double A, invA;         // q(x) = x * A + C[b]; invA = 1 / A (0 before the first column: prev is the single bin x = 0)
    double tol;             // certified bound on |q - d| (see above)
    double delta;           // 2 tol / A + 1e-9: how close to an integer (j - C) / A may be for the boundary to need no test
    // The four bases in RUN ORDER k = 0..3: descending C, i.e. ascending start of their runs in every output bin (a run
    // starts at about (j - C) / A); base[k] is the base's index in A,C,G,T, g[k] its background frequency.
    double C[4], g[4];
    int base[4];
    int nbin;
    int xlo, xhi;           // first / last non-zero bin of prev
};

CDF_HD double cdf_scfn(const CdfScFn& f, int x) { return f.is_const ? 0.0 : ((double)x + 0.5) * f.step + f.lo; }
CDF_HD double cdf_abs(double v) { return v < 0 ? -v : v; }

// plen = length of prev
CDF_HD void cdf_fast_setup(CdfFast& q, const CdfExact& ex, const double bg[4], int nbin, int xlo, int xhi, int plen) {
    const double sp = ex.f.is_const ? 0.0 : ex.f.step, lop = ex.f.is_const ? 0.0 : ex.f.lo;
    q.A = sp / ex.step;
    q.invA = ex.f.is_const ? 0.0 : ex.step / sp;
    double c0 = (((0.5 * sp + lop) + ex.xb[0]) - ex.lo) / ex.step, c1 = (((0.5 * sp + lop) + ex.xb[1]) - ex.lo) / ex.step;
    double c2 = (((0.5 * sp + lop) + ex.xb[2]) - ex.lo) / ex.step, c3 = (((0.5 * sp + lop) + ex.xb[3]) - ex.lo) / ex.step;
    double g0 = bg[0], g1 = bg[1], g2 = bg[2], g3 = bg[3];
    int b0 = 0, b1 = 1, b2 = 2, b3 = 3;
    // stable sorting network, descending C (equal C keep base order)
#define CDF_SW(ca, cb_, ga, gb, ba, bb) if (ca < cb_) { double t_ = ca; ca = cb_; cb_ = t_; t_ = ga; ga = gb; gb = t_; int u_ = ba; ba = bb; bb = u_; }
    CDF_SW(c0, c1, g0, g1, b0, b1) CDF_SW(c2, c3, g2, g3, b2, b3) CDF_SW(c1, c2, g1, g2, b1, b2)
    CDF_SW(c0, c1, g0, g1, b0, b1) CDF_SW(c2, c3, g2, g3, b2, b3) CDF_SW(c1, c2, g1, g2, b1, b2)
#undef CDF_SW
    q.C[0] = c0; q.C[1] = c1; q.C[2] = c2; q.C[3] = c3;
    q.g[0] = g0; q.g[1] = g1; q.g[2] = g2; q.g[3] = g3;
    q.base[0] = b0; q.base[1] = b1; q.base[2] = b2; q.base[3] = b3;
    const double a0 = cdf_abs(ex.xb[0]), a1 = cdf_abs(ex.xb[1]), a2 = cdf_abs(ex.xb[2]), a3 = cdf_abs(ex.xb[3]);
    const double m01 = a0 > a1 ? a0 : a1, m23 = a2 > a3 ? a2 : a3, am = m01 > m23 ? m01 : m23;
    const double S = cdf_abs(lop) + ((double)plen + 1.0) * sp + am + cdf_abs(ex.lo) + 1.0;
    q.tol = 9.094947017729282e-13 * S / ex.step;           // 2^-40 S / step
    q.delta = ex.f.is_const ? 2.0 : 2.0 * q.tol * q.invA + 1e-9;   // before the first column every boundary takes the verified search
    q.nbin = nbin; q.xlo = xlo; q.xhi = xhi;
}

// truncate ((scFn x + x_b - lo) / step) >= j, the reference's own expression (j >= 1)
CDF_NOINLINE bool cdf_idx_ge_exact(const CdfExact* ex, int b, int x, int j) {
    const double d = ((cdf_scfn(ex->f, x) + ex->xb[b]) - ex->lo) / ex->step;
    return d >= (double)j;                                            // trunc d >= j  <=>  d >= j   (j >= 1; false for NaN)
}
CDF_HD bool cdf_idx_ge(double A, double tol, double Cb, const CdfExact* ex, int b, int x, int j) {
    const double v = (double)x * A + Cb, jd = (double)j;
    if (v >= jd + tol) return true;
    if (v < jd - tol) return false;
    return cdf_idx_ge_exact(ex, b, x, j);
}

// B_b(j): the first x in [xlo, xhi + 1] with idx(x) >= j   (j in [0, nbin]; B(0) = xlo, B(nbin) = xhi + 1: the top clamp).
// With e = (j - C) / A: q(x) >= j + tol for every x >= e + tol / A and q(x) < j - tol for every x < e - tol / A, so unless
// an integer lies within delta = 2 tol / A + 1e-9 of e (the 1e-9 covers the rounding of e itself, < 4 u (j + |C|) / A)
// the boundary is floor(e) + 1 — no evaluation of the index function at all.  Otherwise: the verified search below.
CDF_NOINLINE int cdf_boundary_slow(double A, double tol, double Cb, int xlo, int xhi, const CdfExact* ex, int b, int j, double e) {
    int x = e <= (double)xlo ? xlo : (e >= (double)xhi + 1.0 ? xhi + 1 : (int)e);
    if ((double)x < e) ++x;                                           // ceil
    if (x > xhi + 1) x = xhi + 1;
    const bool below_ok = x == xlo || !cdf_idx_ge(A, tol, Cb, ex, b, x - 1, j);
    const bool here_ok = x == xhi + 1 || cdf_idx_ge(A, tol, Cb, ex, b, x, j);
    if (below_ok && here_ok) return x;
    // the estimate is one off when (j - C) / A falls next to an integer; anything further away: binary search
    if (!below_ok) { --x; if (x == xlo || !cdf_idx_ge(A, tol, Cb, ex, b, x - 1, j)) return x; }
    else { ++x; if (x == xhi + 1 || cdf_idx_ge(A, tol, Cb, ex, b, x, j)) return x; }
    int lo = xlo, hi = xhi + 1;
    if (below_ok) lo = x + 1; else hi = x - 1;
    while (lo < hi) { const int mid = lo + ((hi - lo) >> 1); if (cdf_idx_ge(A, tol, Cb, ex, b, mid, j)) hi = mid; else lo = mid + 1; }
    return lo;
}
CDF_HD int cdf_boundary(const CdfFast& q, const CdfExact* ex, int k /* run-order index */, int j) {
    if (j <= 0) return q.xlo;
    if (j >= q.nbin) return q.xhi + 1;
    const double e = ((double)j - q.C[k]) * q.invA;                   // real-valued inverse of q
    if (e > -1.0 && e < 2.0e9) {
        const double ef = e + 1.0;                                    // > 0: truncation is floor
        const int fl = (int)ef;
        const double fr = ef - (double)fl;
        if (fr > q.delta && fr < 1.0 - q.delta) {                     // floor(e) + 1 = fl, clamped to the non-zero range of prev
            const int lo = q.xlo, hi = q.xhi + 1;
            return fl < lo ? lo : (fl > hi ? hi : fl);
        }
    }
    return cdf_boundary_slow(q.A, q.tol, q.C[k], q.xlo, q.xhi, ex, q.base[k], j, e);
}

// Element-wise merge of the four runs by (x, base) — the rare case of overlapping runs.
CDF_NOINLINE double cdf_merge_runs(const double* __restrict__ prev, double g0, double g1, double g2, double g3, int b0, int b1, int b2, int b3,
                                   int s0, int e0, int s1, int e1, int s2, int e2, int s3, int e3) {
    double acc = 0.0;
    for (;;) {
        int key = 0x7FFFFFFF, ks = -1;                               // (x, base) of the next element
        if (s0 < e0) { key = (s0 << 2) | b0; ks = 0; }
        if (s1 < e1 && ((s1 << 2) | b1) < key) { key = (s1 << 2) | b1; ks = 1; }
        if (s2 < e2 && ((s2 << 2) | b2) < key) { key = (s2 << 2) | b2; ks = 2; }
        if (s3 < e3 && ((s3 << 2) | b3) < key) { key = (s3 << 2) | b3; ks = 3; }
        if (ks < 0) break;
        const double g = ks == 0 ? g0 : ks == 1 ? g1 : ks == 2 ? g2 : g3;
        acc = g * prev[key >> 2] + acc;
        if (ks == 0) ++s0; else if (ks == 1) ++s1; else if (ks == 2) ++s2; else ++s3;
    }
    return acc;
}

// new[j] from the four runs [s_k, e_k) of prev (k = run order), merged by (x ascending, base = A,C,G,T).  In run order the
// runs start in ascending x; when they do not overlap they are simply added one after the other (their first elements
// are fetched together: four independent loads), overlapping runs (similar x_b) take the element-wise merge.
// last = index of the last element of prev (clamps the speculative loads).
CDF_HD double cdf_gather_bin(const double* __restrict__ prev, const CdfFast& q, int last, int s0, int e0, int s1, int e1, int s2, int e2, int s3, int e3) {
    const int m01 = e0 > e1 ? e0 : e1, m012 = m01 > e2 ? m01 : e2;
    double acc = 0.0;
    if (e0 <= s1 && m01 <= s2 && m012 <= s3) {
        const double v0 = prev[s0 < last ? s0 : last], v1 = prev[s1 < last ? s1 : last], v2 = prev[s2 < last ? s2 : last], v3 = prev[s3 < last ? s3 : last];
        // Motif.hs:293-296: (bg_b * prob +) applied to new[idx]; a run is one or two bins long almost always
#define CDF_RUN(k, v, s_, e_) if (s_ < e_) { acc = q.g[k] * v + acc; CDF_PRAGMA_UNROLL1 for (int x = s_ + 1; x < e_; ++x) acc = q.g[k] * prev[x] + acc; }
        CDF_RUN(0, v0, s0, e0) CDF_RUN(1, v1, s1, e1) CDF_RUN(2, v2, s2, e2) CDF_RUN(3, v3, s3, e3)
#undef CDF_RUN
        return acc;
    }
    return cdf_merge_runs(prev, q.g[0], q.g[1], q.g[2], q.g[3], q.base[0], q.base[1], q.base[2], q.base[3], s0, e0, s1, e1, s2, e2, s3, e3);
}

}  // namespace pwmscan
