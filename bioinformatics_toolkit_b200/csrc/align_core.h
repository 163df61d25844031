// align_core.h — alignmentBy jsd pFn combFn (Motif/Alignment.hs:83-132) for one PWM pair, written
// once as __host__ __device__ code.  The device build evaluates log2 of the mixture columns with
// the CUDA libm (last-ulp differences from glibc are possible); every comparison that decides the
// result (argmin over offsets, early exit, d1 < d2, forward <= reverse) raises an `ambiguous` flag
// when its operands differ by less than kAmbigRel relative without being bit-equal, and the host
// build of the SAME function (glibc log) re-evaluates exactly those pairs.  Structural ties
// (palindromes, duplicated motifs) produce bit-equal operands in both builds and are not flagged.
//
// Reference arithmetic restated (paths relative to bioinformatics-toolkit/src/Bio/):
//   kld / jsd                      Utils/Functions.hs:114-128   (logBase 2 v = log v / log 2)
//   lin/quad/cub/expPenal          Motif/Alignment.hs:47-64
//   l1 (= Statistics.Sample.mean, KBN-compensated sum / n), l2, l3, lInf   Alignment.hs:66-80
//   loop / optimalSc / alignmentBy Alignment.hs:87-131
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define ALN_HD __host__ __device__ __forceinline__
#else
#define ALN_HD inline
#endif

namespace pwmscan {

constexpr double kAmbigRel = 1e-10;

struct AlignCfg {
    int penal_kind;     // 0 lin, 1 quad, 2 cub, 3 exp
    int comb_kind;      // 0 l1, 1 l2, 2 l3, 3 lInf
    double gap;
    double ln2;         // host libm log(2.0)
};

ALN_HD double aln_penal(const AlignCfg& c, int64_t nGap, int64_t nMatch) {
    double k;
    switch (c.penal_kind) {
        case 0: k = (double)nGap; break;
        case 1: k = (double)(nGap * nGap); break;
        case 2: k = (double)(nGap * nGap * nGap); break;
        default: {                                  // fromIntegral (2^nGap - 1 :: Int): wraps mod 2^64
            const uint64_t p = nGap >= 64 ? 0ull : (1ull << nGap);
            k = (double)(int64_t)(p - 1ull);
        }
    }
    return k * c.gap / (double)nMatch;
}

// log2 of the PWM entries comes precomputed from the host (glibc): lx = log x / log 2.
ALN_HD double aln_kld(const double* x, const double* lx, const double* lz) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (x[i] == 0) continue;
        acc = acc + x[i] * (lx[i] - lz[i]);
    }
    return acc;
}

ALN_HD double aln_jsd(const double* x, const double* lx, const double* y, const double* ly, double ln2) {
    double lz[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) lz[i] = log((x[i] + y[i]) / 2) / ln2;
    return 0.5 * aln_kld(x, lx, lz) + 0.5 * aln_kld(y, ly, lz);
}

ALN_HD bool aln_near(double a, double b) {
    if (a == b) return false;
    const double m = fmax(fabs(a), fabs(b));
    if (!(m < HUGE_VAL)) return false;                 // an infinite operand (the initial minimum) is never "near"
    return fabs(a - b) <= kAmbigRel * m;
}

// The distance of ONE offset i of `loop` (Alignment.hs:105-111): combFn over zipWith jsd a (drop i b), plus the gap penalty.
// It does not depend on the running minimum, so offsets can be evaluated in any order (or by different lanes).
ALN_HD double aln_offset_value(const AlignCfg& c, const double* a, const double* la, int na, const double* b, const double* lb, int nb,
                               int n1, int n2, int i) {
    const int nMatch = na < nb - i ? na : nb - i;
    // combFn over sc = zipWith jsd a (drop i b), streamed
    double s = 0.0, cc = 0.0, mx = 0.0;
    for (int k = 0; k < nMatch; ++k) {
        double v = aln_jsd(a + 4 * k, la + 4 * k, b + 4 * (i + k), lb + 4 * (i + k), c.ln2);
        if (c.comb_kind == 3) { mx = (k == 0 || v > mx) ? v : mx; continue; }
        if (c.comb_kind == 1) v = pow(v, 2.0);
        else if (c.comb_kind == 2) v = pow(v, 3.0);
        const double s1 = s + v;                               // Kahan-Babuska-Neumaier (Numeric.Sum.kbn)
        if (fabs(s) >= fabs(v)) cc = cc + ((s - s1) + v); else cc = cc + ((v - s1) + s);
        s = s1;
    }
    double comb;
    if (c.comb_kind == 3) comb = mx;
    else {
        const double mean = (s + cc) / (double)nMatch;
        comb = c.comb_kind == 0 ? mean : (c.comb_kind == 1 ? sqrt(mean) : pow(mean, 1.0 / 3.0));
    }
    const int nGap = n1 + n2 - 2 * nMatch;
    return comb + aln_penal(c, nGap, nMatch);
}

// loop opti (1/0,-1) a b 0   (Alignment.hs:103-112); a has na rows, b has nb rows.
ALN_HD void aln_loop(const AlignCfg& c, const double* opti, const double* a, const double* la, int na, const double* b,
                     const double* lb, int nb, int n1, int n2, double* dmin, int* imin, int* ambiguous) {
    double mn = HUGE_VAL;
    int mi = -1;
    for (int i = 0; i < nb; ++i) {
        const double o = opti[i];
        if (aln_near(o, mn)) *ambiguous = 1;
        if (o >= mn) break;                                        // Alignment.hs:104
        const double d = aln_offset_value(c, a, la, na, b, lb, nb, n1, n2, i);
        if (aln_near(d, mn)) *ambiguous = 1;
        if (d < mn) { mn = d; mi = i; }
    }
    *dmin = mn; *imin = mi;
}

// The decisions of `loop` replayed over offset distances computed beforehand (dv[i], i < nb): same early exit, same strict <,
// same near-tie flags as aln_loop.
ALN_HD void aln_loop_replay(const double* opti, const double* dv, int nb, double* dmin, int* imin, int* ambiguous) {
    double mn = HUGE_VAL;
    int mi = -1;
    for (int i = 0; i < nb; ++i) {
        const double o = opti[i];
        if (aln_near(o, mn)) *ambiguous = 1;
        if (o >= mn) break;
        const double d = dv[i];
        if (aln_near(d, mn)) *ambiguous = 1;
        if (d < mn) { mn = d; mi = i; }
    }
    *dmin = mn; *imin = mi;
}

// m1/l1: n1 x 4 (probabilities, log2); m2/l2 and the reverse complement m2r/l2r: n2 x 4.
// opti1 = optimalSc n1 n2 (n2 entries), opti2 = optimalSc n2 n1 (n1 entries).
ALN_HD void align_pair(const AlignCfg& c, const double* m1, const double* l1, int n1, const double* m2, const double* l2,
                       const double* m2r, const double* l2r, int n2, const double* opti1, const double* opti2,
                       double* d_out, int* same_dir, int* pos, int* ambiguous) {
    double df, dr; int pf, pr;
    {
        double d1, d2; int i1, i2;
        aln_loop(c, opti2, m2, l2, n2, m1, l1, n1, n1, n2, &d1, &i1, ambiguous);      // loop opti2 .. s2 s1
        aln_loop(c, opti1, m1, l1, n1, m2, l2, n2, n1, n2, &d2, &i2, ambiguous);      // loop opti1 .. s1 s2
        if (aln_near(d1, d2)) *ambiguous = 1;
        if (d1 < d2) { df = d1; pf = i1; } else { df = d2; pf = -i2; }               // Alignment.hs:92-93
    }
    {
        double d1, d2; int i1, i2;
        aln_loop(c, opti2, m2r, l2r, n2, m1, l1, n1, n1, n2, &d1, &i1, ambiguous);
        aln_loop(c, opti1, m1, l1, n1, m2r, l2r, n2, n1, n2, &d2, &i2, ambiguous);
        if (aln_near(d1, d2)) *ambiguous = 1;
        if (d1 < d2) { dr = d1; pr = i1; } else { dr = d2; pr = -i2; }               // Alignment.hs:97-98
    }
    if (aln_near(df, dr)) *ambiguous = 1;
    if (df <= dr) { *d_out = df; *same_dir = 1; *pos = pf; } else { *d_out = dr; *same_dir = 0; *pos = pr; }
}

// optimalSc x y (Alignment.hs:117-124): y entries.
inline void aln_optimal_sc(const AlignCfg& c, int x, int y, double* out) {
    for (int i = 0; i < y; ++i) {
        const int nM = x < y - i ? x : y - i;
        const int nG = i + (x > (y - i) ? x - (y - i) : (y - i) - x);
        out[i] = aln_penal(c, nG, nM);
    }
    for (int i = y - 2; i >= 0; --i) out[i] = out[i] < out[i + 1] ? out[i] : out[i + 1];   // scanr1 min
}

}  // namespace pwmscan
