// cdf.cu — scoreCDF / cdf' / pValueToScore on the device (Motif.hs:213-214, 252-270, 282-326).
//
// One persistent CTA per motif runs the whole column-by-column DP.  The reference scatters
// `new[idx] += bg_b * prob` in the order (x ascending, b = A,C,G,T); fp64 addition is not
// associative, so to be bit-identical every output bin is computed by ONE thread that *gathers*
// its contributors in exactly that order (idx is monotone in x for a fixed base, so the
// contributors of a bin are four contiguous runs of the compacted non-zero list, merged by
// (x, b)).  No atomics.  The prefix sum of toCDF (scanl1 (+)) is order-sensitive too and is done
// sequentially by one thread per motif out of shared memory.  All fp64 arithmetic is IEEE (file is
// compiled with -fmad=false); log' p - log bg_b comes from the host libm.
//
// Work per motif: n columns x (<= 200 000 bins x 4 bases) index evaluations (one fp64 divide each)
// + the gather; ~12 MB of L2/HBM-resident workspace per CTA.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "internal.h"

using namespace pwmscan;

namespace {

constexpr int kCdfThreads = 1024;
constexpr int kMaxBin = 200000;        // Motif.hs:305
constexpr int kScanChunk = 2048;

struct CdfWork {
    double* buf[2];   // prev / new                      kMaxBin doubles each
    double* pk;       // compacted probabilities / P     kMaxBin
    int* xk;          // compacted bin indices           kMaxBin
    int* J;           // [4][kMaxBin] bin index per (base, k)
    int* start;       // [4][kMaxBin] first k of each run
};

constexpr size_t kWorkBytes = (size_t)kMaxBin * (8 * 3 + 4 + 16 + 16) + 256;

__device__ __forceinline__ CdfWork carve(unsigned char* base) {
    CdfWork w;
    w.buf[0] = reinterpret_cast<double*>(base);
    w.buf[1] = w.buf[0] + kMaxBin;
    w.pk = w.buf[1] + kMaxBin;
    w.xk = reinterpret_cast<int*>(w.pk + kMaxBin);
    w.J = w.xk + kMaxBin;
    w.start = w.J + 4 * kMaxBin;
    return w;
}

// exclusive block scan of one int per thread; returns the exclusive prefix, *total = block sum
__device__ __forceinline__ int block_scan_excl(int v, int* total, int* s_warp /*[33]*/) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = s_warp[lane];
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, winc, o); if (lane >= o) winc += t; }
        s_warp[lane] = winc - w;
        if (lane == 31) s_warp[32] = winc;
    }
    __syncthreads();
    const int res = s_warp[wid] + inc - v;
    *total = s_warp[32];
    __syncthreads();
    return res;
}

struct ScFn { int is_const; double step, lo; };
__device__ __forceinline__ double scfn(const ScFn& f, int x) {
    return f.is_const ? 0.0 : __dadd_rn(__dmul_rn(__dadd_rn((double)x, 0.5), f.step), f.lo);   // (x + 0.5) * step + lo
}

// mode 0: thres_out[m] = cdf' cdf q ; mode 1: also write the compressed CDF of motif m0 to out_x/out_p.
__global__ void __launch_bounds__(kCdfThreads, 1)
cdf_kernel(int M, int m0, const int32_t* __restrict__ ncol, const int64_t* __restrict__ xat_off,
           const double* __restrict__ xat /* [col][4] = log' p - log bg */, const double* __restrict__ bg4,
           unsigned char* work, double q, int mode, double* __restrict__ thres_out, int* __restrict__ status_out,
           double* out_x, double* out_p, long long* out_n, double trunc) {
    __shared__ int s_warp[33];
    __shared__ double s_chunk[kScanChunk];
    __shared__ int s_K, s_nbin, s_skip, s_plen;
    __shared__ double s_lo, s_step, s_run;
    const int tid = threadIdx.x;
    const CdfWork w = carve(work + (size_t)blockIdx.x * kWorkBytes);
    const double bg[4] = {bg4[0], bg4[1], bg4[2], bg4[3]};
    const double m_epsilon = 2.220446049250313e-16;

    for (int m = m0 + blockIdx.x; m < M; m += gridDim.x) {
        const int n = ncol[m];
        const double* xa = xat + 4 * xat_off[m];
        int cur = 0;                       // prev = buf[cur]
        ScFn f; f.is_const = 1; f.step = 0.0; f.lo = 0.0;
        if (tid == 0) { w.buf[0][0] = 1.0; s_plen = 1; }
        __syncthreads();
        for (int col = 0; col < n; ++col) {
            const int plen = s_plen;
            const double* prev = w.buf[cur];
            double* nw = w.buf[cur ^ 1];
            // ---- compact the non-zero bins of prev (the reference skips prob == 0, Motif.hs:291)
            int base = 0;
            for (int x0 = 0; x0 < plen; x0 += kCdfThreads) {
                const int x = x0 + tid;
                const double pv = x < plen ? prev[x] : 0.0;
                const int flag = pv != 0.0;
                int tot;
                const int off = block_scan_excl(flag, &tot, s_warp);
                if (flag) { w.xk[base + off] = x; w.pk[base + off] = pv; }
                base += tot;
            }
            const int K = base;
            __syncthreads();
            // ---- column scalars (Motif.hs:303-310)
            if (tid == 0) {
                const double x0 = xa[4 * col], x1 = xa[4 * col + 1], x2 = xa[4 * col + 2], x3 = xa[4 * col + 3];
                double mn = x0, mx = x0;
                if (x1 < mn) mn = x1; if (x1 > mx) mx = x1;
                if (x2 < mn) mn = x2; if (x2 > mx) mx = x2;
                if (x3 < mn) mn = x3; if (x3 > mx) mx = x3;
                const double lo1 = scfn(f, w.xk[0]), hi1 = scfn(f, w.xk[K - 1]);
                const double lo = __dadd_rn(lo1, mn), hi = __dadd_rn(hi1, mx);
                if (lo < hi) {
                    const double nbd = ceil(__ddiv_rn(__dsub_rn(hi, lo), 1e-5));
                    const int nbin = nbd > (double)kMaxBin ? kMaxBin : (int)nbd;
                    s_nbin = nbin; s_lo = lo; s_step = __ddiv_rn(__dsub_rn(hi, lo), (double)nbin); s_skip = 0;
                } else s_skip = 1;
                s_K = K;
            }
            __syncthreads();
            if (s_skip) continue;                                      // Motif.hs:299
            const int nbin = s_nbin;
            const double lo = s_lo, step = s_step;
            // ---- map: bin index of every (k, base)   (getIdx, Motif.hs:301-302)
#pragma unroll
            for (int b = 0; b < 4; ++b)
                for (int j = tid; j < nbin; j += kCdfThreads) w.start[b * kMaxBin + j] = -1;
            for (int k = tid; k < K; k += kCdfThreads) {
                const double sc = scfn(f, w.xk[k]);
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const double v = __ddiv_rn(__dsub_rn(__dadd_rn(sc, xa[4 * col + b]), lo), step);
                    long long j = (long long)v;                        // truncate
                    if (j >= nbin) j = nbin - 1;
                    if (j < 0) j = 0;                                  // cannot happen (v >= 0 up to rounding towards 0)
                    w.J[b * kMaxBin + k] = (int)j;
                }
            }
            __syncthreads();
            for (int k = tid; k < K; k += kCdfThreads)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int j = w.J[b * kMaxBin + k];
                    if (k == 0 || w.J[b * kMaxBin + k - 1] != j) w.start[b * kMaxBin + j] = k;
                }
            __syncthreads();
            // ---- gather: one thread per output bin, contributors merged in (x, base) order
            for (int j = tid; j < nbin; j += kCdfThreads) {
                int kb[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) kb[b] = w.start[b * kMaxBin + j];
                double acc = 0.0;
                for (;;) {
                    int bsel = -1, ksel = 0x7FFFFFFF;
#pragma unroll
                    for (int b = 0; b < 4; ++b) if (kb[b] >= 0 && kb[b] < ksel) { ksel = kb[b]; bsel = b; }
                    if (bsel < 0) break;
                    acc = __dadd_rn(__dmul_rn(bg[bsel], w.pk[ksel]), acc);            // bg_b * prob + new[idx]
                    const int kn = ksel + 1;
                    kb[bsel] = (kn < K && w.J[bsel * kMaxBin + kn] == j) ? kn : -1;
                }
                nw[j] = acc;
            }
            __syncthreads();
            cur ^= 1;
            f.is_const = 0; f.step = step; f.lo = lo;                  // Motif.hs:298
            if (tid == 0) s_plen = nbin;
            __syncthreads();
        }
        // ---- toCDF: sequential scanl1 (+) (Motif.hs:315), P kept in w.pk
        const int len = s_plen;
        const double* v = w.buf[cur];
        double* P = w.pk;
        for (int c0 = 0; c0 < len; c0 += kScanChunk) {
            const int cl = min(kScanChunk, len - c0);
            for (int i = tid; i < cl; i += kCdfThreads) s_chunk[i] = v[c0 + i];
            __syncthreads();
            if (tid == 0) {
                double run = c0 == 0 ? 0.0 : s_run;
                int i = 0;
                if (c0 == 0) { run = s_chunk[0]; i = 1; }
                for (; i < cl; ++i) { run = __dadd_rn(run, s_chunk[i]); s_chunk[i] = run; }
                s_run = run;
            }
            __syncthreads();
            for (int i = tid; i < cl; i += kCdfThreads) P[c0 + i] = s_chunk[i];
            __syncthreads();
        }
        // ---- compressCDF (Motif.hs:316-321) into (cx, cp) = (buf[cur^1], buf[cur] reused after P is final)
        double* cx = w.buf[cur ^ 1];
        double* cp = reinterpret_cast<double*>(w.J);                   // 4*kMaxBin ints = 2*kMaxBin doubles
        int base = 0;
        for (int i0 = 0; i0 < len; i0 += kCdfThreads) {
            const int i = i0 + tid;
            int keep = 0;
            double pi = 0.0;
            if (i < len) {
                pi = P[i];
                if (i == 0) keep = 1;
                else {
                    keep = __dsub_rn(pi, P[i - 1]) > m_epsilon;
                    // right neighbour: the reference reads index i+1 unchecked; for the last element
                    // that is out of bounds (undefined) and is treated as "no difference"
                    if (!keep && i + 1 < len) keep = __dsub_rn(P[i + 1], pi) > m_epsilon;
                }
            }
            int tot;
            const int off = block_scan_excl(keep, &tot, s_warp);
            if (keep) { cx[base + off] = scfn(f, i); cp[base + off] = pi; }
            base += tot;
        }
        const int nc = base;
        __syncthreads();
        if (mode == 1 && m == m0) {
            for (int i = tid; i < nc; i += kCdfThreads) { out_x[i] = cx[i]; out_p[i] = cp[i]; }
            if (tid == 0) *out_n = nc;
        }
        if (mode == 2) {
            // truncateCDF trunc (Motif.hs:273-274): points with P >= trunc; P is non-decreasing, so a suffix
            int lo_i = 0, hi_i = nc;
            while (lo_i < hi_i) { const int mid = (lo_i + hi_i) >> 1; if (cp[mid] >= trunc) hi_i = mid; else lo_i = mid + 1; }
            const int first = lo_i;
            double* ox = out_x + (size_t)(m - m0) * kMaxBin;
            double* op = out_p + (size_t)(m - m0) * kMaxBin;
            for (int i = first + tid; i < nc; i += kCdfThreads) { ox[i - first] = cx[i]; op[i - first] = cp[i]; }
            if (tid == 0) out_n[m - m0] = nc - first;
        }
        // ---- cdf' (Motif.hs:252-270)
        if (tid == 0) {
            int st = 0;
            double res = 0.0;
            if (q > 1 || q < 0) st = 1;
            else {
                int lo_i = 0, hi_i = nc;                                // lower_bound of q in cp
                while (lo_i < hi_i) { const int mid = (lo_i + hi_i) >> 1; if (cp[mid] < q) lo_i = mid + 1; else hi_i = mid; }
                const int i = lo_i;
                if (i >= nc) res = __longlong_as_double(0x7FF0000000000000LL);
                else if (i == 0) { if (q == cp[0]) res = cx[0]; else st = 2; }
                else {
                    const double a = cx[i - 1], a1 = cp[i - 1], b = cx[i], b1 = cp[i];
                    const double al = __ddiv_rn(__dsub_rn(b1, q), __dsub_rn(b1, a1));
                    const double be = __ddiv_rn(__dsub_rn(q, a1), __dsub_rn(b1, a1));
                    res = __dadd_rn(__dmul_rn(al, a), __dmul_rn(be, b));
                }
            }
            thres_out[m] = res;
            status_out[m] = st;
        }
        __syncthreads();
    }
}

int run_cdf(pwmscan_ctx* ctx, const pwmscan_motifs* mo, int m0, int mcount, double q, int mode, double* thres_out,
            double** x_out, double** p_out, int64_t* n_out, double trunc = 0.0) {
    PWM_CUDA(cudaSetDevice(ctx->device));
    const int M = mo->M;
    // host: x_at = log' p - log bg   (Motif.hs:311-314,324-325), libm
    std::vector<int32_t> h_n(M);
    std::vector<int64_t> h_off(M);
    int64_t cols = 0;
    for (int m = 0; m < M; ++m) { h_n[m] = mo->mh[m].n; h_off[m] = cols; cols += mo->mh[m].n; }
    std::vector<double> h_xat((size_t)cols * 4);
    double lbg[4];
    for (int b = 0; b < 4; ++b) lbg[b] = std::log(mo->bg[b]);
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < h_n[m]; ++k)
            for (int b = 0; b < 4; ++b) {
                const double p = mo->mh[m].mat[0][4 * k + b];
                h_xat[(size_t)(h_off[m] + k) * 4 + b] = (p == 0 ? std::log(0.001) : std::log(p)) - lbg[b];
            }
    const int grid = std::min(mcount, ctx->sm_count);
    const size_t b_xat = h_xat.size() * 8, b_off = (size_t)M * 8, b_n = (size_t)M * 4, b_th = (size_t)M * 8, b_st = (size_t)M * 4;
    unsigned char* d_small = (unsigned char*)scratch_get(ctx, 1, b_xat + b_off + b_th + 32 + b_n + b_st + 64 + 8);
    unsigned char* d_work = (unsigned char*)scratch_get(ctx, 2, kWorkBytes * (size_t)grid);
    if (!d_small || !d_work) return PWMSCAN_ECUDA;
    double* d_xat = (double*)d_small;
    int64_t* d_off = (int64_t*)(d_small + b_xat);
    double* d_th = (double*)(d_small + b_xat + b_off);
    double* d_bg = (double*)(d_small + b_xat + b_off + b_th);
    int32_t* d_n = (int32_t*)(d_small + b_xat + b_off + b_th + 32 + 8);
    int* d_st = (int*)(d_small + b_xat + b_off + b_th + 32 + 8 + b_n);
    long long* d_nout = (long long*)scratch_get(ctx, 5, (size_t)(mode == 2 ? mcount : 1) * 8);
    if (!d_nout) return PWMSCAN_ECUDA;
    double *d_ox = nullptr, *d_op = nullptr;
    if (mode >= 1) {
        const size_t cnt = mode == 2 ? (size_t)mcount : 1;
        d_ox = (double*)scratch_get(ctx, 3, cnt * kMaxBin * 8);
        d_op = (double*)scratch_get(ctx, 4, cnt * kMaxBin * 8);
        if (!d_ox || !d_op) return PWMSCAN_ECUDA;
    }
    PWM_CUDA(cudaMemcpyAsync(d_xat, h_xat.data(), b_xat, cudaMemcpyHostToDevice, ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(d_off, h_off.data(), b_off, cudaMemcpyHostToDevice, ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(d_n, h_n.data(), b_n, cudaMemcpyHostToDevice, ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(d_bg, mo->bg, 32, cudaMemcpyHostToDevice, ctx->stream));
    PWM_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    cdf_kernel<<<grid, kCdfThreads, 0, ctx->stream>>>(m0 + mcount, m0, d_n, d_off, d_xat, d_bg, d_work, q, mode, d_th, d_st, d_ox, d_op, d_nout, trunc);
    PWM_CUDA(cudaGetLastError());
    PWM_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    std::vector<double> h_th(M);
    std::vector<int> h_st(M);
    long long nout = 0;
    PWM_CUDA(cudaMemcpyAsync(h_th.data() + m0, d_th + m0, (size_t)mcount * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(h_st.data() + m0, d_st + m0, (size_t)mcount * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (mode == 1) PWM_CUDA(cudaMemcpyAsync(&nout, d_nout, 8, cudaMemcpyDeviceToHost, ctx->stream));
    PWM_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->timing[10] = ms;
    if (mode == 1) {
        double* hx = (double*)std::malloc((size_t)std::max<long long>(nout, 1) * 8);
        double* hp = (double*)std::malloc((size_t)std::max<long long>(nout, 1) * 8);
        if (!hx || !hp) { std::free(hx); std::free(hp); return set_err(ctx, PWMSCAN_EINVAL, "out of host memory"); }
        cudaError_t e1 = cudaMemcpy(hx, d_ox, (size_t)nout * 8, cudaMemcpyDeviceToHost);
        cudaError_t e2 = cudaMemcpy(hp, d_op, (size_t)nout * 8, cudaMemcpyDeviceToHost);
        if (e1 != cudaSuccess || e2 != cudaSuccess) { std::free(hx); std::free(hp); return cuda_fail(ctx, e1 != cudaSuccess ? e1 : e2, "D2H cdf"); }
        *x_out = hx; *p_out = hp; *n_out = nout;
        return PWMSCAN_OK;
    }
    for (int m = m0; m < m0 + mcount; ++m) {
        if (h_st[m] == 1) return set_err(ctx, PWMSCAN_EDOMAIN, "p must be in [0,1]");
        if (h_st[m] == 2) return set_err(ctx, PWMSCAN_EDOMAIN, "cdf': impossible!");
        thres_out[m] = h_th[m];
    }
    if (mode == 2) {
        // gather the truncated CDFs of this batch: n_out[0..mcount) counts, *x_out/*p_out appended by the caller
        std::vector<long long> cnt(mcount);
        PWM_CUDA(cudaMemcpy(cnt.data(), d_nout, (size_t)mcount * 8, cudaMemcpyDeviceToHost));
        long long tot = 0;
        for (int k = 0; k < mcount; ++k) { n_out[k] = cnt[k]; tot += cnt[k]; }
        double* hx = (double*)std::malloc((size_t)std::max<long long>(tot, 1) * 8);
        double* hp = (double*)std::malloc((size_t)std::max<long long>(tot, 1) * 8);
        if (!hx || !hp) { std::free(hx); std::free(hp); return set_err(ctx, PWMSCAN_EINVAL, "out of host memory"); }
        long long o = 0;
        for (int k = 0; k < mcount; ++k) {
            if (cnt[k] > 0) {
                cudaMemcpyAsync(hx + o, d_ox + (size_t)k * kMaxBin, (size_t)cnt[k] * 8, cudaMemcpyDeviceToHost, ctx->stream);
                cudaMemcpyAsync(hp + o, d_op + (size_t)k * kMaxBin, (size_t)cnt[k] * 8, cudaMemcpyDeviceToHost, ctx->stream);
            }
            o += cnt[k];
        }
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { std::free(hx); std::free(hp); return cuda_fail(ctx, e, "D2H truncated cdf"); }
        *x_out = hx; *p_out = hp;
    }
    return PWMSCAN_OK;
}

}  // namespace

extern "C" int pwmscan_score_cdf(pwmscan_ctx* ctx, const pwmscan_motifs* mo, int32_t m, double** x, double** P, int64_t* npts) {
    if (!ctx || !mo || !x || !P || !npts) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_score_cdf: NULL argument");
    if (m < 0 || m >= mo->M) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_score_cdf: bad motif index");
    std::lock_guard<std::mutex> lk(ctx->mu);
    return run_cdf(ctx, mo, m, 1, 0.5, 1, nullptr, x, P, npts);
}

extern "C" int pwmscan_pvalue_to_score(pwmscan_ctx* ctx, const pwmscan_motifs* mo, double p, double* thres_out) {
    if (!ctx || !mo || !thres_out) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_pvalue_to_score: NULL argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    return run_cdf(ctx, mo, 0, mo->M, 1 - p, 0, thres_out, nullptr, nullptr, nullptr);
}

// mkCutoffMotif for a whole motif set (Data/Bed/Utils.hs:88-98): cutoff = cdf' cdf (1 - p) on the
// full CDF, plus truncateCDF (1 - 10 p) cdf for the per-hit p-values of scanMotif.
extern "C" int pwmscan_cutoff_cdfs(pwmscan_ctx* ctx, const pwmscan_motifs* mo, double p, double* thres_out, double** x, double** P,
                                   int64_t* offsets /* M+1 */) {
    if (!ctx || !mo || !thres_out || !x || !P || !offsets) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_cutoff_cdfs: NULL argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    // one launch per batch; a batch is a multiple of the SM count (one persistent CTA per SM walks its motifs),
    // bounded by the output staging (2 x batch x 1.6 MB of device memory)
    const int M = mo->M, batch = 6 * std::max(1, ctx->sm_count);
    std::vector<double> ax, ap;
    std::vector<int64_t> cnt(M);
    offsets[0] = 0;
    for (int m0 = 0; m0 < M; m0 += batch) {
        const int mc = std::min(batch, M - m0);
        double *bx = nullptr, *bp = nullptr;
        int rc = run_cdf(ctx, mo, m0, mc, 1 - p, 2, thres_out, &bx, &bp, cnt.data() + m0, 1 - p * 10);
        if (rc) return rc;
        int64_t tot = 0;
        for (int k = 0; k < mc; ++k) tot += cnt[m0 + k];
        ax.insert(ax.end(), bx, bx + tot);
        ap.insert(ap.end(), bp, bp + tot);
        std::free(bx); std::free(bp);
    }
    for (int m = 0; m < M; ++m) offsets[m + 1] = offsets[m] + cnt[m];
    double* hx = (double*)std::malloc(std::max<size_t>(ax.size(), 1) * 8);
    double* hp = (double*)std::malloc(std::max<size_t>(ap.size(), 1) * 8);
    if (!hx || !hp) { std::free(hx); std::free(hp); return set_err(ctx, PWMSCAN_EINVAL, "out of host memory"); }
    std::memcpy(hx, ax.data(), ax.size() * 8);
    std::memcpy(hp, ap.data(), ap.size() * 8);
    *x = hx; *P = hp;
    return PWMSCAN_OK;
}
