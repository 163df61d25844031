// cdf.cu — scoreCDF / cdf' / pValueToScore on the device (Motif.hs:213-214, 252-270, 282-326).
//
// A CTA runs the whole column-by-column DP of one motif (two CTAs of 512 threads per SM; motifs are handed out longest
// first from a work counter).  The reference scatters `new[idx] += bg_b * prob` in the order (x ascending, b = A,C,G,T);
// fp64 addition is not associative, so to be bit-identical every output bin is computed by ONE thread that *gathers* its
// contributors in exactly that order.  idx is monotone in x for a fixed base, so the contributors of a bin are four
// contiguous runs of `prev`; their boundaries come from inverting the index function (cdf_core.h: a two-operation
// approximation with a certified error bound decides, the IEEE expression itself only within that bound of a bin edge).
// Per column the kernel reads `prev` and writes `new` — 2 x 1.6 MB — and nothing else: no compaction, no index planes,
// no atomics, no shared-memory staging (round 1 moved 12.8 MB of workspace per column through L2/HBM).  The prefix sum of toCDF (scanl1 (+)) is
// order-sensitive too and is done sequentially by one thread per motif out of shared memory; the other CTA of the SM
// keeps the SM busy meanwhile.  All fp64 arithmetic is IEEE (-fmad=false); log' p - log bg_b comes from the host libm.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include "internal.h"
#include "cdf_core.h"

using namespace pwmscan;

namespace {

constexpr int kCdfThreads = 512;
constexpr int kCdfCtasPerSm = 2;
constexpr int kMaxBin = 200000;        // Motif.hs:305
constexpr int kScanChunk = 2048;

constexpr size_t kWorkBytes = (size_t)kMaxBin * 8 * 3 + 256;      // prev / new, compressed P

// exclusive block scan of one int per thread; returns the exclusive prefix, *total = block sum
__device__ __forceinline__ int block_scan_excl(int v, int* total, int* s_warp /*[33]*/) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < kCdfThreads / 32 ? s_warp[lane] : 0;
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

// mode 0: thres_out[m] = cdf' cdf q ; mode 1: also the compressed CDF of the (single) motif to out_x/out_p ;
// mode 2: also truncateCDF trunc of every motif, appended to the pool out_x/out_p (pool[0] = next free entry, pool[1] =
// capacity; out_off[m - m0] = where, -1 if it did not fit — the caller then repeats with the exact capacity).
__global__ void __launch_bounds__(kCdfThreads, kCdfCtasPerSm)
cdf_kernel(int nwork, int m0, const int32_t* __restrict__ order, unsigned int* next_work, const int32_t* __restrict__ ncol,
           const int64_t* __restrict__ xat_off, const double* __restrict__ xat /* [col][4] = log' p - log bg */,
           const double* __restrict__ bg4, unsigned char* work, double q, int mode, double* __restrict__ thres_out,
           int* __restrict__ status_out, double* out_x, double* out_p, long long* out_n, long long* out_off,
           unsigned long long* pool, double trunc) {
    __shared__ int s_warp[33];
    __shared__ double s_chunk[kScanChunk];
    __shared__ int s_nbin, s_skip, s_xlo, s_xhi, s_nlo, s_nhi, s_m;
    __shared__ long long s_off;
    __shared__ double s_run;
    __shared__ CdfExact s_ex;
    const int tid = threadIdx.x, lane = tid & 31;
    double* const buf0 = reinterpret_cast<double*>(work + (size_t)blockIdx.x * kWorkBytes);
    double* const bufs[2] = {buf0, buf0 + kMaxBin};
    double* const cp = buf0 + 2 * (size_t)kMaxBin;
    const double bg[4] = {bg4[0], bg4[1], bg4[2], bg4[3]};
    const double m_epsilon = 2.220446049250313e-16;

    for (;;) {
        if (tid == 0) { const unsigned int k = atomicAdd(next_work, 1u); s_m = k < (unsigned)nwork ? order[k] : -1; }
        __syncthreads();
        const int m = s_m;
        if (m < 0) break;
        const int n = ncol[m];
        const double* xa = xat + 4 * xat_off[m];
        int cur = 0;                       // prev = bufs[cur]
        CdfScFn f; f.is_const = 1; f.step = 0.0; f.lo = 0.0;
        int plen = 1;
        if (tid == 0) { bufs[0][0] = 1.0; s_xlo = 0; s_xhi = 0; }
        __syncthreads();
        for (int col = 0; col < n; ++col) {
            const double* prev = bufs[cur];
            double* nw = bufs[cur ^ 1];
            // ---- column scalars (Motif.hs:303-310); lo' / hi' = scores of the first / last non-zero bin of prev
            if (tid == 0) {
                const double x0 = xa[4 * col], x1 = xa[4 * col + 1], x2 = xa[4 * col + 2], x3 = xa[4 * col + 3];
                double mn = x0, mx = x0;
                if (x1 < mn) mn = x1; if (x1 > mx) mx = x1;
                if (x2 < mn) mn = x2; if (x2 > mx) mx = x2;
                if (x3 < mn) mn = x3; if (x3 > mx) mx = x3;
                const double lo = cdf_scfn(f, s_xlo) + mn, hi = cdf_scfn(f, s_xhi) + mx;
                if (lo < hi) {
                    const double nbd = ceil((hi - lo) / 1e-5);
                    const int nbin = nbd > (double)kMaxBin ? kMaxBin : (int)nbd;
                    s_nbin = nbin; s_skip = 0;
                    s_ex.f = f; s_ex.xb[0] = x0; s_ex.xb[1] = x1; s_ex.xb[2] = x2; s_ex.xb[3] = x3;
                    s_ex.lo = lo; s_ex.step = (hi - lo) / (double)nbin;
                } else s_skip = 1;
                s_nlo = 0x7FFFFFFF; s_nhi = -1;
            }
            __syncthreads();
            if (s_skip) { __syncthreads(); continue; }                 // Motif.hs:299
            CdfFast q;
            cdf_fast_setup(q, s_ex, bg, s_nbin, s_xlo, s_xhi, plen);
            const CdfExact* ex = &s_ex;
            const int last = plen - 1;
            int my_lo = 0x7FFFFFFF, my_hi = -1;
            // ---- gather: a thread owns four adjacent output bins: five run boundaries per base, then the four runs of
            //      each bin merged in (x, base) order
#pragma unroll 1
            for (int j = 4 * tid; j < q.nbin; j += 4 * kCdfThreads) {
                int B[4][5];
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int g = 0; g < 5; ++g) B[k][g] = cdf_boundary(q, ex, k, j + g);
                double acc[4];
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    acc[g] = j + g < q.nbin ? cdf_gather_bin(prev, q, last, B[0][g], B[0][g + 1], B[1][g], B[1][g + 1], B[2][g], B[2][g + 1], B[3][g], B[3][g + 1]) : 0.0;
                if (j + 3 < q.nbin) {
                    *reinterpret_cast<double2*>(nw + j) = make_double2(acc[0], acc[1]);
                    *reinterpret_cast<double2*>(nw + j + 2) = make_double2(acc[2], acc[3]);
                } else {
#pragma unroll
                    for (int g = 0; g < 4; ++g) if (j + g < q.nbin) nw[j + g] = acc[g];
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) if (acc[g] != 0.0) { my_lo = min(my_lo, j + g); my_hi = max(my_hi, j + g); }
            }
            my_lo = __reduce_min_sync(0xFFFFFFFFu, my_lo);
            my_hi = __reduce_max_sync(0xFFFFFFFFu, my_hi);
            if (lane == 0) { atomicMin(&s_nlo, my_lo); atomicMax(&s_nhi, my_hi); }
            __syncthreads();
            if (tid == 0) { s_xlo = s_nlo; s_xhi = s_nhi; }
            cur ^= 1;
            f.is_const = 0; f.step = s_ex.step; f.lo = s_ex.lo;        // Motif.hs:298
            plen = q.nbin;
            __syncthreads();
        }
        // ---- toCDF: sequential scanl1 (+) (Motif.hs:315), in place
        const int len = plen;
        double* P = bufs[cur];
        for (int c0 = 0; c0 < len; c0 += kScanChunk) {
            const int cl = min(kScanChunk, len - c0);
            for (int i = tid; i < cl; i += kCdfThreads) s_chunk[i] = P[c0 + i];
            __syncthreads();
            if (tid == 0) {
                double run = c0 == 0 ? 0.0 : s_run;
                int i = 0;
                if (c0 == 0) { run = s_chunk[0]; i = 1; }
                for (; i < cl; ++i) { run = run + s_chunk[i]; s_chunk[i] = run; }
                s_run = run;
            }
            __syncthreads();
            for (int i = tid; i < cl; i += kCdfThreads) P[c0 + i] = s_chunk[i];
            __syncthreads();
        }
        // ---- compressCDF (Motif.hs:316-321) into (cx, cp)
        double* cx = bufs[cur ^ 1];
        int base = 0;
        for (int i0 = 0; i0 < len; i0 += kCdfThreads) {
            const int i = i0 + tid;
            int keep = 0;
            double pi = 0.0;
            if (i < len) {
                pi = P[i];
                if (i == 0) keep = 1;
                else {
                    keep = (pi - P[i - 1]) > m_epsilon;
                    // right neighbour: the reference reads index i+1 unchecked; for the last element
                    // that is out of bounds (undefined) and is treated as "no difference"
                    if (!keep && i + 1 < len) keep = (P[i + 1] - pi) > m_epsilon;
                }
            }
            int tot;
            const int off = block_scan_excl(keep, &tot, s_warp);
            if (keep) { cx[base + off] = cdf_scfn(f, i); cp[base + off] = pi; }
            base += tot;
        }
        const int nc = base;
        __syncthreads();
        if (mode == 1) {
            for (int i = tid; i < nc; i += kCdfThreads) { out_x[i] = cx[i]; out_p[i] = cp[i]; }
            if (tid == 0) *out_n = nc;
        }
        if (mode == 2) {
            // truncateCDF trunc (Motif.hs:273-274): points with P >= trunc; P is non-decreasing, so a suffix
            int lo_i = 0, hi_i = nc;
            while (lo_i < hi_i) { const int mid = (lo_i + hi_i) >> 1; if (cp[mid] >= trunc) hi_i = mid; else lo_i = mid + 1; }
            const int first = lo_i, cnt = nc - first;
            if (tid == 0) {
                const unsigned long long o = atomicAdd(&pool[0], (unsigned long long)cnt);
                s_off = o + (unsigned long long)cnt <= pool[1] ? (long long)o : -1;
                out_n[m - m0] = cnt; out_off[m - m0] = s_off;
            }
            __syncthreads();
            const long long o = s_off;
            if (o >= 0) for (int i = first + tid; i < nc; i += kCdfThreads) { out_x[o + i - first] = cx[i]; out_p[o + i - first] = cp[i]; }
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
                    const double al = (b1 - q) / (b1 - a1);
                    const double be = (q - a1) / (b1 - a1);
                    res = al * a + be * b;
                }
            }
            thres_out[m] = res;
            status_out[m] = st;
        }
        __syncthreads();
    }
}

// Runs the kernel over the motifs [m0, m0 + mcount).  mode 2: *x_out / *p_out receive the truncated CDFs back to back
// (malloc'ed), n_out[k] their lengths.
int run_cdf(pwmscan_ctx* ctx, const pwmscan_motifs* mo, int m0, int mcount, double q, int mode, double* thres_out,
            double** x_out, double** p_out, int64_t* n_out, double trunc = 0.0) {
    PWM_CUDA(cudaSetDevice(ctx->device));
    const int M = mo->M;
    // host: x_at = log' p - log bg   (Motif.hs:311-314,324-325), libm
    std::vector<int32_t> h_n(M), h_order(mcount);
    std::vector<int64_t> h_off(M);
    int64_t cols = 0;
    for (int m = 0; m < M; ++m) { h_n[m] = mo->mh[m].n; h_off[m] = cols; cols += mo->mh[m].n; }
    std::iota(h_order.begin(), h_order.end(), m0);
    std::stable_sort(h_order.begin(), h_order.end(), [&](int a, int b) { return h_n[a] > h_n[b]; });      // longest first
    std::vector<double> h_xat((size_t)cols * 4);
    double lbg[4];
    for (int b = 0; b < 4; ++b) lbg[b] = std::log(mo->bg[b]);
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < h_n[m]; ++k)
            for (int b = 0; b < 4; ++b) {
                const double p = mo->mh[m].mat[0][4 * k + b];
                h_xat[(size_t)(h_off[m] + k) * 4 + b] = (p == 0 ? std::log(0.001) : std::log(p)) - lbg[b];
            }
    for (double v : h_xat) if (!std::isfinite(v)) return set_err(ctx, PWMSCAN_EDOMAIN, "scoreCDF: a background frequency or matrix entry gives a non-finite log-odds score");
    static const int grid_env = std::getenv("PWMSCAN_CDF_CTAS") ? std::max(1, std::atoi(std::getenv("PWMSCAN_CDF_CTAS"))) : 0;     // experiments
    const int grid = std::min(mcount, grid_env ? std::min(grid_env, kCdfCtasPerSm * ctx->sm_count) : kCdfCtasPerSm * ctx->sm_count);
    auto al = [](size_t b) { return (b + 63) & ~(size_t)63; };
    const size_t b_xat = al(h_xat.size() * 8), b_off = al((size_t)M * 8), b_n = al((size_t)M * 4), b_th = al((size_t)M * 8), b_st = al((size_t)M * 4);
    const size_t b_ord = al((size_t)mcount * 4), b_cnt = al((size_t)mcount * 8);
    unsigned char* d_small = (unsigned char*)scratch_get(ctx, 1, b_xat + b_off + b_th + 64 + b_n + b_st + b_ord + 2 * b_cnt + 64);
    unsigned char* d_work = (unsigned char*)scratch_get(ctx, 2, kWorkBytes * (size_t)grid);
    if (!d_small || !d_work) return PWMSCAN_ECUDA;
    unsigned char* q0 = d_small;
    double* d_xat = (double*)q0; q0 += b_xat;
    int64_t* d_off = (int64_t*)q0; q0 += b_off;
    double* d_th = (double*)q0; q0 += b_th;
    double* d_bg = (double*)q0; q0 += 64;
    int32_t* d_n = (int32_t*)q0; q0 += b_n;
    int* d_st = (int*)q0; q0 += b_st;
    int32_t* d_order = (int32_t*)q0; q0 += b_ord;
    long long* d_nout = (long long*)q0; q0 += b_cnt;
    long long* d_oout = (long long*)q0; q0 += b_cnt;
    unsigned long long* d_pool = (unsigned long long*)q0;              // [0] next free entry [1] capacity [2] (32-bit) work counter
    // Once an asynchronous copy into a local buffer is enqueued, every exit first waits for the stream.
    auto fail = [&](int code) { cudaStreamSynchronize(ctx->stream); return code; };
#define CDF_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail(cuda_fail(ctx, e__, #call)); } while (0)
    std::vector<double> h_th(M);
    std::vector<int> h_st(M);
    std::vector<long long> cnt(mcount), off(mcount);
    // output pool of mode 2: the truncated CDFs are a small suffix of the points (scores above the (1 - trunc) quantile);
    // 48 k entries per motif on average, the exact size on the (rare) second pass
    unsigned long long pool_cap = mode == 2 ? (unsigned long long)mcount * 49152ull : (mode == 1 ? (unsigned long long)kMaxBin : 0ull);
    for (int pass = 0; pass < 2; ++pass) {
        double *d_ox = nullptr, *d_op = nullptr;
        if (mode >= 1) {
            d_ox = (double*)scratch_get(ctx, 3, (size_t)pool_cap * 8);
            d_op = (double*)scratch_get(ctx, 4, (size_t)pool_cap * 8);
            if (!d_ox || !d_op) return fail(PWMSCAN_ECUDA);
        }
        const unsigned long long h_pool[3] = {0ull, pool_cap, 0ull};
        CDF_CUDA(cudaMemcpyAsync(d_xat, h_xat.data(), h_xat.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        CDF_CUDA(cudaMemcpyAsync(d_off, h_off.data(), (size_t)M * 8, cudaMemcpyHostToDevice, ctx->stream));
        CDF_CUDA(cudaMemcpyAsync(d_n, h_n.data(), (size_t)M * 4, cudaMemcpyHostToDevice, ctx->stream));
        CDF_CUDA(cudaMemcpyAsync(d_order, h_order.data(), (size_t)mcount * 4, cudaMemcpyHostToDevice, ctx->stream));
        CDF_CUDA(cudaMemcpyAsync(d_bg, mo->bg, 32, cudaMemcpyHostToDevice, ctx->stream));
        CDF_CUDA(cudaMemcpyAsync(d_pool, h_pool, 24, cudaMemcpyHostToDevice, ctx->stream));
        CDF_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
        cdf_kernel<<<grid, kCdfThreads, 0, ctx->stream>>>(mcount, m0, d_order, (unsigned int*)(d_pool + 2), d_n, d_off, d_xat, d_bg, d_work, q, mode, d_th, d_st,
                                                          d_ox, d_op, d_nout, d_oout, d_pool, trunc);
        CDF_CUDA(cudaGetLastError());
        CDF_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
        CDF_CUDA(cudaMemcpyAsync(h_th.data() + m0, d_th + m0, (size_t)mcount * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CDF_CUDA(cudaMemcpyAsync(h_st.data() + m0, d_st + m0, (size_t)mcount * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (mode >= 1) CDF_CUDA(cudaMemcpyAsync(cnt.data(), d_nout, (size_t)(mode == 2 ? mcount : 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (mode == 2) CDF_CUDA(cudaMemcpyAsync(off.data(), d_oout, (size_t)mcount * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CDF_CUDA(cudaStreamSynchronize(ctx->stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
        ctx->timing[10] = pass == 0 ? ms : ctx->timing[10] + ms;
        if (mode == 1) {
            const long long nout = cnt[0];
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
        if (mode == 0) return PWMSCAN_OK;
        // mode 2: gather the truncated CDFs of this batch
        long long tot = 0;
        bool fits = true;
        for (int k = 0; k < mcount; ++k) { tot += cnt[k]; fits = fits && off[k] >= 0; }
        if (!fits) {                                      // more points than the pool holds: once more with the exact size
            if (pass == 1) return set_err(ctx, PWMSCAN_ERANGE, "pwmscan_cutoff_cdfs: internal: output pool overflow");
            pool_cap = (unsigned long long)tot + 64;
            continue;
        }
        double* hx = (double*)std::malloc((size_t)std::max<long long>(tot, 1) * 8);
        double* hp = (double*)std::malloc((size_t)std::max<long long>(tot, 1) * 8);
        if (!hx || !hp) { std::free(hx); std::free(hp); return set_err(ctx, PWMSCAN_EINVAL, "out of host memory"); }
        // one copy of the used part of the pool, then put the motifs in order on the host (the pool is in completion order)
        std::vector<double> px((size_t)std::max<long long>(tot, 1)), pp((size_t)std::max<long long>(tot, 1));
        cudaError_t e1 = cudaMemcpy(px.data(), d_ox, (size_t)tot * 8, cudaMemcpyDeviceToHost);
        cudaError_t e2 = cudaMemcpy(pp.data(), d_op, (size_t)tot * 8, cudaMemcpyDeviceToHost);
        if (e1 != cudaSuccess || e2 != cudaSuccess) { std::free(hx); std::free(hp); return cuda_fail(ctx, e1 != cudaSuccess ? e1 : e2, "D2H truncated cdf"); }
        long long o = 0;
        for (int k = 0; k < mcount; ++k) {
            std::memcpy(hx + o, px.data() + off[k], (size_t)cnt[k] * 8);
            std::memcpy(hp + o, pp.data() + off[k], (size_t)cnt[k] * 8);
            n_out[k] = cnt[k];
            o += cnt[k];
        }
        *x_out = hx; *p_out = hp;
        return PWMSCAN_OK;
    }
    return PWMSCAN_OK;
#undef CDF_CUDA
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
    // one launch per batch of up to 4096 motifs (the CTAs draw motifs from a work counter, longest first); the truncated
    // CDFs land compacted in a pool (run_cdf), which is released again when it grew large
    const int M = mo->M, batch = 4096;
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
