// affinity.cu — per-hit p-value and BED score of scanMotif (Data/Bed/Utils.hs:109-123):
//     p     = 1 - cdf (motif's truncated CDF) score            cdf: Motif.hs:236-249
//     score = round (1000 / (1 + exp (-(-log10 (max 1e-20 p) - 5))))     toAffinity, round half even
// One thread per hit.  The cdf interpolation is plain IEEE fp64 (+, -, *, / — no FMA, same operation
// order as the reference), so p is bit-identical; log/exp are the CUDA libm, so a score whose
// pre-rounding value lies within 1e-6 of a rounding tie is flagged and settled on the host with glibc.
// HBM traffic: 12 B in + 8 (+8) B out per hit; the CDF points (a few hundred per motif) stay in L1/L2.
#include <cmath>
#include <vector>

#include "internal.h"

using namespace pwmscan;

namespace {

__device__ __forceinline__ double cdf_at(const double* __restrict__ xs, const double* __restrict__ ps, int n, double x) {
    if (n == 0) return 0.0;
    int lo = 0, hi = n;                          // first i with xs[i] >= x
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (xs[mid] < x) lo = mid + 1; else hi = mid; }
    if (lo >= n) return ps[n - 1];
    if (lo == 0) return x == xs[0] ? ps[0] : 0.0;
    const double a = xs[lo - 1], a1 = ps[lo - 1], b = xs[lo], b1 = ps[lo];
    const double al = __ddiv_rn(__dsub_rn(b, x), __dsub_rn(b, a));
    const double be = __ddiv_rn(__dsub_rn(x, a), __dsub_rn(b, a));
    return __dadd_rn(__dmul_rn(al, a1), __dmul_rn(be, b1));
}

__global__ void affinity_kernel(long long n, const int32_t* __restrict__ motif, const double* __restrict__ score,
                                const long long* __restrict__ off, const double* __restrict__ cx, const double* __restrict__ cy,
                                double ln10, double* __restrict__ out_p, long long* __restrict__ out_aff, unsigned char* __restrict__ near_tie) {
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        const int m = motif[k];
        const long long o = off[m];
        const double p = __dsub_rn(1.0, cdf_at(cx + o, cy + o, (int)(off[m + 1] - o), score[k]));
        const double x = -(log(fmax(1e-20, p)) / ln10);
        const double v = 1.0 / (1.0 + exp(-(x - 5.0))) * 1000.0;
        out_p[k] = p;
        out_aff[k] = (long long)rint(v);
        near_tie[k] = fabs(fabs(v - floor(v)) - 0.5) < 1e-6 ? 1 : 0;
    }
}

inline long long to_affinity_host(double p) {    // Utils.hs:120-123 with glibc, round half even
    const double x = -(std::log(std::fmax(1e-20, p)) / std::log(10.0));
    const double sc = 1.0 / (1.0 + std::exp(-(x - 5.0)));
    return (long long)std::nearbyint(sc * 1000.0);
}

}  // namespace

extern "C" int pwmscan_hits_affinity(pwmscan_ctx* ctx, const pwmscan_hits* hits, int32_t M, const int64_t* cdf_off,
                                     const double* cdf_x, const double* cdf_y, double* out_pvalue, int64_t* out_affinity) {
    if (!ctx || !hits || !cdf_off || M < 0) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_hits_affinity: NULL argument");
    const int64_t n = hits->count;
    if (n == 0) return PWMSCAN_OK;
    const int64_t npts = cdf_off[M];
    if (npts < 0 || (npts > 0 && (!cdf_x || !cdf_y))) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_hits_affinity: bad CDF arrays");
    {   // the per-hit motif column is one of the wide columns: expanded from the compact ones on first use
        int rc = hits_expand_wide(const_cast<pwmscan_hits*>(hits));
        if (rc) return rc;
    }
    for (int64_t k = 0; k < n; ++k)
        if (hits->motif[k] < 0 || hits->motif[k] >= M) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_hits_affinity: hit of a motif outside the CDF list");
    std::lock_guard<std::mutex> lk(ctx->mu);
    PWM_CUDA(cudaSetDevice(ctx->device));
    // inputs: [off (M+1) i64][cdf_x][cdf_y][score f64 x n][motif i32 x n]; outputs: [p f64 x n][aff i64 x n][flag u8 x n]
    const size_t b_off = (size_t)(M + 1) * 8, b_pts = (size_t)npts * 8, b_sc = (size_t)n * 8, b_mo = (size_t)n * 4;
    unsigned char* d_in = (unsigned char*)scratch_get(ctx, 1, b_off + 2 * b_pts + b_sc + b_mo + 64);
    unsigned char* d_out = (unsigned char*)scratch_get(ctx, 2, (size_t)n * 17 + 64);
    if (!d_in || !d_out) return PWMSCAN_ECUDA;
    long long* d_off = (long long*)d_in;
    double* d_cx = (double*)(d_in + b_off);
    double* d_cy = (double*)(d_in + b_off + b_pts);
    double* d_sc = (double*)(d_in + b_off + 2 * b_pts);
    int32_t* d_mo = (int32_t*)(d_in + b_off + 2 * b_pts + b_sc);
    double* d_p = (double*)d_out;
    long long* d_aff = (long long*)(d_out + (size_t)n * 8);
    unsigned char* d_flag = d_out + (size_t)n * 16;
    PWM_CUDA(cudaMemcpyAsync(d_off, cdf_off, b_off, cudaMemcpyHostToDevice, ctx->stream));
    if (npts > 0) {
        PWM_CUDA(cudaMemcpyAsync(d_cx, cdf_x, b_pts, cudaMemcpyHostToDevice, ctx->stream));
        PWM_CUDA(cudaMemcpyAsync(d_cy, cdf_y, b_pts, cudaMemcpyHostToDevice, ctx->stream));
    }
    PWM_CUDA(cudaMemcpyAsync(d_sc, hits->score, b_sc, cudaMemcpyHostToDevice, ctx->stream));      // the hit columns are pinned
    PWM_CUDA(cudaMemcpyAsync(d_mo, hits->motif, b_mo, cudaMemcpyHostToDevice, ctx->stream));
    PWM_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    const long long grid = std::min<long long>((n + 255) / 256, (long long)ctx->sm_count * 16);
    affinity_kernel<<<(unsigned)grid, 256, 0, ctx->stream>>>((long long)n, d_mo, d_sc, d_off, d_cx, d_cy, std::log(10.0), d_p, d_aff, d_flag);
    PWM_CUDA(cudaGetLastError());
    PWM_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    std::vector<double> h_p;
    std::vector<unsigned char> h_flag((size_t)n);
    double* p_host = out_pvalue;
    if (!p_host) { h_p.resize((size_t)n); p_host = h_p.data(); }
    PWM_CUDA(cudaMemcpyAsync(p_host, d_p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_affinity) PWM_CUDA(cudaMemcpyAsync(out_affinity, d_aff, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(h_flag.data(), d_flag, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    PWM_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->timing[10] = ms;
    if (out_affinity)
        for (int64_t k = 0; k < n; ++k)
            if (h_flag[(size_t)k]) out_affinity[k] = to_affinity_host(p_host[k]);
    return PWMSCAN_OK;
}
