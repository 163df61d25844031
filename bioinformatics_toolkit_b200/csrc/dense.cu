// dense.cu — dense window scoring: scores / scores' / score (Motif.hs:127-149, scoreHelp 164-193)
// and maxMatchSc (Motif/Search.hs:98-109).  IUPAC-aware ('N' scores 0, any other non-IUPAC byte is
// the reference's `error`).  fp64, one sequential left-to-right sum per window (never fused).
//
// Roofline: scores writes 8 B per window (HBM-write bound: <= hbm_gbs/8 windows/s) and does n fp64
// adds per window; maxMatchSc is fp64-add bound (n adds per window, no output traffic).
#include <cstring>
#include <new>

#include "internal.h"

using namespace pwmscan;

namespace {

__constant__ unsigned char c_lut[256];
int ensure_lut(pwmscan_ctx* ctx) {      // ctx->mu held
    if (ctx->dense_lut_ready) return PWMSCAN_OK;
    unsigned char lut[256];
    for (int i = 0; i < 256; ++i) lut[i] = (unsigned char)iupac_code((unsigned char)i);
    PWM_CUDA(cudaMemcpyToSymbol(c_lut, lut, 256));
    ctx->dense_lut_ready = true;
    return PWMSCAN_OK;
}

__device__ __forceinline__ double window_score(const uint8_t* __restrict__ ascii, int64_t i, int n, const double* stab, bool* bad) {
    double acc = 0.0;
    for (int k = 0; k < n; ++k) {
        const unsigned code = c_lut[ascii[i + k]];
        if (code == 15u) *bad = true;
        acc = __dadd_rn(acc, stab[16 * k + code]);
    }
    return acc;
}

__global__ void scores_kernel(const uint8_t* __restrict__ ascii, int n, const double* __restrict__ tab16,
                              double* __restrict__ out, int64_t nwin, unsigned long long* counters) {
    __shared__ double stab[PWMSCAN_MAX_COLS * 16];
    for (int i = threadIdx.x; i < n * 16; i += blockDim.x) stab[i] = tab16[i];
    __syncthreads();
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwin; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = window_score(ascii, i, n, stab, &bad);
    if (bad) atomicOr(&counters[2], 1ull);
}

__device__ __forceinline__ unsigned long long order_key(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// grid.y = motif; every block reduces the maximum window score of its stripe of windows.
__global__ void max_kernel(const uint8_t* __restrict__ ascii, int64_t len, const int32_t* __restrict__ ncol,
                           const int64_t* __restrict__ tab_off, const double* __restrict__ tab16,
                           unsigned long long* __restrict__ out_keys, unsigned long long* counters) {
    __shared__ double stab[PWMSCAN_MAX_COLS * 16];
    __shared__ unsigned long long s_max[32];
    const int m = blockIdx.y;
    const int n = ncol[m];
    const double* tb = tab16 + tab_off[2 * m];
    for (int i = threadIdx.x; i < n * 16; i += blockDim.x) stab[i] = tb[i];
    __syncthreads();
    const int64_t nwin = len - n + 1;
    bool bad = false;
    unsigned long long best = order_key(__longlong_as_double((long long)0xFFF0000000000000ull));
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwin; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long k = order_key(window_score(ascii, i, n, stab, &bad));
        best = k > best ? k : best;
    }
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long v = __shfl_xor_sync(0xFFFFFFFFu, best, o); best = v > best ? v : best; }
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        best = threadIdx.x < (blockDim.x >> 5) ? s_max[threadIdx.x] : 0ull;
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long v = __shfl_xor_sync(0xFFFFFFFFu, best, o); best = v > best ? v : best; }
        if (threadIdx.x == 0) atomicMax(&out_keys[m], best);
    }
    if (bad) atomicOr(&counters[2], 1ull);
}

}  // namespace

extern "C" int pwmscan_scores(pwmscan_ctx* ctx, const pwmscan_seq* seq, const pwmscan_motifs* mo, int32_t m, int strand,
                              double* out, int64_t out_len) {
    if (!ctx || !seq || !mo || (!out && out_len > 0)) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_scores: NULL argument");
    if (m < 0 || m >= mo->M || strand < 0 || strand > 1) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_scores: bad motif / strand");
    if (!seq->d_ascii) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_scores: sequence must be uploaded with PWMSCAN_SEQ_KEEP_ASCII");
    if (seq->nseg != 1) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_scores: single-segment sequences only");
    const int n = mo->mh[m].n;
    const int64_t nwin = seq->len - n + 1 > 0 ? seq->len - n + 1 : 0;
    if (out_len < nwin) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_scores: output buffer too small");
    if (nwin == 0) return PWMSCAN_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    PWM_CUDA(cudaSetDevice(ctx->device));
    int rc = ensure_lut(ctx);
    if (rc) return rc;
    double* d_out = (double*)scratch_get(ctx, 1, (size_t)nwin * 8);
    if (!d_out) return PWMSCAN_ECUDA;
    PWM_CUDA(cudaMemsetAsync(ctx->d_counters, 0, 4 * sizeof(unsigned long long), ctx->stream));
    PWM_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    const int bs = 256;
    const long long grid = std::min<long long>((nwin + bs - 1) / bs, (long long)ctx->sm_count * 32);
    scores_kernel<<<(unsigned)grid, bs, 0, ctx->stream>>>(seq->d_ascii, n, mo->d_tab16 + mo->tab_off[2 * m + strand], d_out, nwin, ctx->d_counters);
    PWM_CUDA(cudaGetLastError());
    PWM_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(out, d_out, (size_t)nwin * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PWM_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->timing[10] = ms;
    if (ctx->h_counters[2] != 0) return set_err(ctx, PWMSCAN_ENUCL, "Bio.Motif.score: invalid nucleotide");
    return PWMSCAN_OK;
}

extern "C" int pwmscan_max_match_sc(pwmscan_ctx* ctx, const pwmscan_seq* seq, const pwmscan_motifs* mo, double* out) {
    if (!ctx || !seq || !mo || !out) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_max_match_sc: NULL argument");
    if (!seq->d_ascii) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_max_match_sc: sequence must be uploaded with PWMSCAN_SEQ_KEEP_ASCII");
    if (seq->nseg != 1) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_max_match_sc: single-segment sequences only");
    std::lock_guard<std::mutex> lk(ctx->mu);
    PWM_CUDA(cudaSetDevice(ctx->device));
    int rc = ensure_lut(ctx);
    if (rc) return rc;
    const int M = mo->M;
    std::vector<int32_t> h_n(M);
    for (int m = 0; m < M; ++m) h_n[m] = mo->mh[m].n;
    const size_t b_n = (size_t)M * 4, b_off = (size_t)2 * M * 8, b_key = (size_t)M * 8;
    unsigned char* d = (unsigned char*)scratch_get(ctx, 1, b_n + b_off + b_key + 64);
    if (!d) return PWMSCAN_ECUDA;
    int64_t* d_off = (int64_t*)d;
    unsigned long long* d_key = (unsigned long long*)(d + b_off);
    int32_t* d_n = (int32_t*)(d + b_off + b_key);
    PWM_CUDA(cudaMemcpyAsync(d_off, mo->tab_off.data(), b_off, cudaMemcpyHostToDevice, ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(d_n, h_n.data(), b_n, cudaMemcpyHostToDevice, ctx->stream));
    PWM_CUDA(cudaMemsetAsync(d_key, 0, b_key, ctx->stream));   // key 0 sorts below every double
    PWM_CUDA(cudaMemsetAsync(ctx->d_counters, 0, 4 * sizeof(unsigned long long), ctx->stream));
    PWM_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    const int bs = 256;
    long long gx = std::min<long long>((seq->len + bs - 1) / bs, (long long)ctx->sm_count * 8);
    if (gx < 1) gx = 1;
    for (int m0 = 0; m0 < M; m0 += 65535) {
        // grid.y is limited to 65535; PWMSCAN_MAX_MOTIFS keeps this to one launch
        dim3 grid((unsigned)gx, (unsigned)std::min(M - m0, 65535));
        max_kernel<<<grid, bs, 0, ctx->stream>>>(seq->d_ascii, seq->len, d_n + m0, d_off + 2 * m0, mo->d_tab16, d_key + m0, ctx->d_counters);
        PWM_CUDA(cudaGetLastError());
    }
    PWM_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    std::vector<unsigned long long> h_key(M);
    PWM_CUDA(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(h_key.data(), d_key, b_key, cudaMemcpyDeviceToHost, ctx->stream));
    PWM_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->timing[10] = ms;
    if (ctx->h_counters[2] != 0) return set_err(ctx, PWMSCAN_ENUCL, "Bio.Motif.Search.lookAheadSearch: invalid nucleotide");
    for (int m = 0; m < M; ++m) {
        const unsigned long long k = h_key[m];
        unsigned long long b;
        if (k == 0) b = 0xFFF0000000000000ull;                       // no window: loop (-1/0) returns -inf
        else b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
        double v;
        std::memcpy(&v, &b, 8);
        out[m] = v;
    }
    return PWMSCAN_OK;
}
