// dense.cu — dense window scoring: scores / scores' / score (Motif.hs:127-149, scoreHelp 164-193)
// and maxMatchSc (Motif/Search.hs:98-109).  IUPAC-aware ('N' scores 0, any other non-IUPAC byte is
// the reference's `error`).  fp64, one sequential left-to-right sum per window (never fused).
//
// Roofline: scores writes 8 B per window (HBM-write bound: <= hbm_gbs/8 windows/s) and does n fp64
// adds per window; maxMatchSc is fp64-add bound (n adds per window, no output traffic).
#include <algorithm>
#include <cstdlib>
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

__device__ __forceinline__ double key_to_double(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)b);
}

// grid.y = motif; every block reduces the maximum window score of its stripe of the windows [w0, w1).
// PRUNE = false: every window summed to the end.  PRUNE = true (the windows behind a first stretch whose maximum is already in
// out_keys): the reference's own device (Search.hs:98-109: lookAheadSearch against the running maximum) — a window is left as
// soon as its partial sum plus the best its remaining columns can still add stays below the best seen so far.  "So far" is
// only ever the first stretch and this thread's own earlier windows, i.e. windows BEFORE the current one, so every column the
// reference's sequential loop visits is visited here too (an invalid letter it would hit is hit); the bound carries a margin
// of 1e-9 (1 + |best| + sum of the column maxima), far above the rounding of the sums, so no window that could be the maximum
// is left early and the result is the same maximum of the same fp64 sums.
template <bool PRUNE>
__global__ void max_kernel(const uint8_t* __restrict__ ascii, int64_t len, int64_t w0, int64_t w1, const int32_t* __restrict__ ncol,
                           const int64_t* __restrict__ tab_off, const double* __restrict__ tab16, const unsigned long long* __restrict__ bound_keys,
                           unsigned long long* __restrict__ out_keys, unsigned long long* counters) {
    __shared__ double stab[PWMSCAN_MAX_COLS * 16];
    __shared__ double ssuf[PWMSCAN_MAX_COLS];
    __shared__ double s_abs;
    __shared__ unsigned long long s_max[32];
    const int m = blockIdx.y;
    const int n = ncol[m];
    const double* tb = tab16 + tab_off[2 * m];
    for (int i = threadIdx.x; i < n * 16; i += blockDim.x) stab[i] = tb[i];
    __syncthreads();
    if (PRUNE && threadIdx.x == 0) {           // best the columns behind k can still add (any of the 15 letter codes), and the margin's scale
        double suf = 0.0, a = 0.0;
        for (int k = n - 1; k >= 0; --k) {
            ssuf[k] = suf;
            double mx = stab[16 * k];
            for (int c = 1; c < 15; ++c) mx = fmax(mx, stab[16 * k + c]);
            suf += mx; a += fabs(mx);
        }
        s_abs = a;
    }
    __syncthreads();
    int64_t nwin = len - n + 1;
    if (nwin > w1) nwin = w1;
    bool bad = false;
    const double ninf = __longlong_as_double((long long)0xFFF0000000000000ull);
    double best = ninf, cut = ninf;
    if (PRUNE) {
        best = bound_keys[m] == 0ull ? ninf : key_to_double(bound_keys[m]);       // maximum of the first stretch (a snapshot: not what later windows add)
        cut = best - 1e-9 * (1.0 + fabs(best) + s_abs);
    }
    for (int64_t i = w0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwin; i += (int64_t)gridDim.x * blockDim.x) {
        if (!PRUNE) {
            const double v = window_score(ascii, i, n, stab, &bad);
            best = v > best ? v : best;
        } else {
            double acc = 0.0;
            bool alive = true;
            for (int k = 0; k < n; ++k) {
                const unsigned code = c_lut[ascii[i + k]];
                if (code == 15u) bad = true;
                acc = __dadd_rn(acc, stab[16 * k + code]);
                if (acc + ssuf[k] < cut) { alive = false; break; }
            }
            if (alive && acc > best) { best = acc; cut = best - 1e-9 * (1.0 + fabs(best) + s_abs); }
        }
    }
    unsigned long long bk = best == ninf ? 0ull : order_key(best);                // key 0 sorts below every double ("no window")
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long v = __shfl_xor_sync(0xFFFFFFFFu, bk, o); bk = v > bk ? v : bk; }
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = bk;
    __syncthreads();
    if (threadIdx.x < 32) {
        bk = threadIdx.x < (blockDim.x >> 5) ? s_max[threadIdx.x] : 0ull;
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long v = __shfl_xor_sync(0xFFFFFFFFu, bk, o); bk = v > bk ? v : bk; }
        if (threadIdx.x == 0 && bk != 0ull) atomicMax(&out_keys[m], bk);
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
    unsigned char* d = (unsigned char*)scratch_get(ctx, 1, b_n + b_off + 2 * b_key + 64);
    if (!d) return PWMSCAN_ECUDA;
    int64_t* d_off = (int64_t*)d;
    unsigned long long* d_key = (unsigned long long*)(d + b_off);
    unsigned long long* d_key0 = (unsigned long long*)(d + b_off + b_key);      // the first stretch's maxima, frozen
    int32_t* d_n = (int32_t*)(d + b_off + 2 * b_key);
    PWM_CUDA(cudaMemcpyAsync(d_off, mo->tab_off.data(), b_off, cudaMemcpyHostToDevice, ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(d_n, h_n.data(), b_n, cudaMemcpyHostToDevice, ctx->stream));
    PWM_CUDA(cudaMemsetAsync(d_key, 0, b_key, ctx->stream));   // key 0 sorts below every double
    PWM_CUDA(cudaMemsetAsync(ctx->d_counters, 0, 4 * sizeof(unsigned long long), ctx->stream));
    PWM_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    const int bs = 256;
    long long gx = std::min<long long>((seq->len + bs - 1) / bs, (long long)ctx->sm_count * 8);
    if (gx < 1) gx = 1;
    // a first stretch of windows summed to the end gives every motif a maximum to prune the rest against
    static const bool no_prune = std::getenv("PWMSCAN_MAX_NO_PRUNE") != nullptr;      // experiments: the round-1 form
    const int64_t first = no_prune ? seq->len : std::min<int64_t>(seq->len, 65536);
    for (int m0 = 0; m0 < M; m0 += 65535) {
        // grid.y is limited to 65535; PWMSCAN_MAX_MOTIFS keeps this to one launch
        const unsigned gy = (unsigned)std::min(M - m0, 65535);
        const long long g1 = std::min<long long>((first + bs - 1) / bs, gx);
        max_kernel<false><<<dim3((unsigned)std::max<long long>(g1, 1), gy), bs, 0, ctx->stream>>>(seq->d_ascii, seq->len, 0, first, d_n + m0, d_off + 2 * m0, mo->d_tab16,
                                                                                                     nullptr, d_key + m0, ctx->d_counters);
        PWM_CUDA(cudaGetLastError());
        if (seq->len > first) {
            PWM_CUDA(cudaMemcpyAsync(d_key0 + m0, d_key + m0, (size_t)gy * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            max_kernel<true><<<dim3((unsigned)gx, gy), bs, 0, ctx->stream>>>(seq->d_ascii, seq->len, first, seq->len, d_n + m0, d_off + 2 * m0, mo->d_tab16,
                                                                             d_key0 + m0, d_key + m0, ctx->d_counters);
            PWM_CUDA(cudaGetLastError());
        }
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
