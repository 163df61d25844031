// align.cu — PWM-PWM alignment for many pairs on the device (Motif/Alignment.hs:83-132; the
// all-pairs initialisation and row refresh of iterativeMerge, Motif/Merge.hs:138-145,161-165;
// `merge_motifs dist`, app/MergeMotifs.hs:158-169).  One thread per pair runs alignmentBy
// (align_core.h); pairs whose decisive comparisons were within 1e-10 relative are re-evaluated on
// the host with the same function compiled against glibc, so (same_dir, pos) follow the
// reference's tie rules and the distance of those pairs is the reference's bit for bit.
// Bound: fp64 / SFU (four `log` per column pair); bytes are negligible.
#include <cmath>
#include <cstring>

#include "align_core.h"
#include "internal.h"

using namespace pwmscan;

namespace {

struct AlignDev {
    const double* mats;      // per (motif, strand): n x 4 probabilities
    const double* logs;      // same layout: log2 of the entries (host glibc)
    const int64_t* col_off;  // [2m+s] column offset
    const int32_t* ncol;     // [m]
    const double* opti;      // [(x * 65 + y) * 64 + i] = optimalSc x y !! i
};

__device__ __forceinline__ void run_pair(const AlignCfg& cfg, const AlignDev& d, int i, int j, double* dist,
                                         uint8_t* same_dir, int32_t* pos, uint8_t* amb, int64_t k) {
    const int n1 = d.ncol[i], n2 = d.ncol[j];
    const int64_t o1 = 4 * d.col_off[2 * i], o2 = 4 * d.col_off[2 * j], o2r = 4 * d.col_off[2 * j + 1];
    double dd; int sd, pp, am = 0;
    align_pair(cfg, d.mats + o1, d.logs + o1, n1, d.mats + o2, d.logs + o2, d.mats + o2r, d.logs + o2r, n2,
               d.opti + ((size_t)n1 * 65 + n2) * 64, d.opti + ((size_t)n2 * 65 + n1) * 64, &dd, &sd, &pp, &am);
    dist[k] = dd; same_dir[k] = (uint8_t)sd; pos[k] = pp; amb[k] = (uint8_t)am;
}

__global__ void align_all_kernel(AlignCfg cfg, AlignDev d, int M, double* dist, uint8_t* same_dir, int32_t* pos, uint8_t* amb) {
    const int i = blockIdx.y;
    const int j = i + 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    const int64_t k = (int64_t)i * M - (int64_t)i * (i + 1) / 2 + (j - i - 1);
    run_pair(cfg, d, i, j, dist, same_dir, pos, amb, k);
}

__global__ void align_list_kernel(AlignCfg cfg, AlignDev d, int64_t npairs, const int32_t* a, const int32_t* b, double* dist,
                                  uint8_t* same_dir, int32_t* pos, uint8_t* amb) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= npairs) return;
    run_pair(cfg, d, a[k], b[k], dist, same_dir, pos, amb, k);
}

// Launches the pair kernel on prepared device inputs, copies the results back and settles near-ties on the host with
// the same function (host mirrors h_mats / h_logs / h_opti in the device layout).  pa == nullptr: all i < j.
int launch_and_collect(pwmscan_ctx* ctx, const AlignCfg& cfg, const AlignDev& dev, int M, const double* h_mats, const double* h_logs,
                       const int64_t* col_off, const int32_t* h_n, const double* h_opti, int64_t npairs, const int32_t* pa,
                       const int32_t* pb, double* dist, uint8_t* same_dir, int32_t* pos) {
    const bool all = (pa == nullptr);
    const size_t b_out = (size_t)npairs * (8 + 4 + 1 + 1) + 64;
    unsigned char* d_out = (unsigned char*)scratch_get(ctx, 2, b_out + (all ? 0 : (size_t)npairs * 8));
    if (!d_out) return PWMSCAN_ECUDA;
    double* d_dist = (double*)d_out;
    int32_t* d_pos = (int32_t*)(d_out + (size_t)npairs * 8);
    uint8_t* d_sd = (uint8_t*)(d_out + (size_t)npairs * 12);
    uint8_t* d_amb = d_sd + npairs;
    int32_t *d_pa = nullptr, *d_pb = nullptr;
    if (!all) {
        d_pa = (int32_t*)(d_out + ((b_out + 7) & ~(size_t)7));
        d_pb = d_pa + npairs;
        PWM_CUDA(cudaMemcpyAsync(d_pa, pa, (size_t)npairs * 4, cudaMemcpyHostToDevice, ctx->stream));
        PWM_CUDA(cudaMemcpyAsync(d_pb, pb, (size_t)npairs * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    PWM_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    const int bs = 128;
    if (all) {
        dim3 grid((unsigned)((M - 1 + bs - 1) / bs), (unsigned)(M - 1));
        align_all_kernel<<<grid, bs, 0, ctx->stream>>>(cfg, dev, M, d_dist, d_sd, d_pos, d_amb);
    } else {
        align_list_kernel<<<(unsigned)((npairs + bs - 1) / bs), bs, 0, ctx->stream>>>(cfg, dev, npairs, d_pa, d_pb, d_dist, d_sd, d_pos, d_amb);
    }
    PWM_CUDA(cudaGetLastError());
    PWM_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    std::vector<uint8_t> h_amb((size_t)npairs);
    PWM_CUDA(cudaMemcpyAsync(dist, d_dist, (size_t)npairs * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(pos, d_pos, (size_t)npairs * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(same_dir, d_sd, (size_t)npairs, cudaMemcpyDeviceToHost, ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(h_amb.data(), d_amb, (size_t)npairs, cudaMemcpyDeviceToHost, ctx->stream));
    PWM_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    // near-ties: same function, host libm
    int64_t namb = 0;
    auto redo = [&](int64_t k, int i, int j) {
        const int n1 = h_n[i], n2 = h_n[j];
        const size_t o1 = 4 * (size_t)col_off[2 * i], o2 = 4 * (size_t)col_off[2 * j], o2r = 4 * (size_t)col_off[2 * j + 1];
        double dd; int sd, pp, am = 0;
        align_pair(cfg, h_mats + o1, h_logs + o1, n1, h_mats + o2, h_logs + o2, h_mats + o2r, h_logs + o2r, n2,
                   h_opti + ((size_t)n1 * 65 + n2) * 64, h_opti + ((size_t)n2 * 65 + n1) * 64, &dd, &sd, &pp, &am);
        dist[k] = dd; same_dir[k] = (uint8_t)sd; pos[k] = pp;
        ++namb;
    };
    if (all) {
        int64_t k = 0;
        for (int i = 0; i < M; ++i) for (int j = i + 1; j < M; ++j, ++k) if (h_amb[k]) redo(k, i, j);
    } else {
        for (int64_t k = 0; k < npairs; ++k) if (h_amb[k]) redo(k, pa[k], pb[k]);
    }
    ctx->timing[10] = ms; ctx->timing[11] = (double)namb;
    return PWMSCAN_OK;
}

int run_align(pwmscan_ctx* ctx, const pwmscan_motifs* mo, int64_t npairs, const int32_t* pa, const int32_t* pb,
              int penalty_kind, double gap, int combine_kind, double* dist, uint8_t* same_dir, int32_t* pos) {
    if (penalty_kind < 0 || penalty_kind > 3 || combine_kind < 0 || combine_kind > 3)
        return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_align: unknown penalty / combine kind");
    const int M = mo->M;
    const bool all = (pa == nullptr);
    if (all) npairs = (int64_t)M * (M - 1) / 2;
    if (npairs == 0) return PWMSCAN_OK;
    if (!all) for (int64_t k = 0; k < npairs; ++k)
        if (pa[k] < 0 || pa[k] >= M || pb[k] < 0 || pb[k] >= M) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_align_pairs: motif index out of range");
    PWM_CUDA(cudaSetDevice(ctx->device));
    AlignCfg cfg;
    cfg.penal_kind = penalty_kind; cfg.comb_kind = combine_kind; cfg.gap = gap; cfg.ln2 = std::log(2.0);
    // host: log2 of every entry (glibc), optimalSc table for the lengths present
    const size_t ncols = (size_t)mo->total_cols2;
    std::vector<double> h_mats(ncols * 4), h_logs(ncols * 4);
    std::vector<int32_t> h_n(M);
    bool have_len[65] = {false};
    for (int m = 0; m < M; ++m) {
        h_n[m] = mo->mh[m].n; have_len[h_n[m]] = true;
        for (int s = 0; s < 2; ++s) {
            const size_t o = 4 * (size_t)mo->col_off[2 * m + s];
            for (int e = 0; e < 4 * h_n[m]; ++e) {
                const double x = mo->mh[m].mat[s][e];
                h_mats[o + e] = x;
                h_logs[o + e] = std::log(x) / cfg.ln2;            // logBase 2 x (never read when x == 0)
            }
        }
    }
    std::vector<double> h_opti((size_t)65 * 65 * 64, 0.0);
    for (int x = 1; x <= 64; ++x) if (have_len[x])
        for (int y = 1; y <= 64; ++y) if (have_len[y]) aln_optimal_sc(cfg, x, y, h_opti.data() + ((size_t)x * 65 + y) * 64);
    const size_t b_m = ncols * 32, b_opti = h_opti.size() * 8, b_off = (size_t)2 * M * 8, b_n = (size_t)M * 4;
    unsigned char* d_in = (unsigned char*)scratch_get(ctx, 1, 2 * b_m + b_opti + b_off + b_n + 64);
    if (!d_in) return PWMSCAN_ECUDA;
    double* d_mats = (double*)d_in;
    double* d_logs = (double*)(d_in + b_m);
    double* d_opti = (double*)(d_in + 2 * b_m);
    int64_t* d_off = (int64_t*)(d_in + 2 * b_m + b_opti);
    int32_t* d_n = (int32_t*)(d_in + 2 * b_m + b_opti + b_off);
    PWM_CUDA(cudaMemcpyAsync(d_mats, h_mats.data(), b_m, cudaMemcpyHostToDevice, ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(d_logs, h_logs.data(), b_m, cudaMemcpyHostToDevice, ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(d_opti, h_opti.data(), b_opti, cudaMemcpyHostToDevice, ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(d_off, mo->col_off.data(), b_off, cudaMemcpyHostToDevice, ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(d_n, h_n.data(), b_n, cudaMemcpyHostToDevice, ctx->stream));
    AlignDev dev{d_mats, d_logs, d_off, d_n, d_opti};
    return launch_and_collect(ctx, cfg, dev, M, h_mats.data(), h_logs.data(), mo->col_off.data(), h_n.data(), h_opti.data(), npairs, pa, pb,
                              dist, same_dir, pos);
}

}  // namespace

extern "C" int pwmscan_align_all_pairs(pwmscan_ctx* ctx, const pwmscan_motifs* mo, int penalty_kind, double gap, int combine_kind,
                                       double* dist, uint8_t* same_dir, int32_t* pos) {
    if (!ctx || !mo || !dist || !same_dir || !pos) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_align_all_pairs: NULL argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    return run_align(ctx, mo, 0, nullptr, nullptr, penalty_kind, gap, combine_kind, dist, same_dir, pos);
}

extern "C" int pwmscan_align_pairs(pwmscan_ctx* ctx, const pwmscan_motifs* mo, int64_t npairs, const int32_t* a, const int32_t* b,
                                   int penalty_kind, double gap, int combine_kind, double* dist, uint8_t* same_dir, int32_t* pos) {
    if (!ctx || !mo || npairs < 0 || (npairs > 0 && (!a || !b || !dist || !same_dir || !pos)))
        return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_align_pairs: bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    return run_align(ctx, mo, npairs, a, b, penalty_kind, gap, combine_kind, dist, same_dir, pos);
}

// ------------------------------------------------------------------------------------------------
// Resident alignment set: the PWMs of iterativeMerge stay on the device (Merge.hs:113-170).  The
// all-pairs initialisation (161-165) and every row refresh (138-145) run on the same device arrays;
// a merge replaces one slot in place (its matrix, reverse complement and log2 table) instead of
// re-uploading every motif.  Slots have a fixed stride of 64 columns; the optimalSc table covers
// every length pair up to 64 x 64, so merged PWMs of new lengths need nothing else.
// ------------------------------------------------------------------------------------------------
struct pwmscan_alnset {
    pwmscan_ctx* ctx = nullptr;
    int32_t M = 0;
    AlignCfg cfg;
    std::vector<int32_t> n;           // columns per slot
    std::vector<int64_t> col_off;     // [2m+s] = (2m+s) * 64 (column offsets, fixed stride)
    std::vector<double> mats, logs;   // host mirror, [(2m+s) * 256]
    std::vector<double> opti;         // [(x * 65 + y) * 64 + i]
    double *d_mats = nullptr, *d_logs = nullptr, *d_opti = nullptr;
    int64_t* d_off = nullptr;
    int32_t* d_n = nullptr;
};

namespace {

void alnset_fill_slot(pwmscan_alnset* a, int32_t m, int32_t ncol, const double* mat) {
    a->n[m] = ncol;
    double* f = a->mats.data() + (size_t)(2 * m) * 256;
    double* r = f + 256;
    std::memset(f, 0, 512 * sizeof(double));
    for (int e = 0; e < 4 * ncol; ++e) { f[e] = mat[e]; r[e] = mat[4 * ncol - 1 - e]; }      // rcPWM: reversed flattening (Motif.hs:77-78)
    double* lf = a->logs.data() + (size_t)(2 * m) * 256;
    for (int e = 0; e < 512; ++e) lf[e] = 0.0;
    for (int e = 0; e < 4 * ncol; ++e) { lf[e] = std::log(f[e]) / a->cfg.ln2; lf[256 + e] = std::log(r[e]) / a->cfg.ln2; }
}

}  // namespace

extern "C" int pwmscan_alnset_create(pwmscan_ctx* ctx, int32_t M, const int32_t* ncol, const double* mats, int penalty_kind, double gap,
                                     int combine_kind, pwmscan_alnset** out) {
    if (!ctx || !out || M < 1 || !ncol || !mats) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_alnset_create: bad argument");
    *out = nullptr;
    if (penalty_kind < 0 || penalty_kind > 3 || combine_kind < 0 || combine_kind > 3)
        return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_align: unknown penalty / combine kind");
    for (int m = 0; m < M; ++m)
        if (ncol[m] < 1 || ncol[m] > PWMSCAN_MAX_COLS) return set_err(ctx, PWMSCAN_ERANGE, "pwmscan_alnset_create: motif length outside 1..64");
    std::lock_guard<std::mutex> lk(ctx->mu);
    PWM_CUDA(cudaSetDevice(ctx->device));
    pwmscan_alnset* a = new (std::nothrow) pwmscan_alnset();
    if (!a) return set_err(ctx, PWMSCAN_EINVAL, "out of host memory");
    a->ctx = ctx; a->M = M;
    a->cfg.penal_kind = penalty_kind; a->cfg.comb_kind = combine_kind; a->cfg.gap = gap; a->cfg.ln2 = std::log(2.0);
    a->n.resize(M); a->col_off.resize((size_t)2 * M);
    a->mats.assign((size_t)2 * M * 256, 0.0); a->logs.assign((size_t)2 * M * 256, 0.0);
    const double* p = mats;
    for (int m = 0; m < M; ++m) {
        a->col_off[2 * m] = (int64_t)(2 * m) * 64; a->col_off[2 * m + 1] = (int64_t)(2 * m + 1) * 64;
        alnset_fill_slot(a, m, ncol[m], p);
        p += 4 * ncol[m];
    }
    a->opti.assign((size_t)65 * 65 * 64, 0.0);
    for (int x = 1; x <= 64; ++x)
        for (int y = 1; y <= 64; ++y) aln_optimal_sc(a->cfg, x, y, a->opti.data() + ((size_t)x * 65 + y) * 64);
    cudaError_t e;
    auto bail = [&](cudaError_t err, const char* what) { const int rc = cuda_fail(ctx, err, what); pwmscan_alnset_free(a); return rc; };
    const size_t b_m = a->mats.size() * 8;
    if ((e = cudaMalloc((void**)&a->d_mats, b_m)) != cudaSuccess) return bail(e, "cudaMalloc(alnset mats)");
    if ((e = cudaMalloc((void**)&a->d_logs, b_m)) != cudaSuccess) return bail(e, "cudaMalloc(alnset logs)");
    if ((e = cudaMalloc((void**)&a->d_opti, a->opti.size() * 8)) != cudaSuccess) return bail(e, "cudaMalloc(alnset opti)");
    if ((e = cudaMalloc((void**)&a->d_off, a->col_off.size() * 8)) != cudaSuccess) return bail(e, "cudaMalloc(alnset off)");
    if ((e = cudaMalloc((void**)&a->d_n, (size_t)M * 4)) != cudaSuccess) return bail(e, "cudaMalloc(alnset n)");
    if ((e = cudaMemcpy(a->d_mats, a->mats.data(), b_m, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "upload alnset");
    if ((e = cudaMemcpy(a->d_logs, a->logs.data(), b_m, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "upload alnset");
    if ((e = cudaMemcpy(a->d_opti, a->opti.data(), a->opti.size() * 8, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "upload alnset");
    if ((e = cudaMemcpy(a->d_off, a->col_off.data(), a->col_off.size() * 8, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "upload alnset");
    if ((e = cudaMemcpy(a->d_n, a->n.data(), (size_t)M * 4, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "upload alnset");
    *out = a;
    return PWMSCAN_OK;
}

extern "C" void pwmscan_alnset_free(pwmscan_alnset* a) {
    if (!a) return;
    if (a->ctx) cudaSetDevice(a->ctx->device);
    if (a->d_mats) cudaFree(a->d_mats);
    if (a->d_logs) cudaFree(a->d_logs);
    if (a->d_opti) cudaFree(a->d_opti);
    if (a->d_off) cudaFree(a->d_off);
    if (a->d_n) cudaFree(a->d_n);
    delete a;
}

extern "C" int pwmscan_alnset_update(pwmscan_alnset* a, int32_t slot, int32_t ncol, const double* mat) {
    if (!a || !mat) return PWMSCAN_EINVAL;
    pwmscan_ctx* ctx = a->ctx;
    if (slot < 0 || slot >= a->M) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_alnset_update: slot out of range");
    if (ncol < 1 || ncol > PWMSCAN_MAX_COLS) return set_err(ctx, PWMSCAN_ERANGE, "pwmscan_alnset_update: motif length outside 1..64");
    std::lock_guard<std::mutex> lk(ctx->mu);
    PWM_CUDA(cudaSetDevice(ctx->device));
    alnset_fill_slot(a, slot, ncol, mat);
    const size_t o = (size_t)(2 * slot) * 256;
    PWM_CUDA(cudaMemcpyAsync(a->d_mats + o, a->mats.data() + o, 512 * 8, cudaMemcpyHostToDevice, ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(a->d_logs + o, a->logs.data() + o, 512 * 8, cudaMemcpyHostToDevice, ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(a->d_n + slot, a->n.data() + slot, 4, cudaMemcpyHostToDevice, ctx->stream));
    PWM_CUDA(cudaStreamSynchronize(ctx->stream));        // the host mirror may be rewritten by the next update
    return PWMSCAN_OK;
}

extern "C" int pwmscan_alnset_pairs(pwmscan_alnset* a, int64_t npairs, const int32_t* pa, const int32_t* pb, double* dist,
                                    uint8_t* same_dir, int32_t* pos) {
    if (!a) return PWMSCAN_EINVAL;
    pwmscan_ctx* ctx = a->ctx;
    if (!dist || !same_dir || !pos || (pa == nullptr) != (pb == nullptr)) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_alnset_pairs: bad argument");
    const bool all = pa == nullptr;
    if (all) npairs = (int64_t)a->M * (a->M - 1) / 2;
    if (npairs < 0) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_alnset_pairs: bad argument");
    if (npairs == 0) return PWMSCAN_OK;
    if (!all) for (int64_t k = 0; k < npairs; ++k)
        if (pa[k] < 0 || pa[k] >= a->M || pb[k] < 0 || pb[k] >= a->M) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_alnset_pairs: slot out of range");
    std::lock_guard<std::mutex> lk(ctx->mu);
    PWM_CUDA(cudaSetDevice(ctx->device));
    AlignDev dev{a->d_mats, a->d_logs, a->d_off, a->d_n, a->d_opti};
    return launch_and_collect(ctx, a->cfg, dev, a->M, a->mats.data(), a->logs.data(), a->col_off.data(), a->n.data(), a->opti.data(), npairs,
                              pa, pb, dist, same_dir, pos);
}
