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
    const size_t b_out = (size_t)npairs * (8 + 4 + 1 + 1) + 64;
    unsigned char* d_out = (unsigned char*)scratch_get(ctx, 2, b_out + (all ? 0 : (size_t)npairs * 8));
    if (!d_in || !d_out) return PWMSCAN_ECUDA;
    double* d_mats = (double*)d_in;
    double* d_logs = (double*)(d_in + b_m);
    double* d_opti = (double*)(d_in + 2 * b_m);
    int64_t* d_off = (int64_t*)(d_in + 2 * b_m + b_opti);
    int32_t* d_n = (int32_t*)(d_in + 2 * b_m + b_opti + b_off);
    double* d_dist = (double*)d_out;
    int32_t* d_pos = (int32_t*)(d_out + (size_t)npairs * 8);
    uint8_t* d_sd = (uint8_t*)(d_out + (size_t)npairs * 12);
    uint8_t* d_amb = d_sd + npairs;
    int32_t *d_pa = nullptr, *d_pb = nullptr;
    PWM_CUDA(cudaMemcpyAsync(d_mats, h_mats.data(), b_m, cudaMemcpyHostToDevice, ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(d_logs, h_logs.data(), b_m, cudaMemcpyHostToDevice, ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(d_opti, h_opti.data(), b_opti, cudaMemcpyHostToDevice, ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(d_off, mo->col_off.data(), b_off, cudaMemcpyHostToDevice, ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(d_n, h_n.data(), b_n, cudaMemcpyHostToDevice, ctx->stream));
    if (!all) {
        d_pa = (int32_t*)(d_out + ((b_out + 7) & ~(size_t)7));
        d_pb = d_pa + npairs;
        PWM_CUDA(cudaMemcpyAsync(d_pa, pa, (size_t)npairs * 4, cudaMemcpyHostToDevice, ctx->stream));
        PWM_CUDA(cudaMemcpyAsync(d_pb, pb, (size_t)npairs * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    AlignDev dev{d_mats, d_logs, d_off, d_n, d_opti};
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
        const size_t o1 = 4 * (size_t)mo->col_off[2 * i], o2 = 4 * (size_t)mo->col_off[2 * j], o2r = 4 * (size_t)mo->col_off[2 * j + 1];
        double dd; int sd, pp, am = 0;
        align_pair(cfg, h_mats.data() + o1, h_logs.data() + o1, n1, h_mats.data() + o2, h_logs.data() + o2, h_mats.data() + o2r,
                   h_logs.data() + o2r, n2, h_opti.data() + ((size_t)n1 * 65 + n2) * 64, h_opti.data() + ((size_t)n2 * 65 + n1) * 64,
                   &dd, &sd, &pp, &am);
        dist[k] = dd; same_dir[k] = (uint8_t)sd; pos[k] = pp;
        ++namb;
    };
    if (all) {
        int64_t k = 0;
        for (int i = 0; i < M; ++i) for (int j = i + 1; j < M; ++j, ++k) if (h_amb[k]) redo(k, i, j);
    } else {
        for (int64_t k = 0; k < npairs; ++k) if (h_amb[k]) redo(k, pa[k], pb[k]);
    }
    ctx->timing[0] = ms; ctx->timing[3] = ms; ctx->timing[6] = (double)namb;
    ctx->timing[5] = (double)npairs * 14;
    return PWMSCAN_OK;
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
