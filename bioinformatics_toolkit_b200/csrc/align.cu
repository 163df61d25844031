// align.cu — PWM-PWM alignment for many pairs on the device (Motif/Alignment.hs:83-132; the
// all-pairs initialisation and row refresh of iterativeMerge, Motif/Merge.hs:138-145,161-165;
// `merge_motifs dist`, app/MergeMotifs.hs:158-169).  One thread per pair runs alignmentBy
// (align_core.h); pairs whose decisive comparisons were within 1e-10 relative are re-evaluated on
// the host with the same function compiled against glibc, so (same_dir, pos) follow the
// reference's tie rules and the distance of those pairs is the reference's bit for bit.
// Bound: fp64 / SFU (four `log` per column pair); bytes are negligible.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "align_core.h"
#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "internal.h"

using namespace pwmscan;

// fn(t) for t in [0, n) on up to 16 host threads (independent iterations; small n stays on the calling thread)
template <class F>
static void host_parallel_for(size_t n, F fn) {
    const size_t nt = std::min<size_t>({(size_t)16, (size_t)std::max(1u, std::thread::hardware_concurrency()), n / 64 + 1});
    if (nt <= 1) { for (size_t t = 0; t < n; ++t) fn(t); return; }
    std::atomic<size_t> next(0);
    auto work = [&]() { for (;;) { const size_t t0 = next.fetch_add(32); if (t0 >= n) return; for (size_t t = t0; t < std::min(n, t0 + 32); ++t) fn(t); } };
    std::vector<std::thread> th;
    for (size_t w = 1; w < nt; ++w) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
}

namespace {

struct AlignDev {
    const double* mats;      // per (motif, strand): n x 4 probabilities
    const double* logs;      // same layout: log2 of the entries (host glibc)
    const int64_t* col_off;  // [2m+s] column offset
    const int32_t* ncol;     // [m]
    const double* opti;      // [(x * (maxc + 1) + y) * maxc + i] = optimalSc x y !! i
    int maxc;                // longest PWM the opti table covers (64 for pwmscan_motifs; alignment sets grow)
};

__device__ __forceinline__ void run_pair(const AlignCfg& cfg, const AlignDev& d, int i, int j, double* dist,
                                         uint8_t* same_dir, int32_t* pos, uint8_t* amb, int64_t k) {
    const int n1 = d.ncol[i], n2 = d.ncol[j];
    const int64_t o1 = 4 * d.col_off[2 * i], o2 = 4 * d.col_off[2 * j], o2r = 4 * d.col_off[2 * j + 1];
    double dd; int sd, pp, am = 0;
    align_pair(cfg, d.mats + o1, d.logs + o1, n1, d.mats + o2, d.logs + o2, d.mats + o2r, d.logs + o2r, n2,
               d.opti + ((size_t)n1 * (d.maxc + 1) + n2) * d.maxc, d.opti + ((size_t)n2 * (d.maxc + 1) + n1) * d.maxc, &dd, &sd, &pp, &am);
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
    // near-ties: same function, host libm — collected first, then re-evaluated by all host cores (8 942 of the 1 999 000 pairs
    // of config 5: 49 ms on one core, more than the kernel)
    int64_t namb = 0;
    struct Redo { int64_t k; int i, j; };
    std::vector<Redo> todo;
    auto redo = [&](int64_t k, int i, int j) {
        const int n1 = h_n[i], n2 = h_n[j];
        const size_t o1 = 4 * (size_t)col_off[2 * i], o2 = 4 * (size_t)col_off[2 * j], o2r = 4 * (size_t)col_off[2 * j + 1];
        double dd; int sd, pp, am = 0;
        align_pair(cfg, h_mats + o1, h_logs + o1, n1, h_mats + o2, h_logs + o2, h_mats + o2r, h_logs + o2r, n2,
                   h_opti + ((size_t)n1 * (dev.maxc + 1) + n2) * dev.maxc, h_opti + ((size_t)n2 * (dev.maxc + 1) + n1) * dev.maxc, &dd, &sd, &pp, &am);
        dist[k] = dd; same_dir[k] = (uint8_t)sd; pos[k] = pp;
    };
    if (all) {
        int64_t k = 0;
        for (int i = 0; i < M; ++i) for (int j = i + 1; j < M; ++j, ++k) if (h_amb[k]) todo.push_back({k, i, j});
    } else {
        for (int64_t k = 0; k < npairs; ++k) if (h_amb[k]) todo.push_back({k, pa[k], pb[k]});
    }
    namb = (int64_t)todo.size();
    host_parallel_for(todo.size(), [&](size_t t) { redo(todo[t].k, todo[t].i, todo[t].j); });
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
    AlignDev dev{d_mats, d_logs, d_off, d_n, d_opti, 64};
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
// Resident alignment set: the PWMs AND the distance matrix of iterativeMerge stay on the device
// (Merge.hs:113-170).  The all-pairs initialisation (161-165), every row refresh (138-145) and the
// search for the next pair to merge (the first minimum in row-major order, strict <: 149-158) run on
// the same device arrays; a merge replaces one slot in place (its matrix, reverse complement and log2
// table) and only the five numbers of the next minimum come back.  Slots have a fixed stride of `maxc`
// columns (64 to begin with); a merged PWM that is longer makes the set grow (stride doubled, optimalSc
// table rebuilt) up to kAlnMaxCols columns — the reference has no limit, merged PWMs of real motif
// collections stay far below it.
// ------------------------------------------------------------------------------------------------
constexpr int kAlnMaxCols = 256;

struct pwmscan_alnset {
    pwmscan_ctx* ctx = nullptr;
    int32_t M = 0;
    int maxc = 64;
    AlignCfg cfg;
    std::vector<int32_t> n;           // columns per slot
    std::vector<int64_t> col_off;     // [2m+s] = (2m+s) * maxc (column offsets, fixed stride)
    std::vector<double> mats, logs;   // host mirror, [(2m+s) * 4 * maxc]
    std::vector<double> opti;         // [(x * (maxc + 1) + y) * maxc + i]
    double *d_mats = nullptr, *d_logs = nullptr, *d_opti = nullptr;
    int64_t* d_off = nullptr;
    int32_t* d_n = nullptr;
    // the matrix of iterativeMerge (pwmscan_alnset_matrix_init / _merge_step): entry (r, c), r < c, both slots alive;
    // +inf everywhere else
    double* d_D = nullptr;
    uint8_t* d_SD = nullptr;
    int32_t* d_POS = nullptr;
    uint8_t* d_alive = nullptr;
    unsigned long long* d_amb = nullptr;      // [0] count, [1..] r * M + c of entries whose decision was a near-tie
    unsigned long long* d_part = nullptr;     // argmin partials: (value bits, index) per CTA + the final one
    std::vector<uint8_t> alive;
};

namespace {

void alnset_fill_slot(pwmscan_alnset* a, int32_t m, int32_t ncol, const double* mat) {
    const size_t stride = 4 * (size_t)a->maxc;
    a->n[m] = ncol;
    double* f = a->mats.data() + (size_t)(2 * m) * stride;
    double* r = f + stride;
    std::memset(f, 0, 2 * stride * sizeof(double));
    for (int e = 0; e < 4 * ncol; ++e) { f[e] = mat[e]; r[e] = mat[4 * ncol - 1 - e]; }      // rcPWM: reversed flattening (Motif.hs:77-78)
    double* lf = a->logs.data() + (size_t)(2 * m) * stride;
    for (size_t e = 0; e < 2 * stride; ++e) lf[e] = 0.0;
    for (int e = 0; e < 4 * ncol; ++e) { lf[e] = std::log(f[e]) / a->cfg.ln2; lf[stride + e] = std::log(r[e]) / a->cfg.ln2; }
}

void alnset_free_device(pwmscan_alnset* a) {
    if (a->d_mats) cudaFree(a->d_mats);
    if (a->d_logs) cudaFree(a->d_logs);
    if (a->d_opti) cudaFree(a->d_opti);
    if (a->d_off) cudaFree(a->d_off);
    if (a->d_n) cudaFree(a->d_n);
    a->d_mats = a->d_logs = a->d_opti = nullptr; a->d_off = nullptr; a->d_n = nullptr;
}

// (Re)builds the optimalSc table and the device copies for the set's current maxc from the host mirror.
int alnset_upload(pwmscan_alnset* a) {
    pwmscan_ctx* ctx = a->ctx;
    const int mc = a->maxc;
    a->opti.assign((size_t)(mc + 1) * (mc + 1) * mc, 0.0);
    for (int x = 1; x <= mc; ++x)
        for (int y = 1; y <= mc; ++y) aln_optimal_sc(a->cfg, x, y, a->opti.data() + ((size_t)x * (mc + 1) + y) * mc);
    for (int m = 0; m < 2 * a->M; ++m) a->col_off[m] = (int64_t)m * mc;
    alnset_free_device(a);
    const size_t b_m = a->mats.size() * 8;
    PWM_CUDA(cudaMalloc((void**)&a->d_mats, b_m));
    PWM_CUDA(cudaMalloc((void**)&a->d_logs, b_m));
    PWM_CUDA(cudaMalloc((void**)&a->d_opti, a->opti.size() * 8));
    PWM_CUDA(cudaMalloc((void**)&a->d_off, a->col_off.size() * 8));
    PWM_CUDA(cudaMalloc((void**)&a->d_n, (size_t)a->M * 4));
    PWM_CUDA(cudaMemcpy(a->d_mats, a->mats.data(), b_m, cudaMemcpyHostToDevice));
    PWM_CUDA(cudaMemcpy(a->d_logs, a->logs.data(), b_m, cudaMemcpyHostToDevice));
    PWM_CUDA(cudaMemcpy(a->d_opti, a->opti.data(), a->opti.size() * 8, cudaMemcpyHostToDevice));
    PWM_CUDA(cudaMemcpy(a->d_off, a->col_off.data(), a->col_off.size() * 8, cudaMemcpyHostToDevice));
    PWM_CUDA(cudaMemcpy(a->d_n, a->n.data(), (size_t)a->M * 4, cudaMemcpyHostToDevice));
    return PWMSCAN_OK;
}

// Makes room for PWMs of `need` columns: new stride, host mirrors re-laid, everything re-uploaded.
int alnset_grow(pwmscan_alnset* a, int need) {
    int mc = a->maxc;
    while (mc < need) mc *= 2;
    if (mc > kAlnMaxCols) mc = kAlnMaxCols;
    const size_t os = 4 * (size_t)a->maxc, ns = 4 * (size_t)mc;
    std::vector<double> m2((size_t)2 * a->M * ns, 0.0), l2((size_t)2 * a->M * ns, 0.0);
    for (int m = 0; m < 2 * a->M; ++m) {
        std::memcpy(m2.data() + (size_t)m * ns, a->mats.data() + (size_t)m * os, os * 8);
        std::memcpy(l2.data() + (size_t)m * ns, a->logs.data() + (size_t)m * os, os * 8);
    }
    a->mats.swap(m2); a->logs.swap(l2);
    a->maxc = mc;
    return alnset_upload(a);
}

AlignDev alnset_dev(const pwmscan_alnset* a) { return AlignDev{a->d_mats, a->d_logs, a->d_off, a->d_n, a->d_opti, a->maxc}; }

// ---- the resident matrix ------------------------------------------------------------------------
__device__ __forceinline__ void mat_store(const AlignCfg& cfg, const AlignDev& d, int M, int r, int c, double* D, uint8_t* SD, int32_t* POS,
                                          unsigned long long* amb) {
    const int n1 = d.ncol[r], n2 = d.ncol[c];
    const int64_t o1 = 4 * d.col_off[2 * r], o2 = 4 * d.col_off[2 * c], o2r = 4 * d.col_off[2 * c + 1];
    double dd; int sd, pp, am = 0;
    align_pair(cfg, d.mats + o1, d.logs + o1, n1, d.mats + o2, d.logs + o2, d.mats + o2r, d.logs + o2r, n2,
               d.opti + ((size_t)n1 * (d.maxc + 1) + n2) * d.maxc, d.opti + ((size_t)n2 * (d.maxc + 1) + n1) * d.maxc, &dd, &sd, &pp, &am);
    const size_t k = (size_t)r * M + c;
    D[k] = dd; SD[k] = (uint8_t)sd; POS[k] = pp;
    if (am) { const unsigned long long q = atomicAdd(&amb[0], 1ull); amb[1 + q] = (unsigned long long)k; }
}

// all r < c; the rest of the matrix is +inf
__global__ void alnmat_fill_kernel(AlignCfg cfg, AlignDev d, int M, double* D, uint8_t* SD, int32_t* POS, unsigned long long* amb) {
    const int r = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= M) return;
    if (c <= r) { D[(size_t)r * M + c] = __longlong_as_double(0x7FF0000000000000LL); return; }
    mat_store(cfg, d, M, r, c, D, SD, POS, amb);
}

// after slot j was merged into slot i: column and row j are blanked (Merge.hs:135: the matrix is symmetric), slot i is
// re-aligned against every slot still alive, argument order by index (138-145)
__global__ void alnmat_refresh_kernel(AlignCfg cfg, AlignDev d, int M, int i, int j, const uint8_t* __restrict__ alive, double* D, uint8_t* SD,
                                      int32_t* POS, unsigned long long* amb) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= M) return;
    const double inf = __longlong_as_double(0x7FF0000000000000LL);
    D[(size_t)q * M + j] = inf;
    D[(size_t)j * M + q] = inf;
    if (q == i || q == j || !alive[q]) return;
    if (i < q) mat_store(cfg, d, M, i, q, D, SD, POS, amb); else mat_store(cfg, d, M, q, i, D, SD, POS, amb);
}

// The same refresh with ONE WARP per pair.  A merge re-aligns slot i against every live slot — a few thousand pairs, far
// fewer than the device has threads — so what a merge costs is the LATENCY of one alignment, and a thread that walks all
// offsets of all four `loop`s alone needs ~0.9 ms for it.  Here the lanes of a warp evaluate the offsets of a loop in
// parallel (an offset's distance does not depend on the running minimum: aln_offset_value) into shared memory and the
// decision sequence is replayed in offset order (aln_loop_replay) — the same comparisons on the same numbers.
constexpr int kRefreshWarps = 4;
__device__ __forceinline__ void aln_loop_warp(const AlignCfg& c, const double* opti, const double* a, const double* la, int na, const double* b,
                                              const double* lb, int nb, int n1, int n2, double* dv, double* dmin, int* imin, int* ambiguous) {
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < nb; i += 32) dv[i] = aln_offset_value(c, a, la, na, b, lb, nb, n1, n2, i);
    __syncwarp();
    aln_loop_replay(opti, dv, nb, dmin, imin, ambiguous);       // every lane, same result
    __syncwarp();
}

__global__ void __launch_bounds__(32 * kRefreshWarps)
alnmat_refresh_warp_kernel(AlignCfg cfg, AlignDev d, int M, int i, int j, const uint8_t* __restrict__ alive, double* D, uint8_t* SD,
                           int32_t* POS, unsigned long long* amb) {
    extern __shared__ double s_dv[];                                           // [kRefreshWarps][maxc]
    const int q = blockIdx.x * kRefreshWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= M) return;
    const double inf = __longlong_as_double(0x7FF0000000000000LL);
    if (lane == 0) { D[(size_t)q * M + j] = inf; D[(size_t)j * M + q] = inf; }
    if (q == i || q == j || !alive[q]) return;
    const int r = i < q ? i : q, c = i < q ? q : i;
    double* dv = s_dv + (size_t)(threadIdx.x >> 5) * d.maxc;
    const int n1 = d.ncol[r], n2 = d.ncol[c];
    const double *m1 = d.mats + 4 * d.col_off[2 * r], *l1 = d.logs + 4 * d.col_off[2 * r];
    const double *m2 = d.mats + 4 * d.col_off[2 * c], *l2 = d.logs + 4 * d.col_off[2 * c];
    const double *m2r = d.mats + 4 * d.col_off[2 * c + 1], *l2r = d.logs + 4 * d.col_off[2 * c + 1];
    const double* opti1 = d.opti + ((size_t)n1 * (d.maxc + 1) + n2) * d.maxc;
    const double* opti2 = d.opti + ((size_t)n2 * (d.maxc + 1) + n1) * d.maxc;
    int am = 0;
    double df, dr; int pf, pr;
    {   // align_pair (align_core.h), loop by loop
        double d1, d2; int i1, i2;
        aln_loop_warp(cfg, opti2, m2, l2, n2, m1, l1, n1, n1, n2, dv, &d1, &i1, &am);
        aln_loop_warp(cfg, opti1, m1, l1, n1, m2, l2, n2, n1, n2, dv, &d2, &i2, &am);
        if (aln_near(d1, d2)) am = 1;
        if (d1 < d2) { df = d1; pf = i1; } else { df = d2; pf = -i2; }
    }
    {
        double d1, d2; int i1, i2;
        aln_loop_warp(cfg, opti2, m2r, l2r, n2, m1, l1, n1, n1, n2, dv, &d1, &i1, &am);
        aln_loop_warp(cfg, opti1, m1, l1, n1, m2r, l2r, n2, n1, n2, dv, &d2, &i2, &am);
        if (aln_near(d1, d2)) am = 1;
        if (d1 < d2) { dr = d1; pr = i1; } else { dr = d2; pr = -i2; }
    }
    if (aln_near(df, dr)) am = 1;
    if (lane == 0) {
        const size_t k = (size_t)r * M + c;
        if (df <= dr) { D[k] = df; SD[k] = 1; POS[k] = pf; } else { D[k] = dr; SD[k] = 0; POS[k] = pr; }
        if (am) { const unsigned long long t = atomicAdd(&amb[0], 1ull); amb[1 + t] = (unsigned long long)k; }
    }
}

// entries the host re-evaluated (near-ties): k = r * M + c
__global__ void alnmat_patch_kernel(int n, const unsigned long long* __restrict__ k, const double* __restrict__ dd, const int32_t* __restrict__ pp,
                                    const uint8_t* __restrict__ sd, double* D, uint8_t* SD, int32_t* POS) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    D[k[t]] = dd[t]; SD[k[t]] = sd[t]; POS[k[t]] = pp[t];
}

// first minimum in row-major order (strict <): the smallest (value, index) pair.  Distances are >= 0, so the bit pattern of
// a double orders like the value and one 128-bit comparison does both.
__device__ __forceinline__ void argmin_combine(unsigned long long& v, unsigned long long& k, unsigned long long v2, unsigned long long k2) {
    if (v2 < v || (v2 == v && k2 < k)) { v = v2; k = k2; }
}
__global__ void __launch_bounds__(256) alnmat_argmin_kernel(const double* __restrict__ D, unsigned long long n, unsigned long long* __restrict__ part) {
    __shared__ unsigned long long sv[8], sk[8];
    unsigned long long v = ~0ull, k = ~0ull;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (unsigned long long)gridDim.x * blockDim.x)
        argmin_combine(v, k, (unsigned long long)__double_as_longlong(D[t]), t);
    for (int o = 16; o > 0; o >>= 1) argmin_combine(v, k, __shfl_xor_sync(0xFFFFFFFFu, v, o), __shfl_xor_sync(0xFFFFFFFFu, k, o));
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = v; sk[threadIdx.x >> 5] = k; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) argmin_combine(v, k, sv[w], sk[w]);
        part[2 * blockIdx.x] = v; part[2 * blockIdx.x + 1] = k;
    }
}
__global__ void alnmat_argmin_final_kernel(int nparts, unsigned long long* part, const uint8_t* __restrict__ SD, const int32_t* __restrict__ POS) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long v = ~0ull, k = ~0ull;
    for (int p = 0; p < nparts; ++p) argmin_combine(v, k, part[2 * p], part[2 * p + 1]);
    unsigned long long* out = part + 2 * nparts;              // [value bits, index, same_dir, pos]
    out[0] = v; out[1] = k;
    const bool none = v >= 0x7FF0000000000000ull;
    out[2] = none ? 0ull : (unsigned long long)SD[k];
    out[3] = none ? 0ull : (unsigned long long)(long long)POS[k];
}

constexpr int kArgminParts = 592;

// Resolves the near-ties recorded by the last fill / refresh on the host (same function, glibc), patches the matrix and
// returns the first minimum.  ctx->mu held.
int alnmat_settle_and_argmin(pwmscan_alnset* a, int32_t* oi, int32_t* oj, double* od, uint8_t* osd, int32_t* opos) {
    pwmscan_ctx* ctx = a->ctx;
    const int M = a->M;
    unsigned long long namb = 0;
    PWM_CUDA(cudaMemcpyAsync(&namb, a->d_amb, 8, cudaMemcpyDeviceToHost, ctx->stream));
    PWM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (namb > 0) {
        std::vector<unsigned long long> ks((size_t)namb);
        PWM_CUDA(cudaMemcpy(ks.data(), a->d_amb + 1, (size_t)namb * 8, cudaMemcpyDeviceToHost));
        std::vector<double> dd((size_t)namb);
        std::vector<int32_t> pp((size_t)namb);
        std::vector<uint8_t> sd((size_t)namb);
        const int mc = a->maxc;
        host_parallel_for((size_t)namb, [&](size_t t) {
            const int r = (int)(ks[t] / (unsigned long long)M), c = (int)(ks[t] % (unsigned long long)M);
            const int n1 = a->n[r], n2 = a->n[c];
            const size_t o1 = 4 * (size_t)a->col_off[2 * r], o2 = 4 * (size_t)a->col_off[2 * c], o2r = 4 * (size_t)a->col_off[2 * c + 1];
            double d; int s, p, am = 0;
            align_pair(a->cfg, a->mats.data() + o1, a->logs.data() + o1, n1, a->mats.data() + o2, a->logs.data() + o2, a->mats.data() + o2r,
                       a->logs.data() + o2r, n2, a->opti.data() + ((size_t)n1 * (mc + 1) + n2) * mc, a->opti.data() + ((size_t)n2 * (mc + 1) + n1) * mc,
                       &d, &s, &p, &am);
            dd[t] = d; sd[t] = (uint8_t)s; pp[t] = p;
        });
        const size_t nb = (size_t)namb;
        unsigned char* d_p = (unsigned char*)scratch_get(ctx, 1, nb * 24 + 64);
        if (!d_p) return PWMSCAN_ECUDA;
        unsigned long long* d_k = (unsigned long long*)d_p;
        double* d_dd = (double*)(d_p + nb * 8);
        int32_t* d_pp = (int32_t*)(d_p + nb * 16);
        uint8_t* d_sd = d_p + nb * 20;
        PWM_CUDA(cudaMemcpyAsync(d_k, ks.data(), nb * 8, cudaMemcpyHostToDevice, ctx->stream));
        PWM_CUDA(cudaMemcpyAsync(d_dd, dd.data(), nb * 8, cudaMemcpyHostToDevice, ctx->stream));
        PWM_CUDA(cudaMemcpyAsync(d_pp, pp.data(), nb * 4, cudaMemcpyHostToDevice, ctx->stream));
        PWM_CUDA(cudaMemcpyAsync(d_sd, sd.data(), nb, cudaMemcpyHostToDevice, ctx->stream));
        alnmat_patch_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, ctx->stream>>>((int)nb, d_k, d_dd, d_pp, d_sd, a->d_D, a->d_SD, a->d_POS);
        PWM_CUDA(cudaGetLastError());
        PWM_CUDA(cudaStreamSynchronize(ctx->stream));        // the host vectors go out of scope
    }
    ctx->timing[11] += (double)namb;
    alnmat_argmin_kernel<<<kArgminParts, 256, 0, ctx->stream>>>(a->d_D, (unsigned long long)M * M, a->d_part);
    alnmat_argmin_final_kernel<<<1, 32, 0, ctx->stream>>>(kArgminParts, a->d_part, a->d_SD, a->d_POS);
    PWM_CUDA(cudaGetLastError());
    unsigned long long res[4];
    PWM_CUDA(cudaMemcpyAsync(res, a->d_part + 2 * kArgminParts, 32, cudaMemcpyDeviceToHost, ctx->stream));
    PWM_CUDA(cudaStreamSynchronize(ctx->stream));
    double d;
    std::memcpy(&d, &res[0], 8);
    if (res[0] >= 0x7FF0000000000000ull) { *oi = -1; *oj = -1; *od = d; *osd = 0; *opos = 0; return PWMSCAN_OK; }
    *oi = (int32_t)(res[1] / (unsigned long long)M); *oj = (int32_t)(res[1] % (unsigned long long)M);
    *od = d; *osd = (uint8_t)res[2]; *opos = (int32_t)(long long)res[3];
    return PWMSCAN_OK;
}

}  // namespace

extern "C" int pwmscan_alnset_create(pwmscan_ctx* ctx, int32_t M, const int32_t* ncol, const double* mats, int penalty_kind, double gap,
                                     int combine_kind, pwmscan_alnset** out) {
    if (!ctx || !out || M < 1 || !ncol || !mats) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_alnset_create: bad argument");
    *out = nullptr;
    if (penalty_kind < 0 || penalty_kind > 3 || combine_kind < 0 || combine_kind > 3)
        return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_align: unknown penalty / combine kind");
    int longest = 1;
    for (int m = 0; m < M; ++m) {
        if (ncol[m] < 1 || ncol[m] > kAlnMaxCols) return set_err(ctx, PWMSCAN_ERANGE, "pwmscan_alnset_create: motif length outside 1..256");
        longest = std::max(longest, (int)ncol[m]);
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    PWM_CUDA(cudaSetDevice(ctx->device));
    pwmscan_alnset* a = new (std::nothrow) pwmscan_alnset();
    if (!a) return set_err(ctx, PWMSCAN_EINVAL, "out of host memory");
    a->ctx = ctx; a->M = M;
    while (a->maxc < longest) a->maxc *= 2;
    a->cfg.penal_kind = penalty_kind; a->cfg.comb_kind = combine_kind; a->cfg.gap = gap; a->cfg.ln2 = std::log(2.0);
    a->n.resize(M); a->col_off.resize((size_t)2 * M);
    a->mats.assign((size_t)2 * M * 4 * a->maxc, 0.0); a->logs.assign((size_t)2 * M * 4 * a->maxc, 0.0);
    a->alive.assign((size_t)M, 1);
    const double* p = mats;
    for (int m = 0; m < M; ++m) {
        alnset_fill_slot(a, m, ncol[m], p);
        p += 4 * ncol[m];
    }
    int rc = alnset_upload(a);
    if (rc) { pwmscan_alnset_free(a); return rc; }
    *out = a;
    return PWMSCAN_OK;
}

extern "C" void pwmscan_alnset_free(pwmscan_alnset* a) {
    if (!a) return;
    if (a->ctx) cudaSetDevice(a->ctx->device);
    alnset_free_device(a);
    if (a->d_D) cudaFree(a->d_D);
    if (a->d_SD) cudaFree(a->d_SD);
    if (a->d_POS) cudaFree(a->d_POS);
    if (a->d_alive) cudaFree(a->d_alive);
    if (a->d_amb) cudaFree(a->d_amb);
    if (a->d_part) cudaFree(a->d_part);
    delete a;
}

// Replaces the PWM of a slot (host mirror + device); grows the set if the PWM is longer than the current stride.  ctx->mu held.
static int alnset_update_locked(pwmscan_alnset* a, int32_t slot, int32_t ncol, const double* mat) {
    pwmscan_ctx* ctx = a->ctx;
    if (slot < 0 || slot >= a->M) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_alnset_update: slot out of range");
    if (ncol < 1 || ncol > kAlnMaxCols) return set_err(ctx, PWMSCAN_ERANGE, "pwmscan_alnset_update: motif length outside 1..256");
    PWM_CUDA(cudaSetDevice(ctx->device));
    if (ncol > a->maxc) {
        PWM_CUDA(cudaStreamSynchronize(ctx->stream));
        int rc = alnset_grow(a, ncol);
        if (rc) return rc;
    }
    alnset_fill_slot(a, slot, ncol, mat);
    const size_t stride = 4 * (size_t)a->maxc, o = (size_t)(2 * slot) * stride;
    PWM_CUDA(cudaMemcpyAsync(a->d_mats + o, a->mats.data() + o, 2 * stride * 8, cudaMemcpyHostToDevice, ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(a->d_logs + o, a->logs.data() + o, 2 * stride * 8, cudaMemcpyHostToDevice, ctx->stream));
    PWM_CUDA(cudaMemcpyAsync(a->d_n + slot, a->n.data() + slot, 4, cudaMemcpyHostToDevice, ctx->stream));
    PWM_CUDA(cudaStreamSynchronize(ctx->stream));        // the host mirror may be rewritten by the next update
    return PWMSCAN_OK;
}

extern "C" int pwmscan_alnset_update(pwmscan_alnset* a, int32_t slot, int32_t ncol, const double* mat) {
    if (!a || !mat) return PWMSCAN_EINVAL;
    std::lock_guard<std::mutex> lk(a->ctx->mu);
    return alnset_update_locked(a, slot, ncol, mat);
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
    return launch_and_collect(ctx, a->cfg, alnset_dev(a), a->M, a->mats.data(), a->logs.data(), a->col_off.data(), a->n.data(), a->opti.data(), npairs,
                              pa, pb, dist, same_dir, pos);
}

extern "C" int pwmscan_alnset_matrix_init(pwmscan_alnset* a, int32_t* i, int32_t* j, double* d, uint8_t* same_dir, int32_t* pos) {
    if (!a || !i || !j || !d || !same_dir || !pos) return PWMSCAN_EINVAL;
    pwmscan_ctx* ctx = a->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    PWM_CUDA(cudaSetDevice(ctx->device));
    const size_t MM = (size_t)a->M * a->M;
    if (!a->d_D) {
        PWM_CUDA(cudaMalloc((void**)&a->d_D, MM * 8));
        PWM_CUDA(cudaMalloc((void**)&a->d_SD, MM));
        PWM_CUDA(cudaMalloc((void**)&a->d_POS, MM * 4));
        PWM_CUDA(cudaMalloc((void**)&a->d_alive, (size_t)a->M));
        PWM_CUDA(cudaMalloc((void**)&a->d_amb, (MM / 2 + 2) * 8));
        PWM_CUDA(cudaMalloc((void**)&a->d_part, (size_t)(2 * kArgminParts + 4) * 8));
    }
    std::fill(a->alive.begin(), a->alive.end(), (uint8_t)1);
    PWM_CUDA(cudaMemsetAsync(a->d_alive, 1, (size_t)a->M, ctx->stream));
    PWM_CUDA(cudaMemsetAsync(a->d_amb, 0, 8, ctx->stream));
    PWM_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    dim3 grid((unsigned)((a->M + 127) / 128), (unsigned)a->M);
    alnmat_fill_kernel<<<grid, 128, 0, ctx->stream>>>(a->cfg, alnset_dev(a), a->M, a->d_D, a->d_SD, a->d_POS, a->d_amb);
    PWM_CUDA(cudaGetLastError());
    PWM_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    ctx->timing[11] = 0;
    int rc = alnmat_settle_and_argmin(a, i, j, d, same_dir, pos);
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->timing[10] = ms;
    return rc;
}

extern "C" int pwmscan_alnset_merge_step(pwmscan_alnset* a, int32_t i, int32_t j, int32_t ncol, const double* mat, int32_t* ni, int32_t* nj,
                                         double* d, uint8_t* same_dir, int32_t* pos) {
    if (!a || !mat || !ni || !nj || !d || !same_dir || !pos) return PWMSCAN_EINVAL;
    pwmscan_ctx* ctx = a->ctx;
    if (!a->d_D) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_alnset_merge_step: call pwmscan_alnset_matrix_init first");
    if (i < 0 || j < 0 || i >= a->M || j >= a->M || i == j || !a->alive[i] || !a->alive[j])
        return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_alnset_merge_step: slots must be two different live slots");
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = alnset_update_locked(a, i, ncol, mat);
    if (rc) return rc;
    a->alive[j] = 0;
    PWM_CUDA(cudaMemsetAsync(a->d_alive + j, 0, 1, ctx->stream));
    PWM_CUDA(cudaMemsetAsync(a->d_amb, 0, 8, ctx->stream));
    static const bool thread_per_pair = std::getenv("PWMSCAN_ALN_REFRESH_THREAD") != nullptr;       // the round-1 form, for comparison
    if (thread_per_pair)
        alnmat_refresh_kernel<<<(unsigned)((a->M + 127) / 128), 128, 0, ctx->stream>>>(a->cfg, alnset_dev(a), a->M, i, j, a->d_alive, a->d_D, a->d_SD,
                                                                                       a->d_POS, a->d_amb);
    else
        alnmat_refresh_warp_kernel<<<(unsigned)((a->M + kRefreshWarps - 1) / kRefreshWarps), 32 * kRefreshWarps, (size_t)kRefreshWarps * a->maxc * 8,
                                     ctx->stream>>>(a->cfg, alnset_dev(a), a->M, i, j, a->d_alive, a->d_D, a->d_SD, a->d_POS, a->d_amb);
    PWM_CUDA(cudaGetLastError());
    return alnmat_settle_and_argmin(a, ni, nj, d, same_dir, pos);
}
