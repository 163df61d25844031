// pwmscan.cu — context, sequence packing, motif tables, plans, and the thresholded scan
// (findTFBS / findTFBSWith / scanMotif body) of libpwmscan.so.  sm_100a only.
//
// Data layout in HBM:
//   seq2   2 bits/base, 16 bases per 32-bit word, stored base index u = position + 32
//   mask   1 bit/base, 1 = not one of ACGT (or padding)
//   ascii  optional upper-cased copy (IUPAC-aware entry points only)
//   tables k-mer mask tables per 128-motif block (km_prefilter.cu; default first stage), built on the device;
//          the other first stages keep their own (shared-memory 4-mer tables / int8 GEMM tiles)
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <thread>
#include <cstring>
#include <new>

#include "internal.h"
#include "scan_core.cuh"

using namespace pwmscan;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
constexpr int kDefaultMode = PWM_MODE_KMER;             // default first stage (PWMSCAN_PREFILTER=kmer|lds|mma overrides)
constexpr int kMaxChunks = 120;                      // host->device chunks of one pipelined scan
constexpr int kCounterWords = 8 + kMaxChunks + 8;    // [0..7] scalars, [8..] one work counter per launch

namespace pwmscan {

int set_err(pwmscan_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    g_err = msg;
    return code;
}

int cuda_fail(pwmscan_ctx* ctx, cudaError_t e, const char* what) {
    std::string m = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    cudaGetLastError();
    // Every error path returns through here, often with asynchronous copies into caller / stack / std::vector memory already
    // enqueued: nothing may still be in flight when that memory goes away.  (Errors are rare; the wait costs nothing that matters.)
    if (ctx) { cudaDeviceSynchronize(); cudaGetLastError(); }
    return set_err(ctx, PWMSCAN_ECUDA, m);
}

void* scratch_get(pwmscan_ctx* ctx, int slot, size_t bytes) {
    if (ctx->scratch_bytes[slot] >= bytes && ctx->scratch[slot]) return ctx->scratch[slot];
    if (ctx->scratch[slot]) { cudaFree(ctx->scratch[slot]); ctx->scratch[slot] = nullptr; ctx->scratch_bytes[slot] = 0; }
    size_t want = bytes + bytes / 4 + 256;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { cuda_fail(ctx, e, "cudaMalloc(scratch)"); return nullptr; }
    ctx->scratch[slot] = p; ctx->scratch_bytes[slot] = want;
    return p;
}

void* lane_scratch_get(pwmscan_ctx* ctx, ScanLane* lane, int slot, size_t bytes) {
    if (lane->scratch_bytes[slot] >= bytes && lane->scratch[slot]) return lane->scratch[slot];
    if (lane->scratch[slot]) { cudaFree(lane->scratch[slot]); lane->scratch[slot] = nullptr; lane->scratch_bytes[slot] = 0; }
    size_t want = bytes + bytes / 4 + 256;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { cuda_fail(ctx, e, "cudaMalloc(lane scratch)"); return nullptr; }
    lane->scratch[slot] = p; lane->scratch_bytes[slot] = want;
    return p;
}

void* pinned_get(pwmscan_ctx* ctx, int slot, size_t bytes) {
    if (ctx->pinned_bytes[slot] >= bytes && ctx->pinned[slot]) return ctx->pinned[slot];
    if (ctx->pinned[slot]) { cudaFreeHost(ctx->pinned[slot]); ctx->pinned[slot] = nullptr; ctx->pinned_bytes[slot] = 0; }
    size_t want = bytes + bytes / 4 + 256;
    void* p = nullptr;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) { cuda_fail(ctx, e, "cudaMallocHost"); return nullptr; }
    ctx->pinned[slot] = p; ctx->pinned_bytes[slot] = want;
    return p;
}

// Smallest cached buffer that holds `bytes` and wastes at most `slack` times the request (+16 MB).  A miss means a
// cudaMalloc / cudaMallocHost, and a full pool a cudaFree / cudaFreeHost — both synchronise the whole device and stall every
// scan in flight, so the match is generous.
static void* pool_take(std::vector<pwmscan_ctx::PoolBuf>& pool, size_t bytes, size_t* bytes_out, size_t slack) {
    int best = -1;
    for (size_t i = 0; i < pool.size(); ++i)
        if (pool[i].bytes >= bytes && pool[i].bytes <= slack * bytes + ((size_t)16 << 20) && (best < 0 || pool[i].bytes < pool[best].bytes)) best = (int)i;
    if (best < 0) return nullptr;
    void* p = pool[best].p;
    *bytes_out = pool[best].bytes;
    pool.erase(pool.begin() + best);
    return p;
}

void* dev_pool_get(pwmscan_ctx* ctx, size_t bytes, size_t* bytes_out) {
    {
        std::lock_guard<std::mutex> lk(ctx->pool_mu);
        if (void* p = pool_take(ctx->dev_pool, bytes, bytes_out, 4)) return p;
    }
    void* p = nullptr;
    const size_t want = bytes + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {                       // release the cache and retry once
        std::vector<pwmscan_ctx::PoolBuf> drop;
        { std::lock_guard<std::mutex> lk(ctx->pool_mu); drop.swap(ctx->dev_pool); }
        for (auto& b : drop) cudaFree(b.p);
        cudaGetLastError();
        e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) { cuda_fail(ctx, e, "cudaMalloc"); return nullptr; }
    *bytes_out = want;
    return p;
}

void dev_pool_put(pwmscan_ctx* ctx, void* p, size_t bytes) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(ctx->pool_mu);
    if (ctx->dev_pool.size() >= 24) { cudaFree(ctx->dev_pool.front().p); ctx->dev_pool.erase(ctx->dev_pool.begin()); }
    ctx->dev_pool.push_back({p, bytes});
}

// >= n zero int32 values that nobody ever writes (calloc: untouched pages cost nothing)
static const int32_t* zero_column(pwmscan_ctx* ctx, size_t n) {
    std::lock_guard<std::mutex> lk(ctx->pool_mu);
    if (n > ctx->zero_count || ctx->zero_blocks.empty()) {
        const size_t want = std::max<size_t>(n + n / 2, 1 << 20);
        void* p = std::calloc(want, sizeof(int32_t));
        if (!p) return nullptr;
        ctx->zero_blocks.push_back(p);
        ctx->zero_count = want;
    }
    return (const int32_t*)ctx->zero_blocks.back();
}

void* pin_pool_get(pwmscan_ctx* ctx, size_t bytes, size_t* bytes_out) {
    {
        std::lock_guard<std::mutex> lk(ctx->pool_mu);
        if (void* p = pool_take(ctx->pin_pool, bytes, bytes_out, 64)) return p;
    }
    void* p = nullptr;
    const size_t want = bytes + bytes / 8 + 4096;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) { cuda_fail(ctx, e, "cudaMallocHost"); return nullptr; }
    *bytes_out = want;
    return p;
}

void pin_pool_put(pwmscan_ctx* ctx, void* p, size_t bytes) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(ctx->pool_mu);
    if (ctx->pin_pool.size() >= 16) { cudaFreeHost(ctx->pin_pool.front().p); ctx->pin_pool.erase(ctx->pin_pool.begin()); }
    ctx->pin_pool.push_back({p, bytes});
}

}  // namespace pwmscan

extern "C" int pwmscan_abi_version(void) { return PWMSCAN_ABI_VERSION; }

extern "C" const char* pwmscan_last_error(pwmscan_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" int pwmscan_ctx_create(int device, pwmscan_ctx** out) {
    if (!out) return set_err(nullptr, PWMSCAN_EINVAL, "pwmscan_ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return set_err(nullptr, PWMSCAN_ECUDA,
                       std::string("pwmscan_ctx_create: no CUDA device available (") +
                           (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                           "); this library has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return set_err(nullptr, PWMSCAN_EINVAL, "pwmscan_ctx_create: bad device index");
    pwmscan_ctx* ctx = new (std::nothrow) pwmscan_ctx();
    if (!ctx) return set_err(nullptr, PWMSCAN_EINVAL, "out of host memory");
    ctx->device = device;
    auto fail = [&](cudaError_t err, const char* what) { int rc = cuda_fail(nullptr, err, what); delete ctx; return rc; };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(e, "cudaSetDevice");
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    if ((e = cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_greatest)) != cudaSuccess) return fail(e, "cudaStreamCreate");
    {
        const char* cp = std::getenv("PWMSCAN_COPY_PRIO");                    // experiments: 0 = lowest (default priority), 1 = highest
        const int copy_prio = (cp && std::atoi(cp) == 0) ? prio_least : prio_greatest;
        if ((e = cudaStreamCreateWithPriority(&ctx->copy_stream, cudaStreamNonBlocking, copy_prio)) != cudaSuccess) return fail(e, "cudaStreamCreate");
        if ((e = cudaStreamCreateWithPriority(&ctx->d2h_stream, cudaStreamNonBlocking, copy_prio)) != cudaSuccess) return fail(e, "cudaStreamCreate");
    }
    for (auto& ev : ctx->ev) if ((e = cudaEventCreate(&ev)) != cudaSuccess) return fail(e, "cudaEventCreate");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(e, "cudaGetDeviceProperties");
    if (prop.major < 10) { delete ctx; return set_err(nullptr, PWMSCAN_ECUDA, "pwmscan: device is not sm_100 (Blackwell B200) class"); }
    ctx->sm_count = prop.multiProcessorCount;
    if ((e = cudaMalloc(&ctx->d_counters, kCounterWords * sizeof(unsigned long long))) != cudaSuccess) return fail(e, "cudaMalloc(counters)");
    if ((e = cudaMallocHost(&ctx->h_counters, 8 * sizeof(unsigned long long))) != cudaSuccess) return fail(e, "cudaMallocHost(counters)");
    if ((e = cudaEventCreateWithFlags(&ctx->prefilter_chain, cudaEventDisableTiming)) != cudaSuccess) return fail(e, "cudaEventCreate");
    ScanLane* l0 = new (std::nothrow) ScanLane();
    if (!l0) { delete ctx; return set_err(nullptr, PWMSCAN_EINVAL, "out of host memory"); }
    l0->stream = ctx->stream; l0->copy_stream = ctx->copy_stream; l0->d_counters = ctx->d_counters; l0->h_counters = ctx->h_counters;
    ctx->lanes.push_back(l0);
    if ((e = cudaStreamCreateWithPriority(&l0->pstream, cudaStreamNonBlocking, prio_least)) != cudaSuccess) return fail(e, "cudaStreamCreate");
    for (auto& ev : l0->ev) if ((e = cudaEventCreate(&ev)) != cudaSuccess) return fail(e, "cudaEventCreate");
    *out = ctx;
    return PWMSCAN_OK;
}

// Lane `idx` of the context (created on first use; ctx->mu held).
static int lane_get(pwmscan_ctx* ctx, int idx, ScanLane** out) {
    while ((int)ctx->lanes.size() <= idx) {
        ScanLane* l = new (std::nothrow) ScanLane();
        if (!l) return set_err(ctx, PWMSCAN_EINVAL, "out of host memory");
        ctx->lanes.push_back(l);
        l->owns_streams = true;
        int prio_least = 0, prio_greatest = 0;
        PWM_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
        PWM_CUDA(cudaStreamCreateWithPriority(&l->stream, cudaStreamNonBlocking, prio_greatest));
        PWM_CUDA(cudaStreamCreateWithPriority(&l->pstream, cudaStreamNonBlocking, prio_least));
        // ONE host->device copy stream for all lanes: uploads land in the order the items were started.  (With a copy stream per
        // lane the copy engine served the queued uploads of scan_host_many in an order of its own — item 1's bytes sixth of seven
        // — and the first stages, which run in item order, waited for them.)
        l->copy_stream = ctx->copy_stream;
        for (auto& ev : l->ev) PWM_CUDA(cudaEventCreate(&ev));
        PWM_CUDA(cudaMalloc(&l->d_counters, kCounterWords * sizeof(unsigned long long)));
        PWM_CUDA(cudaMallocHost(&l->h_counters, 8 * sizeof(unsigned long long)));
    }
    *out = ctx->lanes[idx];
    return PWMSCAN_OK;
}

extern "C" void pwmscan_ctx_destroy(pwmscan_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (ScanLane* l : ctx->lanes) {
        if (l->owns_streams && l->stream) cudaStreamSynchronize(l->stream);
        if (l->pstream) { cudaStreamSynchronize(l->pstream); cudaStreamDestroy(l->pstream); }
        for (auto& p : l->scratch) if (p) cudaFree(p);
        for (auto& ev : l->ev) if (ev) cudaEventDestroy(ev);
        for (auto& ev : l->chunk_ev) cudaEventDestroy(ev);
        if (l->owns_streams) {
            if (l->d_counters) cudaFree(l->d_counters);
            if (l->h_counters) cudaFreeHost(l->h_counters);
            if (l->stream) cudaStreamDestroy(l->stream);
        }
        delete l;
    }
    if (ctx->prefilter_chain) cudaEventDestroy(ctx->prefilter_chain);
    for (auto& p : ctx->scratch) if (p) cudaFree(p);
    for (auto& p : ctx->pinned) if (p) cudaFreeHost(p);
    for (auto& b : ctx->dev_pool) cudaFree(b.p);
    for (auto& b : ctx->pin_pool) cudaFreeHost(b.p);
    for (void* z : ctx->zero_blocks) std::free(z);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->d2h_stream) { cudaStreamSynchronize(ctx->d2h_stream); cudaStreamDestroy(ctx->d2h_stream); }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int pwmscan_last_timing(pwmscan_ctx* ctx, double* out, int n) {
    if (!ctx || !out) return PWMSCAN_EINVAL;
    for (int i = 0; i < n && i < 12; ++i) out[i] = ctx->timing[i];
    return PWMSCAN_OK;
}

// ------------------------------------------------------------------------------------------------
// sequence upload + packing
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int kThreads = 1024;                         // prefilter CTA: 32 warps, 1 CTA per SM
constexpr int kWarps = kThreads / 32;
constexpr int kIters = 16;                             // x 32 anchors per warp per work item
constexpr int64_t kChunk = (int64_t)kWarps * 32 * kIters;   // anchors per work item (16384) x replication
constexpr int64_t kAnchorRound = kChunk * 32;               // anchor ranges are rounded to the largest work item

__constant__ unsigned char c_iupac_lut[256];

// Four ASCII bytes of one 32-bit word -> their 2-bit codes packed into 8 bits (byte k at bits 2k..2k+1) and a byte-wise
// "not one of ACGT" flag (0x80 per offending byte), without a table: with x = bits 1..2 of a letter (A 0, C 1, T 2, G 3)
// the letter an ACGT byte must be is 0x41 + 2x (+ 15 for x == 2), the code is x ^ (x >> 1), and a multiply gathers the
// four 2-bit fields.  `fold` = 0xDFDFDFDF upper-cases letters (fromBS's toUpper, Seq.hs:86-95), 0xFFFFFFFF leaves them.
__device__ __forceinline__ uint32_t pack4(uint32_t w, uint32_t fold, uint32_t* bad) {
    const uint32_t v = w & fold;
    const uint32_t x = (v >> 1) & 0x03030303u;
    const uint32_t m = (x >> 1) & ~x & 0x01010101u;
    const uint32_t expect = 0x41414141u + (x << 1) + (m << 4) - m;
    const uint32_t d = v ^ expect;
    *bad |= (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;
    const uint32_t code = x ^ ((x >> 1) & 0x01010101u);
    return (code * 0x01041040u) >> 24;
}

// One thread per 32 stored bases: two packed words + one mask word.  Fast path (32 aligned bytes inside the sequence, all
// of them A/C/G/T, no ASCII copy wanted): two 16-byte loads and ~4 integer operations per base; anything else (sequence
// ends, N runs and other IUPAC letters, invalid bytes, KEEP_ASCII) takes the per-byte path below.
__global__ void __launch_bounds__(256) pack_kernel(const uint8_t* ascii, int64_t len, int uppercase, uint32_t* __restrict__ seq2,
                            uint32_t* __restrict__ mask, uint8_t* ascii_out /* may alias ascii */, int64_t t0, int64_t t1,
                            unsigned long long* first_bad /* [0] non-ACGT, [1] non-IUPAC */) {
    const int64_t t = t0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= t1) return;
    const int64_t u0 = t * 32;
    uint32_t w0 = 0, w1 = 0, mk = 0;
    unsigned long long bad_acgt = ~0ull, bad_iupac = ~0ull;
    const int64_t p0 = u0 - kPadFront;
    __align__(16) uint8_t buf[32];
    const bool interior = (p0 >= 0) && (p0 + 32 <= len) && ((((uintptr_t)(ascii + p0)) & 15) == 0);
    if (interior) {
        const uint4 a = __ldcs(reinterpret_cast<const uint4*>(ascii + p0));          // streamed once: do not keep it in the caches
        const uint4 b = __ldcs(reinterpret_cast<const uint4*>(ascii + p0 + 16));
        if (!ascii_out) {
            const uint32_t fold = uppercase ? 0xDFDFDFDFu : 0xFFFFFFFFu;
            uint32_t bad = 0;
            const uint32_t c0 = pack4(a.x, fold, &bad), c1 = pack4(a.y, fold, &bad), c2 = pack4(a.z, fold, &bad), c3 = pack4(a.w, fold, &bad);
            const uint32_t c4 = pack4(b.x, fold, &bad), c5 = pack4(b.y, fold, &bad), c6 = pack4(b.z, fold, &bad), c7 = pack4(b.w, fold, &bad);
            if (bad == 0u) {
                seq2[2 * t] = c0 | (c1 << 8) | (c2 << 16) | (c3 << 24);
                seq2[2 * t + 1] = c4 | (c5 << 8) | (c6 << 16) | (c7 << 24);
                mask[t] = 0u;
                return;
            }
        }
        *reinterpret_cast<uint4*>(buf) = a;
        *reinterpret_cast<uint4*>(buf + 16) = b;
    } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) { const int64_t p = p0 + k; buf[k] = (p >= 0 && p < len) ? ascii[p] : 0; }
    }
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        const int64_t p = p0 + k;
        unsigned char ch = buf[k];
        if (uppercase && ch >= 'a' && ch <= 'z') ch = (unsigned char)(ch - 32);
        buf[k] = ch;
        const bool inside = (p >= 0 && p < len);
        const unsigned code = inside ? c_iupac_lut[ch] : 15u;
        uint32_t c2;
        if (code < 4) c2 = code;
        else {
            c2 = pad_hash(u0 + k);
            mk |= 1u << k;
            if (inside) {
                if (bad_acgt == ~0ull) bad_acgt = (unsigned long long)p;
                if (code == 15 && bad_iupac == ~0ull) bad_iupac = (unsigned long long)p;
            }
        }
        if (k < 16) w0 |= c2 << (2 * k); else w1 |= c2 << (2 * (k - 16));
    }
    seq2[2 * t] = w0;
    seq2[2 * t + 1] = w1;
    mask[t] = mk;
    if (ascii_out) {
        if (interior) {
            *reinterpret_cast<uint4*>(ascii_out + p0) = *reinterpret_cast<uint4*>(buf);
            *reinterpret_cast<uint4*>(ascii_out + p0 + 16) = *reinterpret_cast<uint4*>(buf + 16);
        } else {
            for (int k = 0; k < 32; ++k) { const int64_t p = p0 + k; if (p >= 0 && p < len) ascii_out[p] = buf[k]; }
        }
    }
    if (bad_acgt != ~0ull) atomicMin(&first_bad[0], bad_acgt);
    if (bad_iupac != ~0ull) atomicMin(&first_bad[1], bad_iupac);
}

int ensure_lut(pwmscan_ctx* ctx) {      // ctx->mu held
    if (ctx->lut_ready) return PWMSCAN_OK;
    unsigned char lut[256];
    for (int i = 0; i < 256; ++i) lut[i] = (unsigned char)iupac_code((unsigned char)i);
    PWM_CUDA(cudaMemcpyToSymbol(c_iupac_lut, lut, 256));
    ctx->lut_ready = true;
    return PWMSCAN_OK;
}

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

}  // namespace

static int seq_check_args(pwmscan_ctx* ctx, const uint8_t* ascii, const int64_t* seg_off, int32_t nseg) {
    if (nseg < 1 || nseg > PWMSCAN_MAX_SEGMENTS) return set_err(ctx, PWMSCAN_ERANGE, "pwmscan_seq_upload: segment count out of range");
    const int64_t len = seg_off[nseg];
    if (len < 0 || (len > 0 && !ascii)) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_seq_upload: bad length / NULL data");
    if (len >= ((int64_t)1 << 39)) return set_err(ctx, PWMSCAN_ERANGE, "pwmscan_seq_upload: sequence too long");
    for (int32_t i = 0; i < nseg; ++i) {
        if (seg_off[i + 1] < seg_off[i] || seg_off[0] != 0) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_seq_upload: seg_off must be ascending from 0");
        if (seg_off[i + 1] - seg_off[i] >= ((int64_t)1 << 32)) return set_err(ctx, PWMSCAN_ERANGE, "pwmscan_seq_upload: segment longer than 2^32");
    }
    return PWMSCAN_OK;
}

// Allocates the device planes of a sequence (from the pool) and uploads the segment table.
// *d_raw_out is where the ASCII bytes must land (the kept copy, or scratch slot 0).  ctx->mu held.
// lane_owned: the planes live in the lane's grow-only scratch (the sequence of a pwmscan_scan_host item exists only for
// the duration of its scan): no allocation or release per item.
static int seq_alloc(pwmscan_ctx* ctx, ScanLane* lane, const int64_t* seg_off, int32_t nseg, int flags, pwmscan_seq** out, uint8_t** d_raw_out,
                     bool lane_owned = false, bool want_raw = true) {
    const int64_t len = seg_off[nseg];
    pwmscan_seq* s = new (std::nothrow) pwmscan_seq();
    if (!s) return set_err(ctx, PWMSCAN_EINVAL, "out of host memory");
    s->ctx = ctx; s->len = len; s->nseg = nseg; s->flags = flags;
    s->seg_off.assign(seg_off, seg_off + nseg + 1);
    const int64_t n_anchor = round_up(len + kPadFront + 64, kAnchorRound);
    s->nwords2 = n_anchor / 16 + 8;
    s->nwords_mask = n_anchor / 32 + 4;
    auto bail = [&](int code) { pwmscan_seq_free(s); return code; };
    cudaError_t e;
    s->borrowed = lane_owned;
    if (lane_owned) {
        s->d_seq2_alloc = (uint32_t*)lane_scratch_get(ctx, lane, 9, (size_t)(s->nwords2 + 1) * 4);
        s->d_mask = (uint32_t*)lane_scratch_get(ctx, lane, 10, (size_t)s->nwords_mask * 4);
        s->d_seg_off = (int64_t*)lane_scratch_get(ctx, lane, 11, (size_t)(nseg + 1) * 8);
        if (!s->d_seq2_alloc || !s->d_mask || !s->d_seg_off) return bail(PWMSCAN_ECUDA);
    } else {
        if (!(s->d_seq2_alloc = (uint32_t*)dev_pool_get(ctx, (size_t)(s->nwords2 + 1) * 4, &s->cap_seq2))) return bail(PWMSCAN_ECUDA);
        if (!(s->d_mask = (uint32_t*)dev_pool_get(ctx, (size_t)s->nwords_mask * 4, &s->cap_mask))) return bail(PWMSCAN_ECUDA);
        if (!(s->d_seg_off = (int64_t*)dev_pool_get(ctx, (size_t)(nseg + 1) * 8, &s->cap_seg))) return bail(PWMSCAN_ECUDA);
    }
    s->d_seq2 = s->d_seq2_alloc + 1;
    if ((e = cudaMemcpyAsync(s->d_seg_off, s->seg_off.data(), (size_t)(nseg + 1) * 8, cudaMemcpyHostToDevice, lane->stream)) != cudaSuccess)
        return bail(cuda_fail(ctx, e, "cudaMemcpyAsync(seg_off)"));
    // the raw bytes: either kept (upper-cased in place by the pack kernel) or a scratch buffer (none for packed input)
    if (!want_raw) {
        *d_raw_out = nullptr;
    } else if ((flags & PWMSCAN_SEQ_KEEP_ASCII) && lane_owned) {
        if (!(s->d_ascii = (uint8_t*)lane_scratch_get(ctx, lane, 0, (size_t)len + 64))) return bail(PWMSCAN_ECUDA);
        *d_raw_out = s->d_ascii;
    } else if (flags & PWMSCAN_SEQ_KEEP_ASCII) {
        if (!(s->d_ascii = (uint8_t*)dev_pool_get(ctx, (size_t)len + 64, &s->cap_ascii))) return bail(PWMSCAN_ECUDA);
        *d_raw_out = s->d_ascii;
    } else {
        *d_raw_out = (uint8_t*)lane_scratch_get(ctx, lane, 0, (size_t)len + 64);
        if (!*d_raw_out) return bail(PWMSCAN_ECUDA);
    }
    // (the fromBS verdict words: for the sequence of a scan job the job's reset kernel sets them — no copy-engine operation on its chain)
    if (!lane_owned && (e = cudaMemsetAsync(lane->d_counters + 4, 0xFF, 16, lane->stream)) != cudaSuccess) return bail(cuda_fail(ctx, e, "cudaMemsetAsync"));
    *out = s;
    return PWMSCAN_OK;
}

static int seq_launch_pack(pwmscan_ctx* ctx, ScanLane* lane, cudaStream_t stream, pwmscan_seq* s, const uint8_t* d_raw, int64_t t0, int64_t t1) {
    if (t1 <= t0) return PWMSCAN_OK;
    const int bs = 256;
    pack_kernel<<<(unsigned)((t1 - t0 + bs - 1) / bs), bs, 0, stream>>>(
        d_raw, s->len, (s->flags & PWMSCAN_SEQ_UPPERCASE) ? 1 : 0, s->d_seq2, s->d_mask,
        (s->flags & PWMSCAN_SEQ_KEEP_ASCII) ? s->d_ascii : nullptr, t0, t1, lane->d_counters + 4);
    PWM_CUDA(cudaGetLastError());
    return PWMSCAN_OK;
}

// Packed input (hostpack.cu): the planes of positions [0, len) arrive ready-made; what is left to do on the device is the
// padding around them (pack_kernel with an empty sequence writes exactly that), the mask bits behind `len` in the last unit
// (a slice of a longer sequence carries real bases there) and the `Basic` verdict, which is the first set mask bit.
namespace {
__global__ void packed_finish_kernel(uint32_t* __restrict__ mask, int64_t units, int64_t len, unsigned long long* first_bad) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= units) return;
    uint32_t m = mask[1 + t];
    const int64_t left = len - 32 * t;
    if (left < 32) { const uint32_t tail = ~0u << (int)left; mask[1 + t] = m | tail; m &= ~tail; }
    if (m) atomicMin(&first_bad[0], (unsigned long long)(32 * t + (__ffs((int)m) - 1)));
}
}  // namespace

static int seq_launch_packed_finish(pwmscan_ctx* ctx, ScanLane* lane, cudaStream_t stream, pwmscan_seq* s) {
    const int64_t units = (s->len + 31) / 32;
    const int bs = 256;
    // padding: stored units [0, 1) in front, [1 + units, nwords_mask) behind
    pack_kernel<<<1, bs, 0, stream>>>(nullptr, 0, 0, s->d_seq2, s->d_mask, nullptr, 0, 1, lane->d_counters + 4);
    const int64_t t0 = 1 + units, t1 = s->nwords_mask;
    if (t1 > t0) pack_kernel<<<(unsigned)((t1 - t0 + bs - 1) / bs), bs, 0, stream>>>(nullptr, 0, 0, s->d_seq2, s->d_mask, nullptr, t0, t1, lane->d_counters + 4);
    if (units > 0) packed_finish_kernel<<<(unsigned)((units + bs - 1) / bs), bs, 0, stream>>>(s->d_mask, units, s->len, lane->d_counters + 4);
    PWM_CUDA(cudaGetLastError());
    return PWMSCAN_OK;
}

static int seq_finish(ScanLane* lane, pwmscan_seq* s) {   // after a stream sync: fetch the fromBS validation result
    s->first_non_acgt = lane->h_counters[4] == ~0ull ? -1 : (int64_t)lane->h_counters[4];
    s->first_non_iupac = lane->h_counters[5] == ~0ull ? -1 : (int64_t)lane->h_counters[5];
    return PWMSCAN_OK;
}

static int seq_upload_impl(pwmscan_ctx* ctx, const uint8_t* ascii, const int64_t* seg_off, int32_t nseg, int flags,
                           pwmscan_seq** out) {
    if (!ctx || !out) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_seq_upload: NULL argument");
    *out = nullptr;
    int rc = seq_check_args(ctx, ascii, seg_off, nseg);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    PWM_CUDA(cudaSetDevice(ctx->device));
    if ((rc = ensure_lut(ctx))) return rc;
    pwmscan_seq* s = nullptr;
    uint8_t* d_raw = nullptr;
    ScanLane* lane = ctx->lanes[0];
    if ((rc = seq_alloc(ctx, lane, seg_off, nseg, flags, &s, &d_raw))) return rc;
    auto bail = [&](int code) { pwmscan_seq_free(s); return code; };
    const int64_t len = s->len;
    cudaError_t e;
    if (len > 0 && (e = cudaMemcpyAsync(d_raw, ascii, (size_t)len, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess)
        return bail(cuda_fail(ctx, e, "cudaMemcpyAsync(ascii)"));
    if ((rc = seq_launch_pack(ctx, lane, lane->stream, s, d_raw, 0, s->nwords_mask))) return bail(rc);
    if ((e = cudaMemcpyAsync(ctx->h_counters + 4, ctx->d_counters + 4, 16, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess)
        return bail(cuda_fail(ctx, e, "cudaMemcpyAsync(first_bad)"));
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return bail(cuda_fail(ctx, e, "pack_kernel"));
    seq_finish(lane, s);
    ctx->timing[4] = (double)len + (double)(nseg + 1) * 8;
    *out = s;
    return PWMSCAN_OK;
}

extern "C" int pwmscan_seq_upload(pwmscan_ctx* ctx, const uint8_t* ascii, int64_t len, int flags, pwmscan_seq** out) {
    const int64_t off[2] = {0, len};
    return seq_upload_impl(ctx, ascii, off, 1, flags, out);
}

extern "C" int pwmscan_seq_upload_segments(pwmscan_ctx* ctx, const uint8_t* ascii, const int64_t* seg_off, int32_t nseg,
                                           int flags, pwmscan_seq** out) {
    if (!seg_off) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_seq_upload_segments: seg_off is NULL");
    return seq_upload_impl(ctx, ascii, seg_off, nseg, flags, out);
}

extern "C" int pwmscan_seq_upload_packed(pwmscan_ctx* ctx, const uint32_t* seq2, const uint32_t* mask, int64_t len, pwmscan_seq** out) {
    if (!ctx || !out) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_seq_upload_packed: NULL argument");
    *out = nullptr;
    if (len < 0 || (len > 0 && (!seq2 || !mask))) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_seq_upload_packed: bad length / NULL planes");
    if (len >= ((int64_t)1 << 32)) return set_err(ctx, PWMSCAN_ERANGE, "pwmscan_seq_upload_packed: sequence longer than 2^32");
    std::lock_guard<std::mutex> lk(ctx->mu);
    PWM_CUDA(cudaSetDevice(ctx->device));
    int rc = ensure_lut(ctx);
    if (rc) return rc;
    pwmscan_seq* s = nullptr;
    uint8_t* d_raw = nullptr;
    ScanLane* lane = ctx->lanes[0];
    const int64_t off[2] = {0, len};
    if ((rc = seq_alloc(ctx, lane, off, 1, 0, &s, &d_raw, false, false))) return rc;
    auto bail = [&](int code) { pwmscan_seq_free(s); return code; };
    const int64_t units = (len + 31) / 32;
    cudaError_t e;
    if (units > 0) {
        if ((e = cudaMemcpyAsync(s->d_seq2 + 2, seq2, (size_t)units * 8, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) return bail(cuda_fail(ctx, e, "cudaMemcpyAsync(seq2)"));
        if ((e = cudaMemcpyAsync(s->d_mask + 1, mask, (size_t)units * 4, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) return bail(cuda_fail(ctx, e, "cudaMemcpyAsync(mask)"));
    }
    if ((rc = seq_launch_packed_finish(ctx, lane, ctx->stream, s))) return bail(rc);
    if ((e = cudaMemcpyAsync(ctx->h_counters + 4, ctx->d_counters + 4, 16, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess)
        return bail(cuda_fail(ctx, e, "cudaMemcpyAsync(first_bad)"));
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return bail(cuda_fail(ctx, e, "packed upload"));
    seq_finish(lane, s);
    s->first_non_iupac = -1;                   // not recoverable from the planes: pwmscan_pack_host reported it when they were made
    ctx->timing[4] = (double)units * 12 + 16;
    *out = s;
    return PWMSCAN_OK;
}

extern "C" int64_t pwmscan_seq_length(const pwmscan_seq* seq) { return seq ? seq->len : -1; }

extern "C" int pwmscan_seq_first_invalid(pwmscan_ctx* ctx, const pwmscan_seq* seq, int alphabet, int64_t* pos_out) {
    if (!seq || !pos_out) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_seq_first_invalid: NULL argument");
    *pos_out = alphabet == 0 ? seq->first_non_acgt : seq->first_non_iupac;
    return PWMSCAN_OK;
}

extern "C" void pwmscan_seq_free(pwmscan_seq* s) {
    if (!s) return;
    if (s->ctx && !s->borrowed) {
        dev_pool_put(s->ctx, s->d_seq2_alloc, s->cap_seq2);
        dev_pool_put(s->ctx, s->d_mask, s->cap_mask);
        dev_pool_put(s->ctx, s->d_ascii, s->cap_ascii);
        dev_pool_put(s->ctx, s->d_seg_off, s->cap_seg);
    }
    delete s;
}

// ------------------------------------------------------------------------------------------------
// motifs
// ------------------------------------------------------------------------------------------------
extern "C" int pwmscan_motifs_create(pwmscan_ctx* ctx, int32_t M, const int32_t* ncol, const double* mats,
                                     const double bg[4], pwmscan_motifs** out) {
    if (!ctx || !out || !ncol || !mats || !bg) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_motifs_create: NULL argument");
    *out = nullptr;
    if (M < 1 || M > PWMSCAN_MAX_MOTIFS) return set_err(ctx, PWMSCAN_ERANGE, "pwmscan_motifs_create: motif count out of range");
    for (int m = 0; m < M; ++m)
        if (ncol[m] < 1 || ncol[m] > PWMSCAN_MAX_COLS) return set_err(ctx, PWMSCAN_ERANGE, "pwmscan_motifs_create: PWM must have 1..64 columns");
    std::lock_guard<std::mutex> lk(ctx->mu);
    PWM_CUDA(cudaSetDevice(ctx->device));
    pwmscan_motifs* mo = new (std::nothrow) pwmscan_motifs();
    if (!mo) return set_err(ctx, PWMSCAN_EINVAL, "out of host memory");
    mo->ctx = ctx; mo->M = M;
    for (int b = 0; b < 4; ++b) mo->bg[b] = bg[b];
    mo->mh.resize(M);
    mo->tab_off.resize(2 * (size_t)M);
    mo->col_off.resize(2 * (size_t)M);
    int64_t off = 0, cols = 0;
    const double* p = mats;
    for (int m = 0; m < M; ++m) {
        motif_prepare(mo->mh[m], ncol[m], p, bg);
        p += 4 * ncol[m];
        for (int s = 0; s < 2; ++s) { mo->tab_off[2 * m + s] = off; off += 16 * (int64_t)ncol[m]; mo->col_off[2 * m + s] = cols; cols += ncol[m]; }
    }
    mo->total_cols2 = cols;
    std::vector<double> h_tab((size_t)off), h_mat((size_t)cols * 4);
    for (int m = 0; m < M; ++m)
        for (int s = 0; s < 2; ++s) {
            std::memcpy(h_tab.data() + mo->tab_off[2 * m + s], mo->mh[m].tab16[s].data(), sizeof(double) * 16 * ncol[m]);
            std::memcpy(h_mat.data() + 4 * mo->col_off[2 * m + s], mo->mh[m].mat[s].data(), sizeof(double) * 4 * ncol[m]);
        }
    cudaError_t e;
    auto bail = [&](int code) { pwmscan_motifs_free(mo); return code; };
    if ((e = cudaMalloc(&mo->d_tab16, h_tab.size() * 8)) != cudaSuccess) return bail(cuda_fail(ctx, e, "cudaMalloc(tab16)"));
    if ((e = cudaMalloc(&mo->d_mats, h_mat.size() * 8)) != cudaSuccess) return bail(cuda_fail(ctx, e, "cudaMalloc(mats)"));
    if ((e = cudaMemcpy(mo->d_tab16, h_tab.data(), h_tab.size() * 8, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(cuda_fail(ctx, e, "cudaMemcpy(tab16)"));
    if ((e = cudaMemcpy(mo->d_mats, h_mat.data(), h_mat.size() * 8, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(cuda_fail(ctx, e, "cudaMemcpy(mats)"));
    {
        std::vector<double> h_tab4((size_t)cols * 4);
        for (int m = 0; m < M; ++m)
            for (int s = 0; s < 2; ++s)
                for (int k = 0; k < ncol[m]; ++k)
                    for (int b = 0; b < 4; ++b) h_tab4[(size_t)(mo->col_off[2 * m + s] + k) * 4 + b] = mo->mh[m].tab16[s][(size_t)16 * k + b];
        if ((e = cudaMalloc(&mo->d_tab4, h_tab4.size() * 8 + 64)) != cudaSuccess) return bail(cuda_fail(ctx, e, "cudaMalloc(tab4)"));
        if ((e = cudaMemcpy(mo->d_tab4, h_tab4.data(), h_tab4.size() * 8, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(cuda_fail(ctx, e, "cudaMemcpy(tab4)"));
    }
    *out = mo;
    return PWMSCAN_OK;
}

extern "C" int pwmscan_motifs_set_sigma(pwmscan_motifs* mo, int32_t m, int strand, const double* sigma) {
    if (!mo || !sigma || m < 0 || m >= mo->M || strand < 0 || strand > 1) return set_err(mo ? mo->ctx : nullptr, PWMSCAN_EINVAL, "pwmscan_motifs_set_sigma: bad argument");
    mo->mh[m].sigma[strand].assign(sigma, sigma + mo->mh[m].n);
    return PWMSCAN_OK;
}

extern "C" int pwmscan_motifs_get_sigma(const pwmscan_motifs* mo, int32_t m, int strand, double* sigma_out) {
    if (!mo || !sigma_out || m < 0 || m >= mo->M || strand < 0 || strand > 1) return set_err(mo ? mo->ctx : nullptr, PWMSCAN_EINVAL, "pwmscan_motifs_get_sigma: bad argument");
    std::memcpy(sigma_out, mo->mh[m].sigma[strand].data(), sizeof(double) * mo->mh[m].n);
    return PWMSCAN_OK;
}

extern "C" int32_t pwmscan_motifs_count(const pwmscan_motifs* mo) { return mo ? mo->M : -1; }

extern "C" void pwmscan_motifs_free(pwmscan_motifs* mo) {
    if (!mo) return;
    if (mo->ctx) cudaSetDevice(mo->ctx->device);
    if (mo->d_tab16) cudaFree(mo->d_tab16);
    if (mo->d_tab4) cudaFree(mo->d_tab4);
    if (mo->d_mats) cudaFree(mo->d_mats);
    delete mo;
}

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
extern "C" int pwmscan_plan_info(const pwmscan_plan* pl, double* out, int n) {
    if (!pl || !out) return PWMSCAN_EINVAL;
    double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    v[0] = pl->nblocks;
    for (const auto& b : pl->blocks) {
        v[1] += (double)b.s1 / b.rep;          // first-stage shared-memory wavefronts (128 B) per sequence position
        v[4] += (double)b.nslot * kSlotBytes;  // shared-memory table bytes
        v[5] += 1.0 / b.rep;                   // passes over the packed genome
        v[6] += b.p_warp;
    }
    v[2] = (double)pl->generic.size();
    v[3] = pl->cand_rate;
    v[7] = pl->mode;
    if (pl->mode == PWM_MODE_KMER) {
        v[0] = pl->km_nblocks;
        v[1] = pl->km_req_rate;                // expected 32-byte table sectors gathered per sequence position
        v[2] = (double)pl->km_generic.size();
        v[3] = pl->km_cand_rate;
        v[4] = 0;
        for (const auto& b : pl->km_blocks) v[4] += (double)b.table_bytes();    // mask table bytes (one block's share is L2-resident at a time)
        v[5] = pl->km_nblocks;
        v[6] = pl->km_surv_rate;               // expected first-stage survivors per position
    } else if (pl->mode == PWM_MODE_MMA) {
        v[0] = pl->mma_nblocks; v[2] = (double)pl->mma_generic.size(); v[3] = pl->mma_cand_rate; v[5] = pl->mma_nblocks;
    }
    for (int i = 0; i < n && i < 8; ++i) out[i] = v[i];
    return PWMSCAN_OK;
}

extern "C" void pwmscan_plan_free(pwmscan_plan* pl) {
    if (!pl) return;
    if (pl->ctx) cudaSetDevice(pl->ctx->device);
    if (pl->d_blocks) cudaFree(pl->d_blocks);
    if (pl->d_combos) cudaFree(pl->d_combos);
    if (pl->d_generic) cudaFree(pl->d_generic);
    if (pl->d_tables) cudaFree(pl->d_tables);
    if (pl->d_th) cudaFree(pl->d_th);
    if (pl->d_mma_blocks) cudaFree(pl->d_mma_blocks);
    if (pl->d_mma_combos) cudaFree(pl->d_mma_combos);
    if (pl->d_mma_generic) cudaFree(pl->d_mma_generic);
    if (pl->d_btiles) cudaFree(pl->d_btiles);
    if (pl->d_km_blocks) cudaFree(pl->d_km_blocks);
    if (pl->d_km_combos) cudaFree(pl->d_km_combos);
    if (pl->d_km_generic) cudaFree(pl->d_km_generic);
    if (pl->d_km_ambig) cudaFree(pl->d_km_ambig);
    if (pl->d_km_tables) { if (pl->ctx) dev_pool_put(pl->ctx, pl->d_km_tables, pl->km_tables_cap); else cudaFree(pl->d_km_tables); }
    if (pl->d_km_q16) cudaFree(pl->d_km_q16);
    delete pl;
}

// Which first stage plans are built for and scans run with (PWMSCAN_PREFILTER=kmer|lds|mma, read once).
static int prefilter_mode() {
    static const int mode = [] {
        const char* e = std::getenv("PWMSCAN_PREFILTER");
        if (!e) return kDefaultMode;
        if (std::strcmp(e, "mma") == 0) return (int)PWM_MODE_MMA;
        if (std::strcmp(e, "lds") == 0) return (int)PWM_MODE_LDS;
        if (std::strcmp(e, "kmer") == 0) return (int)PWM_MODE_KMER;
        return kDefaultMode;
    }();
    return mode;
}

extern "C" int pwmscan_plan_create(pwmscan_ctx* ctx, const pwmscan_motifs* mo, const double* thres, int strands,
                                   int skip_ambiguous, pwmscan_plan** out) {
    if (!ctx || !mo || !thres || !out) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_plan_create: NULL argument");
    *out = nullptr;
    if (strands != 1 && strands != 2) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_plan_create: strands must be 1 or 2");
    std::lock_guard<std::mutex> lk(ctx->mu);
    PWM_CUDA(cudaSetDevice(ctx->device));
    pwmscan_plan* pl = new (std::nothrow) pwmscan_plan();
    if (!pl) return set_err(ctx, PWMSCAN_EINVAL, "out of host memory");
    pl->ctx = ctx; pl->motifs = mo; pl->strands = strands; pl->skip = skip_ambiguous ? 1 : 0;
    pl->mode = prefilter_mode();
    const bool verbose = std::getenv("PWMSCAN_VERBOSE") != nullptr;
    auto dev_combo = [&](int m, int s) {
        DevCombo dc;
        dc.motif = m; dc.strand = s; dc.n = mo->mh[m].n; dc.c = 0;
        dc.tab_off = mo->tab_off[2 * m + s]; dc.th_off = mo->col_off[2 * m + s];
        return dc;
    };
    auto empty_combo = [] { DevCombo dc; dc.motif = -1; dc.strand = 0; dc.n = 0; dc.c = 0; dc.tab_off = 0; dc.th_off = 0; return dc; };
    cudaError_t e;
    auto bail = [&](int code) { pwmscan_plan_free(pl); return code; };
    auto up = [&](void** dst, const void* src, size_t bytes, const char* what) -> int {
        if (bytes == 0) return 0;
        if ((e = cudaMalloc(dst, bytes)) != cudaSuccess) return cuda_fail(ctx, e, what);
        if ((e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(ctx, e, what);
        return 0;
    };
    int rc;
    if (pl->mode == PWM_MODE_LDS) {          // shared-memory 4-mer deficit tables: blocks of 64 motifs
        HostPlan hp;
        const char* fs1 = std::getenv("PWMSCAN_FORCE_S1");
        if (!build_host_plan(mo->mh, mo->col_off, mo->total_cols2, thres, strands, pl->skip, hp, fs1 ? std::atoi(fs1) : 0)) {
            pwmscan_plan_free(pl);
            return set_err(ctx, PWMSCAN_ERANGE, "pwmscan_plan_create: internal: block cannot be placed");
        }
        pl->max_n = hp.max_n; pl->max_nslot = hp.max_nslot;
        for (auto& g : hp.generic) pl->generic.push_back(dev_combo(g.first, g.second));
        pl->blocks = std::move(hp.blocks);
        pl->nblocks = (int)pl->blocks.size();
        pl->cand_rate = hp.cand_rate;
        std::vector<DevBlock> h_blocks(pl->nblocks);
        std::vector<DevCombo> h_combos((size_t)pl->nblocks * kCombosPerBlock);
        std::vector<uint32_t> h_tables;
        for (int b = 0; b < pl->nblocks; ++b) {
            const BlockPlan& bp = pl->blocks[b];
            if (verbose) {
                int nmin = 999, nmax = 0, nen = 0;
                for (int ci = 0; ci < kCombosPerBlock; ++ci) if (bp.combo[ci].motif >= 0) { nmin = std::min(nmin, bp.combo[ci].n); nmax = std::max(nmax, bp.combo[ci].n); nen += bp.combo[ci].enabled; }
                std::fprintf(stderr, "[pwmscan] block %d: s1=%d nl=%d nr=%d nslot=%d Q=%d cols %d..%d enabled=%d rep=%d p_warp=%.4f cost=%.3f\n",
                             b, bp.s1, bp.nl, bp.nr, bp.nslot, bp.Q, nmin, nmax, nen, bp.rep, bp.p_warp, bp.cost);
            }
            h_blocks[b].s1 = bp.s1; h_blocks[b].nl = bp.nl; h_blocks[b].nslot = bp.nslot;
            h_blocks[b].rep_log2 = 0;
            while ((1 << h_blocks[b].rep_log2) < bp.rep) ++h_blocks[b].rep_log2;
            h_blocks[b].table_off = (int64_t)h_tables.size();
            h_tables.insert(h_tables.end(), bp.table.begin(), bp.table.end());
            for (int ci = 0; ci < kCombosPerBlock; ++ci) {
                const ComboPlan& cp = bp.combo[ci];
                DevCombo dc = empty_combo();
                if (cp.motif >= 0) { dc = dev_combo(cp.motif, cp.strand); dc.c = cp.c; }
                h_combos[(size_t)b * kCombosPerBlock + ci] = dc;
            }
        }
        if (verbose) std::fprintf(stderr, "[pwmscan] generic combos=%zu cand_rate=%.3e\n", pl->generic.size(), pl->cand_rate);
        if ((rc = up((void**)&pl->d_th, hp.th.data(), hp.th.size() * 8, "upload th"))) return bail(rc);
        if ((rc = up((void**)&pl->d_blocks, h_blocks.data(), h_blocks.size() * sizeof(DevBlock), "upload blocks"))) return bail(rc);
        if ((rc = up((void**)&pl->d_combos, h_combos.data(), h_combos.size() * sizeof(DevCombo), "upload combos"))) return bail(rc);
        if ((rc = up((void**)&pl->d_tables, h_tables.data(), h_tables.size() * 4, "upload tables"))) return bail(rc);
        if ((rc = up((void**)&pl->d_generic, pl->generic.data(), pl->generic.size() * sizeof(DevCombo), "upload generic"))) return bail(rc);
    } else if (pl->mode == PWM_MODE_MMA) {   // tensor-core first stage: blocks of 128 motifs, any motif length
        MmaHostPlan mp;
        const char* fst = std::getenv("PWMSCAN_FORCE_STEPS");
        build_mma_host_plan(mo->mh, mo->col_off, mo->total_cols2, thres, strands, pl->skip, mp, fst ? std::atoi(fst) : 0);
        pl->max_n = mp.max_n;
        pl->mma_blocks = std::move(mp.blocks);
        pl->mma_nblocks = (int)pl->mma_blocks.size();
        pl->mma_cand_rate = mp.cand_rate;
        pl->cand_rate = mp.cand_rate;
        for (auto& g : mp.generic) pl->mma_generic.push_back(dev_combo(g.first, g.second));
        std::vector<DevMmaBlock> hb(pl->mma_nblocks);
        std::vector<DevCombo> hc((size_t)pl->mma_nblocks * kMmaCombos);
        std::vector<int8_t> ht;
        for (int b = 0; b < pl->mma_nblocks; ++b) {
            const MmaBlockPlan& bp = pl->mma_blocks[b];
            while (ht.size() % 128) ht.push_back(0);
            hb[b].steps = bp.steps; hb[b].ncombo = bp.ncombo; hb[b].b_off = (int64_t)ht.size();
            ht.insert(ht.end(), bp.btiles.begin(), bp.btiles.end());
            for (int ci = 0; ci < kMmaCombos; ++ci) {
                const MmaCombo& mc = bp.combo[ci];
                DevCombo dc = empty_combo();
                if (mc.motif >= 0) { dc = dev_combo(mc.motif, mc.strand); dc.c = mc.c; }
                hc[(size_t)b * kMmaCombos + ci] = dc;
            }
            if (verbose)
                std::fprintf(stderr, "[pwmscan] mma block %d: steps=%d (K1=%d cols) N=%d cand_rate=%.3e cost=%.1f\n", b, bp.steps, 8 * bp.steps, bp.ncombo, bp.cand_rate, bp.cost);
        }
        while (ht.size() % 128) ht.push_back(0);
        if ((rc = up((void**)&pl->d_th, mp.th.data(), mp.th.size() * 8, "upload th"))) return bail(rc);
        if ((rc = up((void**)&pl->d_mma_blocks, hb.data(), hb.size() * sizeof(DevMmaBlock), "upload mma blocks"))) return bail(rc);
        if ((rc = up((void**)&pl->d_mma_combos, hc.data(), hc.size() * sizeof(DevCombo), "upload mma combos"))) return bail(rc);
        if ((rc = up((void**)&pl->d_mma_generic, pl->mma_generic.data(), pl->mma_generic.size() * sizeof(DevCombo), "upload mma generic"))) return bail(rc);
        if ((rc = up((void**)&pl->d_btiles, ht.data(), ht.size(), "upload btiles"))) return bail(rc);
    } else {                                 // k-mer mask tables in L2: blocks of 128 motifs, any motif length
        KmHostPlan kp;
        const char *fnt = std::getenv("PWMSCAN_FORCE_NT"), *fl = std::getenv("PWMSCAN_FORCE_L");
        const auto t_plan0 = std::chrono::steady_clock::now();
        build_km_host_plan(mo->mh, mo->col_off, mo->total_cols2, thres, strands, pl->skip, kp, fnt ? std::atoi(fnt) : 0, fl ? std::atoi(fl) : 0);
        pl->max_n = kp.max_n;
        pl->km_blocks = std::move(kp.blocks);
        pl->km_nblocks = (int)pl->km_blocks.size();
        pl->km_cand_rate = kp.cand_rate; pl->km_surv_rate = kp.surv_rate; pl->km_req_rate = kp.req_rate;
        pl->cand_rate = kp.cand_rate;
        for (auto& g : kp.generic) pl->km_generic.push_back(dev_combo(g.first, g.second));
        for (auto& g : kp.ambig) { pl->km_ambig.push_back(dev_combo(g.first, g.second)); pl->km_ambig_max_n = std::max(pl->km_ambig_max_n, (int)mo->mh[g.first].n); }
        std::vector<DevKmBlock> hb(pl->km_nblocks);
        std::vector<DevCombo> hc((size_t)pl->km_nblocks * kKmCombos);
        std::vector<uint16_t> hq;
        std::vector<double> hw, hd;
        std::vector<size_t> w_off(pl->km_nblocks);
        size_t table_bytes = 0;
        for (int b = 0; b < pl->km_nblocks; ++b) {
            KmBlockPlan& bp = pl->km_blocks[b];
            hb[b].nt = bp.nt; hb[b].max_n = bp.max_n;
            int64_t toff[kKmMaxTables] = {0, 0, 0, 0};
            for (int t = 0; t < bp.nt; ++t) { toff[t] = (int64_t)table_bytes; table_bytes += km_table_bytes(bp.L[t]); }
            for (int s = 0; s < 4; ++s) {
                const int t = s < bp.nt ? bp.order[s] : 0;
                hb[b].shift[s] = 2 * bp.off[t]; hb[b].kmask[s] = (1u << (2 * bp.L[t])) - 1u; hb[b].table_off[s] = toff[t];
            }
            hb[b].bitmap_off = (int64_t)table_bytes; hb[b].bitmap_words = (int32_t)(bp.bitmap_bytes() / 4);
            hb[b].direct = bp.nt == 1 ? 1 : 0;                       // ... if every enabled combo's window [c, c + L) covers its columns [0, n)
            for (int ci = 0; ci < kKmCombos && hb[b].direct; ++ci)
                if (bp.combo[ci].enabled && (bp.combo[ci].c > 0 || bp.combo[ci].c + bp.L[0] < bp.combo[ci].n)) hb[b].direct = 0;
            table_bytes += (bp.bitmap_bytes() + 255) / 256 * 256;
            pl->km_smem_bytes = std::max(pl->km_smem_bytes, bp.q16_bytes() + bp.bitmap_bytes());
            hb[b].q16_off = (int64_t)hq.size();
            hq.insert(hq.end(), bp.q16.begin(), bp.q16.end());
            w_off[b] = hw.size();
            hw.insert(hw.end(), bp.wdef.begin(), bp.wdef.end());
            hd.insert(hd.end(), bp.dsafe.begin(), bp.dsafe.end());
            pl->km_max_n = std::max(pl->km_max_n, bp.max_n);
            int nmin = 999, nmax = 0, nen = 0;
            for (int ci = 0; ci < kKmCombos; ++ci) {
                const KmCombo& kc = bp.combo[ci];
                DevCombo dc = empty_combo();
                if (kc.motif >= 0) { dc = dev_combo(kc.motif, kc.strand); dc.c = kc.c; nmin = std::min(nmin, kc.n); nmax = std::max(nmax, kc.n); nen += kc.enabled; }
                hc[(size_t)b * kKmCombos + ci] = dc;
            }
            if (verbose)
                std::fprintf(stderr, "[pwmscan] kmer block %d: nt=%d L=%d,%d,%d,%d order=%d%d%d%d bitmap=%d cols %d..%d enabled=%d sectors/pos=%.3f survivors/pos=%.4f cand/pos=%.3e cost=%.3f\n",
                             b, bp.nt, bp.L[0], bp.nt > 1 ? bp.L[1] : 0, bp.nt > 2 ? bp.L[2] : 0, bp.nt > 3 ? bp.L[3] : 0, bp.order[0], bp.order[1], bp.order[2], bp.order[3], (int)bp.bitmap,
                             nmin, nmax, nen, bp.req_rate, bp.surv_rate, bp.cand_rate, bp.cost);
            std::vector<double>().swap(bp.wdef);         // only needed for the device build
            std::vector<uint16_t>().swap(bp.q16);
        }
        const auto t_plan1 = std::chrono::steady_clock::now();
        if (verbose) std::fprintf(stderr, "[pwmscan] generic combos=%zu cand_rate=%.3e\n", pl->km_generic.size(), pl->cand_rate);
        if ((rc = up((void**)&pl->d_th, kp.th.data(), kp.th.size() * 8, "upload th"))) return bail(rc);
        if ((rc = up((void**)&pl->d_km_blocks, hb.data(), hb.size() * sizeof(DevKmBlock), "upload kmer blocks"))) return bail(rc);
        if ((rc = up((void**)&pl->d_km_combos, hc.data(), hc.size() * sizeof(DevCombo), "upload kmer combos"))) return bail(rc);
        if ((rc = up((void**)&pl->d_km_generic, pl->km_generic.data(), pl->km_generic.size() * sizeof(DevCombo), "upload kmer generic"))) return bail(rc);
        if ((rc = up((void**)&pl->d_km_ambig, pl->km_ambig.data(), pl->km_ambig.size() * sizeof(DevCombo), "upload kmer ambig"))) return bail(rc);
        if ((rc = up((void**)&pl->d_km_q16, hq.data(), hq.size() * 2, "upload q16"))) return bail(rc);
        if (pl->km_nblocks > 0) {
            // the mask tables are computed on the device from the per-window deficits
            double *d_w = nullptr, *d_d = nullptr;
            // from the context's device pool: cudaMalloc / cudaFree of hundreds of MiB per plan cost tens of ms
            if (!(pl->d_km_tables = (uint8_t*)dev_pool_get(ctx, table_bytes, &pl->km_tables_cap))) return bail(PWMSCAN_ECUDA);
            if ((rc = up((void**)&d_w, hw.data(), hw.size() * 8, "upload window deficits"))) return bail(rc);
            if ((rc = up((void**)&d_d, hd.data(), hd.size() * 8, "upload dsafe"))) { cudaFree(d_w); return bail(rc); }
            for (int b = 0; b < pl->km_nblocks && rc == 0; ++b) {
                const KmBlockPlan& bp = pl->km_blocks[b];
                for (int s = 0; s < bp.nt && rc == 0; ++s) {
                    const int t = bp.order[s];
                    rc = km_build_table(ctx, d_w + w_off[b] + (size_t)t * kKmCombos * kKmMaxL * 4, d_d + (size_t)b * kKmCombos, bp.L[t],
                                        pl->d_km_tables + hb[b].table_off[s]);
                }
                if (rc == 0 && bp.bitmap) rc = km_build_bitmap(ctx, pl->d_km_tables + hb[b].table_off[0], bp.L[bp.order[0]], pl->d_km_tables + hb[b].bitmap_off);
            }
            e = cudaStreamSynchronize(ctx->stream);
            cudaFree(d_w); cudaFree(d_d);
            if (rc) return bail(rc);
            if (e != cudaSuccess) return bail(cuda_fail(ctx, e, "kmer table build"));
        }
        if (verbose) {
            const auto t_plan2 = std::chrono::steady_clock::now();
            std::fprintf(stderr, "[pwmscan] kmer plan: host %.1f ms, upload + device table build (%.0f MiB) %.1f ms\n",
                         std::chrono::duration<double, std::milli>(t_plan1 - t_plan0).count(), table_bytes / 1048576.0,
                         std::chrono::duration<double, std::milli>(t_plan2 - t_plan1).count());
        }
    }
    *out = pl;
    return PWMSCAN_OK;
}

// ------------------------------------------------------------------------------------------------
// kernels of the thresholded scan
// ------------------------------------------------------------------------------------------------
namespace {

// counters: [0] candidates, [1] hits, [2] error flag, [3] work counter
//
// Persistent CTAs (one per SM, 32 warps) pull work items (block, chunk of anchors) from an atomic
// counter; items are block-major so a CTA reloads its shared-memory tables only when it moves to
// the next motif block.  Item b,c covers the anchors [c, c+1) * kChunk * rep_b; warp w takes a
// stretch of 512 * rep_b of them and, when the block is replicated (few motifs), lane replica r
// the r-th 512-anchor part of that stretch.
template <bool USE_TMEM>
__global__ void __launch_bounds__(kThreads, 1)
prefilter_kernel(const uint32_t* __restrict__ seq2, long long anchor0, long long n_anchor /* of this launch */, long long total_items,
                 int iters /* x 32 anchors per lane per item: 16 normally, less for short inputs */, int nblocks, const DevBlock* __restrict__ blocks, const uint32_t* __restrict__ tables,
                 unsigned long long* __restrict__ cand, unsigned long long cand_cap, unsigned long long* counters,
                 unsigned long long* work_counter) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ long long s_item;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t smem_addr = (uint32_t)__cvta_generic_to_shared(smem);
    uint32_t tq = 0;                                             // TMEM address of this warp's 32-lane quarter, column 0
    if (USE_TMEM) {
        if (warp == 0) {                                         // all 512 columns: 2 first-stage slots x 256 k-mers
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_tmem)), "n"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
        tq = s_tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    }
    int cur_blk = -1, blk = 0;
    const long long item_anchors = (long long)kWarps * 32 * iters;            // x replication
    long long blk_first = 0, blk_items = (n_anchor / item_anchors) >> blocks[0].rep_log2;
    int s1 = 2, nl = 0, nslot = 2, rep_log2 = 0;
    for (;;) {
        if (tid == 0) s_item = (long long)atomicAdd(work_counter, 1ull);
        __syncthreads();
        const long long item = s_item;
        if (item >= total_items) break;
        while (item >= blk_first + blk_items) {                 // items are handed out in increasing order
            blk_first += blk_items;
            ++blk;
            blk_items = (n_anchor / item_anchors) >> blocks[blk].rep_log2;
        }
        const long long chunk = item - blk_first;
        if (blk != cur_blk) {
            const DevBlock db = blocks[blk];
            s1 = db.s1; nl = db.nl; nslot = db.nslot; rep_log2 = db.rep_log2;
            const uint4* src = reinterpret_cast<const uint4*>(tables + db.table_off);
            uint4* dst = reinterpret_cast<uint4*>(smem);
            const int n16 = nslot * (kSlotBytes / 16);
            for (int i = tid; i < n16; i += kThreads) dst[i] = __ldg(src + i);
            if (USE_TMEM && rep_log2 == 0 && warp < 4) {
                // first-stage slots nl, nl+1 -> TMEM columns [0,256) and [256,512) of this warp's lane quarter
                const uint32_t* t0 = tables + db.table_off + (size_t)nl * 256 * 32 + lane;
                for (int c = 0; c < 512; ++c) tmem_st32(tq + (uint32_t)c, __ldg(t0 + (size_t)c * 32));
                tmem_wait_st();
            }
            if (USE_TMEM) asm volatile("tcgen05.fence::before_thread_sync;");
            cur_blk = blk;
        }
        __syncthreads();   // tables ready; also protects s_item against the next iteration's write
        if (USE_TMEM) asm volatile("tcgen05.fence::after_thread_sync;");
        const int group_lanes = 32 >> rep_log2;                  // lanes that carry distinct motifs
        const int replica = lane >> (5 - rep_log2);
        const long long U0 = anchor0 + ((chunk * item_anchors + (long long)warp * (32 * iters)) << rep_log2) + (long long)replica * (32 * iters);
        const uint32_t* w = seq2 + (U0 >> 4);
        uint32_t x0 = __ldg(w - 1), x1 = __ldg(w), x2 = __ldg(w + 1), x3 = __ldg(w + 2), x4 = __ldg(w + 3);
        const unsigned long long combo_base = (unsigned long long)blk * kCombosPerBlock + (unsigned)(lane & (group_lanes - 1)) * 4u;
        auto emit = [&](int q, long long u) {
            const unsigned long long slot = atomicAdd(&counters[0], 1ull);
            if (slot < cand_cap) cand[slot] = ((combo_base + (unsigned)q) << 40) | (unsigned long long)u;
        };
#define PWM_RUN(S1)                                                                                   \
        _Pragma("unroll 1") for (int it = 0; it < iters; ++it) {                                      \
            const uint32_t n3 = __ldg(w + 4 + 2 * it), n4 = __ldg(w + 5 + 2 * it);                    \
            process32<S1>(x0, x1, x2, x3, x4, U0 + 32 * it, lane, smem_addr, nl, nslot, emit);        \
            x0 = x2; x1 = x3; x2 = x4; x3 = n3; x4 = n4;                                              \
        }
#define PWM_RUN_T(S1)                                                                                 \
        _Pragma("unroll 1") for (int it = 0; it < iters; ++it) {                                      \
            const uint32_t n3 = __ldg(w + 4 + 2 * it), n4 = __ldg(w + 5 + 2 * it);                    \
            process32_tmem<S1>(x0, x1, x2, x3, x4, U0 + 32 * it, lane, smem_addr, tq, nl, nslot, emit); \
            x0 = x2; x1 = x3; x2 = x4; x3 = n3; x4 = n4;                                              \
        }
        if (USE_TMEM && rep_log2 == 0) {
            if (s1 == 2) { PWM_RUN_T(2) } else if (s1 == 3) { PWM_RUN_T(3) } else { PWM_RUN_T(4) }
        } else {
            if (s1 == 2) { PWM_RUN(2) } else if (s1 == 3) { PWM_RUN(3) } else { PWM_RUN(4) }
        }
#undef PWM_RUN
#undef PWM_RUN_T
        __syncthreads();   // every warp is done with the tables before a possible reload
    }
    if (USE_TMEM) {
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(512));
    }
}

__device__ __forceinline__ int find_segment(const int64_t* __restrict__ seg_off, int nseg, int64_t i) {
    int lo = 0, hi = nseg;           // last seg with seg_off[seg] <= i
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (seg_off[mid] <= i) lo = mid; else hi = mid; }
    return lo;
}

// Sort key of a hit (KeyLayout, internal.h), packed as tightly as this scan allows so that the ordering pass
// (order.cu) needs few buckets: [segment | motif : mb bits | strand : 1 bit | position in segment : pb bits].
__device__ __forceinline__ unsigned long long hit_key(KeyLayout kl, int seg, int motif, int strand, int64_t pos_in_seg) {
    return ((((((unsigned long long)seg << kl.mb) | (unsigned long long)motif) << 1) | (unsigned long long)strand) << kl.pb) |
           (unsigned long long)pos_in_seg;
}

// Stores one hit and counts it in its ordering bucket (order.cu).
__device__ __forceinline__ void sink_store(const HitSink& hs, unsigned long long slot, unsigned long long key, double sc) {
    if (slot < hs.cap) {
        hs.keys[slot] = key; hs.sc[slot] = sc;
        atomicAdd(&hs.bucket_cnt[key >> hs.shift], 1u);
    }
}

// Appends the hits of the calling lanes with one atomicAdd per warp (all lanes of the warp that are in the
// surrounding loop iteration must call it; `hit` selects the ones that have something to append).
__device__ __forceinline__ void append_hits(bool hit, unsigned long long key, double sc, const HitSink& hs) {
    const unsigned act = __activemask();
    const unsigned bal = __ballot_sync(act, hit);
    if (bal == 0u) return;
    const int lane = threadIdx.x & 31, leader = __ffs((int)bal) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(&hs.counters[1], (unsigned long long)__popc(bal));
    base = __shfl_sync(act, base, leader);
    if (hit) sink_store(hs, base + (unsigned)__popc(bal & ((1u << lane) - 1u)), key, sc);
}

// replay_skip (scan_core.cuh: the literal lookAheadSearch' loop, Motif/Search.hs:176-189) with the window's 2-bit codes and
// invalid-base flags fetched up front (three + two 32-bit words for up to 32 columns) instead of one dependent load pair per
// column; what remains per column are the two independent table reads.  Same additions in the same order.
__device__ __forceinline__ bool replay_window(const uint32_t* __restrict__ seq2, const uint32_t* __restrict__ mask, int64_t u0, int n,
                                              const double* __restrict__ tab16, const double* __restrict__ tab4, const double* __restrict__ th, double* score) {
    if (n > 32) return replay_skip(seq2, mask, u0, n, tab16, th, score);
    const int64_t w = u0 >> 4;
    const int s2 = 2 * (int)(u0 & 15);
    const uint32_t a0 = __ldg(seq2 + w), a1 = __ldg(seq2 + w + 1), a2 = __ldg(seq2 + w + 2);
    const unsigned long long x = (unsigned long long)__funnelshift_r(a0, a1, s2) | ((unsigned long long)__funnelshift_r(a1, a2, s2) << 32);
    const int64_t mw = u0 >> 5;
    const uint32_t inv = __funnelshift_r(__ldg(mask + mw), __ldg(mask + mw + 1), (int)(u0 & 31));
    double acc = 0.0;
    for (int k = 0; k < n; ++k) {
        const double sc = ((inv >> k) & 1u) ? PWM_NEG_INF : __ldg(tab4 + 4 * k + (int)((x >> (2 * k)) & 3ull));      // == tab16[16k + base]
        acc = __dadd_rn(acc, sc);
        if (acc < __ldg(th + k)) return false;
    }
    *score = acc;
    return true;
}

__global__ void rescore_kernel(const unsigned long long* __restrict__ cand, unsigned long long cand_cap,
                               const DevCombo* __restrict__ combos, const uint32_t* __restrict__ seq2,
                               const uint32_t* __restrict__ mask, const int64_t* __restrict__ seg_off, int nseg, int64_t len,
                               const double* __restrict__ tab16, const double* __restrict__ tab4, const double* __restrict__ th, HitSink hs, KeyLayout kl) {
    unsigned long long ncand = hs.counters[0];
    if (ncand > cand_cap) ncand = cand_cap;
    for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < ncand;
         k += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long cd = cand[k];
        bool hit = false;
        unsigned long long key = 0;
        double sc = 0.0;
        if (cd != ~0ull) {                               // ~0: unused slot of a warp's chunk (tensor-core prefilter)
            const DevCombo cb = combos[cd >> 40];
            const int64_t u = (int64_t)(cd & ((1ull << 40) - 1));
            const int64_t i = u - kPadFront - cb.c;
            if (cb.motif >= 0 && i >= 0 && i + cb.n <= len) {
                const int seg = nseg == 1 ? 0 : find_segment(seg_off, nseg, i);
                if (i + cb.n <= seg_off[seg + 1] &&
                    replay_window(seq2, mask, i + kPadFront, cb.n, tab16 + cb.tab_off, tab4 + 4 * cb.th_off, th + cb.th_off, &sc)) {
                    hit = true;
                    key = hit_key(kl, seg, cb.motif, cb.strand, i - seg_off[seg]);
                }
            }
        }
        append_hits(hit, key, sc, hs);
    }
}

// Tensor-core prefilter records (anchor, block, 32-combo chunk): one warp per record, lane = combo of the chunk,
// exact fp64 replay as in rescore_kernel (the first stage only said "some combo of this chunk may still hit").
__global__ void rescore_chunks_kernel(const unsigned long long* __restrict__ rec, unsigned long long cap,
                                      const DevMmaBlock* __restrict__ blocks, const int8_t* __restrict__ btiles,
                                      const DevCombo* __restrict__ combos, const uint32_t* __restrict__ seq2,
                                      const uint32_t* __restrict__ mask, const int64_t* __restrict__ seg_off, int nseg, int64_t len,
                                      const double* __restrict__ tab16, const double* __restrict__ th, HitSink hs, KeyLayout kl) {
    unsigned long long nrec = hs.counters[0];
    if (nrec > cap) nrec = cap;
    const int lane = threadIdx.x & 31;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    for (unsigned long long k = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < nrec; k += nwarps) {
        const unsigned long long r = rec[k];
        if (r == ~0ull) continue;
        const int64_t u = (int64_t)(r & ((1ull << 40) - 1));
        const unsigned blk = (unsigned)((r >> 40) & 0xFFFF), chunk = (unsigned)((r >> 56) & 7);
        {   // which of the chunk's 32 combos did the tensor core flag?  Redo this lane's int8 dot product.
            const DevMmaBlock db = blocks[blk];
            const int n = (int)chunk * 32 + lane;
            const int8_t* bt = btiles + db.b_off + (size_t)(n / 8) * 256 + (size_t)(n % 8) * 16;
            int acc = 0;
            for (int k = 0; k < 8 * db.steps; ++k) {
                const int x = 4 * (k & 7) + (int)packed_base(seq2, u + k);
                acc += bt[(size_t)(k >> 3) * db.ncombo * 32 + (size_t)(x >> 4) * 128 + (x & 15)];
            }
            if (acc >= 0) continue;
        }
        const DevCombo cb = combos[(size_t)blk * kMmaCombos + chunk * 32 + lane];
        const int64_t i = u - kPadFront - cb.c;
        if (cb.motif < 0 || i < 0 || i + cb.n > len) continue;
        const int seg = nseg == 1 ? 0 : find_segment(seg_off, nseg, i);
        if (i + cb.n > seg_off[seg + 1]) continue;
        double sc;
        if (replay_skip(seq2, mask, i + kPadFront, cb.n, tab16 + cb.tab_off, th + cb.th_off, &sc)) {
            sink_store(hs, atomicAdd(&hs.counters[1], 1ull), hit_key(kl, seg, cb.motif, cb.strand, i - seg_off[seg]), sc);
        }
    }
}

// Exact kernel: every window of every listed combo replayed in fp64 (IUPAC-aware when skip == 0).
// Used for combos the prefilter cannot take (skip == 0, non-finite cutoffs/tables).
__global__ void exact_kernel(const DevCombo* __restrict__ combos, int ncombo, int skip, const uint32_t* __restrict__ seq2,
                             const uint32_t* __restrict__ mask, const uint8_t* __restrict__ ascii,
                             const int64_t* __restrict__ seg_off, int nseg, int64_t len, const double* __restrict__ tab16,
                             const double* __restrict__ th, HitSink hs, KeyLayout kl) {
    unsigned long long* counters = hs.counters;
  for (int ci = blockIdx.y; ci < ncombo; ci += gridDim.y) {
    const DevCombo cb = combos[ci];
    const double* tb = tab16 + cb.tab_off;
    const double* tk = th + cb.th_off;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i + cb.n <= len; i += (int64_t)gridDim.x * blockDim.x) {
        const int seg = nseg == 1 ? 0 : find_segment(seg_off, nseg, i);
        if (i + cb.n > seg_off[seg + 1]) continue;
        double sc = 0.0;
        bool hit;
        if (skip) {
            hit = replay_skip(seq2, mask, i + kPadFront, cb.n, tb, tk, &sc);
        } else {
            hit = true;
            for (int k = 0; k < cb.n; ++k) {          // lookAheadSearch, Search.hs:133-157
                const unsigned code = c_iupac_lut[ascii[i + k]];
                if (code == 15u) { atomicOr(&counters[2], 1ull); hit = false; break; }
                sc = __dadd_rn(sc, tb[16 * k + code]);
                if (sc < tk[k]) { hit = false; break; }
            }
        }
        if (hit) sink_store(hs, atomicAdd(&counters[1], 1ull), hit_key(kl, seg, cb.motif, cb.strand, i - seg_off[seg]), sc);
    }
  }
}

// Start of a scan on a lane: counters [0, 4) and [6, nwords) and the ordering's bucket counters to zero ([4, 6) hold the fromBS
// verdict of the upload that may just have been packed on this lane).
__global__ void scan_reset_kernel(unsigned long long* __restrict__ counters, int nwords, unsigned int* __restrict__ cnt, size_t nb, int new_upload) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    for (size_t i = t; i < (size_t)nwords; i += step) {
        if (i < 4 || i >= 6) counters[i] = 0ull;
        else if (new_upload) counters[i] = ~0ull;            // "no byte outside the alphabet yet": the pack kernels of this job take the minimum
    }
    for (size_t i = t; i < nb; i += step) cnt[i] = 0u;
}

// IUPAC-aware search (skip == 0) next to the first stage: the k-mer masks and the replay of their candidates cover the
// windows made of A/C/G/T only (lookAheadSearch and lookAheadSearch' agree on those); this kernel takes the rest — every
// window that holds at least one other letter (a set bit of the mask plane among its n positions) — through the literal
// lookAheadSearch loop on the bytes (Search.hs:133-157: IUPAC letters score their expected log-odds, an invalid letter THAT
// THE LOOK-AHEAD VISITS is the reference's `error`).  One thread per window start; the mask bits of the next 64 positions
// decide at once whether any motif has to be looked at.
__global__ void ambig_kernel(const DevCombo* __restrict__ combos, int ncombo, int max_n, const uint32_t* __restrict__ mask,
                             const uint8_t* __restrict__ ascii, const int64_t* __restrict__ seg_off, int nseg, int64_t len,
                             const double* __restrict__ tab16, const double* __restrict__ th, HitSink hs, KeyLayout kl) {
    unsigned long long* counters = hs.counters;
    const unsigned long long all_n = max_n >= 64 ? ~0ull : ((1ull << max_n) - 1ull);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t u = i + kPadFront, mw = u >> 5;
        const int sh = (int)(u & 31);
        const uint32_t m0 = __ldg(mask + mw), m1 = __ldg(mask + mw + 1), m2 = __ldg(mask + mw + 2);
        const unsigned long long inv = (unsigned long long)__funnelshift_r(m0, m1, sh) | ((unsigned long long)__funnelshift_r(m1, m2, sh) << 32);
        if ((inv & all_n) == 0ull) continue;                                   // 64 positions of A/C/G/T ahead: nothing for this pass
        const int seg = nseg == 1 ? 0 : find_segment(seg_off, nseg, i);
        const int64_t seg_end = seg_off[seg + 1];
        for (int ci = 0; ci < ncombo; ++ci) {
            const DevCombo cb = combos[ci];
            if (i + cb.n > seg_end) continue;
            const unsigned long long mine = cb.n >= 64 ? ~0ull : ((1ull << cb.n) - 1ull);
            if ((inv & mine) == 0ull) continue;                                // this motif's window is all A/C/G/T: the first stage has it
            const double* tb = tab16 + cb.tab_off;
            const double* tk = th + cb.th_off;
            double sc = 0.0;
            bool hit = true;
            for (int k = 0; k < cb.n; ++k) {
                const unsigned code = c_iupac_lut[ascii[i + k]];
                if (code == 15u) { atomicOr(&counters[2], 1ull); hit = false; break; }
                sc = __dadd_rn(sc, tb[16 * k + code]);
                if (sc < tk[k]) { hit = false; break; }
            }
            if (hit) sink_store(hs, atomicAdd(&counters[1], 1ull), hit_key(kl, seg, cb.motif, cb.strand, i - seg_off[seg]), sc);
        }
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// scan driver
// ------------------------------------------------------------------------------------------------
// A scan runs in three phases on a lane (its own streams, counters and scratch), so that several scans can be in flight
// on one device (pwmscan_scan_many: the (motif block x chromosome chunk) items of a whole-genome scan):
//   A  enqueue: [pipelined upload + 2-bit packing,] first stage (prefilter), exact fp64 replay of the candidates,
//      the counters to the host
//   B  once the hit count is known: ordering (order.cu) into the compact columns, one device->host copy
//   C  once that copy has landed: the hit list object
// The first-stage kernels of different lanes are chained by an event: they fill every SM and cannot overlap anyway, and
// the chain makes the CUDA-event time of a launch the kernel's own time.
namespace {

// Host->device job of a pipelined scan (pwmscan_scan_host): the ASCII bytes go up in chunks on the
// copy stream; each chunk is packed and scanned on the compute stream as soon as it has landed, so
// the PCIe transfer hides behind the prefilter kernel of the previous chunks.
struct UploadJob {
    const uint8_t* ascii = nullptr;
    const uint32_t *h_seq2 = nullptr, *h_mask = nullptr;     // packed input (hostpack.cu) instead of ascii
    uint8_t* d_raw = nullptr;
    pwmscan_seq* seq = nullptr;      // mutable view of the sequence being filled
};

struct ScanJob {
    ScanLane* lane = nullptr;
    const pwmscan_seq* seq = nullptr;
    const pwmscan_plan* pl = nullptr;
    bool host = false;               // sequence uploaded by this job (scan_host): released in phase C
    bool whole_upload = false;       // one upload + one first-stage launch for the whole sequence (items of a scan_many: the
                                     // other lanes hide the transfer; fewer, larger launches keep the mask tables in L2 longer)
    UploadJob up;
    KeyLayout kl{1, 1};
    int key_bits = 0, shift = 0, attempt = 0;
    unsigned long long cand_cap = 0, hit_cap = 0, nhits = 0, ncand = 0;
    int prefilter_launches = 0, kernels = 0;
    size_t n8 = 0, out_bytes = 0;
    size_t col_cap = 0;              // entries per device output column (= hit_cap rounded up)
    unsigned char* d_out = nullptr;
    pwmscan_hits* hits = nullptr;
    bool active = false;
};

inline int bits_for(uint64_t max_value) { int b = 1; while (b < 63 && (max_value >> b) != 0) ++b; return b; }

enum { EV_START = 0, EV_PRE = 1, EV_RES = 2, EV_ORD = 3, EV_GATE = 4, EV_A = 5, EV_B = 6, EV_H2D0 = 7, EV_H2D1 = 8, EV_PSTART = 9 };
enum { SL_RAW = 0, SL_CAND = 1, SL_KEYS = 2, SL_SC = 3, SL_WORK = 4, SL_TKEYS = 5, SL_TSC = 6, SL_SKEYS = 7, SL_OUT = 8 };

void job_release(ScanJob& j) {
    if (j.hits) { pwmscan_hits_free(j.hits); j.hits = nullptr; }
    if (j.host && j.up.seq) { pwmscan_seq_free(j.up.seq); j.up.seq = nullptr; }
    j.active = false;
}

// Device output columns of a job (lane scratch SL_OUT), each with room for col_cap entries.
struct OutCols { double* score; uint32_t* pos32; long long* run_off; int32_t* seg32; uint16_t* ms16; };
OutCols job_out_cols(const ScanJob& j) {
    OutCols oc;
    unsigned char* d = j.d_out;
    oc.score = (double*)d;
    oc.pos32 = (uint32_t*)(d + j.col_cap * 8);
    oc.run_off = (long long*)(d + j.col_cap * 12);
    oc.seg32 = (int32_t*)(d + j.col_cap * 12);
    oc.ms16 = (uint16_t*)(d + j.col_cap * 16);
    return oc;
}

// (Re)allocates what the ordering pass needs for the job's current shift and hit capacity.
int job_order_buffers(pwmscan_ctx* ctx, ScanJob& j) {
    ScanLane* lane = j.lane;
    const size_t nb = order_bucket_count(j.key_bits, j.shift);
    if (nb > ((size_t)1 << 28)) return set_err(ctx, PWMSCAN_ERANGE, "pwmscan_scan: hit density too uneven to order");
    j.col_cap = ((size_t)j.hit_cap + 7) & ~(size_t)7;
    const int M = j.pl->motifs->M;
    const size_t out_bytes = j.seq->nseg == 1 ? j.col_cap * 12 + (size_t)(2 * M + 1) * 8 : j.col_cap * 18;
    if (!lane_scratch_get(ctx, lane, SL_WORK, order_work_bytes(nb, (size_t)j.hit_cap)) || !lane_scratch_get(ctx, lane, SL_TKEYS, j.col_cap * 8) ||
        !lane_scratch_get(ctx, lane, SL_TSC, j.col_cap * 8) || !lane_scratch_get(ctx, lane, SL_SKEYS, j.col_cap * 8) ||
        !(j.d_out = (unsigned char*)lane_scratch_get(ctx, lane, SL_OUT, out_bytes)))
        return PWMSCAN_ECUDA;
    return PWMSCAN_OK;
}

// scan -> scatter -> per-bucket sort -> run table, behind whatever counted the hits into ow.cnt (stream-ordered).
int job_order_finish(pwmscan_ctx* ctx, ScanJob& j, const OrderWork& ow) {
    ScanLane* lane = j.lane;
    const OutCols oc = job_out_cols(j);
    const bool one = j.seq->nseg == 1;
    const unsigned long long n_est = j.nhits ? j.nhits : (unsigned long long)(j.pl->cand_rate * (double)j.seq->len) + 1024;
    return order_finish(ctx, lane->stream, (const unsigned long long*)lane->scratch[SL_KEYS], (const double*)lane->scratch[SL_SC], n_est, j.kl, j.shift,
                        j.seq->nseg, j.pl->motifs->M, ow, (unsigned long long*)lane->scratch[SL_TKEYS], (double*)lane->scratch[SL_TSC],
                        (unsigned long long*)lane->scratch[SL_SKEYS], oc.score, oc.pos32, one ? nullptr : oc.seg32, one ? nullptr : oc.ms16,
                        one ? oc.run_off : nullptr, lane->d_counters + 6, &j.kernels);
}

// Phase A.  attempt > 0 (a capacity overflow): the sequence is packed already, plain re-run.
int scan_phase_a(pwmscan_ctx* ctx, ScanJob& j) {
    ScanLane* lane = j.lane;
    const pwmscan_seq* seq = j.seq;
    const pwmscan_plan* pl = j.pl;
    const pwmscan_motifs* mo = pl->motifs;
    cudaStream_t st = lane->stream;
    // pst: the stream of the counters' reset, the packing and the first-stage kernels; st: everything behind them.
    // Measured (config 4, one B200): with the first stage on its own low-priority stream and the rest on the high-priority one
    // a genome takes 94 ms, with everything on the lane's one stream 87 ms (the replay grabs SMs at every first-stage boundary
    // and delays the next persistent launch).  The split is what the non-persistent first stage (PWMSCAN_KM_KERNEL=3) needs.
    static const bool split = std::getenv("PWMSCAN_SPLIT_STREAMS") != nullptr || !km_prefilter_is_persistent();
    cudaStream_t pst = (pl->mode == PWM_MODE_KMER && split) ? lane->pstream : lane->stream;
    int rc;
    const int64_t len = seq->len;
    const int64_t n_anchor = round_up(len + kPadFront + 64, kAnchorRound);
    if (j.attempt == 0) {
        int64_t max_seg = 0;
        for (int32_t g = 0; g < seq->nseg; ++g) max_seg = std::max(max_seg, seq->seg_off[g + 1] - seq->seg_off[g]);
        j.kl.pb = bits_for((uint64_t)max_seg); j.kl.mb = bits_for((uint64_t)std::max(1, mo->M - 1));
        j.key_bits = j.kl.pb + 1 + j.kl.mb + (seq->nseg > 1 ? bits_for((uint64_t)(seq->nseg - 1)) : 0);
        if (j.key_bits > 64) return set_err(ctx, PWMSCAN_ERANGE, "pwmscan_scan: segments x motifs x segment length exceed the 64-bit hit key");
        if (j.kl.pb > 32) return set_err(ctx, PWMSCAN_ERANGE, "pwmscan_scan: segment longer than 2^32");
    }
    const bool use_mma = pl->mode == PWM_MODE_MMA, use_km = pl->mode == PWM_MODE_KMER;
    const int n_pre_blocks = use_mma ? pl->mma_nblocks : use_km ? pl->km_nblocks : pl->nblocks;
    const std::vector<DevCombo>& generic = use_mma ? pl->mma_generic : use_km ? pl->km_generic : pl->generic;
    const DevCombo* d_generic = use_mma ? pl->d_mma_generic : use_km ? pl->d_km_generic : pl->d_generic;
    const DevCombo* d_combos = use_mma ? pl->d_mma_combos : use_km ? pl->d_km_combos : pl->d_combos;
    if (j.attempt == 0) {
        // capacities: expected + slack, grown and re-run on overflow
        j.cand_cap = (unsigned long long)(pl->cand_rate * (double)len * 1.5) + (1ull << 20);
        if (use_mma) j.cand_cap += (unsigned long long)ctx->sm_count * 28 * 256 * 2;        // every epilogue warp reserves whole 256-record chunks
        j.hit_cap = (unsigned long long)(pl->cand_rate * (double)len * 1.5) + (1ull << 20);
        if (!generic.empty() && j.hit_cap < (1ull << 22)) j.hit_cap = 1ull << 22;
    }
    unsigned long long* d_cand = (unsigned long long*)lane_scratch_get(ctx, lane, SL_CAND, j.cand_cap * 8);
    unsigned long long* d_keys = (unsigned long long*)lane_scratch_get(ctx, lane, SL_KEYS, j.hit_cap * 8);
    double* d_sc = (double*)lane_scratch_get(ctx, lane, SL_SC, j.hit_cap * 8);
    if (!d_cand || !d_keys || !d_sc) return PWMSCAN_ECUDA;
    // the ordering pass (order.cu) runs right behind the kernels that find the hits: its bucket layout comes from the
    // EXPECTED hit count (a bad guess only costs a retry in phase B), its buffers are sized by the hit capacity
    j.shift = order_pick_shift(std::max<unsigned long long>((unsigned long long)(pl->cand_rate * (double)len) + 1024, j.nhits), j.key_bits);
    if ((rc = job_order_buffers(ctx, j))) return rc;
    const OrderWork ow = order_work_view((unsigned char*)lane->scratch[SL_WORK], order_bucket_count(j.key_bits, j.shift), (size_t)j.hit_cap);
    HitSink hs{d_keys, d_sc, j.hit_cap, lane->d_counters, ow.cnt, j.shift};
    if (pl->nblocks > 0 && !ctx->smem_attr_set) {
        PWM_CUDA(cudaFuncSetAttribute(prefilter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSlots * kSlotBytes));
        PWM_CUDA(cudaFuncSetAttribute(prefilter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSlots * kSlotBytes));
        ctx->smem_attr_set = true;
    }
    auto items_of = [&](long long anchors, int iters) {
        long long t = 0;
        for (int b = 0; b < pl->nblocks; ++b) t += (anchors / ((long long)kWarps * 32 * iters)) / pl->blocks[b].rep;
        return t;
    };
    const size_t smem_bytes = (size_t)pl->max_nslot * kSlotBytes;
    bool chained = false;
    const bool persistent = !use_km || km_prefilter_is_persistent();
    auto launch_prefilter = [&](long long a0, long long a1, int slot) -> int {
        if (n_pre_blocks == 0 || a1 <= a0) return PWMSCAN_OK;
        if (!chained) {               // persistent first-stage kernels of different lanes run one after the other
            if (persistent && ctx->prefilter_chain_valid) PWM_CUDA(cudaStreamWaitEvent(pst, ctx->prefilter_chain, 0));
            if (!j.host || j.attempt > 0) PWM_CUDA(cudaEventRecord(lane->ev[EV_START], pst));
            PWM_CUDA(cudaEventRecord(lane->ev[EV_PSTART], pst));
            chained = true;
        }
        if (use_km) {
            ++j.prefilter_launches;
            return launch_km_prefilter(ctx, lane, pst, seq->d_seq2, a0, a1, len, pl->km_nblocks, pl->km_smem_bytes, pl->d_km_blocks, pl->d_km_tables, pl->d_km_q16,
                                       pl->d_km_combos, d_cand, j.cand_cap, lane->d_counters + 8 + slot);
        }
        if (use_mma) {
            ++j.prefilter_launches;
            return launch_mma_prefilter(ctx, lane, seq->d_seq2, a0, a1, pl->mma_nblocks, pl->d_mma_blocks, pl->d_btiles, d_cand, j.cand_cap);
        }
        int iters = kIters;                                   // short inputs: smaller items so that every SM gets work
        while (iters > 1 && items_of(a1 - a0, iters) < 4LL * ctx->sm_count) iters >>= 1;
        const long long items = items_of(a1 - a0, iters);
        if (items == 0) return PWMSCAN_OK;
        const int grid = (int)std::min<long long>(ctx->sm_count, items);
        // Experimental variant (off by default): first-stage slots 0/1 read through TMEM (tcgen05.ld) next to the
        // shared-memory port.  Correct (same parity tests) but measured slower on C2 (9.7 vs 7.6 ms): the extra
        // R2UR / address instructions cost more issue slots than the second port saves.
        static const bool use_tmem = std::getenv("PWMSCAN_TMEM") != nullptr;
        bool any_full = false;                                        // TMEM serves blocks whose 32 lanes carry distinct motifs
        for (const auto& b : pl->blocks) any_full |= (b.rep == 1);
        if (any_full && use_tmem)
            prefilter_kernel<true><<<grid, kThreads, smem_bytes, pst>>>(seq->d_seq2, a0, a1 - a0, items, iters, pl->nblocks, pl->d_blocks, pl->d_tables,
                                                                       d_cand, j.cand_cap, lane->d_counters, lane->d_counters + 8 + slot);
        else
            prefilter_kernel<false><<<grid, kThreads, smem_bytes, pst>>>(seq->d_seq2, a0, a1 - a0, items, iters, pl->nblocks, pl->d_blocks, pl->d_tables,
                                                                        d_cand, j.cand_cap, lane->d_counters, lane->d_counters + 8 + slot);
        PWM_CUDA(cudaGetLastError());
        ++j.prefilter_launches;
        return PWMSCAN_OK;
    };
    if (pst != st) {                  // whatever the lane's main stream still has queued (segment table, counters of the upload) comes first
        PWM_CUDA(cudaEventRecord(lane->ev[EV_GATE], st));
        PWM_CUDA(cudaStreamWaitEvent(pst, lane->ev[EV_GATE], 0));
    }
    // (a kernel, not three cudaMemsetAsync: memsets are copy-engine operations and wait behind the hit-list copies of the items
    //  before this one — at eight GPUs, where those copies are slow, that delayed every item's first stage and ordering)
    scan_reset_kernel<<<std::max(1u, std::min(64u, (unsigned)((ow.nb + 1023) / 1024))), 256, 0, pst>>>(lane->d_counters, kCounterWords, ow.cnt, ow.nb,
                                                                                                       (j.host && j.attempt == 0) ? 1 : 0);
    PWM_CUDA(cudaGetLastError());
    if (use_mma) PWM_CUDA(cudaMemsetAsync(d_cand, 0xFF, j.cand_cap * 8, pst));     // sentinel: warps reserve whole chunks
    if (j.host && j.attempt == 0 && j.up.h_seq2) {
        // packed planes from the host: two copies (0.375 bytes per base), padding on the device, one first-stage launch
        PWM_CUDA(cudaEventRecord(lane->ev[EV_START], pst));
        if (lane->chunk_ev.empty()) {
            cudaEvent_t ev;
            PWM_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            lane->chunk_ev.push_back(ev);
        }
        PWM_CUDA(cudaEventRecord(lane->ev[EV_GATE], pst));              // copies must not overtake the previous use of the planes
        PWM_CUDA(cudaStreamWaitEvent(lane->copy_stream, lane->ev[EV_GATE], 0));
        PWM_CUDA(cudaEventRecord(lane->ev[EV_H2D0], lane->copy_stream));
        const int64_t units = (len + 31) / 32;
        if (units > 0) {
            PWM_CUDA(cudaMemcpyAsync(j.up.seq->d_seq2 + 2, j.up.h_seq2, (size_t)units * 8, cudaMemcpyHostToDevice, lane->copy_stream));
            PWM_CUDA(cudaMemcpyAsync(j.up.seq->d_mask + 1, j.up.h_mask, (size_t)units * 4, cudaMemcpyHostToDevice, lane->copy_stream));
        }
        PWM_CUDA(cudaEventRecord(lane->chunk_ev[0], lane->copy_stream));
        PWM_CUDA(cudaEventRecord(lane->ev[EV_H2D1], lane->copy_stream));
        PWM_CUDA(cudaStreamWaitEvent(pst, lane->chunk_ev[0], 0));
        if ((rc = seq_launch_packed_finish(ctx, lane, pst, j.up.seq))) return rc;
        j.kernels += 3;
        if ((rc = launch_prefilter(0, n_anchor, 0))) return rc;
        if (!chained) PWM_CUDA(cudaEventRecord(lane->ev[EV_START], pst));
    } else if (j.host && j.attempt == 0) {
        PWM_CUDA(cudaEventRecord(lane->ev[EV_START], pst));
        // chunk boundaries are multiples of the anchor rounding unit, so every launch covers whole work items.
        // 16 Mi bases per chunk (smaller chunks cost more in per-launch overhead than they hide), except that
        // the end is split off as a short last chunk so that little work is left once the last byte has landed.
        static const int chunk_units = std::getenv("PWMSCAN_CHUNK_UNITS") ? std::max(8, std::atoi(std::getenv("PWMSCAN_CHUNK_UNITS"))) : 32;   // experiments
        int64_t cb = chunk_units * kAnchorRound;
        while ((len + cb - 1) / cb > kMaxChunks - 1) cb *= 2;
        const int64_t tail = 4 * kAnchorRound;
        std::vector<int64_t> bounds(1, 0);
        if (!j.whole_upload) {
            while (len - bounds.back() > cb + tail) bounds.push_back(bounds.back() + cb);
            if (len - bounds.back() > 2 * tail) bounds.push_back((len - tail) / kAnchorRound * kAnchorRound);
        }
        bounds.push_back(len);
        const int nch = (int)bounds.size() - 1;
        while ((int)lane->chunk_ev.size() < nch) {
            cudaEvent_t ev;
            PWM_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            lane->chunk_ev.push_back(ev);
        }
        PWM_CUDA(cudaEventRecord(lane->ev[EV_GATE], pst));              // copies must not overtake the previous use of d_raw
        PWM_CUDA(cudaStreamWaitEvent(lane->copy_stream, lane->ev[EV_GATE], 0));
        PWM_CUDA(cudaEventRecord(lane->ev[EV_H2D0], lane->copy_stream));
        // (one copy stream: alternating two streams was measured slower, 6.7 vs 5.3 ms on C2)
        for (int l = 0; l < nch; ++l) {
            const int64_t b0 = bounds[l], b1 = bounds[l + 1];
            if (b1 > b0) PWM_CUDA(cudaMemcpyAsync(j.up.d_raw + b0, j.up.ascii + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, lane->copy_stream));
            PWM_CUDA(cudaEventRecord(lane->chunk_ev[l], lane->copy_stream));
        }
        PWM_CUDA(cudaEventRecord(lane->ev[EV_H2D1], lane->copy_stream));
        long long scanned = 0;
        for (int l = 0; l < nch; ++l) {
            const bool last = (l == nch - 1);
            const int64_t b0 = bounds[l], b1 = bounds[l + 1];
            PWM_CUDA(cudaStreamWaitEvent(pst, lane->chunk_ev[l], 0));
            // pack threads own 32 stored bases: positions [32t-32, 32t)
            const int64_t t0 = l == 0 ? 0 : b0 / 32 + 1, t1 = last ? j.up.seq->nwords_mask : b1 / 32 + 1;
            if ((rc = seq_launch_pack(ctx, lane, pst, j.up.seq, j.up.d_raw, t0, t1))) return rc;
            ++j.kernels;
            // anchors whose windows (and the prefetch of the next words) lie inside what is packed so far
            const long long a1 = last ? n_anchor : std::max<long long>(scanned, b1 - kAnchorRound);
            if (a1 > scanned) { if ((rc = launch_prefilter(scanned, a1, l))) return rc; scanned = a1; }
        }
    } else {
        if ((rc = launch_prefilter(0, n_anchor, 0))) return rc;
        if (!chained) PWM_CUDA(cudaEventRecord(lane->ev[EV_START], pst));
    }
    PWM_CUDA(cudaEventRecord(lane->ev[EV_PRE], pst));
    // What the next lane's first stage waits for.  A persistent first stage holds every SM until it ends, so kernels of OTHER
    // lanes that become ready while it runs (the ordering of the previous item) only get SMs at its end — where the next
    // first stage, if it waits for nothing but its predecessor, takes them all again: with several items enqueued ahead the
    // orderings (and the hit copies behind them) of all of them were starved until the last enqueued first stage had ended
    // (8 GPUs, 7 items of 62 Mbp per rank: every copy started after 6.7 ms; 19.1 ms per genome).  The chain event is therefore
    // recorded behind the item's ORDERING: first stage, replay and ordering of an item run back to back, its hit copy
    // overlaps the next item's first stage.  (PWMSCAN_CHAIN=pre: the old behaviour, for experiments.)
    static const bool chain_pre = std::getenv("PWMSCAN_CHAIN") != nullptr && std::string(std::getenv("PWMSCAN_CHAIN")) == "pre";
    const bool chain_here = chained && persistent;
    if (chain_here && (chain_pre || pst != st)) { PWM_CUDA(cudaEventRecord(ctx->prefilter_chain, pst)); ctx->prefilter_chain_valid = true; }
    if (pst != st) PWM_CUDA(cudaStreamWaitEvent(st, lane->ev[EV_PRE], 0));
    j.kernels += j.prefilter_launches;
    if (n_pre_blocks > 0) {
        if (use_mma)
            rescore_chunks_kernel<<<ctx->sm_count * 16, 256, 0, st>>>(d_cand, j.cand_cap, pl->d_mma_blocks, pl->d_btiles, d_combos, seq->d_seq2, seq->d_mask,
                                                                      seq->d_seg_off, seq->nseg, len, mo->d_tab16, pl->d_th, hs, j.kl);
        else
            rescore_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(d_cand, j.cand_cap, d_combos, seq->d_seq2, seq->d_mask,
                                                              seq->d_seg_off, seq->nseg, len, mo->d_tab16, mo->d_tab4, pl->d_th, hs, j.kl);
        PWM_CUDA(cudaGetLastError());
        ++j.kernels;
    }
    if (!generic.empty() && len > 0) {
        const int bs = 256;
        const long long gx = std::min<long long>((len + bs - 1) / bs, 65535LL * 16);
        dim3 grid((unsigned)gx, (unsigned)std::min<size_t>(generic.size(), 65535));
        exact_kernel<<<grid, bs, 0, st>>>(d_generic, (int)generic.size(), pl->skip, seq->d_seq2, seq->d_mask, seq->d_ascii, seq->d_seg_off,
                                          seq->nseg, len, mo->d_tab16, pl->d_th, hs, j.kl);
        PWM_CUDA(cudaGetLastError());
        ++j.kernels;
    }
    if (use_km && !pl->skip && !pl->km_ambig.empty() && len > 0) {
        const int bs = 256;
        const long long gx = std::min<long long>((len + bs - 1) / bs, (long long)ctx->sm_count * 64);
        ambig_kernel<<<(unsigned)gx, bs, 0, st>>>(pl->d_km_ambig, (int)pl->km_ambig.size(), pl->km_ambig_max_n, seq->d_mask, seq->d_ascii, seq->d_seg_off,
                                                  seq->nseg, len, mo->d_tab16, pl->d_th, hs, j.kl);
        PWM_CUDA(cudaGetLastError());
        ++j.kernels;
    }
    PWM_CUDA(cudaEventRecord(lane->ev[EV_RES], st));
    if ((rc = job_order_finish(ctx, j, ow))) return rc;
    PWM_CUDA(cudaEventRecord(lane->ev[EV_ORD], st));
    if (chain_here && !(chain_pre || pst != st)) { PWM_CUDA(cudaEventRecord(ctx->prefilter_chain, st)); ctx->prefilter_chain_valid = true; }
    // (all small copies behind the kernels: a copy in front of the replay kernel cost the stream ~80 us per item)
    if (j.host && j.attempt == 0)     // the fromBS verdict of the packed bytes rides along: counters 0..5 in one copy
        PWM_CUDA(cudaMemcpyAsync(lane->h_counters, lane->d_counters, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    else
        PWM_CUDA(cudaMemcpyAsync(lane->h_counters, lane->d_counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    PWM_CUDA(cudaMemcpyAsync(lane->h_counters + 6, lane->d_counters + 6, 8, cudaMemcpyDeviceToHost, st));
    PWM_CUDA(cudaEventRecord(lane->ev[EV_A], st));
    return PWMSCAN_OK;
}

// Phase B: waits for the counters of phase A, repeats A with larger buffers if they overflowed (and the ordering with
// smaller buckets if one did not fit the sorting kernel's shared memory), then enqueues the copies of the compact columns.
int scan_phase_b(pwmscan_ctx* ctx, ScanJob& j) {
    ScanLane* lane = j.lane;
    const pwmscan_plan* pl = j.pl;
    cudaStream_t st = lane->stream;
    int rc;
    for (;;) {
        PWM_CUDA(cudaEventSynchronize(lane->ev[EV_A]));
        if (j.host && j.attempt == 0) {
            seq_finish(lane, j.up.seq);
            if (j.up.h_seq2) j.up.seq->first_non_iupac = -1;        // see pwmscan_seq_upload_packed
            ctx->last_host_non_acgt = j.up.seq->first_non_acgt; ctx->last_host_non_iupac = j.up.seq->first_non_iupac;
        }
        j.ncand = lane->h_counters[0]; j.nhits = lane->h_counters[1];
        if (pl->mode == PWM_MODE_MMA && lane->h_counters[3] != 0) j.ncand = std::max(j.ncand, j.cand_cap) + 2 * lane->h_counters[3];   // candidates that did not fit
        if (lane->h_counters[2] & 2ull) return set_err(ctx, PWMSCAN_ECUDA, "pwmscan_scan: tensor-core prefilter pipeline aborted (bounded wait expired)");
        if (lane->h_counters[2] != 0) return set_err(ctx, PWMSCAN_ENUCL, "Bio.Motif.Search.lookAheadSearch: invalid nucleotide");
        if (j.ncand <= j.cand_cap && j.nhits <= j.hit_cap) break;
        if (j.attempt == 3) return set_err(ctx, PWMSCAN_ERANGE, "pwmscan_scan: hit buffers kept overflowing");
        j.cand_cap = std::max(j.cand_cap, j.ncand + (j.ncand >> 3) + 1024);
        j.hit_cap = std::max(j.hit_cap, std::max(j.nhits, j.ncand) + (j.nhits >> 3) + 1024);
        ++j.attempt;
        if ((rc = scan_phase_a(ctx, j))) return rc;
    }
    // a bucket did not fit the sorting kernel's shared memory (far more hits than expected, or very uneven): smaller buckets,
    // again — positions are unique per (segment, motif, strand), so 2^kOrderIdxBits positions of one of them always fit
    for (unsigned long long over = lane->h_counters[6]; over != 0; over = lane->h_counters[6]) {
        if (j.shift == 0) return set_err(ctx, PWMSCAN_ERANGE, "pwmscan_scan: internal: ordering bucket overflow at shift 0");
        int drop = 1;
        while (((unsigned long long)kOrderCap << drop) < over) ++drop;
        j.shift = std::max(0, std::min(j.shift - drop - 1, order_pick_shift(j.nhits, j.key_bits)));
        if ((rc = job_order_buffers(ctx, j))) return rc;
        const OrderWork ow = order_work_view((unsigned char*)lane->scratch[SL_WORK], order_bucket_count(j.key_bits, j.shift), (size_t)j.hit_cap);
        PWM_CUDA(cudaMemsetAsync(ow.cnt, 0, ow.nb * 4, st));
        PWM_CUDA(cudaMemsetAsync(lane->d_counters + 6, 0, 8, st));
        if ((rc = order_count(ctx, st, (const unsigned long long*)lane->scratch[SL_KEYS], lane->d_counters + 1, j.hit_cap, j.shift, ow, &j.kernels))) return rc;
        if ((rc = job_order_finish(ctx, j, ow))) return rc;
        PWM_CUDA(cudaMemcpyAsync(lane->h_counters + 6, lane->d_counters + 6, 8, cudaMemcpyDeviceToHost, st));
        PWM_CUDA(cudaStreamSynchronize(st));
    }
    pwmscan_hits* h = new (std::nothrow) pwmscan_hits();
    if (!h) return set_err(ctx, PWMSCAN_EINVAL, "out of host memory");
    j.hits = h;
    h->ctx = ctx; h->count = (int64_t)j.nhits; h->nseg = j.seq->nseg; h->M = pl->motifs->M;
    if (j.host) { h->first_non_acgt = j.up.seq->first_non_acgt; h->first_non_iupac = j.up.seq->first_non_iupac; }
    if (j.nhits == 0) return PWMSCAN_OK;
    const size_t n = (size_t)j.nhits, n8 = (n + 7) & ~(size_t)7;             // keeps every host column 8-byte aligned
    const bool one = h->nseg == 1;
    j.n8 = n8;
    j.out_bytes = one ? n8 * 12 + (size_t)(2 * h->M + 1) * 8 : n8 * 18;
    h->buf = pin_pool_get(ctx, j.out_bytes, &h->buf_bytes);
    if (!h->buf) return PWMSCAN_ECUDA;
    // the compact columns: [score f64][pos32 u32] + [run_off i64 x (2M+1)] or [seg32 i32][ms16 u16]
    unsigned char* hb = (unsigned char*)h->buf;
    const OutCols oc = job_out_cols(j);
    // The hit lists of all lanes leave on ONE device->host stream (the host has seen this item's counters: everything the copies
    // read is final, no event to wait for).  On the lanes' own streams the copies of some lanes completed only together with
    // uploads that had been enqueued ahead on the host->device stream (4.8 instead of 1.2 ms for 25 MB).
    static const bool d2h_on_lane = std::getenv("PWMSCAN_D2H_LANE") != nullptr;          // experiments
    cudaStream_t ds = d2h_on_lane ? st : ctx->d2h_stream;
    PWM_CUDA(cudaMemcpyAsync(hb, oc.score, n * 8, cudaMemcpyDeviceToHost, ds));
    PWM_CUDA(cudaMemcpyAsync(hb + n8 * 8, oc.pos32, n * 4, cudaMemcpyDeviceToHost, ds));
    if (one) {
        PWM_CUDA(cudaMemcpyAsync(hb + n8 * 12, oc.run_off, (size_t)(2 * h->M + 1) * 8, cudaMemcpyDeviceToHost, ds));
    } else {
        PWM_CUDA(cudaMemcpyAsync(hb + n8 * 12, oc.seg32, n * 4, cudaMemcpyDeviceToHost, ds));
        PWM_CUDA(cudaMemcpyAsync(hb + n8 * 16, oc.ms16, n * 2, cudaMemcpyDeviceToHost, ds));
    }
    PWM_CUDA(cudaEventRecord(lane->ev[EV_B], ds));
    return PWMSCAN_OK;
}

// Phase C: waits for the hit list, hands it over (the job keeps nothing), releases a sequence the job owned.
int scan_phase_c(pwmscan_ctx* ctx, ScanJob& j, pwmscan_hits** out, double* timing /* accumulated, 12 entries */) {
    ScanLane* lane = j.lane;
    float ms_pre = 0, ms_res = 0, ms_ord = 0, ms_tot = 0;
    if (j.nhits > 0) {
        PWM_CUDA(cudaEventSynchronize(lane->ev[EV_B]));
        pwmscan_hits* h = j.hits;
        unsigned char* hb = (unsigned char*)h->buf;
        h->score = (const double*)hb;
        h->pos32 = (const uint32_t*)(hb + j.n8 * 8);
        if (h->nseg == 1) h->run_off = (const int64_t*)(hb + j.n8 * 12);
        else { h->seg32 = (const int32_t*)(hb + j.n8 * 12); h->ms16 = (const uint16_t*)(hb + j.n8 * 16); }
    }
    cudaEventElapsedTime(&ms_pre, lane->ev[EV_START], lane->ev[EV_PRE]);
    cudaEventElapsedTime(&ms_res, lane->ev[EV_PRE], lane->ev[EV_RES]);
    cudaEventElapsedTime(&ms_ord, lane->ev[EV_RES], lane->ev[EV_ORD]);
    cudaEventElapsedTime(&ms_tot, lane->ev[EV_START], lane->ev[EV_ORD]);
    timing[0] += ms_pre; timing[1] += ms_res; timing[2] += ms_ord; timing[3] += ms_tot;
    if (j.host) timing[4] += (j.up.h_seq2 ? (double)((j.seq->len + 31) / 32) * 12 : (double)j.seq->len) + (double)(j.seq->nseg + 1) * 8;
    timing[5] += (double)j.out_bytes + 48;
    timing[6] += (double)j.ncand;
    timing[7] += (double)j.prefilter_launches;
    timing[8] += (double)j.kernels;
    timing[9] += (double)j.nhits;
    *out = j.hits;
    j.hits = nullptr;
    job_release(j);
    return PWMSCAN_OK;
}

int scan_one(pwmscan_ctx* ctx, ScanJob& j, pwmscan_hits** out) {
    double t[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    int rc = scan_phase_a(ctx, j);
    if (!rc) rc = scan_phase_b(ctx, j);
    if (!rc) rc = scan_phase_c(ctx, j, out, t);
    if (rc) { cudaStreamSynchronize(j.lane->stream); cudaStreamSynchronize(j.lane->copy_stream); cudaStreamSynchronize(ctx->d2h_stream); job_release(j); return rc; }
    for (int i = 0; i < 12; ++i) ctx->timing[i] = t[i];
    return PWMSCAN_OK;
}

// Sets up the job of a host-bytes scan: allocates the sequence planes, the upload happens inside phase A.
int job_init_host(pwmscan_ctx* ctx, ScanJob& j, ScanLane* lane, const pwmscan_plan* pl, const uint8_t* ascii, const int64_t* seg_off, int32_t nseg, int flags) {
    j = ScanJob();
    j.lane = lane; j.pl = pl; j.host = true; j.active = true;
    if (!pl->skip) flags |= PWMSCAN_SEQ_KEEP_ASCII;
    pwmscan_seq* s = nullptr;
    uint8_t* d_raw = nullptr;
    int rc = seq_alloc(ctx, lane, seg_off, nseg, flags, &s, &d_raw, true);
    if (rc) return rc;
    j.up.ascii = ascii; j.up.d_raw = d_raw; j.up.seq = s;
    j.seq = s;
    return PWMSCAN_OK;
}

// Same for packed planes (no raw-byte buffer on the device).
int job_init_packed(pwmscan_ctx* ctx, ScanJob& j, ScanLane* lane, const pwmscan_plan* pl, const uint32_t* h_seq2, const uint32_t* h_mask, int64_t len) {
    j = ScanJob();
    j.lane = lane; j.pl = pl; j.host = true; j.active = true;
    pwmscan_seq* s = nullptr;
    uint8_t* d_raw = nullptr;
    const int64_t off[2] = {0, len};
    int rc = seq_alloc(ctx, lane, off, 1, 0, &s, &d_raw, true, false);
    if (rc) return rc;
    j.up.h_seq2 = h_seq2; j.up.h_mask = h_mask; j.up.seq = s;
    j.seq = s;
    return PWMSCAN_OK;
}

}  // namespace

extern "C" int pwmscan_scan(pwmscan_ctx* ctx, const pwmscan_seq* seq, const pwmscan_plan* pl, pwmscan_hits** out) {
    if (!ctx || !seq || !pl || !out) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_scan: NULL argument");
    *out = nullptr;
    if (seq->ctx != ctx || pl->ctx != ctx) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_scan: handles belong to another context");
    if (!pl->skip && !seq->d_ascii) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_scan: skip_ambiguous = 0 needs a sequence uploaded with PWMSCAN_SEQ_KEEP_ASCII");
    std::lock_guard<std::mutex> lk(ctx->mu);
    PWM_CUDA(cudaSetDevice(ctx->device));
    int rc = ensure_lut(ctx);
    if (rc) return rc;
    ScanJob j;
    j.lane = ctx->lanes[0]; j.seq = seq; j.pl = pl; j.active = true;
    return scan_one(ctx, j, out);
}

extern "C" int pwmscan_scan_host(pwmscan_ctx* ctx, const pwmscan_plan* pl, const uint8_t* ascii, const int64_t* seg_off, int32_t nseg,
                                 int flags, pwmscan_hits** out) {
    if (!ctx || !pl || !out || !seg_off) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_scan_host: NULL argument");
    *out = nullptr;
    if (pl->ctx != ctx) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_scan_host: plan belongs to another context");
    int rc = seq_check_args(ctx, ascii, seg_off, nseg);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    PWM_CUDA(cudaSetDevice(ctx->device));
    if ((rc = ensure_lut(ctx))) return rc;
    ScanJob j;
    if ((rc = job_init_host(ctx, j, ctx->lanes[0], pl, ascii, seg_off, nseg, flags))) return rc;
    rc = scan_one(ctx, j, out);
    if (rc) cudaDeviceSynchronize();          // no copy may still read the caller's buffer after we return
    return rc;
}

// Many scans of one plan with up to kMaxLanes in flight: while the first stage of one item runs, the exact replay, the
// ordering and the hit-list copy of the previous ones proceed on their own lanes.  Hit lists are handed to `cb` in
// item order (the callee owns them: pwmscan_hits_free); same results as one pwmscan_scan / pwmscan_scan_host per item.
static int scan_many_impl(pwmscan_ctx* ctx, const pwmscan_plan* pl, int32_t n, const pwmscan_seq* const* seqs, const uint8_t* const* ascii,
                          const int64_t* lens, int flags, pwmscan_hits_cb cb, void* user,
                          const uint32_t* const* p_seq2 = nullptr, const uint32_t* const* p_mask = nullptr) {
    if (n == 0) return PWMSCAN_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    PWM_CUDA(cudaSetDevice(ctx->device));
    int rc = ensure_lut(ctx);
    if (rc) return rc;
    const char* env_lanes = std::getenv("PWMSCAN_LANES");                  // read per call: experiments sweep it within one process
    // An item holds its lane from the start of its first stage until its hit list has landed: its replay / ordering wait for
    // SMs behind the next item's first stage, then comes the copy — about four first-stage durations.  Six (seven with the
    // upload in front) keep the next first stage enqueued in time; with fewer the device idles between launches.
    const int want_lanes = env_lanes ? std::max(1, std::min(kMaxLanes, std::atoi(env_lanes))) : (seqs ? 6 : 7);
    const int K = std::min<int>(want_lanes, n);
    ScanJob jobs[kMaxLanes];
    double t[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    auto start = [&](int i) -> int {
        ScanJob& j = jobs[i % K];
        ScanLane* lane = nullptr;
        int r = lane_get(ctx, i % K, &lane);
        if (r) return r;
        if (seqs) {
            j = ScanJob();
            j.lane = lane; j.seq = seqs[i]; j.pl = pl; j.active = true;
        } else if (p_seq2) {
            if ((r = job_init_packed(ctx, j, lane, pl, p_seq2[i], p_mask[i], lens[i]))) return r;
        } else {
            const int64_t off[2] = {0, lens[i]};
            if ((r = job_init_host(ctx, j, lane, pl, ascii[i], off, 1, flags))) return r;
            j.whole_upload = K > 1 && lens[i] <= ((int64_t)1 << 27);
        }
        return scan_phase_a(ctx, j);
    };
    static const bool trace = std::getenv("PWMSCAN_TRACE") != nullptr;       // per-item timeline on stderr (ms since the call started)
    // First-stage launches of different lanes overlap (the next one fills the tail of the running one), so the first-stage
    // time of the call is the length of the UNION of their [start, end] intervals, measured against one reference event.
    cudaEvent_t ref = nullptr;
    PWM_CUDA(cudaEventCreate(&ref));
    PWM_CUDA(cudaEventRecord(ref, ctx->lanes[0]->stream));
    PWM_CUDA(cudaEventSynchronize(ref));
    std::vector<std::pair<float, float>> pre_iv;
    pre_iv.reserve((size_t)n);
    const auto host_t0 = std::chrono::steady_clock::now();
    auto host_ms = [&]() { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - host_t0).count(); };
    std::vector<float> hb_in, hb_out, ha_in, ha_out;          // trace: host times of phase A / phase B of every item
    if (trace) { hb_in.assign((size_t)n, -1.f); hb_out.assign((size_t)n, -1.f); ha_in.assign((size_t)n, -1.f); ha_out.assign((size_t)n, -1.f); }
    auto finish = [&](int i) -> int {
        pwmscan_hits* h = nullptr;
        ScanJob& fj = jobs[i % K];
        const bool had_hits = fj.nhits > 0, was_host = fj.host;
        int r = scan_phase_c(ctx, fj, &h, t);
        if (r) return r;
        if (fj.prefilter_launches > 0) {
            float a = 0, b = 0;
            if (cudaEventElapsedTime(&a, ref, fj.lane->ev[EV_PSTART]) == cudaSuccess && cudaEventElapsedTime(&b, ref, fj.lane->ev[EV_PRE]) == cudaSuccess) pre_iv.emplace_back(a, b);
            else cudaGetLastError();
        }
        if (trace) {
            ScanLane* l = fj.lane;
            auto at = [&](int e) { float ms = -1; cudaEventElapsedTime(&ms, ref, l->ev[e]); return ms; };
            if (!had_hits) cudaEventSynchronize(l->ev[EV_A]);
            std::fprintf(stderr, "[trace] item %3d lane %d: h2d %8.3f..%8.3f first-stage %8.3f..%8.3f rescore ..%8.3f order ..%8.3f d2h ..%8.3f | host: enqueue %7.3f..%7.3f counters %7.3f..%7.3f handed over %7.3f\n", i, i % K,
                         was_host ? at(EV_H2D0) : -1.f, was_host ? at(EV_H2D1) : -1.f, at(EV_PSTART), at(EV_PRE), at(EV_RES), at(EV_ORD), had_hits ? at(EV_B) : -1.f,
                         ha_in[i], ha_out[i], hb_in[i], hb_out[i], host_ms());
        }
        if (cb) cb(user, i, h); else pwmscan_hits_free(h);
        return PWMSCAN_OK;
    };
    // Item j runs on lane j % K, which is free once item j - K has been handed over.  Items are started as far ahead as the
    // lanes allow: their uploads and first stages queue up behind the running one, so the device never waits for the host
    // (which spends its time waiting for the counters of the oldest item) nor for a transfer.
    int next_start = 0, done = 0;
    auto fill = [&]() -> int {
        while (next_start < n && next_start < done + K) {
            if (trace) ha_in[next_start] = host_ms();
            int r = start(next_start);
            if (r) return r;
            if (trace) ha_out[next_start] = host_ms();
            ++next_start;
        }
        return PWMSCAN_OK;
    };
    // The host follows the device by events instead of a fixed order of waits: the hit-list copy of item i is enqueued the
    // moment its counters have landed (phase B), whether or not the copy of item i-1 is still in flight, and a lane is refilled
    // the moment its hit list has been handed over.  (A loop that waited for copy i-1 before looking at the counters of item
    // i+1 enqueued every copy ~2 ms late; end to end from host bytes that was 87 instead of 77 ms per genome.)
    int next_b = 0;                                   // next item whose counters are waited for
    while (!rc && done < n) {
        if ((rc = fill())) break;
        bool progressed = false;
        while (!rc && done < next_b) {                // hit lists that have landed, in item order
            ScanJob& fj = jobs[done % K];
            if (fj.nhits > 0 && cudaEventQuery(fj.lane->ev[EV_B]) != cudaSuccess) { cudaGetLastError(); break; }
            rc = finish(done);
            ++done;
            progressed = true;
            if (!rc) rc = fill();
        }
        if (rc) break;
        if (next_b < next_start) {
            bool ready = done == next_b;              // nothing else to wait for: block inside phase B
            if (!ready) { ready = cudaEventQuery(jobs[next_b % K].lane->ev[EV_A]) == cudaSuccess; if (!ready) cudaGetLastError(); }
            if (ready) {
                if (trace) hb_in[next_b] = host_ms();
                rc = scan_phase_b(ctx, jobs[next_b % K]);
                if (rc) break;
                if (trace) hb_out[next_b] = host_ms();
                ++next_b;
                progressed = true;
            }
        } else if (!progressed && done < next_b) {    // every started item has had its phase B: wait for the oldest hit list
            rc = finish(done);
            ++done;
            progressed = true;
        }
        if (!progressed) std::this_thread::yield();
    }
    if (rc) {                                     // drain: nothing may still touch caller or pooled buffers
        cudaDeviceSynchronize();
        for (int k = 0; k < K; ++k) job_release(jobs[k]);
        cudaEventDestroy(ref);
        return rc;
    }
    cudaEventDestroy(ref);
    if (!pre_iv.empty()) {
        std::sort(pre_iv.begin(), pre_iv.end());
        double busy = 0;
        float lo = pre_iv[0].first, hi = pre_iv[0].second;
        for (size_t k = 1; k < pre_iv.size(); ++k) {
            if (pre_iv[k].first > hi) { busy += hi - lo; lo = pre_iv[k].first; hi = pre_iv[k].second; }
            else hi = std::max(hi, pre_iv[k].second);
        }
        t[0] = busy + (hi - lo);
    }
    for (int i = 0; i < 12; ++i) ctx->timing[i] = t[i];
    return PWMSCAN_OK;
}

extern "C" int pwmscan_scan_many(pwmscan_ctx* ctx, const pwmscan_plan* pl, int32_t n, const pwmscan_seq* const* seqs, pwmscan_hits_cb cb, void* user) {
    if (!ctx || !pl || n < 0 || (n > 0 && !seqs)) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_scan_many: bad argument");
    if (pl->ctx != ctx) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_scan_many: plan belongs to another context");
    for (int32_t i = 0; i < n; ++i) {
        if (!seqs[i] || seqs[i]->ctx != ctx) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_scan_many: sequence belongs to another context");
        if (!pl->skip && !seqs[i]->d_ascii) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_scan_many: skip_ambiguous = 0 needs PWMSCAN_SEQ_KEEP_ASCII sequences");
    }
    return scan_many_impl(ctx, pl, n, seqs, nullptr, nullptr, 0, cb, user);
}

extern "C" int pwmscan_scan_host_many(pwmscan_ctx* ctx, const pwmscan_plan* pl, int32_t n, const uint8_t* const* ascii, const int64_t* lens, int flags,
                                      pwmscan_hits_cb cb, void* user) {
    if (!ctx || !pl || n < 0 || (n > 0 && (!ascii || !lens))) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_scan_host_many: bad argument");
    if (pl->ctx != ctx) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_scan_host_many: plan belongs to another context");
    for (int32_t i = 0; i < n; ++i) {
        const int64_t off[2] = {0, lens[i]};
        int rc = seq_check_args(ctx, ascii[i], off, 1);
        if (rc) return rc;
    }
    return scan_many_impl(ctx, pl, n, nullptr, ascii, lens, flags, cb, user);
}

extern "C" int pwmscan_scan_packed_many(pwmscan_ctx* ctx, const pwmscan_plan* pl, int32_t n, const uint32_t* const* seq2, const uint32_t* const* mask,
                                        const int64_t* lens, pwmscan_hits_cb cb, void* user) {
    if (!ctx || !pl || n < 0 || (n > 0 && (!seq2 || !mask || !lens))) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_scan_packed_many: bad argument");
    if (pl->ctx != ctx) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_scan_packed_many: plan belongs to another context");
    if (!pl->skip) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_scan_packed_many: skip_ambiguous = 0 needs the sequence bytes (IUPAC letters are not kept in the packed planes)");
    for (int32_t i = 0; i < n; ++i) {
        if (lens[i] < 0 || (lens[i] > 0 && (!seq2[i] || !mask[i]))) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_scan_packed_many: bad length / NULL planes");
        if (lens[i] >= ((int64_t)1 << 32)) return set_err(ctx, PWMSCAN_ERANGE, "pwmscan_scan_packed_many: sequence longer than 2^32");
    }
    return scan_many_impl(ctx, pl, n, nullptr, nullptr, lens, 0, cb, user, seq2, mask);
}

extern "C" int pwmscan_scan_host_first_invalid(pwmscan_ctx* ctx, int alphabet, int64_t* pos_out) {
    if (!ctx || !pos_out || alphabet < 0 || alphabet > 1) return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_scan_host_first_invalid: bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    *pos_out = alphabet == 0 ? ctx->last_host_non_acgt : ctx->last_host_non_iupac;
    return PWMSCAN_OK;
}

extern "C" int pwmscan_find_tfbs(pwmscan_ctx* ctx, const pwmscan_seq* seq, const pwmscan_motifs* motifs, const double* thres,
                                 int strands, int skip_ambiguous, pwmscan_hits** out) {
    pwmscan_plan* pl = nullptr;
    int rc = pwmscan_plan_create(ctx, motifs, thres, strands, skip_ambiguous, &pl);
    if (rc) return rc;
    rc = pwmscan_scan(ctx, seq, pl, out);
    pwmscan_plan_free(pl);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// hit lists
// ------------------------------------------------------------------------------------------------
namespace pwmscan {

// Expands the compact columns into the wide ones (host memory).  Idempotent and thread-safe.
int hits_expand_wide(pwmscan_hits* h) {
    std::lock_guard<std::mutex> lk(h->wide_mu);
    if (h->pos || h->count == 0) return PWMSCAN_OK;
    const size_t n = (size_t)h->count, n8 = (n + 7) & ~(size_t)7;
    const bool one = h->nseg == 1;
    unsigned char* w = (unsigned char*)std::malloc(n8 * (8 + 4 + 1 + (one ? 0 : 4)));
    if (!w) return set_err(h->ctx, PWMSCAN_EINVAL, "out of host memory");
    int64_t* pos = (int64_t*)w;
    int32_t* motif = (int32_t*)(w + n8 * 8);
    int32_t* segment = one ? nullptr : (int32_t*)(w + n8 * 12);
    uint8_t* strand = w + n8 * (one ? 12 : 16);
    auto part = [&](size_t i0, size_t i1) {
        if (one) {
            const int64_t* ro = h->run_off;
            int c = (int)(std::upper_bound(ro, ro + 2 * h->M + 1, (int64_t)i0) - ro) - 1;
            for (size_t i = i0; i < i1; ++i) {
                while ((int64_t)i >= ro[c + 1]) ++c;
                pos[i] = (int64_t)h->pos32[i]; motif[i] = c >> 1; strand[i] = (uint8_t)(c & 1);
            }
        } else {
            for (size_t i = i0; i < i1; ++i) {
                pos[i] = (int64_t)h->pos32[i]; motif[i] = (int32_t)(h->ms16[i] >> 1); strand[i] = (uint8_t)(h->ms16[i] & 1); segment[i] = h->seg32[i];
            }
        }
    };
    const unsigned nth = n < (1u << 20) ? 1u : std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    if (nth <= 1) part(0, n);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nth; ++t) th.emplace_back(part, n * t / nth, n * (t + 1) / nth);
        for (auto& x : th) x.join();
    }
    if (one) {
        const int32_t* z = zero_column(h->ctx, n8);
        if (!z) { std::free(w); return set_err(h->ctx, PWMSCAN_EINVAL, "out of host memory"); }
        h->segment = z;
    } else {
        h->segment = segment;
    }
    h->wide = w; h->motif = motif; h->strand = strand;
    h->pos = pos;
    return PWMSCAN_OK;
}

}  // namespace pwmscan

extern "C" int64_t pwmscan_hits_count(const pwmscan_hits* h) { return h ? h->count : -1; }

extern "C" int pwmscan_hits_first_invalid(const pwmscan_hits* h, int alphabet, int64_t* pos_out) {
    if (!h || !pos_out || alphabet < 0 || alphabet > 1) return PWMSCAN_EINVAL;
    *pos_out = alphabet == 0 ? h->first_non_acgt : h->first_non_iupac;
    return PWMSCAN_OK;
}

extern "C" int pwmscan_hits_columns(const pwmscan_hits* h, const int32_t** motif, const uint8_t** strand, const int32_t** segment,
                                    const int64_t** pos, const double** score) {
    if (!h) return PWMSCAN_EINVAL;
    if (motif || strand || segment || pos) {
        int rc = hits_expand_wide(const_cast<pwmscan_hits*>(h));
        if (rc) return rc;
    }
    if (motif) *motif = h->motif;
    if (strand) *strand = h->strand;
    if (segment) *segment = h->segment;
    if (pos) *pos = h->pos;
    if (score) *score = h->score;
    return PWMSCAN_OK;
}

extern "C" int pwmscan_hits_compact(const pwmscan_hits* h, const uint32_t** pos32, const double** score, const int64_t** run_off,
                                    const int32_t** segment, const uint16_t** motif_strand) {
    if (!h) return PWMSCAN_EINVAL;
    if (pos32) *pos32 = h->pos32;
    if (score) *score = h->score;
    if (run_off) *run_off = h->run_off;
    if (segment) *segment = h->seg32;
    if (motif_strand) *motif_strand = h->ms16;
    return PWMSCAN_OK;
}

extern "C" void pwmscan_hits_free(pwmscan_hits* h) {
    if (!h) return;
    if (h->buf && h->ctx) pin_pool_put(h->ctx, h->buf, h->buf_bytes);
    if (h->wide) std::free(h->wide);
    delete h;
}

extern "C" void pwmscan_free(void* p) { std::free(p); }
