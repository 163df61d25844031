// internal.h — handle layouts shared by the translation units of libpwmscan.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/pwmscan.h"
#include "plan.h"
#include "plan_mma.h"
#include "plan_km.h"

// Sort key of a hit: [segment | motif : mb bits | strand : 1 bit | position in segment : pb bits] — the reference's
// output order (Data/Bed/Utils.hs:108-118) is the ascending order of this key.
struct KeyLayout { int pb, mb; };

constexpr int kOrderCap = 4096;        // hits one CTA orders in shared memory (order.cu)
constexpr int kOrderIdxBits = 12;      // log2(kOrderCap)
constexpr int kOrderTarget = 8;        // hits per bucket aimed at (a warp orders spans of whole buckets, up to 256 hits, in registers: order.cu)
constexpr int kMaxLanes = 12;          // scans in flight of pwmscan_scan_many (default: 6 resident items, 7 host items)

// Everything one in-flight scan needs of its own: streams, events, counters, grow-only scratch.  Lane 0 shares the
// context's streams and counters (single scans run on it); lanes 1.. are created on demand by pwmscan_scan_many.
struct ScanLane {
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaStream_t pstream = nullptr;              // low priority: upload-side packing + first stage; `stream` (high priority) carries the exact replay,
                                                 // the ordering and the hit-list copy, which are short and slip in between retiring first-stage CTAs
    bool owns_streams = false;
    std::vector<cudaEvent_t> chunk_ev;           // one per in-flight host->device chunk
    cudaEvent_t ev[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    void* scratch[12] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t scratch_bytes[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    unsigned long long* d_counters = nullptr;    // [0] candidates [1] hits [2] error flag [3] mma overflow [4,5] fromBS verdict
                                                 // [6] largest ordering bucket that did not fit; [8..] work counters
    unsigned long long* h_counters = nullptr;    // pinned mirror of [0..7]
};

struct pwmscan_ctx {
    int device = 0;
    std::vector<ScanLane*> lanes;                // lanes[0] aliases stream / copy_stream / d_counters / h_counters below
    cudaEvent_t prefilter_chain = nullptr;       // completion of the last first-stage kernel of any lane: the PERSISTENT first stages
    bool prefilter_chain_valid = false;          // (lds, mma, kmer kernels 1 and 2) run one after the other; the default one is not chained
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;          // host->device chunks of pwmscan_scan_host
    cudaStream_t d2h_stream = nullptr;           // hit lists of all lanes to the host
    std::vector<cudaEvent_t> chunk_ev;           // one per in-flight chunk (created on demand)
    std::mutex mu;
    std::string err;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double timing[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    int64_t last_host_non_acgt = -1, last_host_non_iupac = -1;   // fromBS verdict of the last pwmscan_scan_host
    int sm_count = 148;
    bool lut_ready = false, dense_lut_ready = false, smem_attr_set = false, mma_attr_set = false;   // per-context one-time device setup
    bool km_attr_set = false, km_build_attr_set = false;
    // grow-only device scratch
    void* scratch[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t scratch_bytes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // pinned host staging (grow-only)
    void* pinned[2] = {nullptr, nullptr};
    size_t pinned_bytes[2] = {0, 0};
    // small caches of freed device / pinned buffers (sequence planes, hit lists): cudaMalloc,
    // cudaFree and cudaMallocHost cost milliseconds and would dominate a scan of a few ms
    struct PoolBuf { void* p; size_t bytes; };
    std::vector<PoolBuf> dev_pool, pin_pool;
    std::mutex pool_mu;
    // all-zero int32 column shared by the hit lists of single-segment scans (their segment column is not transferred);
    // grown by allocating a new zeroed block — older blocks stay alive for the hit lists that point into them
    std::vector<void*> zero_blocks;
    size_t zero_count = 0;
    unsigned long long* d_counters = nullptr;   // [0] candidates [1] hits [2] error flag [3] work counter (32-bit view)
    unsigned long long* h_counters = nullptr;   // pinned mirror
};

struct pwmscan_seq {
    pwmscan_ctx* ctx = nullptr;
    int64_t len = 0;                 // concatenated length
    int32_t nseg = 1;
    std::vector<int64_t> seg_off;    // nseg+1
    int64_t* d_seg_off = nullptr;
    size_t cap_seq2 = 0, cap_mask = 0, cap_ascii = 0, cap_seg = 0;   // allocation sizes (for the pool)
    uint32_t* d_seq2_alloc = nullptr;   // 1 front word + words + tail
    uint32_t* d_seq2 = nullptr;         // = alloc + 1; stored base u at word u>>4
    uint32_t* d_mask = nullptr;         // stored base u at bit u&31 of word u>>5 (1 = not ACGT)
    uint8_t* d_ascii = nullptr;         // optional (KEEP_ASCII): upper-cased bytes, position p at [p]
    int64_t nwords2 = 0, nwords_mask = 0;
    int64_t first_non_acgt = -1, first_non_iupac = -1;
    int flags = 0;
    bool borrowed = false;              // device planes belong to a scan lane's scratch (host scans), not to the pool
};

struct pwmscan_motifs {
    pwmscan_ctx* ctx = nullptr;
    int32_t M = 0;
    double bg[4] = {0.25, 0.25, 0.25, 0.25};
    std::vector<pwmscan::MotifHost> mh;
    std::vector<int64_t> tab_off;       // offset of (motif, strand) in d_tab16 (doubles): tab_off[2m+s]
    std::vector<int64_t> col_off;       // offset of (motif, strand) in per-column arrays: col_off[2m+s]
    double* d_tab16 = nullptr;          // all IUPAC tables
    double* d_tab4 = nullptr;           // the A/C/G/T entries of every column alone, [col_off*4 + 4k + base]: what the exact replay of a
                                        // skip-ambiguous scan reads (32 B per column, four columns per cache line instead of one)
    double* d_mats = nullptr;           // all probability matrices (strand 0 and 1), n x 4, same order as col_off*4
    int64_t total_cols2 = 0;            // sum over (motif,strand) of n
};

struct DevCombo {
    int32_t motif;      // caller's motif index, -1 = empty
    int32_t strand;
    int32_t n;
    int32_t c;          // anchor column
    int64_t tab_off;    // doubles into d_tab16
    int64_t th_off;     // doubles into d_th
};

struct DevBlock {
    int32_t s1, nl, nslot;
    int32_t rep_log2;    // log2 of the lane replication factor (0 = 64 motifs across the 32 lanes)
    int64_t table_off;   // 32-bit words into d_tables
};

struct DevMmaBlock {
    int32_t steps;       // MMAs (8 motif columns each) per phase tile
    int32_t ncombo;      // N of the MMA (multiple of 16)
    int64_t b_off;       // bytes into d_btiles
};

struct DevKmBlock {
    int32_t nt, max_n;   // mask tables of this block; longest motif (row stride of q16)
    // per gather step s (most selective table first): the table's k-mer is bits [shift, shift + 2L) of the
    // 64-bit window (32 bases from the stored position); its masks start table_off[s] bytes into d_km_tables
    int32_t shift[4];
    uint32_t kmask[4];   // 4^L - 1
    int64_t table_off[4];
    int64_t q16_off;     // uint16 elements into d_km_q16 ([256][max_n][4])
    int64_t bitmap_off;  // bytes into d_km_tables of the "first mask non-zero" bitmap (one bit per k-mer of step 0)
    int32_t bitmap_words;  // 32-bit words of it; 0 = this block runs without the bitmap
    int32_t direct;        // 1 = the block's single table sees every column of every combo: a surviving bit is the whole-window test
};

enum { PWM_MODE_LDS = 0, PWM_MODE_MMA = 1, PWM_MODE_KMER = 2 };

struct pwmscan_plan {
    int mode = PWM_MODE_LDS;           // which first stage this plan was built for
    pwmscan_ctx* ctx = nullptr;
    const pwmscan_motifs* motifs = nullptr;
    int strands = 2, skip = 1;
    int nblocks = 0;                   // prefilter blocks
    int max_nslot = 0;
    std::vector<pwmscan::BlockPlan> blocks;
    std::vector<DevCombo> generic;     // combos handled by the exact kernel only
    DevBlock* d_blocks = nullptr;
    DevCombo* d_combos = nullptr;      // nblocks * 128
    DevCombo* d_generic = nullptr;
    uint32_t* d_tables = nullptr;
    double* d_th = nullptr;            // th[k] = thres - sigma[k] per (motif,strand)
    double cand_rate = 0;              // expected candidates per sequence position
    int max_n = 0;
    // tensor-core first stage (mma_prefilter.cu): its own blocks of 128 motifs
    int mma_nblocks = 0;
    std::vector<pwmscan::MmaBlockPlan> mma_blocks;
    std::vector<DevCombo> mma_generic;
    DevMmaBlock* d_mma_blocks = nullptr;
    DevCombo* d_mma_combos = nullptr;  // mma_nblocks * 256
    DevCombo* d_mma_generic = nullptr;
    int8_t* d_btiles = nullptr;
    double mma_cand_rate = 0;
    // k-mer mask first stage (km_prefilter.cu): blocks of 128 motifs
    int km_nblocks = 0, km_max_n = 1;
    size_t km_smem_bytes = 0;          // dynamic shared memory the prefilter kernel needs (largest block)
    std::vector<pwmscan::KmBlockPlan> km_blocks;
    std::vector<DevCombo> km_generic;
    std::vector<DevCombo> km_ambig;              // skip == 0: every prefiltered combo, for the pass over windows with a non-ACGT letter
    DevCombo* d_km_ambig = nullptr;
    int km_ambig_max_n = 0;
    DevKmBlock* d_km_blocks = nullptr;
    DevCombo* d_km_combos = nullptr;   // km_nblocks * 256
    DevCombo* d_km_generic = nullptr;
    uint8_t* d_km_tables = nullptr;
    size_t km_tables_cap = 0;          // pooled allocation size
    uint16_t* d_km_q16 = nullptr;
    double km_cand_rate = 0, km_surv_rate = 0, km_req_rate = 0;
};

// What a scan hands back.  The device transfers the COMPACT columns only (12 bytes per hit for a single-segment scan):
//   score f64[n], pos32 u32[n] and, single segment, run_off i64[2M+1] (hits of (motif m, strand s) are
//   [run_off[2m+s], run_off[2m+s+1])); several segments: seg32 i32[n] + ms16 u16[n] (= 2 * motif + strand).
// The wide columns of pwmscan_hits_columns are expanded on the host the first time they are asked for.
struct pwmscan_hits {
    pwmscan_ctx* ctx = nullptr;
    int64_t count = 0;
    int32_t nseg = 1, M = 0;
    int64_t first_non_acgt = -1, first_non_iupac = -1;   // fromBS verdict of the bytes a host scan uploaded for this list
    void* buf = nullptr;          // pinned
    size_t buf_bytes = 0;
    const double* score = nullptr;
    const uint32_t* pos32 = nullptr;
    const int64_t* run_off = nullptr;     // nseg == 1
    const int32_t* seg32 = nullptr;       // nseg > 1
    const uint16_t* ms16 = nullptr;       // nseg > 1
    // lazily expanded (host memory, owned)
    std::mutex wide_mu;
    void* wide = nullptr;
    const int64_t* pos = nullptr;
    const int32_t* motif = nullptr;
    const int32_t* segment = nullptr;
    const uint8_t* strand = nullptr;
};

namespace pwmscan {

int set_err(pwmscan_ctx* ctx, int code, const std::string& msg);
int cuda_fail(pwmscan_ctx* ctx, cudaError_t e, const char* what);
// grow-only scratch; returns nullptr on failure (error already recorded)
void* scratch_get(pwmscan_ctx* ctx, int slot, size_t bytes);
void* pinned_get(pwmscan_ctx* ctx, int slot, size_t bytes);
// pooled allocations: *bytes_out receives the real capacity to hand back to pool_put
void* dev_pool_get(pwmscan_ctx* ctx, size_t bytes, size_t* bytes_out);
void dev_pool_put(pwmscan_ctx* ctx, void* p, size_t bytes);
void* pin_pool_get(pwmscan_ctx* ctx, size_t bytes, size_t* bytes_out);
void pin_pool_put(pwmscan_ctx* ctx, void* p, size_t bytes);
void* lane_scratch_get(pwmscan_ctx* ctx, ScanLane* lane, int slot, size_t bytes);
int hits_expand_wide(pwmscan_hits* h);     // fills pos / motif / segment / strand (idempotent, thread-safe)
// order.cu
struct OrderWork { size_t nb, big_cap, med_cap; unsigned int* cnt; unsigned long long *start, *cursor, *tiles, *big, *med; };
int order_pick_shift(unsigned long long n, int key_bits);
size_t order_bucket_count(int key_bits, int shift);
size_t order_work_bytes(size_t nb, size_t hit_cap);
OrderWork order_work_view(unsigned char* d_work, size_t nb, size_t hit_cap);
int order_count(pwmscan_ctx* ctx, cudaStream_t st, const unsigned long long* d_keys, const unsigned long long* d_nhits, unsigned long long cap,
                int shift, const OrderWork& w, int* launches);
int order_finish(pwmscan_ctx* ctx, cudaStream_t st, const unsigned long long* d_keys, const double* d_sc, unsigned long long n_est, KeyLayout kl,
                 int shift, int nseg, int M, const OrderWork& w, unsigned long long* d_tkeys, double* d_tsc,
                 unsigned long long* d_skeys, double* d_score, uint32_t* d_pos32, int32_t* d_seg32, uint16_t* d_ms16, long long* d_run_off,
                 unsigned long long* d_flags, int* launches);
// Where the kernels that find hits put them: unordered (key, score) pairs + one counter per ordering bucket (key >> shift).
struct HitSink {
    unsigned long long* keys;
    double* sc;
    unsigned long long cap;
    unsigned long long* counters;     // [1] = number of hits found (may exceed cap: the scan is then repeated with larger buffers)
    unsigned int* bucket_cnt;
    int shift;
};
int launch_mma_prefilter(pwmscan_ctx* ctx, ScanLane* lane, const uint32_t* d_seq2, long long a0, long long a1, int nblocks, const DevMmaBlock* d_blocks,
                         const int8_t* d_btiles, unsigned long long* d_cand, unsigned long long cand_cap);

int km_build_table(pwmscan_ctx* ctx, const double* d_wdef, const double* d_dsafe, int L, uint8_t* d_table);
int km_build_bitmap(pwmscan_ctx* ctx, const uint8_t* d_table, int L, uint8_t* d_bitmap);
bool km_prefilter_is_persistent();          // which k-mer first-stage kernel runs (PWMSCAN_KM_KERNEL; the default one is not persistent)
int launch_km_prefilter(pwmscan_ctx* ctx, ScanLane* lane, cudaStream_t stream, const uint32_t* d_seq2, long long a0, long long a1, long long len, int nblocks, size_t smem_bytes,
                        const DevKmBlock* d_blocks, const uint8_t* d_tables, const uint16_t* d_q16, const DevCombo* d_combos,
                        unsigned long long* d_cand, unsigned long long cand_cap, unsigned long long* work_counter);

#define PWM_CUDA(call)                                                          \
    do {                                                                        \
        cudaError_t e__ = (call);                                               \
        if (e__ != cudaSuccess) return pwmscan::cuda_fail(ctx, e__, #call);     \
    } while (0)

}  // namespace pwmscan
