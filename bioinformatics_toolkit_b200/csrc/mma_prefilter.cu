// mma_prefilter.cu — tensor-core first stage of the thresholded scan (sm_100a: tcgen05 + TMEM).
//
//   D[anchor, combo] = sum_{k<K1} sum_b onehot[anchor + k][b] * W[combo][4k+b]      (int8 x int8 -> int32)
//
// One persistent CTA per SM (31 warps), warp-specialised:
//   warps 0..1   producers: read the 2-bit genome and write the one-hot byte stream of a 512-anchor tile into
//                shared memory, four times (one copy per phase p = anchor mod 4, two copies per warp, so that
//                every copy is 16-byte aligned at its first anchor); double-buffered;
//   warp 2       one elected thread issues tcgen05.mma.kind::i8 (M = 128 anchors of one phase, N = combos,
//                K = 32 bytes = 8 motif columns per instruction).  The A descriptor is a Toeplitz view of
//                the one-hot stream (row r starts 16 bytes = 4 bases after row r-1: LBO = 16 B, SBO = 128 B),
//                so no im2col matrix is ever built.  Accumulators live in TMEM, double-buffered (2 x 256 columns);
//   warps 3..30  epilogue: seven warps per 32-lane quarter of TMEM (= 32 anchors), each pulling one 32-column
//                chunk of its rows with tcgen05.ld.32x32b.x32, OR-reducing it (the sign bit of D says "still
//                possible") and recording (anchor, block, chunk) for the exact fp64 replay of the survivors.
//                (TMEM read: ~900 B/clk/SM with 16+ warps, experiments/tmem/tmem_x32.cu.)
// Stages hand over through mbarriers (tcgen05.commit arrives for the tensor pipe).  Every wait is bounded:
// a broken pipeline sets an abort flag and the kernel ends instead of hanging the GPU.
//
// Roofline: tensor pipe, 128 * N / 256 clocks per MMA (measured 129.4 clk for M128 N256 K32,
// experiments/mma/i8_toeplitz.cu) => (K1/8) * N/256 clocks per anchor per SM; 2 * 4 * K1 int8 MACs... per
// anchor and combo (algorithmic ops = 8 * K1 per bp*combo).
#include <cstdio>
#include <cstdlib>

#include "internal.h"
#include "plan_mma.h"

using namespace pwmscan;

namespace {

constexpr int kMmaThreads = 992;             // 31 warps: 2 producers, 1 MMA issuer, 28 epilogue
constexpr int kEpiWarps = 28;                // 7 chunks of 32 combos x 4 TMEM lane quarters: one tcgen05.ld.x32 per warp and phase tile
constexpr int kFirstEpiWarp = 3;
constexpr int kAccStages = 2;                // TMEM accumulator buffers of 256 columns (N <= 256)
constexpr int kWarpChunk = 256;              // candidate-list entries a warp reserves per global atomic
constexpr int kTile = 512;                   // anchors per one-hot tile (4 phases x 128 rows)
constexpr int kOhWords = 576;                // one-hot words (= bases) per phase copy: 512 + 32 + slack, multiple of 32
constexpr int kItemTiles = 32;               // tiles per work item (16 384 anchors)
constexpr uint32_t kSpinLimit = 1u << 26;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);               // start address           bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;      // leading byte offset     bits [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;      // stride byte offset      bits [32,46)
    d |= (uint64_t)1 << 46;                                // descriptor version (sm_100); no swizzle, base offset 0
    return d;
}

__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
// bounded wait (mbarrier.try_wait lets the hardware park the warp); false = timed out (pipeline broken)
__device__ __forceinline__ bool mbar_wait(uint64_t* b, uint32_t parity, volatile int* abort_flag) {
    const uint32_t a = smem_u32(b);
    for (uint32_t spin = 0; spin < kSpinLimit; ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) return true;
        if ((spin & 1023u) == 1023u && *abort_flag) return false;
    }
    *abort_flag = 1;
    return false;
}
__device__ __forceinline__ void umma_commit(uint64_t* b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}

struct __align__(16) MmaSmem {
    uint8_t onehot[2][4][kOhWords * 4];      // [stage][phase copy][bytes]   2 x 4 x 2304 B
    uint64_t full_oh[2], empty_oh[2], full_acc[kAccStages], empty_acc[kAccStages];
    uint32_t tmem_base;
    int abort_flag;
    unsigned int rec_used[kEpiWarps];        // records written into each epilogue warp's current chunk of the list
    unsigned long long rec_base[kEpiWarps];  // first slot of that chunk
};

// counters: [0] record slots reserved, [1] hits, [2] error flag (bit 1 = pipeline abort), [3] records that did not fit
// candidate record: anchor [0,40) | block [40,56) | 32-combo chunk [56,59); all ones = unused slot
__global__ void __launch_bounds__(kMmaThreads, 1)
mma_prefilter_kernel(const uint32_t* __restrict__ seq2, long long anchor0, long long n_anchor /* of this launch, multiple of 16384 */,
                     int nblocks, const DevMmaBlock* __restrict__ blocks, const int8_t* __restrict__ btiles,
                     unsigned long long* __restrict__ cand, unsigned long long cand_cap, unsigned long long* counters, int dbg) {
    extern __shared__ __align__(128) unsigned char dyn[];
    MmaSmem* sm = reinterpret_cast<MmaSmem*>(dyn);
    int8_t* bsm = reinterpret_cast<int8_t*>(dyn + ((sizeof(MmaSmem) + 127) & ~size_t(127)));      // up to 4 x 8 KB
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&sm->full_oh[i], 2); mbar_init(&sm->empty_oh[i], 1); }
        for (int i = 0; i < kAccStages; ++i) { mbar_init(&sm->full_acc[i], 1); mbar_init(&sm->empty_acc[i], kEpiWarps); }
        sm->abort_flag = 0;
        for (int i = 0; i < kEpiWarps; ++i) sm->rec_used[i] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_base)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = sm->tmem_base;
    volatile int* abort_flag = &sm->abort_flag;

    const long long n_items = n_anchor / ((long long)kItemTiles * kTile);
    uint32_t t_cnt = 0;        // one-hot tiles handled so far by this CTA (same sequence in every role)
    uint32_t g_cnt = 0;        // phase tiles (accumulator buffers) handled so far

    for (int blk = 0; blk < nblocks; ++blk) {
        const DevMmaBlock db = blocks[blk];
        const int steps = db.steps, N = db.ncombo;
        // ---- weights of this block -> shared memory (all threads), visible to the tensor core (async proxy)
        {
            const uint4* src = reinterpret_cast<const uint4*>(btiles + db.b_off);
            uint4* dst = reinterpret_cast<uint4*>(bsm);
            const int n16 = steps * N * 32 / 16;
            for (int i = tid; i < n16; i += kMmaThreads) dst[i] = __ldg(src + i);
            asm volatile("fence.proxy.async.shared::cta;");
        }
        __syncthreads();
        const uint32_t idesc = (2u << 4) /* D = s32 */ | (0u << 7) /* A = u8 */ | (1u << 10) /* B = s8 */ |
                               ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);      // both operands K-major

        if (warp < 2) {
            // ================= producers: warp w writes the one-hot copies of phases 2w and 2w+1 =================
            const int p0 = 2 * warp;
            for (long long item = blockIdx.x; item < n_items && !*abort_flag; item += gridDim.x) {
                for (int tt = 0; tt < kItemTiles; ++tt, ++t_cnt) {
                    const int s = t_cnt & 1;
                    const long long U = anchor0 + (item * kItemTiles + tt) * (long long)kTile;
                    const uint32_t* w = seq2 + (U >> 4);
                    const uint32_t w0 = __ldg(w + lane);                       // 36 packed words cover the 576 bases of a tile
                    const uint32_t w1 = lane < 4 ? __ldg(w + 32 + lane) : 0u;
                    if (t_cnt >= 2 && !mbar_wait(&sm->empty_oh[s], ((t_cnt >> 1) - 1) & 1, abort_flag)) break;
                    uint8_t* dst0 = &sm->onehot[s][p0][0];
                    uint8_t* dst1 = &sm->onehot[s][p0 + 1][0];
#pragma unroll
                    for (int m = 0; m < kOhWords / 32 && !(dbg & 8); ++m) {
                        const int a = lane + 32 * m;                           // base index relative to U
                        const int widx = a >> 4;                               // = 2m + (lane >> 4)
                        const uint32_t word = widx < 32 ? __shfl_sync(0xFFFFFFFFu, w0, widx & 31) : __shfl_sync(0xFFFFFFFFu, w1, widx & 31);
                        const uint32_t oh = 1u << (8 * ((word >> (2 * (a & 15))) & 3u));
                        if (a >= p0) *reinterpret_cast<uint32_t*>(dst0 + 4 * (a - p0)) = oh;
                        if (a >= p0 + 1) *reinterpret_cast<uint32_t*>(dst1 + 4 * (a - p0 - 1)) = oh;
                    }
                    asm volatile("fence.proxy.async.shared::cta;");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm->full_oh[s]);
                }
            }
        } else if (warp == 2) {
            // ================= MMA issuer =================
            for (long long item = blockIdx.x; item < n_items && !*abort_flag; item += gridDim.x) {
                for (int tt = 0; tt < kItemTiles; ++tt, ++t_cnt) {
                    const int s = t_cnt & 1;
                    if (!mbar_wait(&sm->full_oh[s], (t_cnt >> 1) & 1, abort_flag)) break;
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    bool ok = true;
                    for (int p = 0; p < 4 && ok; ++p, ++g_cnt) {
                        const int a = g_cnt & (kAccStages - 1);
                        if (g_cnt >= kAccStages && !mbar_wait(&sm->empty_acc[a], ((g_cnt / kAccStages) - 1) & 1, abort_flag)) { ok = false; break; }
                        asm volatile("tcgen05.fence::after_thread_sync;");
                        if (lane == 0) {
                            const uint32_t d_tmem = tmem + (uint32_t)(a * (512 / kAccStages));
                            for (int j = 0; j < steps && !(dbg & 4); ++j) {
                                const uint64_t adesc = umma_desc(smem_u32(&sm->onehot[s][p][0]) + 32u * j, 16, 128);   // Toeplitz rows
                                const uint64_t bdesc = umma_desc(smem_u32(bsm) + (uint32_t)(j * N * 32), 128, 256);
                                asm volatile(
                                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                                    ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)(j > 0)), "r"(0u));
                            }
                            umma_commit(&sm->full_acc[a]);
                        }
                        __syncwarp();
                    }
                    if (!ok) break;
                    if (lane == 0) umma_commit(&sm->empty_oh[s]);              // all MMAs reading this one-hot stage are done
                    __syncwarp();
                }
            }
        } else {
            // ================= epilogue: 28 warps, seven per TMEM lane quarter =================
            const int quarter = warp & 3;                 // hardware rule: warp w may touch TMEM lanes 32*(w%4) .. +31
            const int sub = (warp - kFirstEpiWarp) >> 2;  // which 32-combo chunk of the quarter's rows this warp scans
            // Survivors are rare and must not hold the TMEM buffer: a lane whose 32-column chunk still has a
            // possible combo only records (anchor, block, chunk) — one shared-memory atomic + one store — into
            // its warp's private 256-record chunk of the candidate list; the re-score kernel replays the 32
            // combos of such a record exactly.  New chunks are reserved after the buffer has been released.
            const int e = warp - kFirstEpiWarp;
            unsigned long long w_base = 0;
            if (e >= 0) {                                  // (always true here) first chunk of this warp
                unsigned long long nb = 0;
                if (lane == 0 && blk == 0) nb = atomicAdd(&counters[0], (unsigned long long)kWarpChunk);
                if (blk == 0) { w_base = __shfl_sync(0xFFFFFFFFu, nb, 0); if (lane == 0) sm->rec_base[e] = w_base; }
                __syncwarp();
                w_base = sm->rec_base[e];
            }
            for (long long item = blockIdx.x; item < n_items && !*abort_flag; item += gridDim.x) {
                for (int tt = 0; tt < kItemTiles; ++tt) {
                    const long long U = anchor0 + (item * kItemTiles + tt) * (long long)kTile;
                    bool ok = true;
                    for (int p = 0; p < 4; ++p, ++g_cnt) {
                        const int a = g_cnt & (kAccStages - 1);
                        if (!mbar_wait(&sm->full_acc[a], (g_cnt / kAccStages) & 1, abort_flag)) { ok = false; break; }
                        asm volatile("tcgen05.fence::after_thread_sync;");
                        const uint32_t trow = tmem + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(a * (512 / kAccStages));
                        const long long u = U + p + 4 * (32 * quarter + lane);          // this thread's anchor
                        for (int c0 = 32 * sub; c0 < N && !(dbg & 1); c0 += 32 * (kEpiWarps / 4)) {
                            uint32_t v[32];
                            asm volatile(
                                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
                                "tcgen05.wait::ld.sync.aligned;"
                                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                                : "r"(trow + (uint32_t)c0) : "memory");
                            // 3-input ORs (LOP3): 32 values -> 1 in 16 instructions
                            const uint32_t o0 = v[0] | v[1] | v[2], o1 = v[3] | v[4] | v[5], o2 = v[6] | v[7] | v[8], o3 = v[9] | v[10] | v[11];
                            const uint32_t o4 = v[12] | v[13] | v[14], o5 = v[15] | v[16] | v[17], o6 = v[18] | v[19] | v[20], o7 = v[21] | v[22] | v[23];
                            const uint32_t o8 = v[24] | v[25] | v[26], o9 = v[27] | v[28] | v[29], o10 = v[30] | v[31];
                            const uint32_t q0 = o0 | o1 | o2, q1 = o3 | o4 | o5, q2 = o6 | o7 | o8, q3 = o9 | o10;
                            if (dbg & 2) { if ((q0 | q1 | q2 | q3) == 0x12345678u) counters[5] = 1; continue; }
                            if ((int32_t)(q0 | q1 | q2 | q3) < 0) {                     // some combo of this (anchor, chunk) is still possible
                                const unsigned idx = atomicAdd(&sm->rec_used[e], 1u);
                                const unsigned long long slot = w_base + idx;
                                if (slot < cand_cap) cand[slot] = (unsigned long long)u | ((unsigned long long)blk << 40) | ((unsigned long long)(c0 >> 5) << 56);
                                else atomicAdd(&counters[3], 1ull);                     // list full: the host grows it and re-runs
                            }
                        }
                        asm volatile("tcgen05.fence::before_thread_sync;");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&sm->empty_acc[a]);
                        // off the critical path: a warp adds at most 64 records per phase tile, so refilling at 128 keeps idx < 256
                        if (sm->rec_used[e] > (unsigned)(kWarpChunk / 2)) {
                            unsigned long long nb = 0;
                            if (lane == 0) { nb = atomicAdd(&counters[0], (unsigned long long)kWarpChunk); sm->rec_used[e] = 0; sm->rec_base[e] = nb; }
                            w_base = __shfl_sync(0xFFFFFFFFu, nb, 0);
                        }
                        __syncwarp();
                    }
                    if (!ok) break;
                }
            }
        }
        // every role is done with this block: all MMAs completed (the epilogue consumed them), shared memory reusable
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
        if (*abort_flag) break;
    }
    if (*abort_flag && tid == 0) atomicOr(&counters[2], 2ull);
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

}  // namespace

namespace pwmscan {

size_t mma_prefilter_smem_bytes() { return ((sizeof(MmaSmem) + 127) & ~size_t(127)) + (size_t)kMmaMaxSteps * kMmaCombos * 32 + 128; }
static_assert(kMmaCombos * kAccStages <= 512, "accumulator stages must fit the 512 TMEM columns");

int launch_mma_prefilter(pwmscan_ctx* ctx, ScanLane* lane, const uint32_t* d_seq2, long long a0, long long a1, int nblocks, const DevMmaBlock* d_blocks,
                         const int8_t* d_btiles, unsigned long long* d_cand, unsigned long long cand_cap) {
    if (nblocks == 0 || a1 <= a0) return PWMSCAN_OK;
    if (!ctx->mma_attr_set) {
        PWM_CUDA(cudaFuncSetAttribute(mma_prefilter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mma_prefilter_smem_bytes()));
        ctx->mma_attr_set = true;
    }
    const long long n_items = (a1 - a0) / ((long long)kItemTiles * kTile);
    const int grid = (int)std::min<long long>(ctx->sm_count, n_items);
    static const int dbg = std::getenv("PWMSCAN_MMA_DBG") ? std::atoi(std::getenv("PWMSCAN_MMA_DBG")) : 0;
    mma_prefilter_kernel<<<grid, kMmaThreads, mma_prefilter_smem_bytes(), lane->stream>>>(d_seq2, a0, a1 - a0, nblocks, d_blocks, d_btiles,
                                                                                        d_cand, cand_cap, lane->d_counters, dbg);
    PWM_CUDA(cudaGetLastError());
    return PWMSCAN_OK;
}

}  // namespace pwmscan
