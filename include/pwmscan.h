/* pwmscan.h — C ABI of libpwmscan.so, the B200 (sm_100a) implementation of the
 * position-weight-matrix motif-scan path of kaizhang/bioinformatics-toolkit.
 *
 * The reference is pure Haskell and has no FFI; the boundary it exposes for this path is a set of
 * exported Haskell functions.  Each entry point below names the reference function(s) it replaces
 * (paths relative to bioinformatics-toolkit/src/Bio/).  INTEGRATION.md shows the
 * `foreign import ccall` stubs a maintainer adds to keep those Haskell signatures unchanged.
 *
 * Conventions
 *  - every call returns 0 on success or a negative PWMSCAN_E* code; the message is available from
 *    pwmscan_last_error(ctx) (ctx may be NULL for failures of pwmscan_ctx_create);
 *  - handles are opaque; inputs are borrowed for the duration of the call only; outputs are
 *    library-owned buffers released by the matching *_free;
 *  - one ctx per device; calls on one ctx are serialised by an internal mutex; different ctxs may
 *    be used concurrently from different OS threads (import as `ccall safe`);
 *  - results are deterministic: same inputs => bit-identical outputs, in the reference's order;
 *  - there is no CPU fallback: without a CUDA device every compute call fails with PWMSCAN_ECUDA.
 */
#ifndef PWMSCAN_H
#define PWMSCAN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PWMSCAN_ABI_VERSION 1

/* error codes */
#define PWMSCAN_OK        0
#define PWMSCAN_EINVAL   -1   /* bad argument */
#define PWMSCAN_ECUDA    -2   /* CUDA runtime failure (no device, OOM, launch failure) */
#define PWMSCAN_ENUCL    -3   /* invalid nucleotide met where the reference calls `error`
                                 (Motif/Search.hs:157, Motif.hs:185) */
#define PWMSCAN_ERANGE   -4   /* a size exceeds what the library supports (see limits below) */
#define PWMSCAN_EDOMAIN  -5   /* "p must be in [0,1]" / "cdf': impossible!" (Motif.hs:254,261) */

/* limits */
#define PWMSCAN_MAX_COLS      64      /* PWM columns in a motif set: scan, CDF, dense scores, pair lists */
#define PWMSCAN_ALNSET_MAX_COLS 256   /* PWM columns in a slot of an alignment set: merged PWMs grow (iterativeMerge) */
#define PWMSCAN_MAX_MOTIFS    32768   /* motifs per pwmscan_motifs */
#define PWMSCAN_MAX_SEGMENTS  65536   /* sequences per pwmscan_seq */

typedef struct pwmscan_ctx    pwmscan_ctx;
typedef struct pwmscan_seq    pwmscan_seq;
typedef struct pwmscan_motifs pwmscan_motifs;
typedef struct pwmscan_plan   pwmscan_plan;
typedef struct pwmscan_hits   pwmscan_hits;

int pwmscan_abi_version(void);

/* ---- context ------------------------------------------------------------------------------ */
int  pwmscan_ctx_create(int device, pwmscan_ctx** out);
void pwmscan_ctx_destroy(pwmscan_ctx* ctx);
const char* pwmscan_last_error(pwmscan_ctx* ctx);
/* Device time (ms, CUDA events on the scan's stream) of the phases of the last scan call (summed over the items of
 * a pwmscan_scan_many): out[0] first-stage (prefilter) kernels, out[1] exact re-score kernel, out[2] ordering,
 * out[3] device span start .. ordered, out[4] host->device bytes, out[5] device->host bytes, out[6] prefilter
 * candidates, out[7] prefilter kernel launches, out[8] kernel launches in all, out[9] hits.  The other entry points
 * (scoreCDF, alignment, dense scores, per-hit affinity) report their kernel time in out[10] (and alignment its number
 * of host re-evaluations in out[11]) and leave the scan slots alone.  n = number of doubles available in out (<= 12). */
int  pwmscan_last_timing(pwmscan_ctx* ctx, double* out, int n);

/* ---- sequences: Bio.Seq `DNA a` (Seq.hs:41,59-67,86-95) ------------------------------------- */
#define PWMSCAN_SEQ_UPPERCASE  1   /* apply toUpper first, as fromBS does (Seq.hs:86-95) */
#define PWMSCAN_SEQ_KEEP_ASCII 2   /* keep the (upper-cased) bytes on the device: needed by the
                                      IUPAC-aware entry points (skip_ambiguous = 0, scores, maxMatchSc) */
/* One sequence (a `DNA a` ByteString via unsafeUseAsCStringLen).  Packs 2 bits/base + an
 * invalid-base bit plane in HBM. */
int  pwmscan_seq_upload(pwmscan_ctx* ctx, const uint8_t* ascii, int64_t len, int flags, pwmscan_seq** out);
/* nseg sequences laid end to end in `ascii`; seg_off has nseg+1 entries (seg_off[0] = 0).
 * Windows never span two segments.  Used by scanMotif to batch BED regions. */
int  pwmscan_seq_upload_segments(pwmscan_ctx* ctx, const uint8_t* ascii, const int64_t* seg_off,
                                 int32_t nseg, int flags, pwmscan_seq** out);
int64_t pwmscan_seq_length(const pwmscan_seq* seq);
/* fromBS validation (Seq.hs:86-95): position of the first byte (after optional upper-casing)
 * outside "ACGT" (alphabet 0, `DNA Basic`) or "ACGTNVHDBMKWSYR" (alphabet 1, `DNA IUPAC`); -1 if none. */
int  pwmscan_seq_first_invalid(pwmscan_ctx* ctx, const pwmscan_seq* seq, int alphabet, int64_t* pos_out);
void pwmscan_seq_free(pwmscan_seq* seq);

/* ---- motifs: PWM / Bkgd (Motif.hs:59-62,101-104), rcPWM (77-78), optimalScoresSuffix
 *      (Motif/Search.hs:111-124) -------------------------------------------------------------- */
/* mats: the M matrices row-major (ncol[m] rows x 4 columns A,C,G,T), concatenated.  The library
 * derives the reverse-complement PWM, both fp64 log-odds tables (host libm `log`, pseudocount
 * 0.0001: Search.hs:185-196) and both sigma vectors. */
int  pwmscan_motifs_create(pwmscan_ctx* ctx, int32_t M, const int32_t* ncol, const double* mats,
                           const double bg[4], pwmscan_motifs** out);
/* findTFBSWith (Search.hs:38-58) lets the caller pass its own sigma; override motif m's vector
 * for strand 0 (the PWM) or 1 (its rcPWM). */
int  pwmscan_motifs_set_sigma(pwmscan_motifs* motifs, int32_t m, int strand, const double* sigma);
int  pwmscan_motifs_get_sigma(const pwmscan_motifs* motifs, int32_t m, int strand, double* sigma_out);
int32_t pwmscan_motifs_count(const pwmscan_motifs* motifs);
void pwmscan_motifs_free(pwmscan_motifs* motifs);

/* ---- thresholded scan: findTFBS / findTFBSWith (Search.hs:25-58), the per-region body of
 *      scanMotif (Data/Bed/Utils.hs:112-118) ---------------------------------------------------- */
/* A plan is the device-side analogue of CutoffMotif (Data/Bed/Utils.hs:78-98): everything that
 * depends on (motifs, cutoffs) and not on the sequence.  thres[m] is the cutoff of motif m and is
 * used for both strands (Utils.hs:91-98).  strands: 1 = the PWMs only (findTFBS), 2 = PWMs and
 * their rcPWMs (scanMotif).  skip_ambiguous: the reference's `skip` flag — 1 = lookAheadSearch'
 * (non-ACGT => -inf), 0 = IUPAC-aware lookAheadSearch (needs PWMSCAN_SEQ_KEEP_ASCII). */
int  pwmscan_plan_create(pwmscan_ctx* ctx, const pwmscan_motifs* motifs, const double* thres,
                         int strands, int skip_ambiguous, pwmscan_plan** out);
void pwmscan_plan_free(pwmscan_plan* plan);
/* Plan facts for the roofline arithmetic: out[7] the first stage the plan was built for (2 = k-mer
 * mask tables in L2, the default; 0 = shared-memory 4-mer tables; 1 = tensor-core variant),
 * out[0] prefilter blocks, out[2] (motif,strand) combos on the exact kernel, out[3] expected
 * candidates per position, out[4] table bytes, out[5] passes over the packed genome;
 * mode 2: out[1] expected 32-byte table sectors gathered per sequence position (summed over blocks),
 *         out[6] expected first-stage survivors per position;
 * mode 0: out[1] first-stage shared-memory wavefronts (128 B) per position, out[6] sum over blocks
 *         of P(a warp enters the second stage). */
int  pwmscan_plan_info(const pwmscan_plan* plan, double* out, int n);
/* Hit list sorted by (segment, motif, strand, position); position is the 0-based window start
 * within its segment on the forward sequence (also for strand 1: Utils.hs:109-111); score is the
 * reference's fp64 accumulator, bit-identical. */
int  pwmscan_scan(pwmscan_ctx* ctx, const pwmscan_seq* seq, const pwmscan_plan* plan, pwmscan_hits** out);
/* One-shot scan of HOST bytes (what `findTFBS bg pwm dna thres` / one scanMotif batch call): the
 * ASCII goes to the device in chunks and each chunk is packed and scanned while the next one is in
 * flight, so the PCIe transfer overlaps the scan.  seg_off/nseg/flags as pwmscan_seq_upload_segments
 * (pass {0,len}, 1 for a single sequence).  Same result as seq_upload + scan. */
int  pwmscan_scan_host(pwmscan_ctx* ctx, const pwmscan_plan* plan, const uint8_t* ascii, const int64_t* seg_off,
                       int32_t nseg, int flags, pwmscan_hits** out);
/* fromBS verdict (Seq.hs:86-95) of the bytes the last pwmscan_scan_host on this context consumed: offset in
 * the concatenated input of the first byte outside the alphabet (0 = ACGT, 1 = IUPAC; after upper-casing
 * if PWMSCAN_SEQ_UPPERCASE was given), -1 if there is none. */
int  pwmscan_scan_host_first_invalid(pwmscan_ctx* ctx, int alphabet, int64_t* pos_out);
/* Many scans of one plan — the (motif block x chromosome chunk) items of a whole-genome scan (the decomposition of
 * the reference's findTFBS', Motif/Search.hs:60-84) — with several of them in flight on the device: while the first
 * stage of one item runs, the exact replay, the ordering and the hit-list copy of the previous items proceed.  Hit
 * lists are handed to `cb` in item order, on the calling thread; the callee owns them (pwmscan_hits_free).  Same
 * results as one pwmscan_scan / pwmscan_scan_host per item.  pwmscan_scan_host_many takes n single sequences
 * (ascii[i], lens[i]) in host memory (pinned for full PCIe speed). */
typedef void (*pwmscan_hits_cb)(void* user, int32_t item, pwmscan_hits* hits);
int  pwmscan_scan_many(pwmscan_ctx* ctx, const pwmscan_plan* plan, int32_t n, const pwmscan_seq* const* seqs,
                       pwmscan_hits_cb cb, void* user);
int  pwmscan_scan_host_many(pwmscan_ctx* ctx, const pwmscan_plan* plan, int32_t n, const uint8_t* const* ascii,
                            const int64_t* lens, int flags, pwmscan_hits_cb cb, void* user);
/* ---- the packed genome store ------------------------------------------------------------------------------------
 * A genome that is scanned again and again need not cross PCIe as one byte per base each time: next to the bytes that
 * mkIndex writes (Seq/IO.hs:87-106) the index can keep the two planes the device scans — 2 bits per base (16 bases per
 * uint32, base t at bits 2t..2t+1; A C G T = 0 1 2 3) and 1 bit per base "not A/C/G/T" — 0.375 bytes per base, made ONCE by
 * pwmscan_pack_host (host, multi-threaded; same codes as the device packer).  Unit: 32 bases = 2 words + 1 word;
 * units = pwmscan_packed_units(len) = ceil(len / 32); seq2_out holds 2 * units words, mask_out units words.  A slice
 * that starts at a multiple of 32 bases is a valid packed sequence of its own (seq2 + start / 16, mask + start / 32) —
 * what the (motif block x chromosome chunk) items of a genome scan pass to pwmscan_scan_packed_many.
 * first_non_acgt / first_non_iupac: fromBS's verdict (Seq.hs:86-95) of the bytes, -1 if all are inside the alphabet;
 * flags: PWMSCAN_SEQ_UPPERCASE.  nthreads <= 0: all host cores. */
int64_t pwmscan_packed_units(int64_t len);
int  pwmscan_pack_host(const uint8_t* ascii, int64_t len, int flags, uint32_t* seq2_out, uint32_t* mask_out,
                       int64_t* first_non_acgt, int64_t* first_non_iupac, int nthreads);
/* getSeq / getChrom from the packed store -> a resident sequence (hits identical to pwmscan_seq_upload of the bytes;
 * pwmscan_seq_first_invalid answers the `Basic` alphabet from the mask plane and -1 for IUPAC: an index whose bytes are
 * not IUPAC was rejected when it was packed). */
int  pwmscan_seq_upload_packed(pwmscan_ctx* ctx, const uint32_t* seq2, const uint32_t* mask, int64_t len, pwmscan_seq** out);
/* pwmscan_scan_host_many for items given as packed planes (pinned host memory for full PCIe speed); plans made with
 * skip_ambiguous = 0 need the bytes and are refused. */
int  pwmscan_scan_packed_many(pwmscan_ctx* ctx, const pwmscan_plan* plan, int32_t n, const uint32_t* const* seq2,
                              const uint32_t* const* mask, const int64_t* lens, pwmscan_hits_cb cb, void* user);
/* Convenience: plan_create + scan + plan_free. */
int  pwmscan_find_tfbs(pwmscan_ctx* ctx, const pwmscan_seq* seq, const pwmscan_motifs* motifs,
                       const double* thres, int strands, int skip_ambiguous, pwmscan_hits** out);
int64_t pwmscan_hits_count(const pwmscan_hits* hits);
/* fromBS verdict (Seq.hs:86-95) of the bytes that were uploaded for THIS hit list by pwmscan_scan_host /
 * pwmscan_scan_host_many (per item, unlike pwmscan_scan_host_first_invalid): offset of the first byte outside the
 * alphabet (0 = ACGT, 1 = IUPAC), -1 if there is none — also -1 for scans of resident sequences
 * (ask pwmscan_seq_first_invalid). */
int  pwmscan_hits_first_invalid(const pwmscan_hits* hits, int alphabet, int64_t* pos_out);
/* What the device transfers — the compact columns, in the order above: score[n], pos32[n] (window start within
 * its segment) and
 *   - single-segment scans: run_off[2M+1]; the hits of (motif m, strand s) are [run_off[2m+s], run_off[2m+s+1])
 *     (exactly the (Int, Double) stream `findTFBS` yields for that PWM), segment = motif_strand = NULL: 12 bytes/hit;
 *   - several segments: segment[n] and motif_strand[n] = 2 * motif + strand, run_off = NULL: 18 bytes/hit.
 * Read-only pinned host memory, valid until pwmscan_hits_free; any pointer argument may be NULL. */
int  pwmscan_hits_compact(const pwmscan_hits* hits, const uint32_t** pos32, const double** score,
                          const int64_t** run_off, const int32_t** segment, const uint16_t** motif_strand);
/* Wide column views (read-only host memory, valid until pwmscan_hits_free / pwmscan_ctx_destroy); any pointer
 * argument may be NULL.  motif / strand / segment / pos are expanded from the compact columns on the host the
 * first time one of them is asked for. */
int  pwmscan_hits_columns(const pwmscan_hits* hits, const int32_t** motif, const uint8_t** strand,
                          const int32_t** segment, const int64_t** pos, const double** score);
void pwmscan_hits_free(pwmscan_hits* hits);

/* ---- dense scoring: scores / scores' / score (Motif.hs:127-149), maxMatchSc (Search.hs:98-109) */
/* out[i] = scoreHelp of the window at i, i in [0, len-n]; IUPAC-aware ('N' scores 0); needs
 * PWMSCAN_SEQ_KEEP_ASCII and a single-segment seq.  PWMSCAN_ENUCL on any other byte. */
int  pwmscan_scores(pwmscan_ctx* ctx, const pwmscan_seq* seq, const pwmscan_motifs* motifs, int32_t m,
                    int strand, double* out, int64_t out_len);
/* out[m] = maximum window score of motif m (strand 0) over the sequence; -inf if len < n. */
int  pwmscan_max_match_sc(pwmscan_ctx* ctx, const pwmscan_seq* seq, const pwmscan_motifs* motifs, double* out);

/* ---- score distribution: scoreCDF (Motif.hs:282-326), cdf' (252-270), pValueToScore (213-214) */
/* Compressed CDF points of motif m (strand 0) under the motifs' background; *x and *P are
 * library-owned (release with pwmscan_free). */
int  pwmscan_score_cdf(pwmscan_ctx* ctx, const pwmscan_motifs* motifs, int32_t m, double** x, double** P, int64_t* npts);
/* thres_out[m] = cdf' (scoreCDF bg pwm_m) (1 - p) for every motif. */
int  pwmscan_pvalue_to_score(pwmscan_ctx* ctx, const pwmscan_motifs* motifs, double p, double* thres_out);
/* mkCutoffMotif for every motif (Data/Bed/Utils.hs:88-98): thres_out[m] = cdf' cdf_m (1 - p) on the
 * full CDF, and the points of truncateCDF (1 - 10 p) cdf_m concatenated in *x / *P (library-owned,
 * release with pwmscan_free); motif m owns [offsets[m], offsets[m+1]).  offsets has M+1 entries. */
int  pwmscan_cutoff_cdfs(pwmscan_ctx* ctx, const pwmscan_motifs* motifs, double p, double* thres_out,
                         double** x, double** P, int64_t* offsets);
void pwmscan_free(void* p);

/* ---- PWM-PWM alignment: alignmentBy jsd pFn combFn (Motif/Alignment.hs:83-132), the all-pairs
 *      initialisation and row refresh of iterativeMerge (Motif/Merge.hs:138-145,161-165) -------- */
#define PWMSCAN_PENAL_LINEAR 0
#define PWMSCAN_PENAL_QUAD   1
#define PWMSCAN_PENAL_CUBIC  2
#define PWMSCAN_PENAL_EXP    3
#define PWMSCAN_COMB_L1   0
#define PWMSCAN_COMB_L2   1
#define PWMSCAN_COMB_L3   2
#define PWMSCAN_COMB_LINF 3
/* All pairs i<j, packed upper triangle in row-major order (pair (i,j) at
 * i*M - i*(i+1)/2 + (j-i-1)): dist, same_dir (1 = True) and pos as alignmentBy returns them for
 * `align pwm_i pwm_j`. */
int  pwmscan_align_all_pairs(pwmscan_ctx* ctx, const pwmscan_motifs* motifs, int penalty_kind, double gap,
                             int combine_kind, double* dist, uint8_t* same_dir, int32_t* pos);
/* Explicit pair list: result k is `align pwm_{a[k]} pwm_{b[k]}` (argument order preserved —
 * Merge.hs:142-143 relies on it). */
int  pwmscan_align_pairs(pwmscan_ctx* ctx, const pwmscan_motifs* motifs, int64_t npairs, const int32_t* a,
                         const int32_t* b, int penalty_kind, double gap, int combine_kind,
                         double* dist, uint8_t* same_dir, int32_t* pos);

/* ---- iterativeMerge with the PWMs resident on the device (Motif/Merge.hs:113-170): create the set once (all-pairs
 *      initialisation = pwmscan_alnset_pairs with a = b = NULL, packed upper triangle as pwmscan_align_all_pairs), then
 *      per merge replace the merged PWM in its slot (pwmscan_alnset_update) and refresh its row with a pair list
 *      (result k = `align pwm_{a[k]} pwm_{b[k]}`, argument order preserved, Merge.hs:138-145).  Same results as
 *      pwmscan_align_pairs on a freshly built motif set; nothing else is re-uploaded or recomputed. ---------------- */
typedef struct pwmscan_alnset pwmscan_alnset;
int  pwmscan_alnset_create(pwmscan_ctx* ctx, int32_t M, const int32_t* ncol, const double* mats_rowmajor, int penalty_kind,
                           double gap, int combine_kind, pwmscan_alnset** out);
int  pwmscan_alnset_update(pwmscan_alnset* set, int32_t slot, int32_t ncol, const double* mat_rowmajor);
int  pwmscan_alnset_pairs(pwmscan_alnset* set, int64_t npairs, const int32_t* a, const int32_t* b, double* dist,
                          uint8_t* same_dir, int32_t* pos);
/* The matrix of iterativeMerge itself resident on the device (Merge.hs:121-165).  pwmscan_alnset_matrix_init aligns all
 * i < j (161-165) and returns the first minimum in the reference's scan order (i ascending, j ascending, strict <:
 * 149-158) as (*i, *j, *d, *same_dir, *pos); *i = -1 when no pair is left.  After the caller has merged slot j into slot
 * i (mergePWMWeighted stays host code), pwmscan_alnset_merge_step installs the merged PWM in slot i, blanks row and
 * column j (135), re-aligns slot i with every live slot (138-145, argument order by index) and returns the next minimum:
 * one call and 40 bytes back per merge.  A merged PWM longer than the slots makes the set grow (up to 256 columns). */
int  pwmscan_alnset_matrix_init(pwmscan_alnset* set, int32_t* i, int32_t* j, double* d, uint8_t* same_dir, int32_t* pos);
int  pwmscan_alnset_merge_step(pwmscan_alnset* set, int32_t i, int32_t j, int32_t ncol, const double* mat_rowmajor,
                               int32_t* next_i, int32_t* next_j, double* d, uint8_t* same_dir, int32_t* pos);
void pwmscan_alnset_free(pwmscan_alnset* set);

/* ---- scanMotif's per-hit columns (Data/Bed/Utils.hs:109-123), computed on the device for the hits of a scan:
 *      p = 1 - cdf score on the motif's (truncated) CDF (Motif.hs:236-249; bit-identical fp64) and
 *      the BED score toAffinity p = round (1000 / (1 + exp (-(-log10 (max 1e-20 p) - 5)))) (round half even;
 *      values within 1e-6 of a rounding tie are settled on the host with the C library's log/exp).
 *      cdf_off[M+1] indexes cdf_x / cdf_y, the (score, cumulative probability) points of every motif
 *      (pwmscan_cutoff_cdfs' output).  out_pvalue and/or out_affinity may be NULL; hits_count entries. -------- */
int  pwmscan_hits_affinity(pwmscan_ctx* ctx, const pwmscan_hits* hits, int32_t M, const int64_t* cdf_off,
                           const double* cdf_x, const double* cdf_y, double* out_pvalue, int64_t* out_affinity);

/* ---- host-side helper: BED6 text of scanMotif hits (toLine, Data/Bed/Types.hs:139-149, of mkBed,
 *      Data/Bed/Utils.hs:109-111).  Names are concatenated bytes with count+1 offsets; region r has
 *      chromosome region_chrom[r] and start region_start[r]; one '\n'-terminated line per hit in *text
 *      (release with pwmscan_free). ------------------------------------------------------------- */
int  pwmscan_format_bed6(int64_t nhits, const int32_t* segment, const int32_t* motif, const uint8_t* strand,
                         const int64_t* pos, const int64_t* score, const int32_t* region_chrom,
                         const int64_t* region_start, const char* chrom_names, const int64_t* chrom_off,
                         const char* motif_names, const int64_t* motif_off, const int32_t* motif_len,
                         char** text, int64_t* text_len);
/* `merge_motifs dist` text (app/MergeMotifs.hs:158-169): one line "name_a<TAB>name_b<TAB>toShortest d" per pair (host code,
 * multi-threaded).  a == b == NULL: all i < j of the M names in pwmscan_align_all_pairs' packed order; otherwise npairs
 * explicit pairs.  names: concatenated bytes with M+1 offsets; *text is released with pwmscan_free. */
int  pwmscan_format_dist(int64_t npairs, int32_t M, const int32_t* a, const int32_t* b, const double* dist, const char* names,
                         const int64_t* name_off, char** text, int64_t* text_len);

#ifdef __cplusplus
}
#endif
#endif /* PWMSCAN_H */
