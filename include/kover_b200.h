/* kover_b200.h -- C ABI of libkoverb200.so: Kover's rule-scoring hot path on B200 (sm_100a).
 *
 * This is the native boundary a Kover maintainer binds (ctypes stub: INTEGRATION.md).  It sits
 * where the reference's Cython module `kover.learning.common.popcount` sits today, but one level
 * up -- at `KmerRuleClassifications.sum_rows` granularity -- because a call per 800 KB host block
 * (reference rules.py:247-262) cannot feed a GPU.  Each entry point cites the reference interface
 * it replaces (paths relative to /root/reference/core/kover).
 *
 * Conventions
 *   - plain pointers and sizes only; every array is caller-owned, contiguous, host memory and is
 *     not retained after the call returns;
 *   - every function returns 0 on success, a negative KV_E_* code otherwise; the message is in
 *     kv_last_error() (thread-local);
 *   - one kv_shard = one contiguous k-mer-column slice of the packed matrix resident in the HBM of
 *     one GPU.  A single-GPU job has one shard covering all columns; a multi-GPU job has one per
 *     rank (one process per GPU) or one per device (single process);
 *   - calls on a shard are synchronous at the ABI and must come from one host thread at a time;
 *   - no CPU fallback: without a usable CUDA device kv_create() fails.
 *
 * Packed layout (the HDF5 `kmer_matrix` contract, utils.py:133-156, create.py:224-230):
 *   uint64 [W = ceil(n_examples/64), K] row-major; example i is bit 63-(i%64) of word i/64 of its
 *   column; padding bits of the last word are zero.  Row masks use the same bit order
 *   (rules.py:210-222).  Rule index r in [0,K) = presence of k-mer r, r in [K,2K) = absence of
 *   k-mer r-K (rules.py:57-79).
 */
#ifndef KOVER_B200_H
#define KOVER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KV_OK 0
#define KV_E_INVALID (-1)   /* bad argument (Python shim raises ValueError)            */
#define KV_E_CUDA (-2)      /* CUDA runtime error / no device (RuntimeError)             */
#define KV_E_NOMEM (-3)     /* device or pinned allocation failed (MemoryError)          */
#define KV_E_OVERFLOW (-4)  /* caller's tie buffer too small; *n_found holds the need    */
#define KV_E_INDEX (-5)     /* example index outside [0, n_examples) (IndexError)        */

typedef struct kv_shard kv_shard;

typedef struct kv_shard_info {
    int32_t device;          /* CUDA ordinal                                             */
    int32_t sm_count;        /* multiprocessors (148 on B200)                            */
    int64_t n_examples;      /* genomes                                                  */
    int64_t n_words;         /* W = ceil(n_examples / 64)                                */
    int64_t n_kmers_total;   /* K of the whole dataset                                   */
    int64_t col0;            /* first global column held by this shard                   */
    int64_t n_cols;          /* columns held                                             */
    int64_t ld;              /* device row pitch in 64-bit words                         */
    int64_t matrix_bytes;    /* bytes of HBM behind the packed image                     */
    double last_kernel_ms;   /* CUDA-event time of the scan kernel of the last scan call */
    int64_t last_launches;   /* kernels launched by the last call                        */
    int64_t total_launches;  /* kernels launched since kv_create                         */
    double scan_ms_total;    /* sum of the scan kernels' CUDA-event times since kv_create */
    int64_t scan_calls;      /* SCM matrix scans since kv_create                         */
    int64_t rescore_calls;   /* SCM scans served from resident counts (kv_scm_reuse_counts) */
    int64_t exch_steps;      /* multi-rank fused steps (kv_scm_best_rules_p2p) since kv_create          */
    double exch_publish_us;  /* summed over those steps, device clock of the rank's last CTA: packet stores, */
    double exch_wait_us;     /* ... waiting for every peer's packet (rank skew + NVLink latency),          */
    double exch_merge_us;    /* ... merging the packets and publishing the result to the host              */
    int64_t stream_steps;    /* fused SCM steps that streamed their candidates (no count arrays written)   */
    int64_t stream_redos;    /* ... of which were redone with the count arrays (candidate / tie overflow)   */
    int64_t level_nodes_counted;  /* kv_gini_level: nodes whose class masks went through a matrix pass ...          */
    int64_t level_nodes_derived;  /* ... and nodes whose counts were parent - sibling (kv_gini_level_family)        */
} kv_shard_info;

const char *kv_last_error(void);
int kv_abi_version(void);
/* Page-locked host memory for caller-owned result arrays (optional: every entry point accepts pageable pointers;
 * copies into pinned memory run at the PCIe rate instead of the 4-6 GB/s of staged pageable copies). */
int kv_host_alloc(int64_t bytes, void **out);
int kv_host_free(void *p);
int kv_device_count(int *n_devices);

/* ---- lifecycle ---------------------------------------------------------------------------
 * Replaces KmerRuleClassifications.__init__ binding an h5py dataset (rules.py:104-133): the
 * shard owns a zero-initialised [W, n_cols] packed image in HBM. */
int kv_create(int device, int64_t n_examples, int64_t n_kmers_total, int64_t col0, int64_t n_cols,
              kv_shard **out);
int kv_destroy(kv_shard *s);
int kv_get_info(kv_shard *s, kv_shard_info *out);

/* ---- upload ------------------------------------------------------------------------------
 * Replaces the per-call h5py chunk reads `dataset[rows, c0:c1]` (rules.py:253): every HDF5 chunk
 * (1 x <=100000, create.py:230) is pushed once.  `host` is a [n_rows, n_cols] block with row pitch
 * `ld_elems` elements of the packed matrix starting at packed row `row0` and GLOBAL column
 * `gcol0`; columns outside the shard are skipped.  pack_bits = 64 or 32 (rules.py:123-128); 32-bit
 * rows 2w, 2w+1 become the high / low halves of 64-bit row w. */
int kv_upload_block(kv_shard *s, int pack_bits, int64_t row0, int64_t n_rows, int64_t gcol0,
                    int64_t n_cols, const void *host, int64_t ld_elems);
/* The same for gzip-filtered chunks WITHOUT inflating them on the host: the reference's `dataset[rows, c0:c1]`
 * (rules.py:253) makes libhdf5 inflate every chunk (create.py:222-230: chunks (1, <=100000), gzip) on a host core
 * on every call; here the raw zlib streams (RFC 1950) travel to the GPU and one warp inflates one chunk (stored,
 * fixed and dynamic Huffman blocks; header and Adler-32 checked).  src[i] / src_bytes[i]: the i-th chunk's stream
 * as stored in the file; (row0[i], gcol0[i]): dataset coordinates of its first element, rows in pack_bits units;
 * every chunk inflates to chunk_rows x chunk_cols elements (HDF5 stores edge chunks at full size); parts outside the
 * dataset or outside this shard's columns are dropped.  status[i] = 0, or a non-zero code when the stream is
 * malformed: the caller then inflates that chunk on the host (its error-reporting path) and uses kv_upload_block. */
int kv_upload_deflate(kv_shard *s, int pack_bits, int64_t n_chunks, const void *const *src,
                      const int64_t *src_bytes, const int64_t *row0, const int64_t *gcol0,
                      int64_t chunk_rows, int64_t chunk_cols, int32_t *status);
/* Host-clock phases of the last kv_upload_deflate on this shard (measurement aid): buffer set-up, gathering the
 * streams into page-locked memory (overlaps the copies and the kernels), whole call. */
int kv_upload_deflate_stats(kv_shard *s, double *setup_ms, double *gather_ms, double *total_ms);
/* Fill the shard with the deterministic synthetic matrix of kover_b200/synthetic.py (bench and
 * full-size parity tests; the CPU twin regenerates any column slice). */
int kv_generate_synthetic(kv_shard *s, uint64_t seed);

/* Read back a [n_rows, n_cols] block of 64-bit packed rows starting at packed row `row0` and LOCAL
 * column `lcol0` (upload round-trip checks; bench staging). */
int kv_read_block(kv_shard *s, int64_t row0, int64_t n_rows, int64_t lcol0, int64_t n_cols, uint64_t *host,
                  int64_t ld_elems);

/* ---- timing ------------------------------------------------------------------------------
 * CUDA events on the shard's own stream (torch.cuda.Event cannot see it): start, run any number of
 * calls, stop -> device milliseconds between the two marks. */
int kv_timer_start(kv_shard *s);
int kv_timer_stop(kv_shard *s, double *elapsed_ms);

/* ---- a1: popcount.pyx:52-67, 97-112 -------------------------------------------------------
 * arr[i,j] <- popcount(arr[i,j] & row_mask[i]) for arr[i,j] != 0, in place, [m,n] C-contiguous
 * host block.  Stateless (uploads, runs, downloads) -- kept for drop-in completeness. */
int kv_inplace_popcount_64(int device, uint64_t *arr, int64_t m, int64_t n, const uint64_t *row_mask);
int kv_inplace_popcount_32(int device, uint32_t *arr, int64_t m, int64_t n, const uint32_t *row_mask);

/* ---- a3: KmerRuleClassifications.sum_rows (rules.py:201-267) ------------------------------
 * mask: W words built as rules.py:210-222.  out[n_cols] = per local column, the number of masked
 * examples having the k-mer, stored with out_itemsize = 1, 2, 4 or 8 bytes (the reference's
 * _minimum_uint_size dtype, utils.py:117-130).  The absence half is len(rows) - out (rules.py:265)
 * and is left to the caller. */
int kv_sum_rows(kv_shard *s, const uint64_t *mask, void *out, int out_itemsize);
/* Both halves of the reference's return value: out_presence[n_cols] as kv_sum_rows, out_absence[n_cols] =
 * n_rows - presence evaluated in the output dtype (rules.py:265). */
int kv_sum_rows_full(kv_shard *s, const uint64_t *mask, int64_t n_rows, void *out_presence, void *out_absence,
                     int out_itemsize);
/* Several masks in one pass over the matrix (the reference's own TODO, rules.py:200).
 * masks: [n_masks, W]; out: [n_masks, n_cols] uint32. */
int kv_sum_rows_multi(kv_shard *s, int n_masks, const uint64_t *masks, uint32_t *out);

/* ---- dataset split: per-k-mer individual risks (dataset/split.py:171-188, 213-228) ---------
 * A k-mer's risk round(((n_pos - pos_present) + neg_present) / n_total, 5) depends only on its integer
 * error count e in [0, n_total].  kv_risk_errors runs ONE two-mask scan, leaves e per k-mer on the device
 * and ORs into present[n_total + 1] which values of e occur.  The caller evaluates the reference's float
 * formula on those values, np.unique's them, and hands back two tables e -> index into unique_risks
 * (presence rule, absence rule); kv_risk_map applies them: out_kmer / out_anti [n_cols] of out_itemsize bytes
 * (split.py:186-187 stores them with _minimum_uint_size(len(unique_risks))). */
int kv_risk_errors(kv_shard *s, const uint64_t *pos_mask, const uint64_t *neg_mask, int64_t n_pos, int64_t n_total,
                   uint8_t *present);
int kv_risk_map(kv_shard *s, const uint32_t *lut_kmer, const uint32_t *lut_anti, int64_t n_lut, int out_itemsize,
                void *out_kmer, void *out_anti);

/* ---- a4: KmerRuleClassifications.get_columns (rules.py:135-171) ---------------------------
 * cols: n LOCAL presence-column indices.  out: [n, W] packed words (column j's W words are
 * contiguous).  Unpacking / inversion for absence rules is host work on W words. */
int kv_get_columns(kv_shard *s, const int64_t *cols, int64_t n, uint64_t *out);

/* ---- a6: SetCoveringMachine._get_best_utility_rules (scm.py:238-288) ----------------------
 * Blacklist: GLOBAL rule indices in [0, 2K); entries outside the shard are ignored; n = 0 clears
 * (scm.py:239-240, 267). */
int kv_set_blacklist(kv_shard *s, const int64_t *rule_idx, int64_t n);
/* Number of utility blocks of `util_block_size` rules over the 2K rule space (scm.py:262). */
int64_t kv_scm_n_blocks(int64_t n_kmers_total, int64_t util_block_size);
/* Phase 1: one pass over the shard.  Counts for both masks, utility
 * u = neg_cover - p * pos_err (two rounded fp64 ops, scm.py:263-264), blacklisted -> -inf, and the
 * maximum of every utility block this shard touches.  block_max[n_blocks] receives -inf for
 * untouched blocks, so shards merge with an element-wise max. */
int kv_scm_scan(kv_shard *s, const uint64_t *pos_mask, const uint64_t *neg_mask, int64_t n_pos,
                int64_t n_neg, double p, int64_t util_block_size, double *block_max, int64_t n_blocks);
/* Arms the NEXT kv_scm_scan / kv_scm_scan_packet / kv_scm_best_rules / kv_scm_best_rules_p2p call on this
 * shard to skip the matrix pass and re-score the counts the previous scan left on the device: same two
 * example masks, another p, and -- swap_roles = 1 -- positives and negatives exchanged relative to the
 * masks of the scan that produced the counts (the disjunction swaps the labels, scm.py:69-73).  The armed
 * call ignores its mask arguments (they may be NULL) and reads 8 bytes per k-mer instead of 8*W: every p of
 * the hyper-parameter grid and both model types share the first iteration's counts. */
int kv_scm_reuse_counts(kv_shard *s, int swap_roles);
/* Host-only replay of the sequential block merge (scm.py:270-286) over the MERGED block maxima:
 * selected[b] = 1 for blocks whose ties survive; *best_utility as the reference returns it. */
int kv_scm_select_blocks(const double *block_max, int64_t n_blocks, double *best_utility,
                         uint8_t *selected);
/* Phase 2: rules of the selected blocks with isclose(u, block_max[b]) (scm.py:273), GLOBAL rule
 * indices ascending, with their pos_err / neg_cover counts.  Uses the counts left on the device by
 * the last kv_scm_scan.  Returns KV_E_OVERFLOW (and the needed size in *n_found) when capacity is
 * too small. */
int kv_scm_ties(kv_shard *s, const double *block_max, const uint8_t *selected, int64_t n_blocks,
                int64_t capacity, int64_t *rule_idx, int64_t *pos_err, int64_t *neg_cover,
                int64_t *n_found);
/* Job queue (single shard): kv_scm_queue_push enqueues scan (reuse = 0) or re-score (reuse = 1: same masks as
 * the previous push, 2: roles swapped; masks may then be NULL) + block merge + tie collection for job `slot` of
 * `n_slots` WITHOUT synchronising; kv_scm_queue_collect waits once and returns, per job, best_utility[j],
 * n_found[j] and its first min(n_found[j], 256) ties at rule_idx / pos_err / neg_cover [j * 256 ...]
 * (ascending).  n_found[j] > 256 = truncated: run that job alone with kv_scm_best_rules.  This is how the
 * lock-step fits of a hyper-parameter grid keep the GPU busy between jobs (north-star item 4). */
int kv_scm_queue_push(kv_shard *s, int64_t slot, int64_t n_slots, const uint64_t *pos_mask, const uint64_t *neg_mask,
                      int64_t n_pos, int64_t n_neg, double p, int64_t util_block_size, int reuse);
int kv_scm_queue_collect(kv_shard *s, int64_t n_slots, double *best_utility, int64_t *n_found, int64_t *rule_idx,
                         int64_t *pos_err, int64_t *neg_cover);
/* Both in one call: masks [n_jobs][2][W] (positive mask, negative mask per job), per-job n_pos / n_neg / p / reuse. */
int kv_scm_queue_run(kv_shard *s, int64_t n_jobs, const uint64_t *masks, const int64_t *n_pos, const int64_t *n_neg,
                     const double *p, const int32_t *reuse, int64_t util_block_size, double *best_utility,
                     int64_t *n_found, int64_t *rule_idx, int64_t *pos_err, int64_t *neg_cover);
/* The same lock-step grid on a PARTIAL shard (several shards per process, or one process per GPU): distinct mask
 * pairs counted several per matrix pass, every job scored from the resident counts, each job ending as a packet
 * (format below) at packets + job * kv_scm_packet_bytes(n_blocks, capacity) -- host memory, or device memory of
 * the shard's GPU when packets_on_device.  The caller gathers the packets of all shards / ranks with ONE exchange
 * for the whole grid and merges them job by job with kv_scm_merge_packets; a job whose count exceeds `capacity`
 * on some shard is redone alone.  The first job must have reuse = 0. */
int kv_scm_queue_packets(kv_shard *s, int64_t n_jobs, const uint64_t *masks, const int64_t *n_pos, const int64_t *n_neg,
                         const double *p, const int32_t *reuse, int64_t util_block_size, int64_t capacity, void *packets,
                         int packets_on_device);
/* Multi-rank, one exchange per scan: phase 1 plus a LOCAL superset of the ties, written as one packet
 *   double block_max[n_blocks] | uint64 count | double threshold | {int64 rule, uint32 pos_err,
 *   uint32 neg_cover} records[capacity]
 * (records of blacklisted rules carry bit 62 of `rule`).  Every rule that can survive the merge over
 * ranks has u >= local_max - 5 * (1e-8 + 1e-5 * |local_max|), so the ranks all-gather their packets
 * once and finish on the host (max-merge, kv_scm_select_blocks, exact isclose filter).  count >
 * capacity means the packet is truncated: fall back to kv_scm_scan / kv_scm_ties.  `packet` is host
 * memory, or -- packet_on_device = 1, the one device pointer in this ABI -- memory on the shard's
 * GPU (e.g. the torch tensor handed to NCCL); the call returns after the stream has drained. */
int64_t kv_scm_packet_bytes(int64_t n_blocks, int64_t capacity);
int kv_scm_scan_packet(kv_shard *s, const uint64_t *pos_mask, const uint64_t *neg_mask, int64_t n_pos,
                       int64_t n_neg, double p, int64_t util_block_size, int64_t n_blocks, int64_t capacity,
                       void *packet, int packet_on_device);
/* Host-only merge of n_packets gathered packets [n_packets][kv_scm_packet_bytes(n_blocks, packet_capacity)]:
 * element-wise max of the block maxima (-> merged_block_max[n_blocks]), kv_scm_select_blocks (-> selected[n_blocks],
 * *best_utility) and the exact isclose filter (scm.py:273) over the candidates, ascending rule index.
 * *overflowed = 1 (and no ties) when some packet was truncated. */
int kv_scm_merge_packets(const void *rows, int64_t n_packets, int64_t n_blocks, int64_t packet_capacity, double p,
                         int64_t util_block_size, double *best_utility, int64_t capacity, int64_t *rule_idx,
                         int64_t *pos_err, int64_t *neg_cover, int64_t *n_found, double *merged_block_max,
                         uint8_t *selected, int *overflowed);
/* Peer mailbox (one process per GPU, all GPUs of one node): the packets are all-gathered by remote stores
 * over NVLink into every peer's mailbox from the same kernel that waits for the peers' packets, so a
 * multi-rank step costs one host synchronisation and no NCCL call.  kv_p2p_create returns this rank's
 * 64-byte CUDA IPC handle; exchange the handles out of band (torch.distributed all_gather) and pass all
 * `world` of them, rank order, to kv_p2p_connect.  Every rank must then call kv_scm_best_rules_p2p in the
 * same order; outputs as kv_scm_merge_packets. */
int kv_p2p_create(kv_shard *s, int rank, int world, int64_t slot_bytes, void *ipc_handle_out);
int kv_p2p_connect(kv_shard *s, const void *all_handles);
int kv_scm_best_rules_p2p(kv_shard *s, const uint64_t *pos_mask, const uint64_t *neg_mask, int64_t n_pos,
                          int64_t n_neg, double p, int64_t util_block_size, int64_t packet_capacity,
                          double *best_utility, int64_t capacity, int64_t *rule_idx, int64_t *pos_err,
                          int64_t *neg_cover, int64_t *n_found, double *merged_block_max, uint8_t *selected,
                          int *overflowed);
/* kv_scm_best_rules_p2p from example index arrays (masks built inside the call, see kv_scm_best_rules_idx). */
int kv_scm_best_rules_p2p_idx(kv_shard *s, const int64_t *pos_idx, int64_t n_pos, const int64_t *neg_idx,
                              int64_t n_neg, double p, int64_t util_block_size, int64_t packet_capacity,
                              double *best_utility, int64_t capacity, int64_t *rule_idx, int64_t *pos_err,
                              int64_t *neg_cover, int64_t *n_found, double *merged_block_max,
                              uint8_t *selected, int *overflowed);
/* Single-shard step, the whole of _get_best_utility_rules (scm.py:238-288) in one call and ONE kernel launch: the
 * scan kernel ends in a grid-wide barrier, replays the block merge, collects the ties of its own tiles into mapped
 * page-locked memory and raises an epoch flag the host polls (no copy, no extra launch; KOVER_B200_FUSED=0 restores
 * the scan + select + ties sequence).  More than 4 096 ties fall back to the tie kernel + device sort. */
int kv_scm_best_rules(kv_shard *s, const uint64_t *pos_mask, const uint64_t *neg_mask, int64_t n_pos,
                      int64_t n_neg, double p, int64_t util_block_size, double *best_utility,
                      int64_t capacity, int64_t *rule_idx, int64_t *pos_err, int64_t *neg_cover,
                      int64_t *n_found);
/* The same from the learner's example index arrays (scm.py:238-253 passes index lists to sum_rows; the masks of
 * rules.py:210-222 are built inside the call).  KV_E_INDEX for an index outside [0, n_examples). */
int kv_scm_best_rules_idx(kv_shard *s, const int64_t *pos_idx, int64_t n_pos, const int64_t *neg_idx,
                          int64_t n_neg, double p, int64_t util_block_size, double *best_utility,
                          int64_t capacity, int64_t *rule_idx, int64_t *pos_err, int64_t *neg_cover,
                          int64_t *n_found);
/* ---- a7: the greedy loop's example bookkeeping on the device (scm.py:133-139) -------------------
 * After a pick the reference fetches the rule's column (get_columns) and keeps the examples it accepts.  Here the
 * shard keeps the remaining positive / negative masks resident: kv_scm_state_set uploads them once per fit,
 * kv_mask_and_column ANDs both with the chosen column (invert = 1 for an absence rule: mask &= ~column) and returns
 * what is left.  local_col >= 0: the column is read from this shard's matrix and its W raw words are returned in
 * col_words_out (may be NULL) for the shards / ranks that do not hold it; local_col < 0: apply col_words_in.
 * The scan entry points take pos_mask = neg_mask = NULL to scan the resident state (n_pos / n_neg are then ignored:
 * the state's counts are used). */
int kv_scm_state_set(kv_shard *s, const uint64_t *pos_mask, const uint64_t *neg_mask, int64_t *n_pos, int64_t *n_neg);
int kv_mask_and_column(kv_shard *s, int64_t local_col, const uint64_t *col_words_in, int invert,
                       uint64_t *col_words_out, int64_t *n_pos, int64_t *n_neg);
/* CUDA events around every scan kernel (kv_shard_info.last_kernel_ms / scan_ms_total are only maintained while
 * this is on; off by default -- two event records per scan are a visible share of a 0.4 ms step).  Also
 * KOVER_B200_PROFILE=1 at kv_create. */
int kv_set_profiling(kv_shard *s, int on);

/* ---- a9 / a11: DecisionTreeClassifier _gini_rule_score + _find_best_split (cart.py:112-161,
 * 219-250) --------------------------------------------------------------------------------
 * class_masks: [n_classes, W] masks of the node's examples per class, in the caller's dict order;
 * altered_prior, n_total (cart.py:71-77), n_node = len(example_idx[c]).  Phase 1 returns the
 * minimum split score over the shard's presence rules (blacklisted / empty-child -> +inf). */
int kv_gini_scan(kv_shard *s, int n_classes, const uint64_t *class_masks, const double *altered_prior,
                 const double *n_total, const int64_t *n_node, double *min_score);
/* Phase 2: rules with score == min_score exactly (cart.py:238), GLOBAL indices ascending. */
int kv_gini_ties(kv_shard *s, double min_score, int64_t capacity, int64_t *rule_idx, int64_t *n_found);
/* Single-shard convenience: both phases in one call with one host synchronisation. */
int kv_gini_best_split(kv_shard *s, int n_classes, const uint64_t *class_masks, const double *altered_prior,
                       const double *n_total, const int64_t *n_node, double *min_score, int64_t capacity,
                       int64_t *rule_idx, int64_t *n_found);
/* A whole tree level at once (cart.py:260-339 visits the nodes of a level one after the other; their example
 * sets are disjoint and their scans independent): the split search of n_nodes two-class nodes, class masks
 * [n_nodes][2][W], altered_prior / n_total / n_node [n_nodes][2] (per node, so that the nodes of SEVERAL trees --
 * the fold trees and the master tree of a cross-validation -- can share passes), the class masks of up to 8
 * nodes counted per matrix pass.  Returns per
 * node the minimum score over the shard's presence rules and the number of rules achieving it exactly;
 * kv_gini_level_fetch then copies node i's rule indices (GLOBAL, ascending) into a caller buffer of at least
 * n_found[i] entries.  The results stay valid until the next kv_gini_level call on the shard. */
int kv_gini_level(kv_shard *s, int64_t n_nodes, const uint64_t *class_masks, const double *altered_prior,
                  const double *n_total, const int64_t *n_node, const int32_t *occ_table, double *min_score,
                  int64_t *n_found, uint32_t *occ_max);
/* Sibling subtraction across tree levels.  The children of a split partition their parent's examples
 * (cart.py:300-312: rule true / rule false), so the class counts of one child are the parent's minus the other
 * child's.  kv_gini_level_family arms the NEXT kv_gini_level call: node_key[i] >= 0 names node i, parent_key[i] the
 * node it was split from (-1: none; parent_key may be NULL).  The counts of the armed call's nodes stay resident
 * (slots of one arena sized from the free memory; nodes that find no slot are simply not remembered); when two nodes
 * of the next armed call name the same resident parent and their n_node add up to the parent's, only the SMALLER one
 * is counted in the matrix pass and the other one's counts are computed as parent - sibling (exact: integers).
 * Results are identical with or without the arming; an unarmed kv_gini_level ends the family.  n_nodes = 0 forgets
 * the resident level, n_nodes < 0 also frees the arena.  KOVER_B200_LEVEL_SUB=0 disables the subtraction. */
int kv_gini_level_family(kv_shard *s, int64_t n_nodes, const int64_t *node_key, const int64_t *parent_key);
/* The reference's CART tiebreaker on the device (experiment_cart.py:82-94, 264: among the best splits keep the
 * k-mers whose occurrence count in the tree's training set is numpy.isclose to the largest one).
 * kv_occ_table_set(table 0..63, mask) computes the presence counts of the shard's k-mers under `mask` into a
 * device-resident table (mask = NULL frees it).  A node of kv_gini_level with occ_table[i] >= 0 (occ_table may be
 * NULL) then reports only the ties that survive the tiebreaker WITHIN THIS SHARD, and their largest occurrence in
 * occ_max[i] (occ_max may be NULL) so that several shards can be merged: a shard's survivors stand when its
 * maximum is the global one; kv_occ_table_gather returns the occurrences of given rules for the rare re-filter
 * (maxima of 100 000 and more, where the tolerance exceeds 1). */
int kv_occ_table_set(kv_shard *s, int table, const uint64_t *mask);
int kv_occ_table_gather(kv_shard *s, int table, const int64_t *rule_idx, int64_t n, uint32_t *out);
int kv_gini_level_fetch(kv_shard *s, int64_t node, int64_t *rule_idx, int64_t capacity);
/* Zero-copy variant: *rule_idx points at node i's n_found indices inside the shard's pinned result arena (tie sets
 * of small nodes run to millions of rules; a second copy into fresh pageable memory costs more than the scan).
 * READ-ONLY, valid until the next kv_gini_level / kv_destroy on the shard -- the one library-owned pointer this
 * ABI hands out. */
int kv_gini_level_result(kv_shard *s, int64_t node, const int64_t **rule_idx, int64_t *n_found);
/* The split scores themselves (testing / callers that want rules_criterion): out[n_cols]. */
int kv_gini_scores(kv_shard *s, double *out);

#ifdef __cplusplus
}
#endif
#endif /* KOVER_B200_H */
