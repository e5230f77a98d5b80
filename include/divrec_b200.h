/*
 * divrec_b200.h -- C ABI of libdivrec_b200.so (B200 / sm_100a only).
 *
 * The reference (amtsyplov/diversity-recommendations, "divrec") has no FFI: its boundary is
 * the Python API.  Each entry point below replaces the arithmetic behind one reference call
 * site; the citation after "replaces:" is the reference file:line (relative to the reference
 * repository root).  The Python shim in diversity-recommendations_b200/divrec/ binds these
 * with ctypes (see INTEGRATION.md for the binding a reference maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into caller-owned memory unless marked "host";
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - functions enqueue work on `stream` and return; they never synchronise the device,
 *     never allocate outputs and never free anything the caller owns;
 *   - return value: DIVREC_OK (0) or a negative divrec_status; divrec_last_error() gives the
 *     message of the last failure on the calling thread;
 *   - there is no CPU path: a box without an sm_100 device gets DIVREC_ERR_DEVICE.
 */
#ifndef DIVREC_B200_H
#define DIVREC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIVREC_B200_VERSION 100 /* 0.1.0 */

typedef enum divrec_status {
  DIVREC_OK = 0,
  DIVREC_ERR_ARG = -1,       /* bad argument (shape, alignment, unsupported size) */
  DIVREC_ERR_WORKSPACE = -2, /* workspace too small */
  DIVREC_ERR_CUDA = -3,      /* a CUDA runtime / driver call failed */
  DIVREC_ERR_DEVICE = -4     /* current device is not sm_100 */
} divrec_status;

typedef enum divrec_dtype {
  DIVREC_F32 = 0, /* fp32 embedding tables (the reference's nn.Embedding default) */
  DIVREC_BF16 = 1 /* bf16 embedding tables (BASELINE.json config 4) */
} divrec_dtype;

typedef enum divrec_score_mode {
  /* tensor-core filter + exact fp32 re-scoring: the tables are rounded to ONE 16-bit plane (fp16 with a
   * power-of-two scale for fp32 tables; bf16 tables as stored), S = U V^T runs on tcgen05 with fp32
   * accumulation and is used only to FILTER; every candidate within 2*delta of the k-th best filter
   * score is re-scored exactly in fp32 from the original tables, so the reported top-k and scores are
   * those of the exact dot products.  Rows the filter cannot certify are re-done by the SIMT kernel. */
  DIVREC_MODE_TENSOR = 0,
  /* exact fp32 SIMT scoring + radix select for every row (no tensor cores; any d <= 8192) */
  DIVREC_MODE_SIMT = 2
} divrec_score_mode;

typedef enum divrec_ild_mode {
  DIVREC_ILD_SUMVEC = 0,  /* sum_{a<b} cos = (|sum_a e_a|^2 - sum_a |e_a|^2) / 2 : O(k d)  */
  DIVREC_ILD_PAIRWISE = 1 /* explicit k(k-1)/2 pairwise cosines in shared memory: O(k^2 d) */
} divrec_ild_mode;

int divrec_version(void);
const char* divrec_last_error(void);
/* number of CUDA kernels this library has launched in this process (bench.py "gpu_launches") */
int64_t divrec_launch_count(void);

/* Measurement aid (bench.py roofline): when on, the library records CUDA events on the launching
 * stream around its phases; divrec_profile_read waits for them and fills ms_out (host, [n]) with
 * the last durations in ms, -1 where nothing was recorded.  Slots: 0 operand preparation,
 * 1 score+filter kernel, 2 select kernel, 3 ILD, 4 coverage histogram, 5 rescue kernel.  Host-only calls. */
void divrec_set_profiling(int on);
/* Measurement aid: 0 = normal; 1 = score+filter kernel runs its TMA/MMA pipeline but skips the filter;
 * 3 = as 1 with half of the tensor-memory read traffic.  Results are meaningless unless mode == 0. */
void divrec_debug_filter_mode(int mode);
/* Measurement aid: device buffer int64 [148][8] (caller-zeroed) receiving, per CTA of the score+filter kernel, the
 * MMA-issuing thread's cycles: [0] total, [1] blocked on the A operand, [2] on a free accumulator stage, [3] on B operand
 * data; [4] list-tightening events and [5] the cycles spent in them (summed over warps).  NULL = off. */
void divrec_debug_filter_stats(void* dev_buf);
int divrec_profile_read(float* ms_out /*host*/, int n);

/* ---------------------------------------------------------------------------------------
 * score + seen-mask + per-user top-k
 *   replaces: divrec/train/utils.py:53-77 (get_model_recommendations) together with
 *             divrec/datasets/base_datasets.py:136-171 (RankingDataset.__iter__, the mask)
 *             and divrec/models/matrix_factorization.py:26-28 (MatrixFactorization.forward)
 *
 *   U [nU,d], V [nI,d] row-major, dtype f32 or bf16.  seen_rowptr [nU+1] / seen_items
 *   (global item ids, ascending inside a row) is the CSR of the `frozen` interactions; pass
 *   NULL/NULL for "no mask".  item_id_offset is added to local row indices of V to form the
 *   reported (global) item ids and subtracted from seen_items (item-sharded catalogue).
 *   Output row u = the k best unmasked items by (score desc, item id asc); rows with fewer
 *   than k unmasked items are padded with score -inf / item -1.
 *   Limits of this version: k <= 512; d <= 512 in DIVREC_MODE_TENSOR; nI < 2^31; tables 16-byte aligned;
 *   finite inputs.  Users are processed in internal passes of <= 151 552 rows, which bounds the workspace.
 *   divrec_score_topk_rescued: (host call, synchronises `stream`) host_out[0] = how many rows of the last call on
 *   this workspace were re-done by the exact SIMT kernel; host_out[1..4] = why (candidate-list overflow or forced,
 *   too many candidates, threshold hint not certified, too few candidates captured).  host_out holds 5 ints.
 * ------------------------------------------------------------------------------------- */
size_t divrec_score_topk_workspace(int64_t nU, int64_t nI, int d, int dtype, int k);

int divrec_score_topk(const void* U, const void* V, int64_t nU, int64_t nI, int d, int dtype,
                      int mode, const int64_t* seen_rowptr, const int32_t* seen_items,
                      int64_t item_id_offset, int k, float* out_scores /*[nU,k]*/,
                      int64_t* out_items /*[nU,k]*/, void* workspace, size_t ws_bytes,
                      void* stream);

/* The same pass with the list metrics fused into the selection kernel (the top-k item rows are already
 * in flight there for the exact re-scoring, so the ILD costs no second gather):
 *   out_ild [nU] (or NULL)          per-user cosine ILD of the row's top-k over the item table V, the value of
 *                                   divrec_ild_cosine on out_items (replaces
 *                                   divrec/losses/intra_list_diversity_score.py:20-42 with D = 1 - cos)
 *   coverage_counts [nI] (or NULL)  += histogram of the recommended ids, caller-zeroed (replaces torch.unique at
 *                                   divrec/metrics/entropy_diversity_score.py:22)
 * Whole catalogue only (no item_id_offset); the fused ILD needs d % 4 == 0, d <= 256, k >= 2. */
int divrec_score_topk_metrics(const void* U, const void* V, int64_t nU, int64_t nI, int d, int dtype,
                              int mode, const int64_t* seen_rowptr, const int32_t* seen_items, int k,
                              float* out_scores, int64_t* out_items, float* out_ild,
                              int32_t* coverage_counts, void* workspace, size_t ws_bytes, void* stream);

int divrec_score_topk_rescued(const void* workspace, int64_t nU, int64_t nI, int d, int dtype,
                                int32_t* host_out /*host [5]*/, void* stream);

/* ---------------------------------------------------------------------------------------
 * merge per-shard top-k lists after the all-gather (item-sharded catalogue)
 *   replaces: the argsort at divrec/train/utils.py:73-75 across catalogue shards
 *   gathered_* are [n_shards, nU, k]; every shard list is sorted (score desc, id asc) and
 *   padded with (-inf, -1).  Output: global top-k per user in the same order.
 * ------------------------------------------------------------------------------------- */
int divrec_topk_merge(const float* gathered_scores, const int64_t* gathered_items, int n_shards,
                      int64_t nU, int k, float* out_scores, int64_t* out_items, void* stream);

/* The same exchange in packed form: one 64-bit composite per (score, id) pair --
 *   (ordered fp32 key of the score) << 32 | ~(uint32_t)id, padding = 0 -- so that descending integer
 *   order is exactly (score desc, id asc).  divrec_topk_pack converts a shard's [nU,k] lists (n = nU*k
 *   elements, global ids < 2^32), ONE all-gather of 8 B per element replaces two of 4 + 8 B, and
 *   divrec_topk_merge_packed merges [n_shards, nU, k] composites (n_shards * k <= 6144) into the
 *   reference's layout (fp32 scores, int64 ids, padded with (-inf, -1)). */
int divrec_topk_pack(const float* scores, const int64_t* items, int64_t n, uint64_t* out_packed, void* stream);
int divrec_topk_merge_packed(const uint64_t* gathered, int n_shards, int64_t nU, int k, float* out_scores,
                             int64_t* out_items, void* stream);

/* ---------------------------------------------------------------------------------------
 * intra-list diversity with cosine distance over the item table
 *   replaces: divrec/losses/intra_list_diversity_score.py:20-42 fed with
 *             D = 1 - En En^T  (D is never materialised)
 *   out_per_user[u] = sum_{a<b} (1 - cos(V[rec_a], V[rec_b])) / (k (k-1)); items < 0 ignored.
 *   inv_norm [nI] fp32 scratch: 1/|V_i| (filled by this call).
 * ------------------------------------------------------------------------------------- */
int divrec_ild_cosine(const void* V, int64_t nI, int d, int dtype, const int64_t* recs /*[nU,k]*/,
                      int64_t nU, int k, int ild_mode, float* inv_norm /*[nI]*/,
                      float* out_per_user /*[nU]*/, void* stream);

/* ILD over an explicit distance matrix D [nI,nI] (fp32 or int32), rank order D[rec_a,rec_b], a<b
 *   replaces: divrec/losses/intra_list_diversity_score.py:36-42 (user_ild) for a caller-supplied D */
int divrec_ild_matrix(const void* D, int d_is_int32, int64_t nI, const int64_t* recs, int64_t nU,
                      int k, float* out_per_user, void* stream);

/* ---------------------------------------------------------------------------------------
 * coverage counts: histogram of recommended item ids
 *   replaces: torch.unique(recommendations, return_counts=True) at
 *             divrec/metrics/entropy_diversity_score.py:22
 *   counts [nI] int32 must be zeroed by the caller (so shards can accumulate); ids < 0 skipped.
 *   divrec_entropy_from_counts: out[0] = -sum p log p with p = counts / sum(counts)
 *   (entropy_diversity_score.py:23-24), out[1] = number of non-zero counts (as float).
 * ------------------------------------------------------------------------------------- */
int divrec_coverage_hist(const int64_t* recs, int64_t n, int64_t nI, int32_t* counts, void* stream);
int divrec_entropy_from_counts(const int32_t* counts, int64_t nI, float* out2, void* stream);

/* ---------------------------------------------------------------------------------------
 * BPR loss: gather + forward, and scatter-add backward
 *   replaces: divrec/models/matrix_factorization.py:26-28 called twice at
 *             divrec/train/utils.py:145-146, divrec/losses/log_sigmoid_difference_loss.py:14,
 *             divrec/metrics/auc_score.py:10, and autograd's embedding backward behind
 *             divrec/train/utils.py:149.
 *   fwd: pos[b] = <U[u_b],V[p_b]>, neg[b] = <U[u_b],V[n_b]>, loss_sum += -logsigmoid(pos-neg),
 *        auc_sum += (pos >= neg).  loss_sum [1] fp32 and auc_sum [1] int32 are caller-zeroed.
 *   bwd: with g_b = dL/dx_b supplied per triple (x = pos-neg):
 *        gU[u] += g (V[p]-V[n]);  gV[p] += g U[u];  gV[n] -= g U[u]   (duplicates accumulate)
 *        If g == NULL it is recomputed as -sigmoid(-(pos-neg)) * grad_scale.
 *   fp32 tables; any d <= 1024 (rows move as 128-bit vectors when d % 4 == 0 and the tables are 16-byte aligned,
 *   one float at a time otherwise -- nn.Embedding takes any embedding_dim).
 *   nU / nI are the table heights: a triple with an id outside [0, nU) x [0, nI)^2 is SKIPPED (no out-of-range
 *   access; its pos / neg read 0) and counted -- the reference's nn.Embedding raises IndexError there, and the
 *   shim turns a non-zero divrec_bad_index_count into that IndexError at its next synchronisation point.
 * ------------------------------------------------------------------------------------- */
int divrec_bpr_fwd(const float* U, const float* V, int64_t nU, int64_t nI, const int64_t* u, const int64_t* p,
                   const int64_t* n, int64_t B, int d, float* pos /*[B]*/, float* neg /*[B]*/,
                   float* loss_sum, int32_t* auc_sum, void* stream);

int divrec_bpr_bwd(const float* U, const float* V, int64_t nU, int64_t nI, const int64_t* u, const int64_t* p,
                   const int64_t* n, const float* pos, const float* neg, const float* g,
                   float grad_scale, int64_t B, int d, float* gU, float* gV, void* stream);

/* Forward and backward of one batch in ONE pass (the loop's `loss_value.backward()` at
 *   divrec/train/utils.py:149 has upstream gradient 1, so g_b = -sigmoid(-(pos_b-neg_b)) * grad_scale
 *   is known as soon as the scores are; grad_scale = 1/B for the mean loss): rows are read once, DRAM
 *   traffic = 3 row reads + 3 row read-modify-writes per triple.  pos / neg may be NULL (required when d % 4 != 0:
 *   that shape runs forward then backward); loss_sum and auc_sum as in divrec_bpr_fwd; gU / gV accumulate (dense
 *   tables, caller-owned, either may be NULL).
 *   item_rows (nullable, [B,d]): the SPARSE form of the item gradient -- row b = g_b U[u_b], to be added to
 *   gV[p_b] and subtracted from gV[n_b].  Multi-GPU steps whose batch touches few rows all-gather these rows (+ ids)
 *   instead of all-reducing the dense [nI,d] table and apply them with divrec_bpr_apply_rows. */
int divrec_bpr_step(const float* U, const float* V, int64_t nU, int64_t nI, const int64_t* u, const int64_t* p,
                    const int64_t* n, int64_t B, int d, float grad_scale, float* pos, float* neg,
                    float* loss_sum, int32_t* auc_sum, float* gU, float* gV, float* item_rows,
                    void* stream);

/* The same one-pass step with the item gradient SHARDED BY OWNER across the GPUs of a node (BASELINE.json config 3's
 *   gradient exchange fused into the kernel): gV_slices [n_slices] (device array of device pointers) holds, for every
 *   rank r, the address of the [slice_rows, d] gradient slice of items [r slice_rows, (r + 1) slice_rows) -- this rank's
 *   own slice and the peer-mapped slices of the others (divrec_peer_alloc / divrec_peer_open).  Every item-row update
 *   is scattered straight to its owner with red.global.add.v4.f32 over NVLink, overlapped with the gathers of the
 *   following triples: no packing, no all-to-all, no second apply pass; after the ranks' kernels and ONE barrier
 *   every rank holds the globally summed gradient of the items it owns (reduce-scatter semantics; the dense
 *   replicated table of autograd is the concatenation of the slices).  d % 4 == 0, 16-byte aligned tables.
 *   The two step statistics come out as ONE float pair so that a single all-reduce (which is also the barrier behind
 *   the remote reductions) carries both. */
int divrec_bpr_step_owner(const float* U, const float* V, int64_t nU, int64_t nI, const int64_t* u, const int64_t* p,
                          const int64_t* n, int64_t B, int d, float grad_scale, float* pos, float* neg,
                          float* stats2 /*[2] caller-zeroed: += loss sum, += AUC hit count (as float, exact below 2^24)*/,
                          float* gU, float* const* gV_slices, int n_slices, int64_t slice_rows, void* stream);

/* (host call, synchronises `stream`) triples / pairs skipped since the last reset because an id was out of range */
int divrec_bad_index_count(int64_t* host_out /*host*/, int reset, void* stream);

/* Peer buffers: device memory that the other processes of the node can map (CUDA IPC, NVLink / NVSwitch peer access).
 *   divrec_peer_alloc: cudaMalloc + export; handle_out is DIVREC_PEER_HANDLE_BYTES opaque bytes to hand to the peers
 *   (e.g. one all_gather); divrec_peer_open maps a peer's buffer into this process; close / free undo them.
 *   These four are the library's only allocating entry points. */
#define DIVREC_PEER_HANDLE_BYTES 64
int divrec_peer_alloc(size_t bytes, void** ptr_out, unsigned char* handle_out /*host [64]*/);
int divrec_peer_open(const unsigned char* handle /*host [64]*/, void** ptr_out);
int divrec_peer_close(void* ptr);
int divrec_peer_free(void* ptr);

/* ---------------------------------------------------------------------------------------
 * seen-item mask builder: interaction pairs -> CSR (rows ascending and unique)
 *   replaces: the per-user scan `interactions[:, 0] == user_id` + set construction of
 *             divrec/datasets/base_datasets.py:139-149 (RankingDataset.__iter__; 78 % of the reference's evaluation
 *             time at the MovieLens-20M shape) and the same scan in divrec/metrics/precision_at_k.py:18 etc.
 *   pairs [nnz, 2] int64 (user, item), any order, duplicates allowed; pairs with a user outside [0, nU) or an item
 *   outside [0, nI) are dropped.  rowptr [nU + 1] int64 and items [>= nnz] int32 are outputs: row u holds the distinct
 *   items of user u in ascending order, rowptr[nU] of the entries of `items` are valid.
 *   Sort-free: count per user, scan, scatter into buckets, per-bucket sort (bitonic or id-window bitmap) + de-duplication,
 *   second scan, compaction -- all on the device, no synchronisation.
 * ------------------------------------------------------------------------------------- */
size_t divrec_build_csr_workspace(int64_t nnz, int64_t nU);
int divrec_build_csr(const int64_t* pairs, int64_t nnz, int64_t nU, int64_t nI, int64_t* rowptr /*[nU+1]*/,
                     int32_t* items /*[nnz]*/, void* workspace, size_t ws_bytes, void* stream);

/* BPR triple sampler on the device
 *   replaces: PairWiseDataset.__iter__, divrec/datasets/base_datasets.py:70-107 -- per user max_sampled positives drawn
 *             with replacement from the user's items, max_sampled negatives with replacement from
 *             items - positives - frozen (random.choices, :88-90), cross product positive-major (:92-107).
 *   pos_* : CSR of the users' positive items; excl_* : CSR of positives + frozen items (the exclusion list of the
 *   negatives), both sorted and unique per row.  Outputs are [nU * max_sampled^2] in the reference's iteration order;
 *   users without a positive or without a negative get (-1,-1,-1) rows (the reference raises IndexError).
 *   Draws are counter-based on (seed, user, draw): reproducible, not bit-compatible with CPython's generator. */
int divrec_bpr_sample_triples(const int64_t* pos_rowptr, const int32_t* pos_items, const int64_t* excl_rowptr,
                              const int32_t* excl_items, int64_t nU, int64_t nI, int max_sampled, uint64_t seed,
                              int64_t* out_u, int64_t* out_p, int64_t* out_n, void* stream);

/* gV[p_j] += rows_j;  gV[n_j] -= rows_j   for j < N (duplicates accumulate) */
int divrec_bpr_apply_rows(const float* rows, const int64_t* p, const int64_t* n, int64_t N, int64_t nI, int d,
                          float* gV, void* stream);

/* paired scoring  out[l] = <U[user_id[l]], V[item_id[l]]>
 *   replaces: divrec/models/matrix_factorization.py:19-28 (MatrixFactorization.forward) */
int divrec_pair_scores(const float* U, const float* V, int64_t nU, int64_t nI, const int64_t* user_id,
                       const int64_t* item_id, int64_t L, int d, float* out, void* stream);

/* backward of divrec_pair_scores: gU[user_id[l]] += g[l] V[item_id[l]]; gV[item_id[l]] += g[l] U[user_id[l]]
 *   replaces: autograd's embedding_dense_backward behind loss_value.backward() at
 *             divrec/train/utils.py:149 for a single model(...) call; gU / gV may be NULL. */
int divrec_pair_scores_bwd(const float* U, const float* V, int64_t nU, int64_t nI, const int64_t* user_id,
                           const int64_t* item_id, const float* g, int64_t L, int d, float* gU,
                           float* gV, void* stream);

/* ILD with the label-equality distance D[i,j] = (labels[i] == labels[j]) without the [I,I] matrix
 *   replaces: divrec/losses/intra_list_diversity_score.py:45-63 (IntraListBinaryUnfairnessScore) */
int divrec_ild_labels(const int64_t* labels /*[nI]*/, int64_t nI, const int64_t* recs, int64_t nU, int k,
                      float* out_per_user, void* stream);

/* coverage counts plus the per-item sum of list positions (avg_rank)
 *   replaces: divrec/metrics/popularity_rank_correlation_for_items.py:27-41
 *   counts [nI] int32 and rank_sums [nI] uint64 are caller-zeroed. */
int divrec_coverage_hist_ranks(const int64_t* recs, int64_t nU, int k, int64_t nI, int32_t* counts,
                               uint64_t* rank_sums, void* stream);

/* relevance metrics @k against the CSR of each user's positive items
 *   replaces: divrec/metrics/precision_at_k.py:6-22, recall_at_k.py:6-22,
 *             average_precision_at_k.py:6-24, normalized_discounted_cumulative_gain.py:6-23
 *   pos_rowptr [nU+1], pos_items ascending per row; raw_counts [nU] = the number of interaction rows
 *   of the user (duplicates included: the reference's recall denominator).
 *   out [4, nU] fp32: precision, recall, average precision, ndcg. */
int divrec_relevance_metrics(const int64_t* recs, int64_t nU, int k, const int64_t* pos_rowptr,
                             const int32_t* pos_items, const int32_t* raw_counts, float* out, void* stream);

/* elementwise BPR loss on precomputed scores; loss_values / hits may be NULL
 *   replaces: divrec/losses/log_sigmoid_difference_loss.py:14 and divrec/metrics/auc_score.py:10
 *   bwd: gpos[b] = -sigmoid(-(pos[b]-neg[b])) * g[b], gneg[b] = -gpos[b] (autograd of the same line) */
int divrec_bpr_loss_fwd(const float* pos, const float* neg, int64_t B, float* loss_values, int32_t* hits,
                        void* stream);
int divrec_bpr_loss_bwd(const float* pos, const float* neg, const float* g, int64_t B, float* gpos,
                        float* gneg, void* stream);

/* ---------------------------------------------------------------------------------------
 * greedy MMR-style re-rank of each user's top-C candidates (BASELINE.json config 5)
 *   The reference has NO implementation of this (README.md:3-5 only announces "tools for increasing ...
 *   diversity"; divrec/train/utils.py:167-216 cannot run): semantics are those of
 *   oracle/divrec_oracle.py::mmr_rerank -- parity is pinned on the oracle, not on the reference.
 *   cand_items [nU,C] (sorted by (score desc, id asc), padded with -1) / cand_scores [nU,C] are the output of
 *   divrec_score_topk with k = C.  First pick = candidate 0, then k-1 times
 *       argmax_c  lambda * score_c - (1 - lambda) * max_{j picked} cos(V_c, V_j)     (ties -> lowest rank).
 *   out_items [nU,k] (padded with -1), out_scores [nU,k] (relevance of the picks) may be NULL.
 * ------------------------------------------------------------------------------------- */
size_t divrec_mmr_workspace(int C, int d);
int divrec_mmr_rerank(const void* V, int64_t nI, int d, int dtype, const int64_t* cand_items,
                      const float* cand_scores, int64_t nU, int C, int k, float lambda,
                      int64_t* out_items, float* out_scores, void* workspace, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIVREC_B200_H */
