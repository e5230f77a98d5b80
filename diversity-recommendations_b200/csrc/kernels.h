// kernels.h -- internal launcher interface between the .cu files and the C ABI (api.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>

namespace divrec {

constexpr int LIST_CAP = 256;        // candidate groups per list; a user row owns LISTS_PER_ROW lists
constexpr int LIST_CAP_MAX = 512;    // ... and for catalogues that tighten lists online (fewer, larger tightenings)
constexpr int LISTS_PER_ROW = 4;     // (column half of a tile) x (tile parity)
constexpr int GROUP = 8;             // consecutive item columns appended together
constexpr int REC_FLOATS = GROUP;    // one candidate record = 32 bytes = ONE 256-bit store: the GROUP fp32 scores, whose low mantissa
                                     // bytes carry the group's index (first item id / GROUP): byte j of the index in score j, for
                                     // j < id_bytes (2 below 2^19 items, else 4).  The epilogue of the filter is instruction- and
                                     // store-bound: a separate id store per append cost a second L2 transaction and its address
                                     // arithmetic.  The carried scores are off by < 2^-15 relative, which the certification
                                     // margin absorbs (REC_ID_SLACK, select_topk.cu).
constexpr float REC_ID_SLACK = 3.0f / 32768.0f;  // 3 eps, eps = 2^-15 |u||v|: added to 2 delta wherever record scores are compared
constexpr int N_SAMPLE = 16;         // item tiles scored first to seed the per-row threshold hint
constexpr int HINT_MIN_TILES = 32;   // catalogues with fewer 128-item tiles keep every group (no hint)
constexpr int TIGHTEN_MIN_ITEMS = 500000;  // catalogues at least this large use the filter variant that tightens lists online
constexpr int MAX_K = 512;          // largest k of score_topk (the select / rescue kernels switch to larger buffers above MAX_K_SMALL)
constexpr int MAX_K_SMALL = 128;

void count_launch();

// Debugging / measurement knobs read from the environment exist only in builds made with -DDIVREC_DEBUG_KNOBS; the release
// library never consults the environment (a stray variable must not be able to change what a call returns).
#ifdef DIVREC_DEBUG_KNOBS
inline const char* debug_env(const char* name) { return getenv(name); }
#else
inline const char* debug_env(const char*) { return nullptr; }
#endif

struct FilterParams {
  int64_t nU;
  int32_t nI;
  int kblocks;  // K / 64
  int fmt;      // 0 = fp16 plane, 1 = bf16
  int k;
  int use_hint;
  float hint_sigma;  // safety of the sampled threshold hint in standard deviations (set by the launcher)
  int hint_extra;    // extra ranks on top
  float hint_eps;    // > 0 (large catalogues): exact binomial tail bound on a too-high hint per row, instead of sigma + extra
  int debug_mode;  // 0 = normal; 1 = skip the filter (pipeline timing); 2 = filter without stores
  int cap;      // = LIST_CAP
  // filled by the launcher
  int n_ublocks, n_tiles, b_stages, acc_stages, n_sample, sample_stride;
  int sample_tpg;   // sample tiles folded into one hint maximum (large catalogues sample more tiles)
  const void* a16;             // [nU, K] 16-bit user plane (read row-wise into TMEM)
  long long* stats;            // [grid][8] MMA-issuer cycles: total, blocked on A, on accumulators, on B; [4] tightening events, [5] their cycles (null = off)
  const int64_t* seen_rowptr;  // [nU+1] or null (only the per-row counts are used here)
  float* lists;                // [nU_pad][4][cap][REC_FLOATS]  candidate records: tensor-core scores (scaled units) carrying the group index
  int32_t* list_count;         // [nU_pad][4]              groups appended; cap + 1 = overflow
  float* thr_out;              // [nU_pad][4]              final capture threshold of each list (scaled units, -inf = none)
  // error-bound inputs (the 2 delta margin of list tightening; same values as SelectParams)
  const float* norms_u;
  const uint32_t* nvmax_bits;
  const uint32_t* amax_bits;
  const int32_t* scale_exps;
  const int32_t* row_exps;
  float c_margin;
  int d;
  // two-block kernel only: seen-item pattern of the rows over the sample tiles (hint_mask_kernel), or null
  const unsigned long long* hint_mask;
  int64_t hint_mask_stride;    // rows (padded) between consecutive words of a row
  int mask_words;              // set by the launcher: 64-bit words of the dirty-block bitmap per (row, column half)
  int ladder_stride;           // set by the launcher: tiles between threshold-ladder checkpoints (0 = no ladder)
};

__host__ __device__ inline int rec_id_bytes(int64_t nI) { return nI < (int64_t)GROUP * 65536 ? 2 : 4; }
int score_filter_b_stages(int kblocks);
size_t score_filter_smem_bytes(int kblocks);
cudaError_t launch_score_filter(const void* A16, const void* B16, FilterParams p, int sm_count,
                                cudaStream_t stream, const char** err);
// two user blocks per CTA (d_pad <= 128, nI < TIGHTEN_MIN_ITEMS); a row's lists come out merged pairwise:
// slots (h,0)+(h,1) hold ONE list of 2 * cap groups, list_count[h*2+1] = 0
cudaError_t launch_score_filter2(const void* A16, const void* B16, FilterParams p, int sm_count,
                                 cudaStream_t stream, const char** err);
size_t hint_mask_bytes(int64_t rows_padded, int32_t nI);
cudaError_t launch_hint_mask(const int64_t* seen_rowptr, const int32_t* seen_items, int64_t nU, int64_t rows_padded, int32_t nI,
                             int64_t item_id_offset, unsigned long long* out, cudaStream_t stream);

// fp32 [rows,d] -> one fp16 plane [rows,d_pad] of X * 2^scale_exp (max |x| scaled into [2^13, 2^14)),
// plus fp32 row norms of X and (optionally) the maximum row norm as float bits.
// amax_bits must be zeroed by the caller; scale_exp is written.
cudaError_t launch_prep_f32(const float* X, int64_t rows, int d, int d_pad, __half* out, uint32_t* amax_bits,
                            int32_t* scale_exp, float* norms, float* inv_norms, uint32_t* normmax_bits, cudaStream_t stream);
// fp32 user rows -> fp16 plane with a per-ROW power-of-two scale (row_exp[r]) + row norms, in one pass
cudaError_t launch_prep_f32_rows(const float* X, int64_t rows, int d, int d_pad, __half* out, int32_t* row_exp, float* norms,
                                 cudaStream_t stream);
// bf16 [rows,d]: row norms (+ max); zero-padded copy [rows,d_pad] when out != null
cudaError_t launch_prep_bf16(const void* X, int64_t rows, int d, int d_pad, void* out, float* norms, float* inv_norms,
                             uint32_t* normmax_bits, cudaStream_t stream);

struct SelectParams {
  int64_t nU;
  int32_t nI;
  int k, d, dtype, cap;
  const void* U;               // original tables (fp32 or bf16) for the exact re-scoring
  const void* V;
  const float* lists;          // candidate records written by the filter (REC_FLOATS floats each)
  const int32_t* list_count;
  const float* thr;
  const int64_t* seen_rowptr;
  const int32_t* seen_items;
  int64_t item_id_offset;
  const int32_t* scale_exps;   // device [2] or null: [1] = exponent of the item plane (tensor-core scores carry 2^(row_exps[row] + scale_exps[1]))
  const int32_t* row_exps;     // device [nU] or null: per-row exponent of the user plane
  const float* norms_u;        // [nU]
  const uint32_t* nvmax_bits;  // device: max item row norm (float bits)
  const float* inv_norm_v;     // [nI] 1 / |V_i| (fused ILD)
  const uint32_t* amax_bits;   // device [2]: [1] = max |element| of the item table (float bits), null for bf16
  float c_margin;              // delta_u = c_margin * |u| * max|v|
  float* out_scores;           // [nU,k]
  int64_t* out_items;          // [nU,k]
  int32_t* fail_rows;          // [nU] rows the tensor-core path could not certify
  int32_t* fail_count;         // device counter (zeroed by the caller)
  int32_t* fail_reasons;       // device [4] statistics by reason (see report_fail), may be null
  int id_bytes;                // low mantissa bytes of a record that carry the group index (2 or 4; set from nI by the launcher)
  int merged_lists;            // 1 = lists written by launch_score_filter2 (slots 0 and 2 hold up to 2 * cap groups)
  int debug_dump;              // 1 = uncertified rows leave their diagnostics in out_scores (env DIVREC_DEBUG_NO_RESCUE)
  int force_fail;              // 1 = send every row to the rescue kernel (DIVREC_MODE_SIMT)
  float* out_ild;              // [nU] fused cosine ILD of the row's top-k (null = off; needs d % 4 == 0, d <= 512)
  int32_t* coverage;           // [>= nI + offset] fused coverage histogram over global item ids (null = off)
};
cudaError_t launch_select_topk(const SelectParams& p, cudaStream_t stream);

struct RescueParams {
  int64_t nU;
  int32_t nI;
  int k, d, dtype;
  const void* U;
  const void* V;
  const int64_t* seen_rowptr;
  const int32_t* seen_items;
  int64_t item_id_offset;
  const int32_t* fail_rows;
  const int32_t* fail_count;
  uint32_t* scratch;           // [grid][nI_pad] ordered score keys
  int grid;
  int prescored;               // set by the launcher: rows f < prescored arrive with their keys already in slot f
  int keys_in_smem;            // set by the launcher: the row's keys fit the CTA's shared memory for the selection passes
  float* out_scores;
  int64_t* out_items;
  float* out_ild;              // fused metrics for the rows re-done here (null = off)
  int32_t* coverage;
  int32_t* total_fail;         // device counter accumulated over the chunks of a call (statistics), may be null
};
size_t rescue_scratch_bytes(int64_t nI, int* grid_out);
cudaError_t launch_rescue(const RescueParams& p, cudaStream_t stream);

cudaError_t launch_topk_merge(const float* gs, const int64_t* gi, int n_shards, int64_t nU, int k,
                              float* out_scores, int64_t* out_items, cudaStream_t stream);

cudaError_t launch_topk_pack(const float* scores, const int64_t* items, int64_t n, uint64_t* out, cudaStream_t stream);
cudaError_t launch_topk_merge_packed(const uint64_t* g, int n_shards, int64_t nU, int k, float* out_scores,
                                     int64_t* out_items, cudaStream_t stream);

cudaError_t launch_inv_norm(const void* V, int64_t nI, int d, int dtype, float* inv_norm, cudaStream_t stream);
cudaError_t launch_ild_sumvec(const void* V, int64_t nI, int d, int dtype, const int64_t* recs, int64_t nU, int k,
                              const float* inv_norm, float* out, cudaStream_t stream);
cudaError_t launch_ild_pairwise(const void* V, int64_t nI, int d, int dtype, const int64_t* recs, int64_t nU, int k,
                                const float* inv_norm, float* out, cudaStream_t stream);
cudaError_t launch_ild_matrix(const void* D, int d_is_int32, int64_t nI, const int64_t* recs, int64_t nU, int k,
                              float* out, cudaStream_t stream);
cudaError_t launch_coverage_hist(const int64_t* recs, int64_t n, int64_t nI, int32_t* counts, cudaStream_t stream);
cudaError_t launch_entropy(const int32_t* counts, int64_t nI, float* out2, cudaStream_t stream);

cudaError_t launch_bpr_fwd(const float* U, const float* V, int64_t nU, int64_t nI, const int64_t* u, const int64_t* p,
                           const int64_t* n, int64_t B, int d, float* pos, float* neg, float* loss_sum, int32_t* auc_sum,
                           cudaStream_t stream);
cudaError_t launch_bpr_bwd(const float* U, const float* V, int64_t nU, int64_t nI, const int64_t* u, const int64_t* p,
                           const int64_t* n, const float* pos, const float* neg, const float* g, float grad_scale, int64_t B,
                           int d, float* gU, float* gV, cudaStream_t stream);
// slices != null: the item gradient goes to per-owner slices of slice_rows items (peer-mapped addresses), gV unused
cudaError_t launch_bpr_step(const float* U, const float* V, int64_t nU, int64_t nI, const int64_t* u, const int64_t* p,
                            const int64_t* n, int64_t B, int d, float grad_scale, float* pos, float* neg, float* loss_sum,
                            int32_t* auc_sum, float* gU, float* gV, float* item_rows, float* const* slices, int64_t slice_rows,
                            float* auc_f, cudaStream_t stream);
cudaError_t launch_bpr_sample(const int64_t* pos_rowptr, const int32_t* pos_items, const int64_t* excl_rowptr,
                              const int32_t* excl_items, int64_t nU, int64_t nI, int m, uint64_t seed, int64_t* out_u,
                              int64_t* out_p, int64_t* out_n, cudaStream_t stream);
cudaError_t launch_bpr_apply_rows(const float* rows, const int64_t* p, const int64_t* n, int64_t N, int64_t nI, int d, float* gV,
                                  cudaStream_t stream);
cudaError_t launch_pair_scores(const float* U, const float* V, int64_t nU, int64_t nI, const int64_t* uid, const int64_t* iid,
                               int64_t L, int d, float* out, cudaStream_t stream);
cudaError_t launch_pair_scores_bwd(const float* U, const float* V, int64_t nU, int64_t nI, const int64_t* uid, const int64_t* iid,
                                   const float* g, int64_t L, int d, float* gU, float* gV, cudaStream_t stream);
cudaError_t read_bad_index_count(unsigned long long* host_out, int reset, cudaStream_t stream);

cudaError_t launch_ild_labels(const int64_t* labels, int64_t nI, const int64_t* recs, int64_t nU, int k, float* out,
                              cudaStream_t stream);
cudaError_t launch_coverage_hist_ranks(const int64_t* recs, int64_t nU, int k, int64_t nI, int32_t* counts,
                                       unsigned long long* rank_sums, cudaStream_t stream);
cudaError_t launch_relevance(const int64_t* recs, int64_t nU, int k, const int64_t* rowptr, const int32_t* items,
                             const int32_t* raw_counts, float* out, cudaStream_t stream);
cudaError_t launch_bpr_loss_fwd(const float* pos, const float* neg, int64_t B, float* lv, int32_t* hit, cudaStream_t stream);
cudaError_t launch_bpr_loss_bwd(const float* pos, const float* neg, const float* g, int64_t B, float* gpos, float* gneg,
                                cudaStream_t stream);

size_t build_csr_workspace_bytes(int64_t nnz, int64_t nU);
cudaError_t launch_build_csr(const int64_t* pairs, int64_t nnz, int64_t nU, int64_t nI, int64_t* rowptr, int32_t* items,
                             void* workspace, cudaStream_t stream);

struct MmrParams {
  const void* V;
  int64_t nI;
  int d, dtype;
  const int64_t* cand_items;   // [nU, C] sorted by (score desc, id asc), padded with -1
  const float* cand_scores;    // [nU, C]
  int64_t nU;
  int C, k;
  float lambda;
  float* scratch;              // mmr_scratch_bytes
  int64_t* out_items;          // [nU, k]
  float* out_scores;           // [nU, k] relevance scores of the picks, or null
};
size_t mmr_scratch_bytes(int C, int d, int* grid_out);
cudaError_t launch_mmr(MmrParams p, cudaStream_t stream);

}  // namespace divrec
