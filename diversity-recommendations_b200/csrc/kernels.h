// kernels.h -- internal launcher interface between the .cu files and the C ABI (api.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace divrec {

constexpr int LIST_CAP = 512;  // candidate slots per (user row, column half)
constexpr int MAX_K = 128;

void count_launch();

struct FilterParams {
  int64_t nU;
  int32_t nI;
  int kblocks;  // K / 64
  int fmt;      // 0 = fp16 planes, 1 = bf16
  int k;
  int n_ublocks;
  int n_tiles;
  const int64_t* seen_rowptr;  // [nU+1] or null
  const int32_t* seen_items;   // global item ids, ascending per row
  int64_t item_id_offset;
  uint2* lists;                // [nU_pad][2][LIST_CAP]  (score bits, local item id)
  int32_t* list_count;         // [nU_pad][2]
  float* dump;                 // dense [nU,nI] output (debug) or null
  const int32_t* scale_exps;   // device [2]: scores are scaled by 2^(e0+e1); null = unscaled
};

size_t score_filter_smem_bytes(int kblocks);
cudaError_t launch_score_filter(const void* A16, const void* B16, FilterParams p, int sm_count,
                                cudaStream_t stream, const char** err);

// fp32 [rows,d] -> fp16 planes [rows, 3*d_pad]; role 0: [hi | hi*2^-11 | lo*2^11], role 1: [hi | lo*2^11 | hi*2^-11]
// amax_bits / scale_exp are device scalars (amax must be zeroed by the caller before the call).
cudaError_t launch_split_f32(const float* X, int64_t rows, int d, int d_pad, int role, __half* out,
                             uint32_t* amax_bits, int32_t* scale_exp, cudaStream_t stream);
// bf16 [rows,d] -> bf16 [rows,d_pad] zero padded (only when d % 64 != 0)
cudaError_t launch_pad_bf16(const void* X, int64_t rows, int d, int d_pad, void* out, cudaStream_t stream);

struct SelectParams {
  int64_t nU;
  int32_t nI;
  int k;
  const uint2* lists;
  const int32_t* list_count;
  const int64_t* seen_rowptr;
  const int32_t* seen_items;
  int64_t item_id_offset;
  const int32_t* scale_exps;
  float* out_scores;    // [nU,k]
  int64_t* out_items;   // [nU,k]
};
cudaError_t launch_select_topk(const SelectParams& p, cudaStream_t stream);

cudaError_t launch_topk_merge(const float* gs, const int64_t* gi, int n_shards, int64_t nU, int k,
                              float* out_scores, int64_t* out_items, cudaStream_t stream);

cudaError_t launch_inv_norm(const void* V, int64_t nI, int d, int dtype, float* inv_norm, cudaStream_t stream);
cudaError_t launch_ild_sumvec(const void* V, int64_t nI, int d, int dtype, const int64_t* recs, int64_t nU, int k,
                              const float* inv_norm, float* out, cudaStream_t stream);
cudaError_t launch_ild_pairwise(const void* V, int64_t nI, int d, int dtype, const int64_t* recs, int64_t nU, int k,
                                const float* inv_norm, float* out, cudaStream_t stream);
cudaError_t launch_ild_matrix(const void* D, int d_is_int32, int64_t nI, const int64_t* recs, int64_t nU, int k,
                              float* out, cudaStream_t stream);
cudaError_t launch_coverage_hist(const int64_t* recs, int64_t n, int64_t nI, int32_t* counts, cudaStream_t stream);
cudaError_t launch_entropy(const int32_t* counts, int64_t nI, float* out2, cudaStream_t stream);

cudaError_t launch_bpr_fwd(const float* U, const float* V, const int64_t* u, const int64_t* p, const int64_t* n,
                           int64_t B, int d, float* pos, float* neg, float* loss_sum, int32_t* auc_sum,
                           cudaStream_t stream);
cudaError_t launch_bpr_bwd(const float* U, const float* V, const int64_t* u, const int64_t* p, const int64_t* n,
                           const float* pos, const float* neg, const float* g, float grad_scale, int64_t B, int d,
                           float* gU, float* gV, cudaStream_t stream);
cudaError_t launch_pair_scores(const float* U, const float* V, const int64_t* uid, const int64_t* iid, int64_t L,
                               int d, float* out, cudaStream_t stream);

cudaError_t launch_pair_scores_bwd(const float* U, const float* V, const int64_t* uid, const int64_t* iid,
                                   const float* g, int64_t L, int d, float* gU, float* gV, cudaStream_t stream);

cudaError_t launch_ild_labels(const int64_t* labels, int64_t nI, const int64_t* recs, int64_t nU, int k, float* out,
                              cudaStream_t stream);
cudaError_t launch_coverage_hist_ranks(const int64_t* recs, int64_t nU, int k, int64_t nI, int32_t* counts,
                                       unsigned long long* rank_sums, cudaStream_t stream);
cudaError_t launch_relevance(const int64_t* recs, int64_t nU, int k, const int64_t* rowptr, const int32_t* items,
                             const int32_t* raw_counts, float* out, cudaStream_t stream);
cudaError_t launch_bpr_loss_fwd(const float* pos, const float* neg, int64_t B, float* lv, int32_t* hit, cudaStream_t stream);
cudaError_t launch_bpr_loss_bwd(const float* pos, const float* neg, const float* g, int64_t B, float* gpos, float* gneg,
                                cudaStream_t stream);

}  // namespace divrec
