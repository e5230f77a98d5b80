// metrics.cu -- intra-list diversity and coverage over the [U,k] recommendation lists.
//
// ILD replaces IntraListDiversityScore.recommendations_loss / user_ild
// (divrec/losses/intra_list_diversity_score.py:20-42), whose per-user value is
//     sum_{a<b} D[rec_a, rec_b] / (k (k-1)).
// For the cosine distance D = 1 - En En^T the [I,I] matrix is never built: rows of the item table
// are gathered with 128-bit loads and either
//   SUMVEC   : sum_{a<b} <e_a,e_b> = (|sum_a e_a|^2 - sum_a |e_a|^2) / 2        (O(k d), HBM/L2 bound)
//   PAIRWISE : all k(k-1)/2 dot products from normalised rows staged in shared memory (O(k^2 d))
// Coverage replaces torch.unique(recommendations, return_counts=True)
// (divrec/metrics/entropy_diversity_score.py:22) with an int32 histogram over item ids.
#include "common.cuh"
#include "kernels.h"

namespace divrec {

namespace {

__device__ __forceinline__ float4 ld_row4(const void* V, int dtype, int64_t row, int d, int c4) {
  // 4 consecutive elements c4*4.. of row `row` as fp32
  if (dtype == 0) {
    return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(V) + row * d) + c4);
  } else {
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(V) + row * d) + c4);
    float4 r;
    r.x = __uint_as_float(raw.x << 16);
    r.y = __uint_as_float(raw.x & 0xffff0000u);
    r.z = __uint_as_float(raw.y << 16);
    r.w = __uint_as_float(raw.y & 0xffff0000u);
    return r;
  }
}

// one warp per item row: inv_norm[i] = 1 / sqrt(sum v^2)   (d % 4 == 0)
__global__ void inv_norm_kernel(const void* __restrict__ V, int64_t nI, int d, int dtype, float* __restrict__ inv_norm) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= nI) return;
  const uint32_t lane = lane_id();
  float s = 0.f;
  for (int c4 = lane; c4 < (d >> 2); c4 += 32) {
    const float4 v = ld_row4(V, dtype, row, d, c4);
    s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
  }
  s = warp_sum(s);
  if (lane == 0) inv_norm[row] = 1.0f / sqrtf(s);
}

// SUMVEC: one warp per user; CH float4 accumulators per lane cover d <= 128*CH
template <int CH>
__global__ void __launch_bounds__(256)
ild_sumvec_kernel(const void* __restrict__ V, int64_t nI, int d, int dtype, const int64_t* __restrict__ recs,
                  int64_t nU, int k, const float* __restrict__ inv_norm, float* __restrict__ out) {
  const int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (u >= nU) return;
  const uint32_t lane = lane_id();
  const int nvec = d >> 2;
  float4 acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  int kv = 0;
  const int64_t* r = recs + u * k;
  for (int base = 0; base < k; base += 32) {
    int64_t my_id = -1;
    float my_inv = 0.f;
    if (base + (int)lane < k) {
      my_id = r[base + lane];
      if (my_id >= nI) my_id = -1;
      if (my_id >= 0) my_inv = __ldg(inv_norm + my_id);
    }
    const int cnt = min(32, k - base);
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      const int64_t id = __shfl_sync(0xffffffffu, my_id, j);
      const float inv = __shfl_sync(0xffffffffu, my_inv, j);
      if (id < 0) continue;
      kv++;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int c4 = c * 32 + lane;
        if (c4 < nvec) {
          const float4 v = ld_row4(V, dtype, id, d, c4);
          acc[c].x = fmaf(v.x, inv, acc[c].x); acc[c].y = fmaf(v.y, inv, acc[c].y);
          acc[c].z = fmaf(v.z, inv, acc[c].z); acc[c].w = fmaf(v.w, inv, acc[c].w);
        }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    s = fmaf(acc[c].x, acc[c].x, s); s = fmaf(acc[c].y, acc[c].y, s);
    s = fmaf(acc[c].z, acc[c].z, s); s = fmaf(acc[c].w, acc[c].w, s);
  }
  s = warp_sum(s);
  if (lane == 0) {
    const float kvf = (float)kv;
    const float sum_cos = 0.5f * (s - kvf);                 // sum_{a<b} cos
    const float pairs = 0.5f * kvf * (kvf - 1.0f);
    out[u] = (pairs - sum_cos) / ((float)k * (float)(k - 1));
  }
}

// PAIRWISE: one CTA (128 threads) per user.  Normalised rows in shared memory with a row stride of
// d+4 floats (conflict-free 128-bit reads); each thread owns 4x4 blocks of the pair matrix.
__global__ void __launch_bounds__(128)
ild_pairwise_kernel(const void* __restrict__ V, int64_t nI, int d, int dtype, const int64_t* __restrict__ recs,
                    int64_t nU, int k, const float* __restrict__ inv_norm, float* __restrict__ out) {
  extern __shared__ __align__(16) float rows_s[];
  __shared__ float red[4];
  const int64_t u = blockIdx.x;
  const int ld = d + 4;
  const int kp = (k + 3) & ~3;
  const int nvec = d >> 2;
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  // stage: warp w loads rows w, w+4, ...
  for (int a = warp; a < kp; a += 4) {
    int64_t id = -1;
    if (a < k) { id = recs[u * k + a]; if (id >= nI) id = -1; }
    const float inv = id >= 0 ? __ldg(inv_norm + id) : 0.f;
    for (int c4 = lane; c4 < nvec; c4 += 32) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (id >= 0) { v = ld_row4(V, dtype, id, d, c4); v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv; }
      *reinterpret_cast<float4*>(rows_s + (size_t)a * ld + c4 * 4) = v;
    }
    if (lane == 0) rows_s[(size_t)a * ld + d] = id >= 0 ? 1.f : 0.f;  // validity flag in the pad
  }
  __syncthreads();
  const int nb = kp >> 2;
  const int nblk = nb * (nb + 1) / 2;
  float total = 0.f;
  for (int blk = threadIdx.x; blk < nblk; blk += blockDim.x) {
    // blk -> (ba <= bb) in row-major upper-triangular order
    int ba = 0, rem = blk;
    while (rem >= nb - ba) { rem -= nb - ba; ++ba; }
    const int bb = ba + rem;
    float accp[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) accp[i][j] = 0.f;
    const float* pa = rows_s + (size_t)(ba * 4) * ld;
    const float* pb = rows_s + (size_t)(bb * 4) * ld;
    for (int c4 = 0; c4 < nvec; ++c4) {
      float4 xa[4], xb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        xa[i] = *reinterpret_cast<const float4*>(pa + (size_t)i * ld + c4 * 4);
        xb[i] = *reinterpret_cast<const float4*>(pb + (size_t)i * ld + c4 * 4);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          accp[i][j] = fmaf(xa[i].x, xb[j].x, accp[i][j]); accp[i][j] = fmaf(xa[i].y, xb[j].y, accp[i][j]);
          accp[i][j] = fmaf(xa[i].z, xb[j].z, accp[i][j]); accp[i][j] = fmaf(xa[i].w, xb[j].w, accp[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int a = ba * 4 + i, b = bb * 4 + j;
        if (a < b && b < k) {
          const float va = pa[(size_t)i * ld + d], vb = pb[(size_t)j * ld + d];
          if (va != 0.f && vb != 0.f) total += 1.0f - accp[i][j];
        }
      }
  }
  total = warp_sum(total);
  if (lane == 0) red[warp] = total;
  __syncthreads();
  if (threadIdx.x == 0) out[u] = (red[0] + red[1] + red[2] + red[3]) / ((float)k * (float)(k - 1));
}

// explicit distance matrix: one warp per user, pairs a<b in rank order D[rec_a, rec_b]
__global__ void __launch_bounds__(256)
ild_matrix_kernel(const void* __restrict__ D, int d_is_int32, int64_t nI, const int64_t* __restrict__ recs,
                  int64_t nU, int k, float* __restrict__ out) {
  const int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (u >= nU) return;
  const uint32_t lane = lane_id();
  const int64_t* r = recs + u * k;
  float s = 0.f;
  const int npairs = k * (k - 1) / 2;
  for (int pidx = lane; pidx < npairs; pidx += 32) {
    // pidx -> (a<b): row-major over the strict upper triangle
    int a = 0, rem = pidx;
    while (rem >= k - 1 - a) { rem -= k - 1 - a; ++a; }
    const int b = a + 1 + rem;
    const int64_t ia = r[a], ib = r[b];
    if (ia < 0 || ib < 0 || ia >= nI || ib >= nI) continue;
    if (d_is_int32) s += (float)__ldg(reinterpret_cast<const int32_t*>(D) + ia * nI + ib);
    else s += __ldg(reinterpret_cast<const float*>(D) + ia * nI + ib);
  }
  s = warp_sum(s);
  if (lane == 0) out[u] = s / ((float)k * (float)(k - 1));
}

__global__ void __launch_bounds__(256)
coverage_hist_kernel(const int64_t* __restrict__ recs, int64_t n, int64_t nI, int32_t* __restrict__ counts) {
  // two ids per 128-bit load
  const int64_t n2 = n >> 1;
  const longlong2* r2 = reinterpret_cast<const longlong2*>(recs);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    const longlong2 v = __ldg(r2 + i);
    if (v.x >= 0 && v.x < nI) atomicAdd(counts + v.x, 1);
    if (v.y >= 0 && v.y < nI) atomicAdd(counts + v.y, 1);
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t v = recs[n - 1];
    if (v >= 0 && v < nI) atomicAdd(counts + v, 1);
  }
}

// label distance D[i,j] = (label_i == label_j) (IntraListBinaryUnfairnessScore,
// divrec/losses/intra_list_diversity_score.py:60-63) without the [I,I] matrix: one warp per user,
// the k labels sit in registers / shared memory and every pair is compared once.
__global__ void __launch_bounds__(256)
ild_labels_kernel(const int64_t* __restrict__ labels, int64_t nI, const int64_t* __restrict__ recs, int64_t nU, int k,
                  float* __restrict__ out) {
  extern __shared__ long long lab_s[];  // [8 warps][k]
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int64_t u = (int64_t)blockIdx.x * 8 + warp;
  if (u >= nU) return;
  long long* lab = lab_s + (size_t)warp * k;
  for (int a = lane; a < k; a += 32) {
    const int64_t id = recs[u * k + a];
    // invalid ids get distinct sentinels so they never compare equal
    lab[a] = (id >= 0 && id < nI) ? (long long)__ldg(labels + id) : (long long)(INT64_MIN + a);
  }
  __syncwarp();
  int same = 0;
  for (int a = 0; a < k - 1; ++a) {
    const long long la = lab[a];
    if (la <= (long long)(INT64_MIN + k)) continue;
    for (int b = a + 1 + lane; b < k; b += 32) same += (lab[b] == la) ? 1 : 0;
  }
  same = __reduce_add_sync(0xffffffffu, same);
  if (lane == 0) out[u] = (float)same / ((float)k * (float)(k - 1));
}

// coverage histogram that also accumulates the sum of list positions per item (avg_rank,
// divrec/metrics/popularity_rank_correlation_for_items.py:27-41)
__global__ void __launch_bounds__(256)
coverage_hist_ranks_kernel(const int64_t* __restrict__ recs, int64_t nU, int k, int64_t nI, int32_t* __restrict__ counts,
                           unsigned long long* __restrict__ rank_sums) {
  const int64_t n = nU * k;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = __ldg(recs + i);
    if (v >= 0 && v < nI) {
      atomicAdd(counts + v, 1);
      atomicAdd(rank_sums + v, (unsigned long long)(i % k));
    }
  }
}

// relevance metrics @k: one warp per user; membership of each recommended id in the user's positives
// (CSR row, ascending) by binary search.  out [4][nU]: precision, recall, "average precision", ndcg
// (divrec/metrics/precision_at_k.py:16-22, recall_at_k.py:16-22, average_precision_at_k.py:18-24,
//  normalized_discounted_cumulative_gain.py:16-23)
__global__ void __launch_bounds__(256)
relevance_kernel(const int64_t* __restrict__ recs, int64_t nU, int k, const int64_t* __restrict__ rowptr,
                 const int32_t* __restrict__ items, const int32_t* __restrict__ raw_counts, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int64_t u = (int64_t)blockIdx.x * 8 + warp;
  if (u >= nU) return;
  const int64_t s0 = rowptr[u], s1 = rowptr[u + 1];
  float idcg = 0.f;
  for (int r = lane; r < k; r += 32) idcg += 1.0f / log2f((float)(r + 2));
  idcg = warp_sum(idcg);
  int hits_before = 0;  // relevant items in earlier 32-chunks
  float ap = 0.f, dcg = 0.f;
  for (int base = 0; base < k; base += 32) {
    const int r = base + lane;
    bool rel = false;
    if (r < k) {
      const int64_t id = recs[u * k + r];
      int64_t lo = s0, hi = s1;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)items[mid] < id) lo = mid + 1; else hi = mid;
      }
      rel = lo < s1 && (int64_t)items[lo] == id;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, rel);
    if (r < k) {
      const int cum = hits_before + __popc(m & (lanemask_lt() | (1u << lane)));  // inclusive cumsum of relevance
      ap += (float)cum / (float)(r + 1);
      if (rel) dcg += 1.0f / log2f((float)(r + 2));
    }
    hits_before += __popc(m);
  }
  ap = warp_sum(ap);
  dcg = warp_sum(dcg);
  if (lane == 0) {
    const float hits = (float)hits_before;
    out[0 * nU + u] = hits / (float)k;
    out[1 * nU + u] = hits / (float)raw_counts[u];
    out[2 * nU + u] = ap / (float)k;
    out[3 * nU + u] = dcg / idcg;
  }
}

// single CTA: out[0] = -sum p log p, p = c / sum(c) (fp32 p as the reference, fp64 accumulation); out[1] = #non-zero
__global__ void __launch_bounds__(1024)
entropy_kernel(const int32_t* __restrict__ counts, int64_t nI, float* __restrict__ out2) {
  __shared__ double sh[32];
  __shared__ double total_s;
  const uint32_t lane = lane_id();
  const int warp = threadIdx.x >> 5;
  double t = 0.0;
  for (int64_t i = threadIdx.x; i < nI; i += blockDim.x) t += (double)counts[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (lane == 0) sh[warp] = t;
  __syncthreads();
  if (threadIdx.x == 0) { double s = 0; for (int i = 0; i < 32; ++i) s += sh[i]; total_s = s; }
  __syncthreads();
  const float total = (float)total_s;
  double h = 0.0, nz = 0.0;
  for (int64_t i = threadIdx.x; i < nI; i += blockDim.x) {
    const int32_t c = counts[i];
    if (c > 0) { const float pr = (float)c / total; h -= (double)(pr * logf(pr)); nz += 1.0; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { h += __shfl_xor_sync(0xffffffffu, h, o); nz += __shfl_xor_sync(0xffffffffu, nz, o); }
  __syncthreads();
  __shared__ double sh2[32];
  if (lane == 0) { sh[warp] = h; sh2[warp] = nz; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int i = 0; i < 32; ++i) { a += sh[i]; b += sh2[i]; }
    out2[0] = (float)a;
    out2[1] = (float)b;
  }
}

}  // namespace

cudaError_t launch_inv_norm(const void* V, int64_t nI, int d, int dtype, float* inv_norm, cudaStream_t stream) {
  const int64_t threads = nI * 32;
  inv_norm_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(V, nI, d, dtype, inv_norm);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_ild_sumvec(const void* V, int64_t nI, int d, int dtype, const int64_t* recs, int64_t nU, int k,
                              const float* inv_norm, float* out, cudaStream_t stream) {
  const unsigned grid = (unsigned)((nU * 32 + 255) / 256);
  const int ch = (d + 127) / 128;
  if (ch <= 1) ild_sumvec_kernel<1><<<grid, 256, 0, stream>>>(V, nI, d, dtype, recs, nU, k, inv_norm, out);
  else if (ch <= 2) ild_sumvec_kernel<2><<<grid, 256, 0, stream>>>(V, nI, d, dtype, recs, nU, k, inv_norm, out);
  else if (ch <= 4) ild_sumvec_kernel<4><<<grid, 256, 0, stream>>>(V, nI, d, dtype, recs, nU, k, inv_norm, out);
  else ild_sumvec_kernel<8><<<grid, 256, 0, stream>>>(V, nI, d, dtype, recs, nU, k, inv_norm, out);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_ild_pairwise(const void* V, int64_t nI, int d, int dtype, const int64_t* recs, int64_t nU, int k,
                                const float* inv_norm, float* out, cudaStream_t stream) {
  const int kp = (k + 3) & ~3;
  const size_t smem = (size_t)kp * (d + 4) * sizeof(float);
  if (smem > 220 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(ild_pairwise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  ild_pairwise_kernel<<<(unsigned)nU, 128, smem, stream>>>(V, nI, d, dtype, recs, nU, k, inv_norm, out);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_ild_matrix(const void* D, int d_is_int32, int64_t nI, const int64_t* recs, int64_t nU, int k,
                              float* out, cudaStream_t stream) {
  const unsigned grid = (unsigned)((nU * 32 + 255) / 256);
  ild_matrix_kernel<<<grid, 256, 0, stream>>>(D, d_is_int32, nI, recs, nU, k, out);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_coverage_hist(const int64_t* recs, int64_t n, int64_t nI, int32_t* counts, cudaStream_t stream) {
  int64_t g = (n / 2 + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  coverage_hist_kernel<<<(unsigned)g, 256, 0, stream>>>(recs, n, nI, counts);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_entropy(const int32_t* counts, int64_t nI, float* out2, cudaStream_t stream) {
  entropy_kernel<<<1, 1024, 0, stream>>>(counts, nI, out2);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_ild_labels(const int64_t* labels, int64_t nI, const int64_t* recs, int64_t nU, int k, float* out,
                              cudaStream_t stream) {
  const size_t smem = (size_t)8 * k * sizeof(long long);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ild_labels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  ild_labels_kernel<<<(unsigned)((nU + 7) / 8), 256, smem, stream>>>(labels, nI, recs, nU, k, out);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_coverage_hist_ranks(const int64_t* recs, int64_t nU, int k, int64_t nI, int32_t* counts,
                                       unsigned long long* rank_sums, cudaStream_t stream) {
  int64_t g = (nU * k + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  coverage_hist_ranks_kernel<<<(unsigned)g, 256, 0, stream>>>(recs, nU, k, nI, counts, rank_sums);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_relevance(const int64_t* recs, int64_t nU, int k, const int64_t* rowptr, const int32_t* items,
                             const int32_t* raw_counts, float* out, cudaStream_t stream) {
  relevance_kernel<<<(unsigned)((nU + 7) / 8), 256, 0, stream>>>(recs, nU, k, rowptr, items, raw_counts, out);
  count_launch();
  return cudaGetLastError();
}

}  // namespace divrec
