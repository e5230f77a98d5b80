// bpr.cu -- BPR (log-sigmoid difference) loss: embedding gather + forward, scatter-add backward.
//
// Replaces, for one batch of (user, positive, negative) triples:
//   forward  : model(u,pos), model(u,neg)  -- divrec/train/utils.py:145-146 through
//              divrec/models/matrix_factorization.py:26-28; loss -logsigmoid(pos-neg) with mean
//              reduction -- divrec/losses/log_sigmoid_difference_loss.py:14, base_losses.py:25-26;
//              AUC (pos >= neg) -- divrec/metrics/auc_score.py:10
//   backward : loss.backward() -- divrec/train/utils.py:149, i.e. autograd's dense embedding
//              backward: gU[u] += g (V[p]-V[n]); gV[p] += g U[u]; gV[n] -= g U[u]
// One warp per triple; rows move as 128-bit vectors; the backward uses vectorised
// red.global.add.v4.f32 so duplicates of a row accumulate in L2 without a read-back.
#include "common.cuh"
#include "kernels.h"

namespace divrec {

namespace {

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ void red_add_f32(float* addr, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}

// Rows move as 128-bit vectors when d % 4 == 0 (V4); any other d takes the same kernels one float at a time
// (the reference's nn.Embedding takes any embedding_dim).  `c` counts vectors (V4) or floats.
template <bool V4>
__device__ __forceinline__ float4 ld_row(const float* row, int c) {
  if (V4) return __ldg(reinterpret_cast<const float4*>(row) + c);
  return make_float4(__ldg(row + c), 0.f, 0.f, 0.f);
}
template <bool V4>
__device__ __forceinline__ void red_row(float* row, int c, float x, float y, float z, float w) {
  if (V4) red_add_v4(row + c * 4, x, y, z, w);
  else red_add_f32(row + c, x);
}

// ids out of [0, n): torch.nn.Embedding (the reference) raises IndexError / a device assert.  Here the triple is
// skipped (no out-of-range access) and counted; divrec_bad_index_count reports the count at the caller's next sync.
__device__ unsigned long long g_bad_ids = 0ull;
__device__ __forceinline__ bool ids_ok(int64_t iu, int64_t nU, int64_t ia, int64_t ib, int64_t nI, uint32_t lane) {
  const bool ok = (uint64_t)iu < (uint64_t)nU && (uint64_t)ia < (uint64_t)nI && (uint64_t)ib < (uint64_t)nI;
  if (!ok && lane == 0) atomicAdd(&g_bad_ids, 1ull);
  return ok;
}

// torch.nn.LogSigmoid: min(x,0) - log1p(exp(-|x|))
__device__ __forceinline__ float neg_logsigmoid(float x) { return log1pf(expf(-fabsf(x))) - fminf(x, 0.f); }

template <bool V4>
__global__ void __launch_bounds__(256)
bpr_fwd_kernel(const float* __restrict__ U, const float* __restrict__ V, int64_t nU, int64_t nI, const int64_t* __restrict__ u,
               const int64_t* __restrict__ p, const int64_t* __restrict__ n, int64_t B, int d,
               float* __restrict__ pos, float* __restrict__ neg, float* __restrict__ loss_sum,
               int32_t* __restrict__ auc_sum) {
  __shared__ float red_loss[8];
  __shared__ int red_auc[8];
  const uint32_t lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int nvec = V4 ? d >> 2 : d;
  float wl = 0.f;
  int wa = 0;
  for (int64_t b = (int64_t)blockIdx.x * 8 + warp; b < B; b += (int64_t)gridDim.x * 8) {
    const int64_t iu = u[b], ip = p[b], in = n[b];
    if (!ids_ok(iu, nU, ip, in, nI, lane)) { if (lane == 0) { pos[b] = 0.f; neg[b] = 0.f; } continue; }
    const float* ur = U + iu * d;
    const float* pr = V + ip * d;
    const float* nr = V + in * d;
    float sp = 0.f, sn = 0.f;
    for (int c = lane; c < nvec; c += 32) {
      const float4 a = ld_row<V4>(ur, c), x = ld_row<V4>(pr, c), y = ld_row<V4>(nr, c);
      sp = fmaf(a.x, x.x, sp); sp = fmaf(a.y, x.y, sp); sp = fmaf(a.z, x.z, sp); sp = fmaf(a.w, x.w, sp);
      sn = fmaf(a.x, y.x, sn); sn = fmaf(a.y, y.y, sn); sn = fmaf(a.z, y.z, sn); sn = fmaf(a.w, y.w, sn);
    }
    sp = warp_sum(sp);
    sn = warp_sum(sn);
    if (lane == 0) {
      pos[b] = sp;
      neg[b] = sn;
      wl += neg_logsigmoid(sp - sn);
      wa += (sp >= sn) ? 1 : 0;
    }
  }
  if (lane == 0) { red_loss[warp] = wl; red_auc[warp] = wa; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float l = 0.f;
    int a = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { l += red_loss[i]; a += red_auc[i]; }
    if (loss_sum) atomicAdd(loss_sum, l);
    if (auc_sum) atomicAdd(auc_sum, a);
  }
}

template <bool V4>
__global__ void __launch_bounds__(256)
bpr_bwd_kernel(const float* __restrict__ U, const float* __restrict__ V, int64_t nU, int64_t nI, const int64_t* __restrict__ u,
               const int64_t* __restrict__ p, const int64_t* __restrict__ n, const float* __restrict__ pos,
               const float* __restrict__ neg, const float* __restrict__ g, float grad_scale, int64_t B, int d,
               float* __restrict__ gU, float* __restrict__ gV) {
  const uint32_t lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int nvec = V4 ? d >> 2 : d;
  for (int64_t b = (int64_t)blockIdx.x * 8 + warp; b < B; b += (int64_t)gridDim.x * 8) {
    const int64_t iu = u[b], ip = p[b], in = n[b];
    if (!ids_ok(iu, nU, ip, in, nI, lane)) continue;
    float c;
    if (g != nullptr) {
      c = g[b];
    } else {
      const float x = pos[b] - neg[b];
      c = -grad_scale / (1.0f + expf(x));  // d/dx [-logsigmoid(x)] = -sigmoid(-x)
    }
    const float* ur = U + iu * d;
    const float* pr = V + ip * d;
    const float* nr = V + in * d;
    for (int v = lane; v < nvec; v += 32) {
      const float4 a = ld_row<V4>(ur, v), x = ld_row<V4>(pr, v), y = ld_row<V4>(nr, v);
      if (gU) red_row<V4>(gU + iu * d, v, c * (x.x - y.x), c * (x.y - y.y), c * (x.z - y.z), c * (x.w - y.w));
      if (gV) {
        red_row<V4>(gV + ip * d, v, c * a.x, c * a.y, c * a.z, c * a.w);
        red_row<V4>(gV + in * d, v, -c * a.x, -c * a.y, -c * a.z, -c * a.w);
      }
    }
  }
}

// Forward AND backward of one batch in ONE pass over the rows (the train loop's loss.backward() has upstream
// gradient 1, so c = -sigmoid(-(pos - neg)) * grad_scale is known as soon as the two dot products are): the three
// rows of a triple are read once into registers, used for the scores, then for the three scatter-adds.  DRAM
// traffic per triple = 3 row reads + 3 row read-modify-writes = the algorithmic figure (the two-kernel form reads
// the rows twice).  NV = float4 chunks of a row per lane (d <= 128 * NV).
//
// OWNER (multi-GPU, item gradient sharded by owner): the item table's gradient is not one local dense table but
// `n_slices` slices of `slice_rows` items, slice r living on rank r; `slices` holds their device addresses -- this rank's
// own slice and the PEER-MAPPED slices of the other ranks (CUDA IPC over NVLink / NVSwitch).  The scatter-add of an item
// row goes straight to its owner with the same vector red: the exchange of the item gradient happens inside this
// kernel, row by row, overlapped with the gathers of the next triples -- no pack, no all-to-all, no second apply pass,
// and every rank ends up holding the globally summed gradient of the items it owns (reduce-scatter semantics).
template <int NV, bool OWNER>
__global__ void __launch_bounds__(256)
bpr_step_kernel(const float* __restrict__ U, const float* __restrict__ V, int64_t nU, int64_t nI, const int64_t* __restrict__ u,
                const int64_t* __restrict__ p, const int64_t* __restrict__ n, int64_t B, int d, float grad_scale,
                float* __restrict__ pos, float* __restrict__ neg, float* __restrict__ loss_sum,
                int32_t* __restrict__ auc_sum, float* __restrict__ gU, float* __restrict__ gV,
                float* __restrict__ item_rows, float* const* __restrict__ slices, int64_t slice_rows, float* __restrict__ auc_f) {
  __shared__ float red_loss[8];
  __shared__ int red_auc[8];
  const uint32_t lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int nvec = d >> 2;
  float wl = 0.f;
  int wa = 0;
  for (int64_t b = (int64_t)blockIdx.x * 8 + warp; b < B; b += (int64_t)gridDim.x * 8) {
    const int64_t iu = u[b], ip = p[b], in = n[b];
    if (!ids_ok(iu, nU, ip, in, nI, lane)) { if (lane == 0) { if (pos) pos[b] = 0.f; if (neg) neg[b] = 0.f; } continue; }
    const float4* ur = reinterpret_cast<const float4*>(U + iu * d);
    const float4* pr = reinterpret_cast<const float4*>(V + ip * d);
    const float4* nr = reinterpret_cast<const float4*>(V + in * d);
    float4 a[NV], x[NV], y[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = (int)lane + 32 * j;
      if (c < nvec) { a[j] = __ldg(ur + c); x[j] = __ldg(pr + c); y[j] = __ldg(nr + c); }
      else { a[j] = x[j] = y[j] = make_float4(0.f, 0.f, 0.f, 0.f); }
    }
    float sp = 0.f, sn = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      sp = fmaf(a[j].x, x[j].x, sp); sp = fmaf(a[j].y, x[j].y, sp); sp = fmaf(a[j].z, x[j].z, sp); sp = fmaf(a[j].w, x[j].w, sp);
      sn = fmaf(a[j].x, y[j].x, sn); sn = fmaf(a[j].y, y[j].y, sn); sn = fmaf(a[j].z, y[j].z, sn); sn = fmaf(a[j].w, y[j].w, sn);
    }
    sp = warp_sum(sp);   // xor-shuffle reduction: every lane holds the same sums (same order as bpr_fwd_kernel)
    sn = warp_sum(sn);
    const float c = -grad_scale / (1.0f + expf(sp - sn));
    if (lane == 0) {
      if (pos) pos[b] = sp;
      if (neg) neg[b] = sn;
      wl += neg_logsigmoid(sp - sn);
      wa += (sp >= sn) ? 1 : 0;
    }
    float* gp = nullptr;
    float* gn = nullptr;
    if (OWNER) {
      const int64_t op = ip / slice_rows, on = in / slice_rows;
      gp = slices[op] + (ip - op * slice_rows) * d;
      gn = slices[on] + (in - on * slice_rows) * d;
    } else if (gV) {
      gp = gV + ip * d;
      gn = gV + in * d;
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int v = (int)lane + 32 * j;
      if (v < nvec) {
        if (gU) red_add_v4(gU + iu * d + v * 4, c * (x[j].x - y[j].x), c * (x[j].y - y[j].y), c * (x[j].z - y[j].z), c * (x[j].w - y[j].w));
        if (gp) {
          red_add_v4(gp + v * 4, c * a[j].x, c * a[j].y, c * a[j].z, c * a[j].w);
          red_add_v4(gn + v * 4, -c * a[j].x, -c * a[j].y, -c * a[j].z, -c * a[j].w);
        }
        // sparse form of the item gradient: row b = c U[u_b], to be added to gV[p_b] and subtracted from gV[n_b]
        if (item_rows) reinterpret_cast<float4*>(item_rows + b * d)[v] = make_float4(c * a[j].x, c * a[j].y, c * a[j].z, c * a[j].w);
      }
    }
  }
  if (lane == 0) { red_loss[warp] = wl; red_auc[warp] = wa; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float l = 0.f;
    int au = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { l += red_loss[i]; au += red_auc[i]; }
    if (loss_sum) atomicAdd(loss_sum, l);
    if (auc_sum) atomicAdd(auc_sum, au);
    if (auc_f) atomicAdd(auc_f, (float)au);   // the same count as a float (exact below 2^24): one all-reduce carries loss and AUC
  }
}

// gV[p_j] += rows_j;  gV[n_j] -= rows_j  -- applies item-gradient rows gathered from all ranks (bpr_step's item_rows)
__global__ void __launch_bounds__(256)
bpr_apply_rows_kernel(const float* __restrict__ rows, const int64_t* __restrict__ p, const int64_t* __restrict__ n,
                      int64_t N, int64_t nI, int d, float* __restrict__ gV) {
  const uint32_t lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int nvec = d >> 2;
  for (int64_t b = (int64_t)blockIdx.x * 8 + warp; b < N; b += (int64_t)gridDim.x * 8) {
    const int64_t ip = p[b], in = n[b];
    if (!ids_ok(0, 1, ip, in, nI, lane)) continue;
    for (int v = lane; v < nvec; v += 32) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(rows + b * d) + v);
      red_add_v4(gV + ip * d + v * 4, r.x, r.y, r.z, r.w);
      red_add_v4(gV + in * d + v * 4, -r.x, -r.y, -r.z, -r.w);
    }
  }
}

template <bool V4>
__global__ void __launch_bounds__(256)
pair_scores_kernel(const float* __restrict__ U, const float* __restrict__ V, int64_t nU, int64_t nI, const int64_t* __restrict__ uid,
                   const int64_t* __restrict__ iid, int64_t L, int d, float* __restrict__ out) {
  const uint32_t lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int nvec = V4 ? d >> 2 : d;
  for (int64_t l = (int64_t)blockIdx.x * 8 + warp; l < L; l += (int64_t)gridDim.x * 8) {
    const int64_t iu = uid[l], ii = iid[l];
    if (!ids_ok(iu, nU, ii, ii, nI, lane)) { if (lane == 0) out[l] = 0.f; continue; }
    const float* ur = U + iu * d;
    const float* vr = V + ii * d;
    float s = 0.f;
    for (int c = lane; c < nvec; c += 32) {
      const float4 a = ld_row<V4>(ur, c), x = ld_row<V4>(vr, c);
      s = fmaf(a.x, x.x, s); s = fmaf(a.y, x.y, s); s = fmaf(a.z, x.z, s); s = fmaf(a.w, x.w, s);
    }
    s = warp_sum(s);
    if (lane == 0) out[l] = s;
  }
}

// backward of pair_scores: gU[uid[l]] += g[l] V[iid[l]];  gV[iid[l]] += g[l] U[uid[l]]
template <bool V4>
__global__ void __launch_bounds__(256)
pair_scores_bwd_kernel(const float* __restrict__ U, const float* __restrict__ V, int64_t nU, int64_t nI, const int64_t* __restrict__ uid,
                       const int64_t* __restrict__ iid, const float* __restrict__ g, int64_t L, int d,
                       float* __restrict__ gU, float* __restrict__ gV) {
  const uint32_t lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int nvec = V4 ? d >> 2 : d;
  for (int64_t l = (int64_t)blockIdx.x * 8 + warp; l < L; l += (int64_t)gridDim.x * 8) {
    const int64_t iu = uid[l], ii = iid[l];
    if (!ids_ok(iu, nU, ii, ii, nI, lane)) continue;
    const float c = g[l];
    const float* ur = U + iu * d;
    const float* vr = V + ii * d;
    for (int v = lane; v < nvec; v += 32) {
      const float4 a = ld_row<V4>(ur, v), x = ld_row<V4>(vr, v);
      if (gU) red_row<V4>(gU + iu * d, v, c * x.x, c * x.y, c * x.z, c * x.w);
      if (gV) red_row<V4>(gV + ii * d, v, c * a.x, c * a.y, c * a.z, c * a.w);
    }
  }
}

// elementwise loss on precomputed scores: lv[b] = -logsigmoid(pos[b] - neg[b]) (log_sigmoid_difference_loss.py:14),
// hit[b] = (pos[b] >= neg[b]) (auc_score.py:10); either output may be null
__global__ void __launch_bounds__(256)
bpr_loss_fwd_kernel(const float* __restrict__ pos, const float* __restrict__ neg, int64_t B, float* __restrict__ lv,
                    int32_t* __restrict__ hit) {
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
    const float p = pos[b], n = neg[b];
    if (lv) lv[b] = neg_logsigmoid(p - n);
    if (hit) hit[b] = (p >= n) ? 1 : 0;
  }
}
// gpos[b] = -sigmoid(-(pos-neg)) g[b];  gneg[b] = -gpos[b]
__global__ void __launch_bounds__(256)
bpr_loss_bwd_kernel(const float* __restrict__ pos, const float* __restrict__ neg, const float* __restrict__ g, int64_t B,
                    float* __restrict__ gpos, float* __restrict__ gneg) {
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
    const float c = -g[b] / (1.0f + expf(pos[b] - neg[b]));
    gpos[b] = c;
    gneg[b] = -c;
  }
}

// ---- device-side BPR triple sampler ----------------------------------------------------------
// Replaces PairWiseDataset.__iter__ (divrec/datasets/base_datasets.py:70-107): per user, max_sampled positives drawn
// with replacement from the user's items and max_sampled negatives drawn with replacement from
// items - positives - frozen (random.choices, :88-90), then the cross product in the reference's order (positive-major,
// :92-107).  The reference scans all interactions per user (O(U nnz)) and builds Python sets; here both item sets come
// as CSR rows (sorted, unique).  Draws are counter-based (splitmix64 of seed, user, draw index): reproducible for a seed,
// not bit-compatible with CPython's Mersenne twister -- parity is distributional.  A negative is drawn WITHOUT
// rejection: r uniform in [0, nI - n_excl), the r-th item not in the exclusion list is r + t with t the smallest index
// such that excl[t] - t > r (binary search).  Users without a positive or without a negative get (-1, -1, -1) rows
// (the reference raises IndexError there).
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ long long uniform_below(unsigned long long h, long long n) {
  return (long long)__umul64hi(h, (unsigned long long)n);  // floor(h n / 2^64)
}

__global__ void __launch_bounds__(256)
bpr_sample_kernel(const int64_t* __restrict__ pos_rowptr, const int32_t* __restrict__ pos_items,
                  const int64_t* __restrict__ excl_rowptr, const int32_t* __restrict__ excl_items, int64_t nU, int64_t nI,
                  int m, unsigned long long seed, int64_t* __restrict__ out_u, int64_t* __restrict__ out_p,
                  int64_t* __restrict__ out_n) {
  extern __shared__ int64_t samp[];  // per warp: [m] positives | [m] negatives
  const uint32_t lane = lane_id();
  const int warp = threadIdx.x >> 5;
  int64_t* sp = samp + (size_t)warp * 2 * m;
  int64_t* sn = sp + m;
  for (int64_t u = (int64_t)blockIdx.x * 8 + warp; u < nU; u += (int64_t)gridDim.x * 8) {
    const int64_t p0 = pos_rowptr[u], np = pos_rowptr[u + 1] - p0;
    const int64_t e0 = excl_rowptr[u], ne = excl_rowptr[u + 1] - e0;
    const int64_t nneg = nI - ne;
    const bool ok = np > 0 && nneg > 0;
    for (int j = lane; j < m; j += 32) {
      int64_t pj = -1, nj = -1;
      if (ok) {
        const unsigned long long base = splitmix64(seed ^ splitmix64((unsigned long long)u));
        pj = (int64_t)pos_items[p0 + uniform_below(splitmix64(base + 2ull * j), np)];
        const int64_t r = uniform_below(splitmix64(base + 2ull * j + 1ull), nneg);
        int64_t lo = 0, hi = ne;  // smallest t in [0, ne] with excl[t] - t > r (t = ne: none)
        while (lo < hi) {
          const int64_t mid = (lo + hi) >> 1;
          if ((int64_t)excl_items[e0 + mid] - mid > r) hi = mid; else lo = mid + 1;
        }
        nj = r + lo;
      }
      sp[j] = pj;
      sn[j] = nj;
    }
    __syncwarp();
    const int64_t out0 = u * (int64_t)m * m;
    for (int t = lane; t < m * m; t += 32) {
      const int a = t / m, b = t - a * m;
      out_u[out0 + t] = ok ? u : -1;
      out_p[out0 + t] = sp[a];
      out_n[out0 + t] = sn[b];
    }
    __syncwarp();
  }
}

inline unsigned warp_grid(int64_t items) {
  int64_t g = (items + 7) / 8;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace

static inline bool rows_v4(int d, const void* a, const void* b) {
  return (d & 3) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
}

cudaError_t launch_bpr_fwd(const float* U, const float* V, int64_t nU, int64_t nI, const int64_t* u, const int64_t* p,
                           const int64_t* n, int64_t B, int d, float* pos, float* neg, float* loss_sum, int32_t* auc_sum,
                           cudaStream_t stream) {
  if (rows_v4(d, U, V)) bpr_fwd_kernel<true><<<warp_grid(B), 256, 0, stream>>>(U, V, nU, nI, u, p, n, B, d, pos, neg, loss_sum, auc_sum);
  else bpr_fwd_kernel<false><<<warp_grid(B), 256, 0, stream>>>(U, V, nU, nI, u, p, n, B, d, pos, neg, loss_sum, auc_sum);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_bpr_bwd(const float* U, const float* V, int64_t nU, int64_t nI, const int64_t* u, const int64_t* p,
                           const int64_t* n, const float* pos, const float* neg, const float* g, float grad_scale, int64_t B,
                           int d, float* gU, float* gV, cudaStream_t stream) {
  const bool v4 = rows_v4(d, U, V) && rows_v4(d, gU, gV);
  if (v4) bpr_bwd_kernel<true><<<warp_grid(B), 256, 0, stream>>>(U, V, nU, nI, u, p, n, pos, neg, g, grad_scale, B, d, gU, gV);
  else bpr_bwd_kernel<false><<<warp_grid(B), 256, 0, stream>>>(U, V, nU, nI, u, p, n, pos, neg, g, grad_scale, B, d, gU, gV);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_bpr_step(const float* U, const float* V, int64_t nU, int64_t nI, const int64_t* u, const int64_t* p,
                            const int64_t* n, int64_t B, int d, float grad_scale, float* pos, float* neg, float* loss_sum,
                            int32_t* auc_sum, float* gU, float* gV, float* item_rows, float* const* slices, int64_t slice_rows,
                            float* auc_f, cudaStream_t stream) {
  // every resident warp loops over several triples, so loads of the next triple overlap the scatter-adds of this one
  int64_t g = (B + 7) / 8;
  if (g > 148 * 8) g = 148 * 8;
#define DIVREC_BPR_STEP(NV)                                                                                                        \
  do {                                                                                                                             \
    if (slices) bpr_step_kernel<NV, true><<<(unsigned)g, 256, 0, stream>>>(U, V, nU, nI, u, p, n, B, d, grad_scale, pos, neg,      \
                                                                            loss_sum, auc_sum, gU, nullptr, item_rows, slices, slice_rows, auc_f); \
    else bpr_step_kernel<NV, false><<<(unsigned)g, 256, 0, stream>>>(U, V, nU, nI, u, p, n, B, d, grad_scale, pos, neg, loss_sum,  \
                                                                     auc_sum, gU, gV, item_rows, nullptr, 1, auc_f);              \
  } while (0)
  if (d <= 128) DIVREC_BPR_STEP(1);
  else if (d <= 256) DIVREC_BPR_STEP(2);
  else if (d <= 512) DIVREC_BPR_STEP(4);
  else DIVREC_BPR_STEP(8);
#undef DIVREC_BPR_STEP
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_bpr_sample(const int64_t* pos_rowptr, const int32_t* pos_items, const int64_t* excl_rowptr,
                              const int32_t* excl_items, int64_t nU, int64_t nI, int m, uint64_t seed, int64_t* out_u,
                              int64_t* out_p, int64_t* out_n, cudaStream_t stream) {
  const size_t smem = (size_t)8 * 2 * m * sizeof(int64_t);
  int64_t g = (nU + 7) / 8;
  if (g > 148 * 8) g = 148 * 8;
  bpr_sample_kernel<<<(unsigned)g, 256, smem, stream>>>(pos_rowptr, pos_items, excl_rowptr, excl_items, nU, nI, m,
                                                        (unsigned long long)seed, out_u, out_p, out_n);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_bpr_apply_rows(const float* rows, const int64_t* p, const int64_t* n, int64_t N, int64_t nI, int d, float* gV,
                                  cudaStream_t stream) {
  int64_t g = (N + 7) / 8;
  if (g > 148 * 8) g = 148 * 8;
  bpr_apply_rows_kernel<<<(unsigned)g, 256, 0, stream>>>(rows, p, n, N, nI, d, gV);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_pair_scores(const float* U, const float* V, int64_t nU, int64_t nI, const int64_t* uid, const int64_t* iid,
                               int64_t L, int d, float* out, cudaStream_t stream) {
  if (rows_v4(d, U, V)) pair_scores_kernel<true><<<warp_grid(L), 256, 0, stream>>>(U, V, nU, nI, uid, iid, L, d, out);
  else pair_scores_kernel<false><<<warp_grid(L), 256, 0, stream>>>(U, V, nU, nI, uid, iid, L, d, out);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_pair_scores_bwd(const float* U, const float* V, int64_t nU, int64_t nI, const int64_t* uid, const int64_t* iid,
                                   const float* g, int64_t L, int d, float* gU, float* gV, cudaStream_t stream) {
  const bool v4 = rows_v4(d, U, V) && rows_v4(d, gU, gV);
  if (v4) pair_scores_bwd_kernel<true><<<warp_grid(L), 256, 0, stream>>>(U, V, nU, nI, uid, iid, g, L, d, gU, gV);
  else pair_scores_bwd_kernel<false><<<warp_grid(L), 256, 0, stream>>>(U, V, nU, nI, uid, iid, g, L, d, gU, gV);
  count_launch();
  return cudaGetLastError();
}

cudaError_t read_bad_index_count(unsigned long long* host_out, int reset, cudaStream_t stream) {
  cudaError_t e = cudaMemcpyFromSymbolAsync(host_out, g_bad_ids, sizeof(unsigned long long), 0, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e == cudaSuccess && reset) {
    const unsigned long long zero = 0ull;
    e = cudaMemcpyToSymbolAsync(g_bad_ids, &zero, sizeof(zero), 0, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  }
  return e;
}

cudaError_t launch_bpr_loss_fwd(const float* pos, const float* neg, int64_t B, float* lv, int32_t* hit, cudaStream_t stream) {
  int64_t g = (B + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  bpr_loss_fwd_kernel<<<(unsigned)g, 256, 0, stream>>>(pos, neg, B, lv, hit);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_bpr_loss_bwd(const float* pos, const float* neg, const float* g, int64_t B, float* gpos, float* gneg,
                                cudaStream_t stream) {
  int64_t gr = (B + 255) / 256;
  if (gr > 148 * 8) gr = 148 * 8;
  bpr_loss_bwd_kernel<<<(unsigned)gr, 256, 0, stream>>>(pos, neg, g, B, gpos, gneg);
  count_launch();
  return cudaGetLastError();
}

}  // namespace divrec
