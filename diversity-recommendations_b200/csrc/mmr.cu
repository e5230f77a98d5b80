// mmr.cu -- greedy maximal-marginal-relevance re-rank of each user's top-C candidates (diversity-increasing path).
//
// The reference ships no re-ranking code (its README.md:3-5 only promises "tools for increasing ... diversity"; the one
// diversity-optimising loop, divrec/train/utils.py:167-216, cannot run).  Semantics are therefore the oracle's
// (oracle/divrec_oracle.py::mmr_rerank): candidates arrive sorted by (score desc, id asc); the first
// pick is candidate 0; then, k - 1 times, pick  argmax_c  lambda * score_c - (1 - lambda) * max_{j picked} cos(V_c, V_j)
// over the unpicked real candidates (fp32, ties -> lowest candidate rank).  Padding ids (< 0) are never picked.
//
// One CTA per user (persistent, grid-stride).  The normalised candidate rows E_c = V_c / |V_c| are written once to a
// per-CTA scratch slab (C x d fp32, L2-resident: 296 CTAs x 256 KB at C = 500, d = 128); every step broadcasts the
// newly picked row through shared memory, each thread updates max-similarity of its candidates with one dot product
// per candidate (128-bit loads) and the block reduces the argmax.  Work per user: k C d FMA + the C row gather.
#include "common.cuh"
#include "kernels.h"

namespace divrec {

namespace {

constexpr int MT = 256;

struct MmrSmem {
  float red_val[MT / 32];
  int red_idx[MT / 32];
  int pick;
};

__device__ __forceinline__ float ldv(const void* V, int dtype, int64_t idx) {
  if (dtype == 0) return __ldg(reinterpret_cast<const float*>(V) + idx);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(V)[idx]);
}

__global__ void __launch_bounds__(MT)
mmr_kernel(const MmrParams p) {
  extern __shared__ __align__(16) float dyn[];  // [d] picked row | [C] score | [C] maxsim | [C] state (0 free, 1 picked, 2 invalid)
  __shared__ MmrSmem sm;
  const int tid = threadIdx.x;
  const uint32_t lane = lane_id();
  const int warp = tid >> 5;
  const int C = p.C, d = p.d, k = p.k;
  float* row_j = dyn;
  float* score = dyn + ((d + 3) & ~3);
  float* maxsim = score + C;
  int* state = reinterpret_cast<int*>(maxsim + C);
  float* E = p.scratch + (size_t)blockIdx.x * (size_t)C * d;
  const float lam = p.lambda, om = 1.0f - p.lambda;
  const bool vec = (d & 3) == 0;

  for (int64_t u = blockIdx.x; u < p.nU; u += gridDim.x) {
    __syncthreads();
    // ---- normalised candidate rows -> scratch (one warp per candidate) ----
    for (int c = warp; c < C; c += MT / 32) {
      const int64_t id = p.cand_items[u * C + c];
      const bool ok = id >= 0 && id < p.nI;
      float ss = 0.f;
      if (ok) for (int x = lane; x < d; x += 32) { const float v = ldv(p.V, p.dtype, id * d + x); ss = fmaf(v, v, ss); }
      ss = warp_sum(ss);
      const float nrm = sqrtf(ss);
      for (int x = lane; x < d; x += 32) E[(size_t)c * d + x] = ok ? ldv(p.V, p.dtype, id * d + x) / nrm : 0.f;
      if (lane == 0) {
        score[c] = p.cand_scores[u * C + c];
        maxsim[c] = -INFINITY;
        state[c] = ok ? 0 : 2;
      }
    }
    __syncthreads();
    __threadfence_block();
    int n_valid = 0;
    for (int c = tid; c < C; c += MT) n_valid += state[c] == 0 ? 1 : 0;
    // block sum of n_valid
    {
      int v = n_valid;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) sm.red_idx[warp] = v;
      __syncthreads();
      int t = 0;
      for (int w2 = 0; w2 < MT / 32; ++w2) t += sm.red_idx[w2];
      n_valid = t;
      __syncthreads();
    }
    const int n_pick = min(k, n_valid);
    for (int r = 0; r < k; ++r) {
      if (r >= n_pick) {  // fewer real candidates than k: pad
        if (tid == 0) { p.out_items[u * k + r] = -1; if (p.out_scores) p.out_scores[u * k + r] = -INFINITY; }
        continue;
      }
      // ---- choose ----
      int j;
      if (r == 0) {
        j = 0;
      } else {
        float best = -INFINITY;
        int bidx = 0x7fffffff;
        for (int c = tid; c < C; c += MT) {
          if (state[c] == 0) {
            const float obj = lam * score[c] - om * maxsim[c];
            if (obj > best || (obj == best && c < bidx)) { best = obj; bidx = c; }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
          if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
        }
        if (lane == 0) { sm.red_val[warp] = best; sm.red_idx[warp] = bidx; }
        __syncthreads();
        if (tid == 0) {
          float b = sm.red_val[0];
          int bi = sm.red_idx[0];
          for (int w2 = 1; w2 < MT / 32; ++w2)
            if (sm.red_val[w2] > b || (sm.red_val[w2] == b && sm.red_idx[w2] < bi)) { b = sm.red_val[w2]; bi = sm.red_idx[w2]; }
          sm.pick = bi;
        }
        __syncthreads();
        j = sm.pick;
      }
      // ---- record, broadcast the picked row ----
      if (tid == 0) {
        state[j] = 1;
        p.out_items[u * k + r] = p.cand_items[u * C + j];
        if (p.out_scores) p.out_scores[u * k + r] = score[j];
      }
      for (int x = tid; x < d; x += MT) row_j[x] = E[(size_t)j * d + x];
      __syncthreads();
      // ---- max-similarity update: one dot product per candidate ----
      for (int c = tid; c < C; c += MT) {
        if (state[c] != 0) continue;
        const float* ec = E + (size_t)c * d;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        if (vec) {
          for (int x = 0; x < d; x += 4) {
            const float4 a = *reinterpret_cast<const float4*>(ec + x);
            const float4 b = *reinterpret_cast<const float4*>(row_j + x);
            s0 = fmaf(a.x, b.x, s0); s1 = fmaf(a.y, b.y, s1); s2 = fmaf(a.z, b.z, s2); s3 = fmaf(a.w, b.w, s3);
          }
        } else {
          for (int x = 0; x < d; ++x) s0 = fmaf(ec[x], row_j[x], s0);
        }
        maxsim[c] = fmaxf(maxsim[c], (s0 + s1) + (s2 + s3));
      }
      __syncthreads();
    }
  }
}

}  // namespace

size_t mmr_scratch_bytes(int C, int d, int* grid_out) {
  int g = 296;
  const size_t per = (size_t)C * d * sizeof(float);
  while (g > 8 && per * g > ((size_t)512 << 20)) g /= 2;
  if (grid_out) *grid_out = g;
  return per * g;
}

cudaError_t launch_mmr(MmrParams p, cudaStream_t stream) {
  int grid = 0;
  mmr_scratch_bytes(p.C, p.d, &grid);
  if ((int64_t)grid > p.nU) grid = (int)p.nU;
  const size_t smem = ((size_t)((p.d + 3) & ~3) + 3 * (size_t)p.C) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(mmr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  mmr_kernel<<<grid, MT, smem, stream>>>(p);
  count_launch();
  return cudaGetLastError();
}

}  // namespace divrec
