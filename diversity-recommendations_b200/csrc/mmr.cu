// mmr.cu -- greedy maximal-marginal-relevance re-rank of each user's top-C candidates (diversity-increasing path).
//
// The reference ships no re-ranking code (its README.md:3-5 only promises "tools for increasing ... diversity"; the one
// diversity-optimising loop, divrec/train/utils.py:167-216, cannot run).  Semantics are therefore the oracle's
// (oracle/divrec_oracle.py::mmr_rerank): candidates arrive sorted by (score desc, id asc); the first
// pick is candidate 0; then, k - 1 times, pick  argmax_c  lambda * score_c - (1 - lambda) * max_{j picked} cos(V_c, V_j)
// over the unpicked real candidates (fp32, ties -> lowest candidate rank).  Padding ids (< 0) are never picked.
//
// mmr_lazy_kernel (the product path): LAZY greedy.  A candidate's objective can only fall as picks are added (its
// max-similarity only grows), so the objective evaluated against the first `seen_c` picks is an upper bound of the
// current one (before any evaluation: cos >= -1).  Per step every warp nominates the candidate with the largest bound
// among those it owns; stale nominees are brought up to date (dot products with the picks they have not seen yet) and
// the nomination repeats until every warp's top candidate is exact; the block then takes the best of the eight exact
// champions, ties -> lowest rank -- the same pick as the eager arg-max, with only the candidates that can still win
// ever being gathered and evaluated (eager costs k C d FMA per user and re-reads every candidate row k times).
// The refreshes of a round are done by the WHOLE block, one thread per (nominee, unseen pick) pair, so a candidate
// that is far behind (first evaluation, or long ignored) costs one pass instead of a long serial loop in one warp
// while seven wait at the barrier (ncu on the per-warp version: 55 % of the stall samples at that barrier).
// One CTA of 8 warps per user (persistent, 4 CTAs per SM).  Candidate c belongs to warp c % 8; its score, bound and
// state live in that warp's registers.  The normalised picked rows sit in shared memory (row stride d + 4 floats:
// conflict-free 128-bit reads with one lane per pick); a warp keeps the raw row of its current champion in registers
// (cos = <V_c, E_t> / |V_c|: one division per refresh) and copies it to shared memory when it asks for a refresh.
// mmr_kernel: the eager form (any d, any C), kept for shapes the lazy kernel does not take.
#include <cstdlib>

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

// ---------------------------------------------------------------------------------------------------------
// lazy greedy
// ---------------------------------------------------------------------------------------------------------
constexpr int LW = 8;             // warps per user

template <int DT>
__device__ __forceinline__ float4 ld_row4(const void* V, int64_t row, int d, int c4) {
  if (DT == 0) {
    return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(V) + row * d) + c4);
  } else {
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(V) + row * d) + c4);
    return make_float4(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u),
                       __uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xffff0000u));
  }
}
// S = candidate slots per lane (C <= 32 LW S); NV = float4 chunks of a row per lane (d <= 128 NV); DT = table dtype.
// cos(c, t) = <V_c, E_t> / |V_c| with E_t the normalised pick row: the candidate's row stays raw (registers) and the
// maximum over the picks is divided by |V_c| once.
template <int S, int NV, int DT>
__global__ void __launch_bounds__(LW * 32, 4)
mmr_lazy_kernel(const MmrParams p) {
  extern __shared__ __align__(16) float dyn[];  // picks [k][ds] normalised | champion rows [LW][d] raw
  __shared__ float ch_obj[LW];
  __shared__ int ch_idx[LW];
  __shared__ int rq_need[LW];        // refresh requests of the current round: picks the nominee has not seen ...
  __shared__ int rq_cs[LW];          // ... starting at this pick
  __shared__ unsigned rq_max[LW];    // max over those picks of <V_c, E_t> as an ordered key
  const int w = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int C = p.C, d = p.d, k = p.k, ds = d + 4, nvec = d >> 2;
  float* picks = dyn;
  float* cache = dyn + (size_t)k * ds + (size_t)w * d;
  const float lam = p.lambda, om = 1.0f - p.lambda;

  for (int64_t u = blockIdx.x; u < p.nU; u += gridDim.x) {
    // ---- this warp's candidates: c = LW (32 s + lane) + w ----
    float sc[S], ms[S];
    int seen[S], id[S];   // id < 0: not a candidate (padding, or picked)
    int nv = 0;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int c = LW * (s * 32 + (int)lane) + w;
      sc[s] = -INFINITY; ms[s] = -1.0001f; seen[s] = 0; id[s] = -1;  // cos >= -1: the bound before any evaluation
      if (c < C) {
        const int64_t it = p.cand_items[u * C + c];
        if (it >= 0 && it < p.nI) { id[s] = (int)it; sc[s] = p.cand_scores[u * C + c]; ++nv; }
      }
    }
    nv = __reduce_add_sync(0xffffffffu, nv);
    if (lane == 0) rq_need[w] = nv;
    __syncthreads();
    int n_valid = 0;
#pragma unroll
    for (int i = 0; i < LW; ++i) n_valid += rq_need[i];
    const int n_pick = min(k, n_valid);
    int cache_c = -1;      // candidate whose raw row is in this warp's cache / in `e` (warp-uniform)
    float cache_nrm = 1.f; // its norm
    float4 e[NV];          // this lane's chunks of that row

    auto load_row = [&](int c, int item) {
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c4 = (int)lane + 32 * j;
        e[j] = c4 < nvec ? ld_row4<DT>(p.V, item, d, c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        ss = fmaf(e[j].x, e[j].x, ss); ss = fmaf(e[j].y, e[j].y, ss); ss = fmaf(e[j].z, e[j].z, ss); ss = fmaf(e[j].w, e[j].w, ss);
      }
      cache_nrm = sqrtf(warp_sum(ss));
      cache_c = c;
    };

    for (int r = 0; r < k; ++r) {
      if (r >= n_pick) {  // fewer real candidates than k: pad
        if (threadIdx.x == 0) { p.out_items[u * k + r] = -1; if (p.out_scores) p.out_scores[u * k + r] = -INFINITY; }
        continue;
      }
      float best = -INFINITY;
      int bidx = 0x7fffffff;
      if (r > 0) {
        // ---- rounds: every warp nominates its top candidate by bound; the stale nominees are brought up to date by
        // the WHOLE block (one thread per (nominee, unseen pick) pair), until every warp's top candidate is exact ----
        bool fresh = true;       // the nominee has to be (re)determined
        int cs = 0, slot = 0, ln = 0;
        float m_old = 0.f;
        for (;;) {
          if (fresh) {
            best = -INFINITY; bidx = 0x7fffffff;
#pragma unroll
            for (int s = 0; s < S; ++s) {
              if (id[s] >= 0) {
                const float ub = lam * sc[s] - om * ms[s];
                const int c = LW * (s * 32 + (int)lane) + w;
                if (ub > best || (ub == best && c < bidx)) { best = ub; bidx = c; }
              }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const float ov = __shfl_xor_sync(0xffffffffu, best, o);
              const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
              if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
            }
            cs = r;  // "nothing to refresh" when the warp has no candidate left
            if (bidx != 0x7fffffff) {
              slot = (bidx / LW) >> 5; ln = (bidx / LW) & 31;
              int my_seen = 0, my_id = 0;
              float my_ms = 0.f;
#pragma unroll
              for (int s = 0; s < S; ++s) if (s == slot) { my_seen = seen[s]; my_id = id[s]; my_ms = ms[s]; }
              cs = __shfl_sync(0xffffffffu, my_seen, ln);
              m_old = __shfl_sync(0xffffffffu, my_ms, ln);
              if (cs < r && cache_c != bidx) load_row(bidx, __shfl_sync(0xffffffffu, my_id, ln));
            }
            fresh = false;
          }
          const int need = r - cs;  // picks the nominee has not seen (0: exact)
          if (need > 0) {
#pragma unroll
            for (int j = 0; j < NV; ++j) {
              const int c4 = (int)lane + 32 * j;
              if (c4 < nvec) reinterpret_cast<float4*>(cache)[c4] = e[j];
            }
          }
          if (lane == 0) { rq_need[w] = need; rq_cs[w] = cs; rq_max[w] = 0u; ch_obj[w] = best; ch_idx[w] = bidx; }
          __syncthreads();
          int pre[LW + 1];
          pre[0] = 0;
#pragma unroll
          for (int i = 0; i < LW; ++i) pre[i + 1] = pre[i] + rq_need[i];
          const int tot = pre[LW];
          if (tot == 0) break;  // every nominee is exact (block-uniform)
          for (int T = threadIdx.x; T < tot; T += LW * 32) {
            int q = 0, base = 0;
#pragma unroll
            for (int i = 1; i < LW; ++i) { if (T >= pre[i]) { q = i; base = pre[i]; } }
            const int t = rq_cs[q] + (T - base);
            const float4* pr = reinterpret_cast<const float4*>(picks + (size_t)t * ds);
            const float4* cr = reinterpret_cast<const float4*>(dyn + (size_t)k * ds + (size_t)q * d);
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 8
            for (int x = 0; x < nvec; ++x) {
              const float4 a4 = cr[x], b4 = pr[x];
              s0 = fmaf(a4.x, b4.x, s0); s1 = fmaf(a4.y, b4.y, s1); s2 = fmaf(a4.z, b4.z, s2); s3 = fmaf(a4.w, b4.w, s3);
            }
            const float dot = (s0 + s1) + (s2 + s3);
            atomicMax(&rq_max[q], f32_to_key(__float_as_uint(dot + 0.0f)));
          }
          __syncthreads();
          if (need > 0) {
            const float mnew = __uint_as_float(key_to_f32(rq_max[w])) / cache_nrm;
            const float m = cs == 0 ? mnew : fmaxf(m_old, mnew);
            if ((int)lane == ln) {
#pragma unroll
              for (int s = 0; s < S; ++s) if (s == slot) { ms[s] = m; seen[s] = r; }
            }
            // bound unchanged: the nominee is still this warp's top candidate, now exact; otherwise nominate again
            fresh = !(cs != 0 && m == m_old);
            cs = r;
          }
        }
        best = ch_obj[0]; bidx = ch_idx[0];
#pragma unroll
        for (int i = 1; i < LW; ++i) {
          const float ov = ch_obj[i];
          const int oi = ch_idx[i];
          if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
        }
      } else {
        bidx = 0;  // the first pick is the best-scored candidate
      }
      // ---- commit pick j: its owner writes the normalised row to the pick list ----
      const int j = bidx;
      if (j != 0x7fffffff && (j & (LW - 1)) == w) {
        const int slot = (j / LW) >> 5, ln = (j / LW) & 31;
        int my_id = -1;
        float my_sc = -INFINITY;
#pragma unroll
        for (int s = 0; s < S; ++s) if (s == slot) { my_id = id[s]; my_sc = sc[s]; }
        const int item = __shfl_sync(0xffffffffu, my_id, ln);
        if (item >= 0 && cache_c != j) load_row(j, item);
#pragma unroll
        for (int jj = 0; jj < NV; ++jj) {
          const int c4 = (int)lane + 32 * jj;
          if (c4 < nvec) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (item >= 0) o = make_float4(e[jj].x / cache_nrm, e[jj].y / cache_nrm, e[jj].z / cache_nrm, e[jj].w / cache_nrm);
            reinterpret_cast<float4*>(picks + (size_t)r * ds)[c4] = o;
          }
        }
        if ((int)lane == ln) {
#pragma unroll
          for (int s = 0; s < S; ++s) if (s == slot) id[s] = -1;  // picked: no longer a candidate
          // outputs straight from the registers (a dependent global load here would sit on every step's critical path)
          p.out_items[u * k + r] = my_id >= 0 ? (int64_t)my_id : p.cand_items[u * C + j];
          if (p.out_scores) p.out_scores[u * k + r] = my_id >= 0 ? my_sc : p.cand_scores[u * C + j];
        }
      }
      __syncthreads();
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// lazy greedy, ONE WARP PER USER (round 2)
// ---------------------------------------------------------------------------------------------------------
// The block-per-user kernel above spends its time at block barriers (ncu: 55 % of the stall samples, ~10k cycles per
// pick for a handful of dot products).  Lazy greedy needs very little arithmetic per pick -- a nomination (arg-max of
// the bounds) and the refresh of one or two stale nominees -- so here a user is handled by a single warp and nothing
// ever waits for another warp: candidate c = 32 s + lane lives in slot s of its lane (score, max-cosine so far, picks
// seen, item id: registers); the picks made so far are a list of (item id, 1 / |V|) in shared memory, and a refresh
// re-gathers the unseen pick rows from the L2-resident item table (coalesced 512-byte rows, four in flight) instead of
// keeping a 50 KB copy per user in shared memory -- which is what capped the old kernel at four users per SM.  Twenty-odd
// users per SM are in flight and hide each other's latencies.
constexpr int WPB = 4;  // warps (users) per CTA

// Four warp sums for the price of six shuffles: two butterfly steps fold the four partials onto lane groups, three more
// finish.  Afterwards every lane with ((lane >> 4) & 1) | (((lane >> 3) & 1) << 1) == q holds the total of part[q].
__device__ __forceinline__ float warp_sum4(const float (&part)[4], uint32_t lane) {
  float v01 = (lane & 16) ? part[1] : part[0], o01 = (lane & 16) ? part[0] : part[1];
  float v23 = (lane & 16) ? part[3] : part[2], o23 = (lane & 16) ? part[2] : part[3];
  v01 += __shfl_xor_sync(0xffffffffu, o01, 16);
  v23 += __shfl_xor_sync(0xffffffffu, o23, 16);
  float v = (lane & 8) ? v23 : v01, o = (lane & 8) ? v01 : v23;
  v += __shfl_xor_sync(0xffffffffu, o, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
__device__ __forceinline__ int sum4_group(uint32_t lane) { return (int)(((lane >> 4) & 1u) | (((lane >> 3) & 1u) << 1)); }

// per-warp shared memory: candidate state [Cp] x (score, max cosine so far, item id, picks seen as u16) + the pick list
// (14 bytes per candidate: the state is what bounds the users in flight per SM, and the kernel is latency-bound)
__host__ __device__ inline size_t mmr_warp_smem_bytes(int Cp, int k) { return (size_t)Cp * 14 + (size_t)k * 8; }

template <int NV, int DT>
__global__ void __launch_bounds__(WPB * 32)
mmr_warp_kernel(const MmrParams p, int Cp) {
  extern __shared__ __align__(16) uint8_t wsm[];
  const int w = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int C = p.C, d = p.d, k = p.k, nvec = d >> 2;
  uint8_t* base = wsm + (size_t)w * mmr_warp_smem_bytes(Cp, k);
  float* sc = reinterpret_cast<float*>(base);      // [Cp] relevance score (-inf: not a candidate)
  float* ms = sc + Cp;                             // [Cp] max cosine with the picks seen so far (-1.0001 before any)
  int* id = reinterpret_cast<int*>(ms + Cp);       // [Cp] item id
  int* pick_id = id + Cp;                          // [k]
  float* pick_inv = reinterpret_cast<float*>(pick_id + k);
  unsigned short* seen = reinterpret_cast<unsigned short*>(pick_inv + k);  // [Cp] picks the candidate has been evaluated against
  const float lam = p.lambda, om = 1.0f - p.lambda;
  const int64_t warp0 = (int64_t)blockIdx.x * WPB + w, nwarps = (int64_t)gridDim.x * WPB;

  for (int64_t u = warp0; u < p.nU; u += nwarps) {
    __syncwarp();
    int nv = 0;
    for (int c = lane; c < Cp; c += 32) {
      float s_ = -INFINITY;
      int i_ = -1;
      if (c < C) {
        const int64_t it = p.cand_items[u * C + c];
        if (it >= 0 && it < p.nI) { i_ = (int)it; s_ = p.cand_scores[u * C + c]; ++nv; }
      }
      sc[c] = s_; id[c] = i_; ms[c] = -1.0001f; seen[c] = 0;   // score -inf: padding (or, later, picked)
    }
    nv = __reduce_add_sync(0xffffffffu, nv);
    __syncwarp();
    const int n_pick = min(k, nv);

    for (int r = 0; r < k; ++r) {
      if (r >= n_pick) {  // fewer real candidates than k: pad
        if (lane == 0) { p.out_items[u * k + r] = -1; if (p.out_scores) p.out_scores[u * k + r] = -INFINITY; }
        continue;
      }
      int j = 0;  // the first pick is the best-scored candidate
      while (r > 0) {
        // ---- nominate: the candidate with the largest bound, ties -> lowest rank ----
        float best = -INFINITY;
        int bidx = 0x7fffffff;
        for (int c = lane; c < Cp; c += 32) {
          const float ub = lam * sc[c] - om * ms[c];  // -inf for padding and picked candidates
          if (ub > best) { best = ub; bidx = c; }     // ascending c per lane: the first maximum is the lowest rank
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
          if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
        }
        j = bidx;
        const int cs = seen[j];
        if (cs >= r) break;   // the nominee's bound is exact: it is the eager arg-max
        // ---- refresh: cosines with the picks the nominee has not seen; its row and the pick rows are re-gathered from
        // the L2-resident item table, all loads of a batch in flight together ----
        const int item = id[j];
        float4 e[NV];
#pragma unroll
        for (int jj = 0; jj < NV; ++jj) {
          const int c4 = (int)lane + 32 * jj;
          e[jj] = c4 < nvec ? ld_row4<DT>(p.V, item, d, c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float nn = 0.f;   // |V_c|^2 from the row that is in flight anyway
#pragma unroll
        for (int jj = 0; jj < NV; ++jj) { nn = fmaf(e[jj].x, e[jj].x, nn); nn = fmaf(e[jj].y, e[jj].y, nn); nn = fmaf(e[jj].z, e[jj].z, nn); nn = fmaf(e[jj].w, e[jj].w, nn); }
        float mx = -INFINITY;
        for (int t0 = cs; t0 < r; t0 += 4) {
          float part[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int t = min(t0 + q, r - 1);
            const int pit = pick_id[t];
            float a = 0.f;
#pragma unroll
            for (int jj = 0; jj < NV; ++jj) {
              const int c4 = (int)lane + 32 * jj;
              if (c4 < nvec) {
                const float4 x = ld_row4<DT>(p.V, pit, d, c4);
                a = fmaf(e[jj].x, x.x, a); a = fmaf(e[jj].y, x.y, a); a = fmaf(e[jj].z, x.z, a); a = fmaf(e[jj].w, x.w, a);
              }
            }
            part[q] = a;
          }
          const float v = warp_sum4(part, lane);
          const int t = t0 + sum4_group(lane);
          if (t < r) mx = fmaxf(mx, v * pick_inv[t]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); nn += __shfl_xor_sync(0xffffffffu, nn, o); }
        if (lane == 0) { ms[j] = fmaxf(ms[j], mx / sqrtf(nn)); seen[j] = (unsigned short)r; }
        __syncwarp();
      }
      // ---- commit pick j ----
      if (lane == 0) {
        const int item = id[j];
        pick_id[r] = item;
        p.out_items[u * k + r] = (int64_t)item;
        if (p.out_scores) p.out_scores[u * k + r] = sc[j];
        sc[j] = -INFINITY;  // picked: no longer a candidate
      }
      __syncwarp();
      if (r > 0) {  // 1 / |V| of the new pick (its row is L2- or L1-hot: it was the last refreshed nominee more often than not)
        const int pit = pick_id[r];
        float n1 = 0.f;
#pragma unroll
        for (int jj = 0; jj < NV; ++jj) {
          const int c4 = (int)lane + 32 * jj;
          if (c4 < nvec) { const float4 x = ld_row4<DT>(p.V, pit, d, c4); n1 = fmaf(x.x, x.x, n1); n1 = fmaf(x.y, x.y, n1); n1 = fmaf(x.z, x.z, n1); n1 = fmaf(x.w, x.w, n1); }
        }
        n1 = warp_sum(n1);
        if (lane == 0) pick_inv[r] = 1.0f / sqrtf(n1);
        __syncwarp();
      }
      if (r == 0) {
        // ---- cosines of all candidates with the FIRST pick, in one batched pass ----
        // Before any evaluation a candidate's bound is lambda score + (1 - lambda) (cos >= -1): higher than every exact
        // objective, so lazy greedy would refresh all C candidates one by one during the first picks -- C serial row
        // loads per user.  Here the rows stream through once, four in flight, against the first pick's row.
        const int item0 = pick_id[0];
        float4 e0[NV];
#pragma unroll
        for (int jj = 0; jj < NV; ++jj) {
          const int c4 = (int)lane + 32 * jj;
          e0[jj] = c4 < nvec ? ld_row4<DT>(p.V, item0, d, c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float n0 = 0.f;
#pragma unroll
        for (int jj = 0; jj < NV; ++jj) { n0 = fmaf(e0[jj].x, e0[jj].x, n0); n0 = fmaf(e0[jj].y, e0[jj].y, n0); n0 = fmaf(e0[jj].z, e0[jj].z, n0); n0 = fmaf(e0[jj].w, e0[jj].w, n0); }
        const float inv0 = 1.0f / sqrtf(warp_sum(n0));
        if (lane == 0) pick_inv[0] = inv0;
        for (int c0 = 0; c0 < C; c0 += 4) {
          float pd[4], ps[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int it = (c0 + q < C && sc[c0 + q] > -INFINITY) ? id[c0 + q] : -1;
            float a = 0.f, b2 = 0.f;
            if (it >= 0) {
#pragma unroll
              for (int jj = 0; jj < NV; ++jj) {
                const int c4 = (int)lane + 32 * jj;
                if (c4 < nvec) {
                  const float4 x = ld_row4<DT>(p.V, it, d, c4);
                  a = fmaf(e0[jj].x, x.x, a); a = fmaf(e0[jj].y, x.y, a); a = fmaf(e0[jj].z, x.z, a); a = fmaf(e0[jj].w, x.w, a);
                  b2 = fmaf(x.x, x.x, b2); b2 = fmaf(x.y, x.y, b2); b2 = fmaf(x.z, x.z, b2); b2 = fmaf(x.w, x.w, b2);
                }
              }
            }
            pd[q] = a; ps[q] = b2;
          }
          const float vd = warp_sum4(pd, lane), vs = warp_sum4(ps, lane);
          const int q = sum4_group(lane);
          if ((lane & 7u) == 0u && c0 + q < C && sc[c0 + q] > -INFINITY) {  // one lane of each group writes its candidate
            ms[c0 + q] = vd * inv0 / sqrtf(vs);
            seen[c0 + q] = 1;
          }
        }
        __syncwarp();
      }
    }
  }
}

template <int NV>
cudaError_t launch_mmr_warp_t(const MmrParams& p, int grid, int Cp, size_t smem, cudaStream_t stream) {
  cudaError_t e;
  if (p.dtype == 0) {
    e = cudaFuncSetAttribute(mmr_warp_kernel<NV, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    mmr_warp_kernel<NV, 0><<<grid, WPB * 32, smem, stream>>>(p, Cp);
  } else {
    e = cudaFuncSetAttribute(mmr_warp_kernel<NV, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    mmr_warp_kernel<NV, 1><<<grid, WPB * 32, smem, stream>>>(p, Cp);
  }
  return cudaGetLastError();
}

template <int S, int NV>
cudaError_t launch_mmr_lazy_t(const MmrParams& p, int grid, size_t smem, cudaStream_t stream) {
  cudaError_t e;
  if (p.dtype == 0) {
    e = cudaFuncSetAttribute(mmr_lazy_kernel<S, NV, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    mmr_lazy_kernel<S, NV, 0><<<grid, LW * 32, smem, stream>>>(p);
  } else {
    e = cudaFuncSetAttribute(mmr_lazy_kernel<S, NV, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    mmr_lazy_kernel<S, NV, 1><<<grid, LW * 32, smem, stream>>>(p);
  }
  return cudaGetLastError();
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
  // warp-per-user lazy greedy: 128-bit rows, per-warp state within shared memory
  const int Cp = (p.C + 31) & ~31;
  const size_t warp_smem = (size_t)WPB * mmr_warp_smem_bytes(Cp, p.k);
  if ((p.d & 3) == 0 && p.d <= 512 && p.k < 65536 && warp_smem <= 100 * 1024 && p.nI < ((int64_t)1 << 31) && !debug_env("DIVREC_DEBUG_MMR_BLOCK") &&
      !debug_env("DIVREC_DEBUG_MMR_EAGER")) {
    const int NV = p.d <= 128 ? 1 : p.d <= 256 ? 2 : 4;
    int per_sm = (int)((size_t)232448 / (warp_smem + 1024));
    if (per_sm > 12) per_sm = 12;
    int64_t g = (int64_t)148 * per_sm;
    if (g * WPB > p.nU) g = (p.nU + WPB - 1) / WPB;
    cudaError_t e = NV == 1 ? launch_mmr_warp_t<1>(p, (int)g, Cp, warp_smem, stream)
                  : NV == 2 ? launch_mmr_warp_t<2>(p, (int)g, Cp, warp_smem, stream)
                            : launch_mmr_warp_t<4>(p, (int)g, Cp, warp_smem, stream);
    count_launch();
    return e;
  }
  // block-per-user lazy greedy (round 1; kept behind a debug knob for comparison): pick list + caches within shared memory
  const size_t lazy_smem = ((size_t)p.k * (size_t)(p.d + 4) + (size_t)LW * p.d) * sizeof(float);
  if ((p.d & 3) == 0 && p.d <= 512 && p.C <= 1024 && lazy_smem <= 200 * 1024 && p.nI < ((int64_t)1 << 31) &&
      !debug_env("DIVREC_DEBUG_MMR_EAGER")) {
    int per_sm = (int)((size_t)232448 / (lazy_smem + 1024 + 160));  // per-SM shared memory / (dynamic + reserved + static)
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    int64_t g = (int64_t)148 * per_sm;
    if (g > p.nU) g = p.nU;
    const int S = p.C <= 256 ? 1 : p.C <= 512 ? 2 : 4;   // candidate slots per lane: C <= 32 LW S
    const int NV = p.d <= 128 ? 1 : p.d <= 256 ? 2 : 4;
    cudaError_t e = cudaErrorInvalidValue;
#define DIVREC_MMR(SS, NN) if (S == SS && NV == NN) e = launch_mmr_lazy_t<SS, NN>(p, (int)g, lazy_smem, stream)
    DIVREC_MMR(1, 1); DIVREC_MMR(2, 1); DIVREC_MMR(4, 1);
    DIVREC_MMR(1, 2); DIVREC_MMR(2, 2); DIVREC_MMR(4, 2);
    DIVREC_MMR(1, 4); DIVREC_MMR(2, 4); DIVREC_MMR(4, 4);
#undef DIVREC_MMR
    count_launch();
    return e;
  }
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
