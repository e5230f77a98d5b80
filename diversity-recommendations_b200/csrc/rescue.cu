// rescue.cu -- exact SIMT scorer + radix-select top-k for single user rows.
//
// The safety net of the tensor-core path (and DIVREC_MODE_SIMT as a whole): rows that
// select_topk.cu could not certify -- threshold hint too high, candidate list overflow (floods of
// near-ties, e.g. an all-zero user vector or thousands of identical item vectors), more candidates
// than fit in shared memory -- are re-done here from scratch, in plain fp32, one CTA per row:
//   1. every item is scored exactly, E = sum_c U[row,c] V[i,c] (the arithmetic of
//      divrec/models/matrix_factorization.py:26-28), and stored as an ordered integer key;
//   2. the row's seen items are removed (divrec/datasets/base_datasets.py:143-151);
//   3. an 8-bit-digit radix select finds the k-th largest key exactly, ties at the k-th key are
//      resolved towards the lowest item id;
//   4. the winners are sorted by (E desc, id asc): divrec/train/utils.py:72-76 in canonical tie order.
// Throughput is irrelevant by design (a handful of rows per call); correctness for ANY input is the point.
#include "common.cuh"
#include "kernels.h"

namespace divrec {

namespace {

constexpr int RT = 1024;       // threads per CTA of the selection kernel (one CTA per row: its passes are latency, so go wide)
constexpr int RT_SCORE = 256;  // threads per CTA of the chip-wide scoring kernel
#ifndef DIVREC_RESCUE_CHUNKS
#define DIVREC_RESCUE_CHUNKS 148  // item ranges per row in the chip-wide scoring kernel: one wave of CTAs
#endif
constexpr int RSORT = 512;     // >= MAX_K
static_assert(RSORT >= MAX_K, "sort buffer covers the largest k");

struct RescueSmem {
  unsigned int hist[256];
  unsigned int warp_tot[RT / 32];
  unsigned int prefix, kmask, k_rem, n_win, tie_base, n_eq;
  unsigned long long srt[RSORT];
  float inv[RSORT];
  float red[RT / 32];
};

__device__ __forceinline__ float rd(const void* T, int dtype, int64_t idx) {
  if (dtype == 0) return __ldg(reinterpret_cast<const float*>(T) + idx);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(T)[idx]);
}

// Exact fp32 scores of items [i0, i1) against the user row staged in u_s -> ordered keys (0 = not recommendable).
// The summation order replays the select kernel's exactly (same partial sums, same butterfly), so a row gets the same
// bits whichever kernel scores it: item-sharded, user-sharded, tensor and SIMT routes agree bit for bit.
template <int NT>
__device__ __forceinline__ void score_range(const RescueParams& p, const float* u_s, uint32_t* keys, int64_t i0, int64_t i1,
                                            int tid) {
  const int d = p.d;
  const bool vec = (d & 3) == 0 && d <= 256;  // select_topk.cu: 8 lanes per candidate, float4 chunk c4 = sub + 8 j
  if (vec) {
    // exactly select_topk.cu's arithmetic and layout: 8 lanes per item (coalesced 128-byte pieces of its row), lane
    // `sub` accumulates the chunks sub, sub + 8, ... in one fma chain, the partial sums are combined by the xor 1, 2, 4
    // butterfly and lane 0 of the group keeps the result
    const int sub = tid & 7, grp = tid >> 3, nvec = d >> 2;
    for (int64_t base = i0; base < i1; base += NT / 8) {  // uniform trip count: the shuffles below need every lane
      const int64_t i = base + grp;
      const bool on = i < i1;
      float s = 0.f;
      for (int c4 = sub; c4 < nvec; c4 += 8) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (on) {
          if (p.dtype == 0) {
            x = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.V) + i * d) + c4);
          } else {
            const uint2 raw = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.V) + i * d) + c4);
            x = make_float4(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u),
                            __uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xffff0000u));
          }
        }
        const float4 y = *reinterpret_cast<const float4*>(u_s + 4 * c4);
        s = fmaf(y.x, x.x, s); s = fmaf(y.y, x.y, s); s = fmaf(y.z, x.z, s); s = fmaf(y.w, x.w, s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (on && sub == 0) {
        uint32_t kk = f32_to_key(__float_as_uint(s + 0.0f));
        if (kk <= KEY_NEG_INF || !(s == s)) kk = 0;
        keys[i] = kk;
      }
    }
    return;
  }
  for (int64_t i = i0 + tid; i < i1; i += NT) {
    float part[32];  // scalar path: lane l takes columns l, l + 32, ...; butterfly xor 16, 8, 4, 2, 1
#pragma unroll
    for (int l = 0; l < 32; ++l) {
      float acc = 0.f;
      for (int c = l; c < d; c += 32) acc = fmaf(u_s[c], rd(p.V, p.dtype, i * d + c), acc);
      part[l] = acc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int l = 0; l < o; ++l) part[l] = part[l] + part[l + o];
    const float s = part[0];
    uint32_t kk = f32_to_key(__float_as_uint(s + 0.0f));
    if (kk <= KEY_NEG_INF || !(s == s)) kk = 0;
    keys[i] = kk;
  }
}

// Scores of the first `gridDim.y` uncertified rows, spread over the whole chip (blockIdx.x = item range): one CTA
// scoring a whole catalogue on its own takes 0.7 ms for 27k items and 65 ms for 1.25M, which made every uncertified
// row expensive; rescue_kernel then only masks, selects and sorts these rows.
__global__ void __launch_bounds__(RT_SCORE)
rescue_score_kernel(const RescueParams p, int64_t items_per_cta) {
  extern __shared__ __align__(16) float u_s[];
  const int f = blockIdx.y;
  if (f >= *p.fail_count) return;
  const int64_t row = p.fail_rows[f];
  for (int c = threadIdx.x; c < p.d; c += RT_SCORE) u_s[c] = rd(p.U, p.dtype, row * p.d + c);
  __syncthreads();
  uint32_t* keys = p.scratch + (size_t)f * (size_t)(((int64_t)p.nI + 31) & ~(int64_t)31);
  const int64_t i0 = (int64_t)blockIdx.x * items_per_cta;
  const int64_t i1 = i0 + items_per_cta < (int64_t)p.nI ? i0 + items_per_cta : (int64_t)p.nI;
  score_range<RT_SCORE>(p, u_s, keys, i0, i1, threadIdx.x);
}

__global__ void __launch_bounds__(RT)
rescue_kernel(const RescueParams p) {
  extern __shared__ __align__(16) float u_s[];  // [d] the user row
  __shared__ RescueSmem sm;
  const int tid = threadIdx.x;
  const uint32_t lane = lane_id();
  const int warp = tid >> 5;
  const int n_fail = *p.fail_count;
  if (blockIdx.x == 0 && tid == 0 && p.total_fail != nullptr) *p.total_fail += n_fail;  // statistics: rows re-done exactly (launches are stream-ordered)
  const int d = p.d, k = p.k;
  const int64_t nI = p.nI;
  uint32_t* keys = p.scratch + (size_t)blockIdx.x * (size_t)((nI + 31) & ~(int64_t)31);

  for (int f = blockIdx.x; f < n_fail; f += gridDim.x) {
    const int64_t row = p.fail_rows[f];
    __syncthreads();
    for (int c = tid; c < d; c += RT) u_s[c] = rd(p.U, p.dtype, row * d + c);
    __syncthreads();
    // 1. exact scores -> ordered keys; the first p.prescored rows were scored by rescue_score_kernel into this slot
    if (f >= p.prescored) score_range<RT>(p, u_s, keys, 0, nI, tid);
    __syncthreads();
    // the selection passes below read every key five times with little work per key: from shared memory when the
    // catalogue fits (independent loads, no synchronisation in the copy), else from the global scratch
    uint32_t* kr = keys;
    if (p.keys_in_smem) {
      uint32_t* ks = reinterpret_cast<uint32_t*>(u_s + ((d + 3) & ~3));
      for (int64_t i = tid; i < nI; i += RT) ks[i] = keys[i];
      kr = ks;
      __syncthreads();
    }
    // 2. seen-item mask
    if (p.seen_rowptr != nullptr) {
      const int64_t s0 = p.seen_rowptr[row], s1 = p.seen_rowptr[row + 1];
      for (int64_t s = s0 + tid; s < s1; s += RT) {
        const int64_t li = (int64_t)p.seen_items[s] - p.item_id_offset;
        if (li >= 0 && li < nI) kr[li] = 0;
      }
    }
    __syncthreads();
    // 3. radix select of the k-th largest non-zero key, most significant digit first
    if (tid == 0) { sm.prefix = 0; sm.kmask = 0; sm.k_rem = (unsigned)k; }
    bool have_k = true;
    for (int pass = 3; pass >= 0; --pass) {
      if (tid < 256) sm.hist[tid] = 0;
      __syncthreads();
      const unsigned prefix = sm.prefix, kmask = sm.kmask;
      for (int64_t i = tid; i < nI; i += RT) {
        const uint32_t kk = kr[i];
        // scores cluster in a few digit values: aggregate equal digits inside the warp before touching shared memory
        const bool on = kk != 0 && (kk & kmask) == prefix;
        const uint32_t digit = (kk >> (8 * pass)) & 255u;
        const uint32_t active = __ballot_sync(__activemask(), on);
        if (on) {
          const uint32_t peers = __match_any_sync(active, digit);
          if ((int)lane == __ffs(peers) - 1) atomicAdd(&sm.hist[digit], (unsigned)__popc(peers));
        }
      }
      __syncthreads();
      if (tid == 0) {
        unsigned rem = sm.k_rem, acc = 0;
        int bin = 255;
        for (; bin >= 0; --bin) {
          if (acc + sm.hist[bin] >= rem) break;
          acc += sm.hist[bin];
        }
        if (bin < 0) {
          sm.k_rem = 0xffffffffu;  // fewer than k recommendable items: take them all
        } else {
          sm.k_rem = rem - acc;
          sm.prefix = prefix | ((unsigned)bin << (8 * pass));
          sm.kmask = kmask | (255u << (8 * pass));
        }
      }
      __syncthreads();
      if (sm.k_rem == 0xffffffffu) { have_k = false; break; }
    }
    const uint32_t T = have_k ? sm.prefix : 1u;             // keep keys > T, and `quota` keys == T (lowest ids)
    const unsigned quota = have_k ? sm.k_rem : 0xffffffffu;  // no limit when everything is taken
    // 4. gather the winners, then sort.  Keys above T are winners in any order (the sort fixes it); keys equal to T
    // compete for `quota` places by lowest id.  Common case: all of them fit -> one unordered sweep with atomic appends.
    int nsort = 32;  // sort buffer in use: the smallest power of two >= k (the winners never exceed k)
    while (nsort < k) nsort <<= 1;
    for (int i = tid; i < nsort; i += RT) sm.srt[i] = 0ull;
    if (tid == 0) { sm.n_win = 0; sm.tie_base = 0; sm.n_eq = 0; }
    __syncthreads();
    {
      unsigned my_eq = 0;
      for (int64_t i = tid; i < nI; i += RT) {
        const uint32_t kk = kr[i];
        if (kk != 0 && (have_k ? kk > T : true)) {
          const unsigned slot = atomicAdd(&sm.n_win, 1u);
          if (slot < (unsigned)nsort) sm.srt[slot] = ((unsigned long long)kk << 32) | (unsigned long long)(~(uint32_t)i);
        }
        my_eq += (have_k && kk == T) ? 1u : 0u;
      }
      if (my_eq) atomicAdd(&sm.n_eq, my_eq);
    }
    __syncthreads();
    const bool ties_fit = sm.n_eq <= quota;
    if (have_k && ties_fit) {
      for (int64_t i = tid; i < nI; i += RT) {
        const uint32_t kk = kr[i];
        if (kk == T) {
          const unsigned slot = atomicAdd(&sm.n_win, 1u);
          if (slot < (unsigned)nsort) sm.srt[slot] = ((unsigned long long)kk << 32) | (unsigned long long)(~(uint32_t)i);
        }
      }
    } else if (have_k) {
      // more ties than places: the `quota` lowest ids win -- ranked by a block-wide ordered scan in id order
      for (int64_t base = 0; base < nI; base += RT) {
        const int64_t i = base + tid;
        const uint32_t kk = i < nI ? kr[i] : 0;
        const bool eq = kk == T;
        const uint32_t eb = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) sm.warp_tot[warp] = __popc(eb);
        __syncthreads();
        unsigned before = sm.tie_base;
        for (int w2 = 0; w2 < warp; ++w2) before += sm.warp_tot[w2];
        const unsigned my_rank = before + __popc(eb & lanemask_lt());
        if (eq && my_rank < quota) {
          const unsigned slot = atomicAdd(&sm.n_win, 1u);
          if (slot < (unsigned)nsort) sm.srt[slot] = ((unsigned long long)kk << 32) | (unsigned long long)(~(uint32_t)i);
        }
        __syncthreads();
        if (tid == 0) { unsigned t = 0; for (int w2 = 0; w2 < RT / 32; ++w2) t += sm.warp_tot[w2]; sm.tie_base += t; }
        __syncthreads();
        if (sm.tie_base >= quota) break;  // block-uniform: every place is taken
      }
    }
    __syncthreads();
    // bitonic sort (descending) of RSORT composites by the whole block
    for (int size = 2; size <= nsort; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        if (tid < nsort / 2) {
          const int a = ((tid & ~(stride - 1)) << 1) | (tid & (stride - 1));
          const int b = a | stride;
          const bool desc = (a & size) == 0;
          const unsigned long long x = sm.srt[a], y = sm.srt[b];
          if (desc ? (x < y) : (x > y)) { sm.srt[a] = y; sm.srt[b] = x; }
        }
        __syncthreads();
      }
    }
    for (int r = tid; r < k; r += RT) {
      const unsigned long long c = sm.srt[r];
      const uint32_t kk = (uint32_t)(c >> 32);
      float sc = -INFINITY;
      int64_t it = -1;
      if (kk != 0) { sc = __uint_as_float(key_to_f32(kk)); it = (int64_t)(~(uint32_t)c) + p.item_id_offset; }
      p.out_scores[row * k + r] = sc;
      p.out_items[row * k + r] = it;
      if (kk != 0 && p.coverage != nullptr) atomicAdd(p.coverage + it, 1);
    }
    if (p.out_ild != nullptr) {
      // cosine ILD of the winners: sum_{a<b} cos = (|sum e|^2 - kv) / 2  (same formula as metrics.cu / select_topk.cu)
      const int n_win = min((int)sm.n_win, k);
      for (int wdx = warp; wdx < n_win; wdx += RT / 32) {
        const int64_t it = (int64_t)(~(uint32_t)sm.srt[wdx]);
        float q = 0.f;
        for (int c = lane; c < d; c += 32) { const float v = rd(p.V, p.dtype, it * d + c); q = fmaf(v, v, q); }
        q = warp_sum(q);
        if (lane == 0) sm.inv[wdx] = 1.0f / sqrtf(q);
      }
      __syncthreads();
      float part = 0.f;
      for (int c = tid; c < d; c += RT) {
        float sc = 0.f;
#pragma unroll 8
        for (int wdx = 0; wdx < n_win; ++wdx) sc = fmaf(rd(p.V, p.dtype, (int64_t)(~(uint32_t)sm.srt[wdx]) * d + c), sm.inv[wdx], sc);
        part = fmaf(sc, sc, part);
      }
      part = warp_sum(part);
      if (lane == 0) sm.red[warp] = part;
      __syncthreads();
      if (tid == 0) {
        float ssum = 0.f;
        for (int w2 = 0; w2 < RT / 32; ++w2) ssum += sm.red[w2];
        const float kvf = (float)n_win;
        p.out_ild[row] = (0.5f * kvf * (kvf - 1.0f) - 0.5f * (ssum - kvf)) / ((float)k * (float)(k - 1));
      }
    }
  }
}

}  // namespace

size_t rescue_scratch_bytes(int64_t nI, int* grid_out) {
  const size_t per = (size_t)((nI + 31) & ~(int64_t)31) * sizeof(uint32_t);
  int64_t g = (int64_t)((size_t)256 << 20) / (int64_t)per;  // at most 256 MB of scratch
  if (g > 296) g = 296;
  if (g < 4) g = 4;
  if (grid_out) *grid_out = (int)g;
  return per * (size_t)g;
}

cudaError_t launch_rescue(const RescueParams& p_in, cudaStream_t stream) {
  RescueParams p = p_in;
  const size_t smem = (size_t)p.d * sizeof(float);
  const size_t keys_bytes = (size_t)(((int64_t)p.nI + 31) & ~(int64_t)31) * sizeof(uint32_t);
  const size_t smem_sel = (size_t)((p.d + 3) & ~3) * sizeof(float) + keys_bytes;
  p.keys_in_smem = smem_sel <= 200 * 1024 ? 1 : 0;
  // the first `slots` uncertified rows are scored by the whole chip (CTAs beyond the device-side count exit at once)
#ifndef DIVREC_RESCUE_SLOTS
#define DIVREC_RESCUE_SLOTS 32   // uncertified rows scored by the whole chip (a step has ~6; the rest fall back to one CTA per row)
#endif
  const int slots = p.grid < DIVREC_RESCUE_SLOTS ? p.grid : DIVREC_RESCUE_SLOTS;
  // item ranges of >= 128 items, about one wave of CTAs per row: one uncertified row (the usual count) is scored by the whole chip
  int64_t chunks = ((int64_t)p.nI + 127) / 128;
  if (chunks > DIVREC_RESCUE_CHUNKS) chunks = DIVREC_RESCUE_CHUNKS;
  const int64_t items_per_cta = ((int64_t)p.nI + chunks - 1) / chunks;
  p.prescored = slots;
  rescue_score_kernel<<<dim3((unsigned)chunks, (unsigned)slots), RT_SCORE, smem, stream>>>(p, items_per_cta);
  count_launch();
  const size_t smem_main = p.keys_in_smem ? smem_sel : smem;
  cudaError_t e = cudaFuncSetAttribute(rescue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem_main > 48 * 1024 ? smem_main : 48 * 1024));
  if (e != cudaSuccess) return e;
  rescue_kernel<<<p.grid, RT, smem_main, stream>>>(p);
  count_launch();
  return cudaGetLastError();
}

}  // namespace divrec
