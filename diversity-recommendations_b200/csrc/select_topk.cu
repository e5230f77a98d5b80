// select_topk.cu -- exact per-user top-k over the candidate groups left by score_filter.cu, and the
// n-way merge of per-shard top-k lists (item-sharded catalogue).
//
// Replaces `negative_item_id[torch.argsort(scores, descending=True)][:k]` (divrec/train/utils.py:72-76)
// and the seen-item mask of RankingDataset.__iter__ (divrec/datasets/base_datasets.py:143-151).
// The reference's argsort is not stable, so its order on exact score ties is arbitrary; the order
// produced here is the canonical (score desc, item id asc), equal to the reference when it is tie-free.
//
// One warp per user row:
//   A  enumerate the captured elements (tensor-core score A >= the row's threshold) into shared memory;
//      rows that captured more than SEL_CAP elements (small catalogues, no hint) are first cut at
//      X - 2 delta, X = the (k + seen_u)-th largest A (at least k of those are unmasked);
//   B  apply the seen-item mask (binary search of each candidate in the row's sorted seen list);
//   C  T = a lower bound of the k-th best unmasked A.  The true top-k (by exact score E, |A-E| <= delta)
//      all have A >= T - 2 delta, so the row is CERTIFIED when T - 2 delta >= threshold, i.e. everything
//      that can matter was captured; otherwise the row is handed to the rescue kernel;
//   D  re-score the candidates with A >= T - 2 delta exactly in fp32 from the original tables;
//   E  exact top-k of those by (E desc, id asc), sorted, written out.
#include "common.cuh"
#include "kernels.h"

namespace divrec {

namespace {

constexpr int SEL_WARPS = 4;
constexpr int SEL_CAP = 1024;             // candidate elements per row held in shared memory
constexpr int SEL_SLOTS = SEL_CAP / 32;
constexpr int SORT_N = 128;               // >= MAX_K, power of two
constexpr int MAX_UCH = 8;                // user row kept in registers: d <= 128 * MAX_UCH

struct SelSmem {
  uint32_t key[SEL_WARPS][SEL_CAP];
  uint32_t id[SEL_WARPS][SEL_CAP];
  uint32_t tie_id[SEL_WARPS][SEL_CAP];
  unsigned long long srt[SEL_WARPS][SORT_N];
};

// descending bitonic sort of 128 64-bit keys held in shared memory by one warp
__device__ __forceinline__ void warp_bitonic_desc_128(unsigned long long* a, uint32_t lane) {
#pragma unroll 1
  for (int size = 2; size <= SORT_N; size <<= 1) {
#pragma unroll 1
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
      for (int r = 0; r < SORT_N / 64; ++r) {
        const int t = r * 32 + lane;                       // 0..63: one compare-exchange each
        const int i = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
        const int j = i | stride;
        const bool desc = ((i & size) == 0);
        const unsigned long long x = a[i], y = a[j];
        const bool swap = desc ? (x < y) : (x > y);
        if (swap) { a[i] = y; a[j] = x; }
      }
      __syncwarp();
    }
  }
}

__device__ __forceinline__ float4 ld4(const void* T, int dtype, int64_t row, int d, int c4) {
  if (dtype == 0) {
    return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(T) + row * d) + c4);
  } else {
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(T) + row * d) + c4);
    return make_float4(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u),
                       __uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xffff0000u));
  }
}
__device__ __forceinline__ float ld1(const void* T, int dtype, int64_t row, int d, int c) {
  if (dtype == 0) return __ldg(reinterpret_cast<const float*>(T) + row * d + c);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(T)[row * d + c]);
}

struct GroupRef { const float* s; uint32_t item0; };
__device__ __forceinline__ GroupRef group_at(const SelectParams& p, int64_t row, int n0, int g) {
  const int h = g < n0 ? 0 : 1;
  const int idx = g - (h ? n0 : 0);
  const size_t base = ((size_t)row * 2 + h) * (size_t)p.cap + idx;
  return GroupRef{p.list_scores + base * GROUP, p.list_item0[base]};
}

// number of captured elements of the row with ordered key >= kmin (and a real item id)
__device__ __forceinline__ int count_ge(const SelectParams& p, int64_t row, int n0, int ng, uint32_t kmin, uint32_t lane) {
  int c = 0;
  for (int g = lane; g < ng; g += 32) {
    const GroupRef gr = group_at(p, row, n0, g);
    const float4 a = *reinterpret_cast<const float4*>(gr.s), b = *reinterpret_cast<const float4*>(gr.s + 4);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; ++j)
      c += (f32_to_key(__float_as_uint(v[j] + 0.0f)) >= kmin && gr.item0 + j < (uint32_t)p.nI) ? 1 : 0;
  }
  return __reduce_add_sync(0xffffffffu, c);
}

// enumerate elements with key >= kmin into shared memory (at most SEL_CAP are stored); returns the total
__device__ __forceinline__ int enumerate_ge(const SelectParams& p, int64_t row, int n0, int ng, uint32_t kmin,
                                            uint32_t* key, uint32_t* id, uint32_t lane) {
  const uint32_t lt = lanemask_lt();
  int total = 0;
  for (int g0 = 0; g0 < ng; g0 += 32) {
    const int g = g0 + lane;
    uint32_t kk[8];
    uint32_t item0 = 0;
    uint32_t m = 0;
    if (g < ng) {
      const GroupRef gr = group_at(p, row, n0, g);
      item0 = gr.item0;
      const float4 a = *reinterpret_cast<const float4*>(gr.s), b = *reinterpret_cast<const float4*>(gr.s + 4);
      const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        kk[j] = f32_to_key(__float_as_uint(v[j] + 0.0f));
        if (kk[j] >= kmin && kk[j] > KEY_NEG_INF && item0 + j < (uint32_t)p.nI) m |= 1u << j;
      }
    }
    // exclusive prefix sum of popc(m) over the warp
    const int mine = __popc(m);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((int)lane >= o) incl += t;
    }
    int pos = total + incl - mine;
    total += __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (m & (1u << j)) {
        if (pos < SEL_CAP) { key[pos] = kk[j]; id[pos] = item0 + j; }
        ++pos;
      }
    }
    (void)lt;
  }
  __syncwarp();
  return total;
}

__device__ __forceinline__ void report_fail(const SelectParams& p, int64_t row, uint32_t lane) {
  if (lane == 0) {
    const int slot = atomicAdd(p.fail_count, 1);
    p.fail_rows[slot] = (int32_t)row;
  }
}

__global__ void __launch_bounds__(SEL_WARPS * 32)
select_topk_kernel(const SelectParams p) {
  extern __shared__ __align__(16) uint8_t sel_smem_raw[];
  SelSmem* sm = reinterpret_cast<SelSmem*>(sel_smem_raw);
  const int w = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const uint32_t lt = lanemask_lt();
  const int64_t row = (int64_t)blockIdx.x * SEL_WARPS + w;
  if (row >= p.nU) return;
  uint32_t* key = sm->key[w];
  uint32_t* id = sm->id[w];
  unsigned long long* srt = sm->srt[w];
  uint32_t* tie_id = sm->tie_id[w];
  const int k = p.k;

  const int n0 = p.list_count[row * 2 + 0];
  const int n1 = p.list_count[row * 2 + 1];
  if (p.force_fail || n0 > p.cap || n1 > p.cap) { report_fail(p, row, lane); return; }
  const int ng = n0 + n1;

  int64_t s0 = 0, s1 = 0;
  if (p.seen_rowptr != nullptr) { s0 = p.seen_rowptr[row]; s1 = p.seen_rowptr[row + 1]; }
  const int64_t seen_u = s1 - s0;

  // 2 delta in the (scaled) units of the tensor-core scores
  int e = 0;
  if (p.scale_exps != nullptr) e = p.scale_exps[0] + p.scale_exps[1];
  const float nu = p.norms_u[row], nvmax = __uint_as_float(*p.nvmax_bits);
  float delta = p.c_margin * nu * nvmax;
  if (p.amax_bits != nullptr) {
    // fp16 subnormals: |fp16(x) - x| <= 2^-11 |x| + 2^-25 in scaled units, the absolute part adds
    // 2^-25 sqrt(d) (|u| 2^-e1 + |v| 2^-e0) <= 2^-38 sqrt(d) (|u| amax_v + |v| amax_u) to delta
    const float au = __uint_as_float(p.amax_bits[0]), av = __uint_as_float(p.amax_bits[1]);
    delta += 3.64e-12f * sqrtf((float)p.d) * (nu * av + nvmax * au);
  }
  const float margin2 = ldexpf(2.0f * delta, e) * 1.0001f;

  const float thr = p.thr[row];
  uint32_t thr_key = (thr == -INFINITY) ? 0u : f32_to_key(__float_as_uint(thr + 0.0f));  // effective capture threshold

  // ---- A: enumerate ----
  int ne = enumerate_ge(p, row, n0, ng, thr_key, key, id, lane);
  if (ne > SEL_CAP) {
    // coarse cut: X = the need-th largest captured score, need = k + seen_u  =>  k-th best unmasked >= X
    const int64_t need64 = (int64_t)k + seen_u;
    if (need64 >= ne) { report_fail(p, row, lane); return; }
    const int need = (int)need64;
    uint32_t lo = thr_key > KEY_NEG_INF ? thr_key : KEY_NEG_INF + 1, hi = 0xff800000u;  // hi: above every finite key
    while (hi - lo > 1) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      const int c = count_ge(p, row, n0, ng, mid, lane);
      if (c >= need) lo = mid; else hi = mid;
      if (c >= need && c <= need + need / 8) break;  // close enough
    }
    const float X = __uint_as_float(key_to_f32(lo));
    const uint32_t cut_key = f32_to_key(__float_as_uint((X - margin2) + 0.0f));
    if (cut_key <= thr_key) { report_fail(p, row, lane); return; }
    thr_key = cut_key;
    ne = enumerate_ge(p, row, n0, ng, thr_key, key, id, lane);
    if (ne > SEL_CAP) { report_fail(p, row, lane); return; }
  }

  // ---- B: seen-item mask ----
  if (seen_u > 0) {
    for (int i = lane; i < ne; i += 32) {
      const int64_t gid = (int64_t)id[i] + p.item_id_offset;
      int64_t lo = s0, hi = s1;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)__ldg(p.seen_items + mid) < gid) lo = mid + 1; else hi = mid;
      }
      if (lo < s1 && (int64_t)__ldg(p.seen_items + lo) == gid) key[i] = 0;
    }
  }
  __syncwarp();

  // ---- C: lower bound T of the k-th best unmasked tensor-core score ----
  uint32_t rk[SEL_SLOTS], rid[SEL_SLOTS];
  int nv_local = 0;
  uint32_t kmin = 0xffffffffu, kmax = 0;
#pragma unroll
  for (int s = 0; s < SEL_SLOTS; ++s) {
    const int idx = s * 32 + lane;
    const uint32_t kk = idx < ne ? key[idx] : 0;
    rk[s] = kk;
    rid[s] = idx < ne ? id[idx] : 0;
    if (kk) { nv_local++; kmin = min(kmin, kk); kmax = max(kmax, kk); }
  }
  const int nv = __reduce_add_sync(0xffffffffu, nv_local);
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  uint32_t cut = 1;  // keep every unmasked candidate
  if (nv >= k) {
    uint32_t lo = kmin, hi = kmax + 1;
    int c_lo = nv;
    while (hi - lo > 1 && c_lo != k) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      int c = 0;
#pragma unroll
      for (int s = 0; s < SEL_SLOTS; ++s) c += (rk[s] >= mid) ? 1 : 0;
      c = __reduce_add_sync(0xffffffffu, c);
      if (c >= k) { lo = mid; c_lo = c; } else { hi = mid; }
    }
    const float T = __uint_as_float(key_to_f32(lo));
    cut = f32_to_key(__float_as_uint((T - margin2) + 0.0f));
    if (thr_key != 0 && cut < thr_key) { report_fail(p, row, lane); return; }  // not certified
  } else if (thr_key != 0) {
    report_fail(p, row, lane);  // fewer than k unmasked candidates captured above a threshold
    return;
  }
  // ---- D: compact the candidates to re-score, then exact fp32 scores ----
  __syncwarp();
  int nr = 0;
#pragma unroll
  for (int s = 0; s < SEL_SLOTS; ++s) {
    const bool keep = rk[s] != 0 && rk[s] >= cut;
    const uint32_t kb = __ballot_sync(0xffffffffu, keep);
    if (keep) id[nr + __popc(kb & lt)] = rid[s];
    nr += __popc(kb);
  }
  __syncwarp();
  const int d = p.d;
  const bool vec = ((d & 3) == 0) && d <= 128 * MAX_UCH;
  if (vec) {
    const int nvec = d >> 2;
    float4 ur[MAX_UCH];
#pragma unroll
    for (int c = 0; c < MAX_UCH; ++c) {
      const int c4 = c * 32 + lane;
      ur[c] = c4 < nvec ? ld4(p.U, p.dtype, row, d, c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int r = 0; r < nr; r += 2) {
      const uint32_t ia = id[r], ib = id[min(r + 1, nr - 1)];
      float sa = 0.f, sb = 0.f;
#pragma unroll
      for (int c = 0; c < MAX_UCH; ++c) {
        const int c4 = c * 32 + lane;
        if (c * 32 < nvec) {
          if (c4 < nvec) {
            const float4 va = ld4(p.V, p.dtype, ia, d, c4), vb = ld4(p.V, p.dtype, ib, d, c4);
            sa = fmaf(ur[c].x, va.x, sa); sa = fmaf(ur[c].y, va.y, sa); sa = fmaf(ur[c].z, va.z, sa); sa = fmaf(ur[c].w, va.w, sa);
            sb = fmaf(ur[c].x, vb.x, sb); sb = fmaf(ur[c].y, vb.y, sb); sb = fmaf(ur[c].z, vb.z, sb); sb = fmaf(ur[c].w, vb.w, sb);
          }
        }
      }
      sa = warp_sum(sa);
      sb = warp_sum(sb);
      if (lane == 0) {
        key[r] = f32_to_key(__float_as_uint(sa + 0.0f));
        if (r + 1 < nr) key[r + 1] = f32_to_key(__float_as_uint(sb + 0.0f));
      }
    }
  } else {
    for (int r = 0; r < nr; ++r) {
      const uint32_t ia = id[r];
      float sa = 0.f;
      for (int c = lane; c < d; c += 32) sa = fmaf(ld1(p.U, p.dtype, row, d, c), ld1(p.V, p.dtype, ia, d, c), sa);
      sa = warp_sum(sa);
      if (lane == 0) key[r] = f32_to_key(__float_as_uint(sa + 0.0f));
    }
  }
  __syncwarp();

  // ---- E: exact top-k by (E desc, id asc) ----
  nv_local = 0; kmin = 0xffffffffu; kmax = 0;
#pragma unroll
  for (int s = 0; s < SEL_SLOTS; ++s) {
    const int idx = s * 32 + lane;
    uint32_t kk = idx < nr ? key[idx] : 0;
    if (kk <= KEY_NEG_INF) kk = 0;  // -inf / NaN scores are not recommendable
    rk[s] = kk;
    rid[s] = idx < nr ? id[idx] : 0;
    if (kk) { nv_local++; kmin = min(kmin, kk); kmax = max(kmax, kk); }
  }
  const int nv2 = __reduce_add_sync(0xffffffffu, nv_local);
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  uint32_t T = 0;
  int quota = 0x7fffffff;
  if (nv2 > k) {
    uint32_t lo = kmin, hi = kmax + 1;
    int c_lo = nv2, c_hi = 0;
    while (hi - lo > 1 && c_lo != k) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      int c = 0;
#pragma unroll
      for (int s = 0; s < SEL_SLOTS; ++s) c += (rk[s] >= mid) ? 1 : 0;
      c = __reduce_add_sync(0xffffffffu, c);
      if (c >= k) { lo = mid; c_lo = c; } else { hi = mid; c_hi = c; }
    }
    T = lo;
    if (c_lo != k) quota = k - c_hi;  // hi == lo + 1: T is the exact k-th key, ties trimmed by id
  }
  // ties on the k-th score: keep the `quota` lowest item ids among them
  uint32_t id_cut = 0xffffffffu;
  if (quota != 0x7fffffff) {
    int nt = 0;
#pragma unroll
    for (int s = 0; s < SEL_SLOTS; ++s) {
      const bool eq = rk[s] == T && rk[s] != 0;
      const uint32_t eb = __ballot_sync(0xffffffffu, eq);
      if (eq) tie_id[nt + __popc(eb & lt)] = rid[s];
      nt += __popc(eb);
    }
    __syncwarp();
    for (int i = lane; i < nt; i += 32) {
      const uint32_t me = tie_id[i];
      int r = 0;
      for (int j = 0; j < nt; ++j) r += (tie_id[j] < me) ? 1 : 0;
      if (r == quota - 1) id_cut = me;
    }
    id_cut = __reduce_min_sync(0xffffffffu, id_cut);
  }
  for (int i = lane; i < SORT_N; i += 32) srt[i] = 0ull;
  __syncwarp();
  int outn = 0;
#pragma unroll
  for (int s = 0; s < SEL_SLOTS; ++s) {
    const uint32_t kk = rk[s];
    const bool keep = kk != 0 && (kk > T || (kk == T && rid[s] <= id_cut));
    const uint32_t kb = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int pos = outn + __popc(kb & lt);
      if (pos < SORT_N) srt[pos] = ((unsigned long long)kk << 32) | (unsigned long long)(~rid[s]);
    }
    outn += __popc(kb);
  }
  __syncwarp();
  warp_bitonic_desc_128(srt, lane);
  for (int r = lane; r < k; r += 32) {
    const unsigned long long c = srt[r];
    const uint32_t kk = (uint32_t)(c >> 32);
    float sc = -INFINITY;
    int64_t it = -1;
    if (kk != 0) {
      sc = __uint_as_float(key_to_f32(kk));
      it = (int64_t)(~(uint32_t)c) + p.item_id_offset;
    }
    p.out_scores[row * k + r] = sc;
    p.out_items[row * k + r] = it;
  }
}

// ---- n-way merge of per-shard sorted top-k lists -------------------------------------------------
// One warp per user.  Element (s, r) of shard s rank r has the total order key (score desc, id asc);
// its output rank is the number of elements of all shards that precede it, found by binary search
// in every shard list.  n_shards * k is small (<= 8 * 128), everything stays in registers / L1.
__global__ void __launch_bounds__(128)
topk_merge_kernel(const float* __restrict__ gs, const int64_t* __restrict__ gi, int n_shards, int64_t nU, int k,
                  float* __restrict__ out_scores, int64_t* __restrict__ out_items) {
  const int w = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int64_t row = (int64_t)blockIdx.x * 4 + w;
  if (row >= nU) return;
  const int total = n_shards * k;
  for (int r = lane; r < k; r += 32) { out_scores[row * k + r] = -INFINITY; out_items[row * k + r] = -1; }
  __syncwarp();
  for (int e = lane; e < total; e += 32) {
    const int s = e / k, r = e - s * k;
    const float sc = gs[((int64_t)s * nU + row) * k + r];
    const int64_t it = gi[((int64_t)s * nU + row) * k + r];
    if (it < 0) continue;  // padding
    int rank = 0;
    for (int s2 = 0; s2 < n_shards; ++s2) {
      if (s2 == s) { rank += r; continue; }
      const float* ps = gs + ((int64_t)s2 * nU + row) * k;
      const int64_t* pi = gi + ((int64_t)s2 * nU + row) * k;
      // count elements of shard s2 that precede (sc, it): score > sc, or == and id < it; padding never precedes
      int lo = 0, hi = k;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const float ms = ps[mid];
        const int64_t mi = pi[mid];
        const bool before = (mi >= 0) && (ms > sc || (ms == sc && mi < it));
        if (before) lo = mid + 1; else hi = mid;
      }
      rank += lo;
    }
    if (rank < k) { out_scores[row * k + rank] = sc; out_items[row * k + rank] = it; }
  }
}

}  // namespace

cudaError_t launch_select_topk(const SelectParams& p, cudaStream_t stream) {
  const size_t smem = sizeof(SelSmem);
  cudaError_t e = cudaFuncSetAttribute(select_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int64_t grid = (p.nU + SEL_WARPS - 1) / SEL_WARPS;
  select_topk_kernel<<<(unsigned)grid, SEL_WARPS * 32, smem, stream>>>(p);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_topk_merge(const float* gs, const int64_t* gi, int n_shards, int64_t nU, int k,
                              float* out_scores, int64_t* out_items, cudaStream_t stream) {
  const int64_t grid = (nU + 3) / 4;
  topk_merge_kernel<<<(unsigned)grid, 128, 0, stream>>>(gs, gi, n_shards, nU, k, out_scores, out_items);
  count_launch();
  return cudaGetLastError();
}

}  // namespace divrec
