// select_topk.cu -- exact per-user top-k over the candidate lists left by score_filter.cu, and
// the n-way merge of per-shard top-k lists (item-sharded catalogue).
//
// Replaces `negative_item_id[torch.argsort(scores, descending=True)][:k]`
// (divrec/train/utils.py:72-76).  The reference's argsort is not stable, so its order on exact
// score ties is arbitrary; the order produced here is the canonical (score desc, item id asc),
// which equals the reference whenever it is tie-free.
#include "common.cuh"
#include "kernels.h"

namespace divrec {

namespace {

constexpr int SEL_WARPS = 4;
constexpr int SEL_CAP = 2 * LIST_CAP;   // both lists of a row
constexpr int SEL_SLOTS = SEL_CAP / 32;
constexpr int SORT_N = 128;             // >= MAX_K, power of two

struct SelSmem {
  uint32_t key[SEL_WARPS][SEL_CAP];
  uint32_t id[SEL_WARPS][SEL_CAP];
  unsigned long long srt[SEL_WARPS][SORT_N];
  uint32_t tie_id[SEL_WARPS][SEL_CAP];
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
  const int n = n0 + n1;
  const uint2* L0 = p.lists + (size_t)row * 2 * LIST_CAP;
  const uint2* L1 = L0 + LIST_CAP;
  for (int i = lane; i < n; i += 32) {
    const uint2 e = i < n0 ? L0[i] : L1[i - n0];
    uint32_t kk = f32_to_key(e.x);
    if (kk <= KEY_NEG_INF) kk = 0;
    key[i] = kk;
    id[i] = e.y;
  }
  __syncwarp();
  // seen-item mask: every seen id of the row is looked up in both id-sorted segments
  if (p.seen_rowptr != nullptr) {
    const int64_t s0 = p.seen_rowptr[row], s1 = p.seen_rowptr[row + 1];
    for (int64_t si = s0 + lane; si < s1; si += 32) {
      const int64_t sid64 = (int64_t)p.seen_items[si] - p.item_id_offset;
      if (sid64 < 0 || sid64 >= p.nI) continue;
      const uint32_t sid = (uint32_t)sid64;
#pragma unroll
      for (int seg = 0; seg < 2; ++seg) {
        int lo = seg ? n0 : 0, hi = seg ? n : n0;
        const int end = hi;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (id[mid] < sid) lo = mid + 1; else hi = mid;
        }
        if (lo < end && id[lo] == sid) key[lo] = 0;
      }
    }
  }
  __syncwarp();
  // registers: slot-major copy of the keys
  uint32_t rk[SEL_SLOTS];
  int nv_local = 0;
  uint32_t kmin = 0xffffffffu, kmax = 0;
#pragma unroll
  for (int s = 0; s < SEL_SLOTS; ++s) {
    const int idx = s * 32 + lane;
    const uint32_t kk = idx < n ? key[idx] : 0;
    rk[s] = kk;
    if (kk) { nv_local++; kmin = min(kmin, kk); kmax = max(kmax, kk); }
  }
  const int nv = __reduce_add_sync(0xffffffffu, nv_local);
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  kmax = __reduce_max_sync(0xffffffffu, kmax);

  uint32_t T = 0;
  int quota = 0x7fffffff;
  if (nv > k) {
    uint32_t lo = kmin, hi = kmax + 1;
    int c_lo = nv, c_hi = 0;
    while (hi - lo > 1 && c_lo != k) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      int c = 0;
#pragma unroll
      for (int s = 0; s < SEL_SLOTS; ++s) c += (rk[s] >= mid) ? 1 : 0;
      c = __reduce_add_sync(0xffffffffu, c);
      if (c >= k) { lo = mid; c_lo = c; } else { hi = mid; c_hi = c; }
    }
    T = lo;
    if (c_lo != k) quota = k - c_hi;
  }
  // ties on the k-th score: keep the `quota` lowest item ids among them
  uint32_t id_cut = 0xffffffffu;  // keep ties with id <= id_cut
  if (quota != 0x7fffffff) {
    int nt = 0;
#pragma unroll
    for (int s = 0; s < SEL_SLOTS; ++s) {
      const int idx = s * 32 + lane;
      const bool eq = rk[s] == T && rk[s] != 0;
      const uint32_t eb = __ballot_sync(0xffffffffu, eq);
      if (eq) tie_id[nt + __popc(eb & lt)] = id[idx];
      nt += __popc(eb);
    }
    __syncwarp();
    // rank of each tie id = number of smaller tie ids; the one with rank quota-1 is the cut
    for (int i = lane; i < nt; i += 32) {
      const uint32_t me = tie_id[i];
      int r = 0;
      for (int j = 0; j < nt; ++j) r += (tie_id[j] < me) ? 1 : 0;
      if (r == quota - 1) id_cut = me;
    }
    id_cut = __reduce_min_sync(0xffffffffu, id_cut);
  }
  // gather winners into the sort buffer: composite = key << 32 | ~id  (descending sort)
  for (int i = lane; i < SORT_N; i += 32) srt[i] = 0ull;
  __syncwarp();
  int outn = 0;
#pragma unroll
  for (int s = 0; s < SEL_SLOTS; ++s) {
    const int idx = s * 32 + lane;
    const uint32_t kk = rk[s];
    bool keep = false;
    uint32_t my_id = 0;
    if (kk != 0) {
      my_id = id[idx];
      keep = kk > T || (kk == T && my_id <= id_cut);
    }
    const uint32_t kb = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int pos = outn + __popc(kb & lt);
      if (pos < SORT_N) srt[pos] = ((unsigned long long)kk << 32) | (unsigned long long)(~my_id);
    }
    outn += __popc(kb);
  }
  __syncwarp();
  warp_bitonic_desc_128(srt, lane);
  float scale = 1.0f;
  int e0 = 0, e1 = 0;
  if (p.scale_exps != nullptr) { e0 = p.scale_exps[0]; e1 = p.scale_exps[1]; }
  (void)scale;
  for (int r = lane; r < k; r += 32) {
    const unsigned long long c = srt[r];
    const uint32_t kk = (uint32_t)(c >> 32);
    float sc = -INFINITY;
    int64_t it = -1;
    if (kk != 0) {
      sc = ldexpf(ldexpf(__uint_as_float(key_to_f32(kk)), -e0), -e1);
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
