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

// Per-warp shared-memory budget by k class.  KB = 0: k <= 128 (the common case, 6 CTAs per SM); KB = 1: k <= 512
// (BASELINE.json config 5 takes top-500 candidate lists; 2 CTAs per SM).
//   SEL_CAP   candidate elements per row held in shared memory
//   SEEN_CAP  seen items of the row staged in shared memory (longer lists stay in global memory)
//   SORT_N    sort chunk, >= the largest k of the class, power of two
// One contiguous region of 32-bit words per warp: [key: SEL_CAP][id: SEL_CAP][seen: SEEN_CAP (phase B; reused as the
// sort chunk of phase E)][top: SORT_N 64-bit, the running top list; phase B borrows it for the seen-item Bloom filter].
// The coarse cut of rows that captured more than SEL_CAP elements uses the whole region as a keys-only buffer.
template <int KB>
struct SelCfg {
#ifdef DIVREC_EXP_SELCAP512
  static constexpr int SEL_CAP = KB ? 2048 : 512;
#else
  static constexpr int SEL_CAP = KB ? 2048 : 768;
#endif
  static constexpr int SEEN_CAP = KB ? 1024 : 512;
  static constexpr int SORT_N = KB ? 512 : 128;
  static constexpr int SEL_WORDS = 2 * SEL_CAP + SEEN_CAP + 2 * SORT_N;
  static constexpr int BLOOM_WORDS = 64;                          // 2048-bit Bloom filter over the row's seen ids
  static constexpr int HIT_CAP = 2 * SORT_N - BLOOM_WORDS;        // candidates the filter flags, verified exactly
  static_assert(SEEN_CAP * sizeof(int32_t) >= SORT_N * sizeof(unsigned long long), "sort chunk aliases the seen list");
  static_assert(((2 * SEL_CAP + SEEN_CAP) * 4) % 8 == 0 && (SEL_WORDS * 4) % 16 == 0, "64-bit alignment of the top list");
};
static_assert(SelCfg<1>::SORT_N >= MAX_K && SelCfg<0>::SORT_N >= MAX_K_SMALL, "sort chunk covers the k class");

// descending bitonic sort of N 64-bit keys held in shared memory by one warp
template <int N>
__device__ __forceinline__ void warp_bitonic_desc(unsigned long long* a, uint32_t lane) {
#pragma unroll 1
  for (int size = 2; size <= N; size <<= 1) {
#pragma unroll 1
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
      for (int r = 0; r < N / 64; ++r) {
        const int t = r * 32 + lane;                       // one compare-exchange each
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
// `a` is bitonic (first descending then ascending, or any rotation): make it descending
template <int N>
__device__ __forceinline__ void warp_bitonic_merge_desc(unsigned long long* a, uint32_t lane) {
#pragma unroll 1
  for (int stride = N >> 1; stride > 0; stride >>= 1) {
#pragma unroll
    for (int r = 0; r < N / 64; ++r) {
      const int t = r * 32 + lane;
      const int i = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
      const int j = i | stride;
      const unsigned long long x = a[i], y = a[j];
      if (x < y) { a[i] = y; a[j] = x; }
    }
    __syncwarp();
  }
}

template <int DT>
__device__ __forceinline__ float4 ld4(const void* T, int64_t row, int d, int c4) {
  if (DT == 0) {
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
__device__ __forceinline__ float dot4(const float4& a, const float4& b, float s) {
  s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); return fmaf(a.w, b.w, s);
}
// sum over the 8 lanes of a sub-group (lanes that share lane >> 3)
__device__ __forceinline__ float sum8(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

// One pass over the row's captured groups: iteration `it` handles group it*32+lane of each of the row's four lists
// (four independent record reads in flight per lane, trivial addressing).
//   MODE 1: elements with score >= fmin are written to key[] / id[] (at most SEL_CAP), any order;
//   MODE 2: their keys only are written to key[0 .. SEL_WORDS) (the warp's whole region);  MODE 0: count only.
//   Returns how many there are in total.  Scores are compared as floats; -inf never qualifies.
template <int MODE, int KB>
__device__ __forceinline__ int scan_groups(const SelectParams& p, int64_t row, const int4 nl, float fmin,
                                           uint32_t* key, uint32_t* id, uint32_t lane) {
  constexpr int SEL_CAP = SelCfg<KB>::SEL_CAP, SEL_WORDS = SelCfg<KB>::SEL_WORDS;
  int total = 0;
  const float lim = fmaxf(fmin, -3.0e38f);
  // four parallel read streams: the row's four lists, or -- when the two-block filter wrote two lists of
  // 2 * cap groups (slots 0 and 2) -- the two halves of each
  int cnt[LISTS_PER_ROW] = {nl.x, nl.y, nl.z, nl.w};
  int start[LISTS_PER_ROW] = {0, p.cap, 2 * p.cap, 3 * p.cap};
  if (p.merged_lists) {
    const int h0 = (nl.x + 1) >> 1, h2 = (nl.z + 1) >> 1;
    cnt[0] = h0; cnt[1] = nl.x - h0; start[1] = h0;
    cnt[2] = h2; cnt[3] = nl.z - h2; start[3] = 2 * p.cap + h2;
  }
  const int nmax = max(max(cnt[0], cnt[1]), max(cnt[2], cnt[3]));
  const size_t row_base = (size_t)row * LISTS_PER_ROW * (size_t)p.cap;
  const uint32_t nI = (uint32_t)p.nI;
  const bool wide_ids = p.id_bytes > 2;
  for (int g = lane; g < nmax + (int)lane; g += 32) {  // uniform trip count: every lane runs ceil(nmax / 32) iterations
    float4 a[LISTS_PER_ROW], b[LISTS_PER_ROW];
    uint32_t item0[LISTS_PER_ROW];
#pragma unroll
    for (int l = 0; l < LISTS_PER_ROW; ++l) {
      item0[l] = 0xffffffffu;
      if (g < cnt[l]) {
        const float4* src = reinterpret_cast<const float4*>(p.lists + (row_base + (size_t)start[l] + g) * REC_FLOATS);
        a[l] = src[0];
        b[l] = src[1];
        // the group index rides in the low mantissa bytes of the first scores
        uint32_t gi = (__float_as_uint(a[l].x) & 0xffu) | ((__float_as_uint(a[l].y) & 0xffu) << 8);
        if (wide_ids) gi |= ((__float_as_uint(a[l].z) & 0xffu) << 16) | ((__float_as_uint(a[l].w) & 0xffu) << 24);
        item0[l] = gi * GROUP;
      }
    }
    uint32_t m[LISTS_PER_ROW];
    int mine = 0;
#pragma unroll
    for (int l = 0; l < LISTS_PER_ROW; ++l) {
      m[l] = 0;
      if (item0[l] < nI) {  // also false for the 0xffffffff "no group" marker and for groups past the catalogue
        m[l] = (a[l].x >= lim ? 1u : 0u) | (a[l].y >= lim ? 2u : 0u) | (a[l].z >= lim ? 4u : 0u) | (a[l].w >= lim ? 8u : 0u) |
               (b[l].x >= lim ? 16u : 0u) | (b[l].y >= lim ? 32u : 0u) | (b[l].z >= lim ? 64u : 0u) | (b[l].w >= lim ? 128u : 0u);
        // only the last tile's groups can run past the catalogue (zero-filled operand rows score 0)
        if (nI - item0[l] < 8u) m[l] &= (1u << (nI - item0[l])) - 1u;
      }
      mine += __popc(m[l]);
    }
    if (MODE == 0) {
      total += mine;
    } else {
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += t;
      }
      int pos = total + incl - mine;
      total += __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
      for (int l = 0; l < LISTS_PER_ROW; ++l) {
        uint32_t mm = m[l];
        while (mm) {
          const int e = __ffs(mm) - 1;
          mm &= mm - 1;
          const float v = e < 4 ? (e == 0 ? a[l].x : e == 1 ? a[l].y : e == 2 ? a[l].z : a[l].w)
                                : (e == 4 ? b[l].x : e == 5 ? b[l].y : e == 6 ? b[l].z : b[l].w);
          if (MODE == 1) {
            if (pos < SEL_CAP) { key[pos] = f32_to_key(__float_as_uint(v + 0.0f)); id[pos] = item0[l] + e; }
          } else {
            if (pos < SEL_WORDS) key[pos] = f32_to_key(__float_as_uint(v + 0.0f));
          }
          ++pos;
        }
      }
    }
  }
  if (MODE == 0) total = __reduce_add_sync(0xffffffffu, total);
  __syncwarp();
  return total;
}

// reason: 0 forced / list overflow, 1 too many candidates, 2 threshold hint not certified, 3 too few captured
__device__ __forceinline__ void report_fail(const SelectParams& p, int64_t row, uint32_t lane, int reason = 0) {
  if (lane == 0) {
    const int slot = atomicAdd(p.fail_count, 1);
    p.fail_rows[slot] = (int32_t)row;
    if (p.fail_reasons != nullptr) atomicAdd(p.fail_reasons + reason, 1);
  }
}

// exact membership of global id g32 in the row's sorted seen list staged in shared memory (padded to seen_pow2 with
// INT_MAX): largest position whose value is <= g32, one conditional add per level (no branches)
__device__ __forceinline__ bool seen_lookup(const int32_t* seen_s, int seen_pow2, int32_t g32) {
  int pos = 0;
  for (int half = seen_pow2 >> 1; half > 0; half >>= 1) pos += (seen_s[pos + half] <= g32) ? half : 0;
  return seen_s[pos] == g32;
}

// NJ = float4 chunks of a table row per lane of an 8-lane sub-group: vector path for d <= 32 * NJ, d % 4 == 0.
// NJ == 0: scalar path for any d (no fused metrics).  NJ == 8 (rows up to d = 256) re-scores one candidate per sub-group
// and iteration instead of two: the second set of row registers is what used to spill.
template <int NJ, int DT, bool FULL, int KB>
#ifdef DIVREC_EXP_SELCAP512
__global__ void __launch_bounds__(SEL_WARPS * 32, KB ? 2 : (NJ == 8 ? 4 : 7))
#else
__global__ void __launch_bounds__(SEL_WARPS * 32, KB ? 2 : (NJ == 8 ? 4 : 6))
#endif
select_topk_kernel(const SelectParams p) {
  using Cfg = SelCfg<KB>;
  constexpr int SEL_CAP = Cfg::SEL_CAP, SEEN_CAP = Cfg::SEEN_CAP, SORT_N = Cfg::SORT_N, SEL_WORDS = Cfg::SEL_WORDS;
  constexpr bool TWO = NJ != 8;  // candidates per sub-group and re-scoring iteration: two (ra, rb) or one
  extern __shared__ __align__(16) uint8_t sel_smem_raw[];
  const int w = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const uint32_t lt = lanemask_lt();
  const int64_t row = (int64_t)blockIdx.x * SEL_WARPS + w;
  if (row >= p.nU) return;
  uint32_t* key = reinterpret_cast<uint32_t*>(sel_smem_raw) + (size_t)w * SEL_WORDS;
  uint32_t* id = key + SEL_CAP;
  int32_t* seen_s = reinterpret_cast<int32_t*>(key + 2 * SEL_CAP);
  unsigned long long* top = reinterpret_cast<unsigned long long*>(key + 2 * SEL_CAP + SEEN_CAP);
  unsigned long long* chunk = reinterpret_cast<unsigned long long*>(seen_s);
  const int k = p.k;

  {
    // start the row's record lists (4 KB of each stream) and its user row toward L2 while the counts, thresholds and seen list
    // are fetched: the first record reads then hit L2 (-1.5 % of the kernel)
    const int n_str = p.merged_lists ? 2 : 4, per = 32 / n_str;
    const size_t list0 = ((size_t)row * LISTS_PER_ROW + (size_t)((int)lane / per) * (p.merged_lists ? 2 : 1)) * (size_t)p.cap;
    const char* a = reinterpret_cast<const char*>(p.lists + list0 * REC_FLOATS) + (size_t)((int)lane % per) * 128;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(a + (size_t)per * 128));
    const size_t row_bytes = (size_t)p.d * (DT == 0 ? 4 : 2);
    if ((size_t)lane * 128 < row_bytes)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.U) + (size_t)row * row_bytes + (size_t)lane * 128));
  }
  const int4 nl = *reinterpret_cast<const int4*>(p.list_count + row * LISTS_PER_ROW);
  const int cap_l = p.merged_lists ? 2 * p.cap : p.cap;  // capacity of a list (an overflowed list reports capacity + 1)
  if (p.force_fail || nl.x > cap_l || nl.y > cap_l || nl.z > cap_l || nl.w > cap_l) { report_fail(p, row, lane); return; }

  int64_t s0 = 0, s1 = 0;
  if (p.seen_rowptr != nullptr) { s0 = p.seen_rowptr[row]; s1 = p.seen_rowptr[row + 1]; }
  const int64_t seen_u = s1 - s0;
  const bool seen_in_smem = seen_u <= SEEN_CAP;
  int seen_pow2 = 1;  // smallest power of two >= seen_u: the shared copy is padded with INT_MAX up to it
  while (seen_pow2 < (int)seen_u && seen_pow2 < SEEN_CAP) seen_pow2 <<= 1;
  if (seen_in_smem)
    for (int i = lane; i < seen_pow2; i += 32) seen_s[i] = i < (int)seen_u ? __ldg(p.seen_items + s0 + i) : 0x7fffffff;

  // 2 delta in the (scaled) units of the tensor-core scores
  int e2 = 0;
  if (p.scale_exps != nullptr) e2 = p.scale_exps[1];
  if (p.row_exps != nullptr) e2 += p.row_exps[row];
  const float nu = p.norms_u[row], nvmax = __uint_as_float(*p.nvmax_bits);
  float delta = p.c_margin * nu * nvmax;
  if (p.amax_bits != nullptr) {
    // fp16 subnormals: |fp16(x) - x| <= 2^-11 |x| + 2^-25 in scaled units, the absolute part adds
    // 2^-25 sqrt(d) (|u| 2^-e_v + |v| 2^-e_u) <= 2^-38 sqrt(d) (|u| amax_v + |v| amax_u), amax_u = the ROW's largest
    // element (its own scale), itself <= |u|
    const float av = __uint_as_float(p.amax_bits[1]);
    delta += 3.64e-12f * sqrtf((float)p.d) * (nu * av + nvmax * nu);
  }
  const float margin2 = ldexpf(2.0f * delta + REC_ID_SLACK * nu * nvmax, e2) * 1.0001f;  // + the index bytes carried by the record scores

  const float4 thr4 = *reinterpret_cast<const float4*>(p.thr + row * LISTS_PER_ROW);
  const float thr = fmaxf(fmaxf(thr4.x, thr4.y), fmaxf(thr4.z, thr4.w));  // everything >= the largest list threshold was captured by all lists
  uint32_t thr_key = (thr == -INFINITY) ? 0u : f32_to_key(__float_as_uint(thr + 0.0f));  // effective capture threshold

  // ---- A: enumerate ----
  int ne = scan_groups<1, KB>(p, row, nl, thr, key, id, lane);
  if (ne > SEL_CAP) {
    // coarse cut: X = a lower bound of the need-th largest captured score, need = k + seen_u  =>  the k-th best
    // unmasked score is >= X, nothing below X - 2 delta can reach the top-k: the capture threshold is raised to it
    const int64_t need64 = (int64_t)k + seen_u;
    if (need64 >= ne) { report_fail(p, row, lane, 1); return; }
    const int need = (int)need64;
    uint32_t cut_key;
    if (ne <= SEL_WORDS) {
      // the captured keys fit the warp's shared region: ONE more pass over the lists, then bisect in shared memory
      // (large catalogues leave up to 4 x 512 groups per row; bisecting with passes over global memory cost 7 ms
      // per 65 536 rows on a 1.25M-item shard)
      scan_groups<2, KB>(p, row, nl, thr, key, id, lane);
      uint32_t kmin = 0xffffffffu, kmax = 0;
      for (int i = lane; i < ne; i += 32) { const uint32_t kk = key[i]; kmin = min(kmin, kk); kmax = max(kmax, kk); }
      kmin = __reduce_min_sync(0xffffffffu, kmin);
      kmax = __reduce_max_sync(0xffffffffu, kmax);
      uint32_t lo = kmin, hi = kmax + 1;
      int c_lo = ne;
      const int c_ok = need + max(2, need >> 3);
      bool exact = false;
      for (;;) {
        while (hi - lo > 1 && (exact || c_lo > c_ok)) {
          const uint32_t mid = lo + ((hi - lo) >> 1);
          int c = 0;
          for (int i = lane; i < ne; i += 32) c += (key[i] >= mid) ? 1 : 0;
          c = __reduce_add_sync(0xffffffffu, c);
          if (c >= need) { lo = mid; c_lo = c; } else { hi = mid; }
        }
        const float X = __uint_as_float(key_to_f32(lo));
        cut_key = f32_to_key(__float_as_uint((X - margin2) + 0.0f));
        if (cut_key > thr_key || exact) break;
        exact = true;  // the loose bound does not raise the threshold: go on to the exact need-th largest
      }
      __syncwarp();
      if (seen_in_smem)  // the keys-only buffer ran over the staged seen list: stage it again
        for (int i = lane; i < seen_pow2; i += 32) seen_s[i] = i < (int)seen_u ? __ldg(p.seen_items + s0 + i) : 0x7fffffff;
    } else {
      uint32_t lo = thr_key > KEY_NEG_INF ? thr_key : KEY_NEG_INF + 1, hi = 0xff800000u;  // hi: above every finite key
      while (hi - lo > 1) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        const int c = scan_groups<0, KB>(p, row, nl, __uint_as_float(key_to_f32(mid)), key, id, lane);
        if (c >= need) lo = mid; else hi = mid;
      }
      const float X = __uint_as_float(key_to_f32(lo));
      cut_key = f32_to_key(__float_as_uint((X - margin2) + 0.0f));
    }
    // lists tightened by the filter were cut at a (need + 8)-th largest, which the exact X dominates: a cut that
    // does not raise the threshold means the row really holds too many near-ties
    if (cut_key <= thr_key) { report_fail(p, row, lane, 1); return; }
    thr_key = cut_key;
    ne = scan_groups<1, KB>(p, row, nl, __uint_as_float(key_to_f32(thr_key)), key, id, lane);
    if (ne > SEL_CAP) { report_fail(p, row, lane, 1); return; }
  }

#if defined(DIVREC_DIAG_SEL_STOP) && DIVREC_DIAG_SEL_STOP == 1
  if (ne >= 0) { if (lane == 0) p.out_scores[row * k] = (float)ne; return; }
#endif
  // ---- B: seen-item mask ----
  // Rows whose seen list is staged in shared memory: a 2048-bit Bloom filter over the seen ids (one bit per id, built in
  // the not-yet-used top-list area) clears most candidates with one shared load; the few it flags (true members plus
  // ~seen_u / 2048 of the rest) are compacted and verified by the exact search.  Before: every candidate ran the full
  // fixed-depth search, 14 % of the kernel's instructions.
  if (seen_u > 0) {
    const int32_t off32 = (int32_t)p.item_id_offset;
    const bool ids32 = p.item_id_offset >= 0 && p.item_id_offset + (int64_t)p.nI < 0x7fffffffLL;  // global ids fit int32 (seen ids do)
    bool done = false;
    if (seen_in_smem && ids32) {
      uint32_t* bloom = reinterpret_cast<uint32_t*>(top);
      uint32_t* hits = bloom + Cfg::BLOOM_WORDS;
      bloom[lane] = 0u;
      bloom[lane + 32] = 0u;
      __syncwarp();
      for (int i = lane; i < (int)seen_u; i += 32) {
        const uint32_t g = (uint32_t)seen_s[i];
        atomicOr(&bloom[(g >> 5) & (Cfg::BLOOM_WORDS - 1)], 1u << (g & 31u));
      }
      __syncwarp();
      int nh = 0;
      for (int base = 0; base < ne; base += 32) {
        const int i = base + lane;
        bool hit = false;
        if (i < ne) {
          const uint32_t g = id[i] + (uint32_t)off32;
          hit = (bloom[(g >> 5) & (Cfg::BLOOM_WORDS - 1)] >> (g & 31u)) & 1u;
        }
        const uint32_t hb = __ballot_sync(0xffffffffu, hit);
        if (hit) { const int slot = nh + __popc(hb & lt); if (slot < Cfg::HIT_CAP) hits[slot] = (uint32_t)i; }
        nh += __popc(hb);
      }
      __syncwarp();
      if (nh <= Cfg::HIT_CAP) {
        for (int j = lane; j < nh; j += 32) {
          const uint32_t i = hits[j];
          if (seen_lookup(seen_s, seen_pow2, (int32_t)(id[i] + (uint32_t)off32))) key[i] = 0;
        }
        done = true;
      }
    }
    if (!done) {
      for (int i = lane; i < ne; i += 32) {
        const int64_t gid = (int64_t)id[i] + p.item_id_offset;
        if (seen_in_smem) {
          if (gid < 0x7fffffffLL && seen_lookup(seen_s, seen_pow2, (int32_t)gid)) key[i] = 0;  // seen ids are int32
        } else {
          int64_t lo = 0, hi = seen_u;
          while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if ((int64_t)__ldg(p.seen_items + s0 + mid) < gid) lo = mid + 1; else hi = mid;
          }
          if (lo < seen_u && (int64_t)__ldg(p.seen_items + s0 + lo) == gid) key[i] = 0;
        }
      }
    }
  }
  __syncwarp();

#if defined(DIVREC_DIAG_SEL_STOP) && DIVREC_DIAG_SEL_STOP == 2
  if (ne >= 0) { if (lane == 0) p.out_scores[row * k] = (float)ne; return; }
#endif
  // ---- C: a lower bound T of the k-th best unmasked tensor-core score (count(>= T) in [k, k + k/8]) ----
  int nv_local = 0;
  uint32_t kmin = 0xffffffffu, kmax = 0;
  for (int i = lane; i < ne; i += 32) {
    const uint32_t kk = key[i];
    if (kk) { nv_local++; kmin = min(kmin, kk); kmax = max(kmax, kk); }
  }
  const int nv = __reduce_add_sync(0xffffffffu, nv_local);
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  uint32_t cut = 1;  // keep every unmasked candidate
  if (nv >= k) {
    uint32_t lo = kmin, hi = kmax + 1;
    int c_lo = nv;
    const int c_ok = k + max(2, k >> 4);  // (k + 2 measured no faster: the re-scoring cost is not in the few extra candidates)
    bool exact = false;
    // a loose T (count up to k + k/16) is enough unless it drops T - 2 delta below the capture threshold: then the
    // bisection continues to the exact k-th key before the row is declared uncertifiable
    for (;;) {
      while (hi - lo > 1 && (exact || c_lo > c_ok)) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        int c = 0;
        for (int i = lane; i < ne; i += 32) c += (key[i] >= mid) ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= k) { lo = mid; c_lo = c; } else { hi = mid; }
      }
      const float T = __uint_as_float(key_to_f32(lo));
      cut = f32_to_key(__float_as_uint((T - margin2) + 0.0f));
      if (thr_key == 0 || cut >= thr_key) break;          // certified
      if (exact) {  // T is the exact k-th key (largest key with at least k candidates at or above it): not certifiable
        if (p.debug_dump && lane == 0) {
          float* o = p.out_scores + row * k;
          o[0] = T; o[1] = margin2; o[2] = thr4.x; o[3] = thr4.y; o[4] = thr4.z; o[5] = thr4.w; o[6] = (float)ne; o[7] = (float)nv;
          o[8] = (float)nl.x; o[9] = (float)nl.y; o[10] = (float)nl.z; o[11] = (float)nl.w; o[12] = (float)c_lo;
          p.out_items[row * k] = -1002;
        }
        report_fail(p, row, lane, 2);
        return;
      }
      exact = true;                                       // bisect down to the exact k-th key and try again
    }
  } else if (thr_key != 0) {
    report_fail(p, row, lane, 3);  // fewer than k unmasked candidates captured above a threshold
    return;
  }

#if defined(DIVREC_DIAG_SEL_STOP) && DIVREC_DIAG_SEL_STOP == 3
  if (ne >= 0) { if (lane == 0) p.out_scores[row * k] = (float)ne; return; }
#endif
  // ---- D: compact the candidates to re-score (in place), then exact fp32 scores ----
  int nr = 0;
  for (int base = 0; base < ne; base += 32) {
    const int i = base + lane;
    const uint32_t kk = i < ne ? key[i] : 0;
    const uint32_t ii = i < ne ? id[i] : 0;
    const bool keep = kk != 0 && kk >= cut;
    const uint32_t kb = __ballot_sync(0xffffffffu, keep);
    __syncwarp();
    if (keep) id[nr + __popc(kb & lt)] = ii;
    nr += __popc(kb);
  }
  __syncwarp();
  const int d = p.d;
  const bool want_ild = NJ > 0 && p.out_ild != nullptr;
  float4 acc[NJ > 0 ? NJ : 1];  // sum of the normalised rows of every re-scored candidate (fused ILD)
  if (NJ > 0) {
    const int nvec = d >> 2;
    const int sub = lane & 7, grp = lane >> 3;
    float4 ur[NJ > 0 ? NJ : 1];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c4 = sub + 8 * j;
      ur[j] = (FULL || c4 < nvec) ? ld4<DT>(p.U, row, d, c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (TWO) {
      for (int r0 = 0; r0 < nr; r0 += 8) {
        const int ra = r0 + grp, rb = r0 + 4 + grp;
        const uint32_t ia = id[min(ra, nr - 1)], ib = id[min(rb, nr - 1)];
        float4 va[NJ > 0 ? NJ : 1], vb[NJ > 0 ? NJ : 1];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int c4 = sub + 8 * j;
          if (FULL || 8 * j < nvec) {
            va[j] = (FULL || c4 < nvec) ? ld4<DT>(p.V, ia, d, c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            vb[j] = (FULL || c4 < nvec) ? ld4<DT>(p.V, ib, d, c4) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        float sa = 0.f, sb = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          if (FULL || 8 * j < nvec) { sa = dot4(ur[j], va[j], sa); sb = dot4(ur[j], vb[j], sb); }
        }
        sa = sum8(sa); sb = sum8(sb);
        if (want_ild) {
          const float ia_n = ra < nr ? __ldg(p.inv_norm_v + ia) : 0.f, ib_n = rb < nr ? __ldg(p.inv_norm_v + ib) : 0.f;
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            if (FULL || 8 * j < nvec) {
              acc[j].x = fmaf(va[j].x, ia_n, acc[j].x); acc[j].y = fmaf(va[j].y, ia_n, acc[j].y);
              acc[j].z = fmaf(va[j].z, ia_n, acc[j].z); acc[j].w = fmaf(va[j].w, ia_n, acc[j].w);
              acc[j].x = fmaf(vb[j].x, ib_n, acc[j].x); acc[j].y = fmaf(vb[j].y, ib_n, acc[j].y);
              acc[j].z = fmaf(vb[j].z, ib_n, acc[j].z); acc[j].w = fmaf(vb[j].w, ib_n, acc[j].w);
            }
          }
        }
        if (sub == 0) {
          if (ra < nr) key[ra] = f32_to_key(__float_as_uint(sa + 0.0f));
          if (rb < nr) key[rb] = f32_to_key(__float_as_uint(sb + 0.0f));
        }
      }
    } else {
      // same lane layout and per-lane fma chain as above (bit-identical scores), one candidate per sub-group
      for (int r0 = 0; r0 < nr; r0 += 4) {
        const int ra = r0 + grp;
        const uint32_t ia = id[min(ra, nr - 1)];
        float4 va[NJ > 0 ? NJ : 1];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int c4 = sub + 8 * j;
          if (FULL || 8 * j < nvec) va[j] = (FULL || c4 < nvec) ? ld4<DT>(p.V, ia, d, c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float sa = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          if (FULL || 8 * j < nvec) sa = dot4(ur[j], va[j], sa);
        }
        sa = sum8(sa);
        if (want_ild) {
          const float ia_n = ra < nr ? __ldg(p.inv_norm_v + ia) : 0.f;
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            if (FULL || 8 * j < nvec) {
              acc[j].x = fmaf(va[j].x, ia_n, acc[j].x); acc[j].y = fmaf(va[j].y, ia_n, acc[j].y);
              acc[j].z = fmaf(va[j].z, ia_n, acc[j].z); acc[j].w = fmaf(va[j].w, ia_n, acc[j].w);
            }
          }
        }
        if (sub == 0 && ra < nr) key[ra] = f32_to_key(__float_as_uint(sa + 0.0f));
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

#if defined(DIVREC_DIAG_SEL_STOP) && DIVREC_DIAG_SEL_STOP == 4
  if (ne >= 0) { if (lane == 0) p.out_scores[row * k] = (float)ne; return; }
#endif
  // ---- E: exact top-k by (E desc, id asc): composite keys are unique, sort chunks of SORT_N and merge ----
  for (int c0 = 0; c0 < nr || c0 == 0; c0 += SORT_N) {
    unsigned long long* dst = c0 == 0 ? top : chunk;
    // (a register-resident shuffle network was measured slower than this shared-memory one: 1.69 vs 1.60 ms, and 1.88 ms
    // fully unrolled -- the kernel is instruction-cache sensitive)
    for (int i = lane; i < SORT_N; i += 32) {
      const int r = c0 + i;
      unsigned long long c = 0ull;
      if (r < nr) {
        uint32_t kk = key[r];
        if (kk > KEY_NEG_INF && kk < 0xff800000u) c = ((unsigned long long)kk << 32) | (unsigned long long)(~id[r]);
      }
      dst[i] = c;
    }
    __syncwarp();
    warp_bitonic_desc<SORT_N>(dst, lane);
    if (c0 > 0) {
      for (int i = lane; i < SORT_N; i += 32) {
        const unsigned long long a = top[i], b = chunk[SORT_N - 1 - i];
        top[i] = a > b ? a : b;
      }
      __syncwarp();
      warp_bitonic_merge_desc<SORT_N>(top, lane);
    }
  }
  int kv = 0;  // valid results of the row
  for (int r = lane; r < k; r += 32) {
    const unsigned long long c = top[r];
    const uint32_t kk = (uint32_t)(c >> 32);
    float sc = -INFINITY;
    int64_t it = -1;
    if (kk != 0) {
      sc = __uint_as_float(key_to_f32(kk));
      it = (int64_t)(~(uint32_t)c) + p.item_id_offset;
      kv++;
      if (p.coverage != nullptr) atomicAdd(p.coverage + it, 1);
    }
    p.out_scores[row * k + r] = sc;
    p.out_items[row * k + r] = it;
  }
  if (NJ > 0 && want_ild) {
    // fused cosine ILD: sum_{a<b} cos = (|sum e|^2 - kv) / 2 over the winners.  acc holds the sum over ALL
    // re-scored candidates; subtract the few that did not make the top-k.
    kv = __reduce_add_sync(0xffffffffu, kv);
    const unsigned long long kth = top[k - 1];  // smallest winning composite (0 when fewer than k)
    const int nvec = d >> 2;
    const int sub = lane & 7, grp = lane >> 3;
    // list the losers (lane per candidate), reusing the front of key[] (the scores are no longer needed)
    int nlose = 0;
    for (int base = 0; base < nr; base += 32) {
      const int r = base + lane;
      bool loser = false;
      uint32_t ia = 0;
      if (r < nr) {
        ia = id[r];
        const uint32_t kk = key[r];
        unsigned long long c = 0ull;
        if (kk > KEY_NEG_INF && kk < 0xff800000u) c = ((unsigned long long)kk << 32) | (unsigned long long)(~ia);
        loser = c < kth || c == 0ull;
      }
      const uint32_t lb = __ballot_sync(0xffffffffu, loser);
      __syncwarp();
      if (loser) key[nlose + __popc(lb & lt)] = ia;  // writes land at or before the slots already read
      nlose += __popc(lb);
    }
    __syncwarp();
    for (int r0 = 0; r0 < nlose; r0 += 4) {
      const int r = r0 + grp;
      const bool on = r < nlose;
      const uint32_t ia = key[on ? r : 0];
      const float inv = on ? __ldg(p.inv_norm_v + ia) : 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int c4 = sub + 8 * j;
        if (on && (FULL || c4 < nvec)) {
          const float4 va = ld4<DT>(p.V, ia, d, c4);
          acc[j].x = fmaf(-va.x, inv, acc[j].x); acc[j].y = fmaf(-va.y, inv, acc[j].y);
          acc[j].z = fmaf(-va.z, inv, acc[j].z); acc[j].w = fmaf(-va.w, inv, acc[j].w);
        }
      }
    }
    // combine the four sub-groups' partial sums (same `sub` => same columns), then |sum e|^2
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      float4 t = acc[j];
      t.x += __shfl_xor_sync(0xffffffffu, t.x, 8); t.y += __shfl_xor_sync(0xffffffffu, t.y, 8);
      t.z += __shfl_xor_sync(0xffffffffu, t.z, 8); t.w += __shfl_xor_sync(0xffffffffu, t.w, 8);
      t.x += __shfl_xor_sync(0xffffffffu, t.x, 16); t.y += __shfl_xor_sync(0xffffffffu, t.y, 16);
      t.z += __shfl_xor_sync(0xffffffffu, t.z, 16); t.w += __shfl_xor_sync(0xffffffffu, t.w, 16);
      s = dot4(t, t, s);
    }
    s = sum8(s);  // lanes 0..7 hold the distinct column chunks
    if (lane == 0) {
      const float kvf = (float)kv;
      const float sum_cos = 0.5f * (s - kvf);
      const float pairs = 0.5f * kvf * (kvf - 1.0f);
      p.out_ild[row] = (pairs - sum_cos) / ((float)k * (float)(k - 1));
    }
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

// ---- packed exchange format (item-sharded catalogue) --------------------------------------------
// A (score, global id) pair travels as ONE 64-bit composite (ordered score key << 32 | ~id): descending
// composite order IS the canonical order (score desc, id asc), padding is 0.  One all-gather of 8 B per
// element instead of two of 4 + 8 B, and the merge compares plain integers.
__global__ void __launch_bounds__(256)
topk_pack_kernel(const float* __restrict__ scores, const int64_t* __restrict__ items, int64_t n,
                 unsigned long long* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int64_t it = items[i];
  unsigned long long c = 0ull;
  if (it >= 0) c = ((unsigned long long)f32_to_key(__float_as_uint(scores[i] + 0.0f)) << 32) | (unsigned long long)(~(uint32_t)it);
  out[i] = c;
}

// One warp per user: the n_shards * k composites of the row are staged in shared memory (coalesced), each
// element's output rank = its rank in its own list + the number of larger composites in every other list
// (composites are unique: ids differ across shards).
__global__ void __launch_bounds__(128)
topk_merge_packed_kernel(const unsigned long long* __restrict__ g, int n_shards, int64_t nU, int k,
                         float* __restrict__ out_scores, int64_t* __restrict__ out_items) {
  extern __shared__ __align__(16) unsigned long long mp_smem[];
  const int w = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int64_t row = (int64_t)blockIdx.x * 4 + w;
  if (row >= nU) return;
  const int total = n_shards * k;
  unsigned long long* a = mp_smem + (size_t)w * total;
  for (int e = lane; e < total; e += 32) {
    const int s = e / k, r = e - s * k;
    a[e] = g[((int64_t)s * nU + row) * k + r];
  }
  for (int r = lane; r < k; r += 32) { out_scores[row * k + r] = -INFINITY; out_items[row * k + r] = -1; }
  __syncwarp();
  for (int e = lane; e < total; e += 32) {
    const unsigned long long c = a[e];
    if (c == 0ull) continue;  // padding
    const int s = e / k, r = e - s * k;
    int rank = r;
    for (int s2 = 0; s2 < n_shards && rank < k; ++s2) {
      if (s2 == s) continue;
      const unsigned long long* l = a + s2 * k;
      int lo = 0, hi = k;  // first position whose composite is < c (lists are descending, padding 0 at the end)
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (l[mid] > c) lo = mid + 1; else hi = mid;
      }
      rank += lo;
    }
    if (rank < k) {
      out_scores[row * k + rank] = __uint_as_float(key_to_f32((uint32_t)(c >> 32)));
      out_items[row * k + rank] = (int64_t)(~(uint32_t)c);
    }
  }
}

}  // namespace

cudaError_t launch_topk_pack(const float* scores, const int64_t* items, int64_t n, uint64_t* out, cudaStream_t stream) {
  topk_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(scores, items, n, reinterpret_cast<unsigned long long*>(out));
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_topk_merge_packed(const uint64_t* g, int n_shards, int64_t nU, int k, float* out_scores,
                                     int64_t* out_items, cudaStream_t stream) {
  const size_t smem = (size_t)4 * n_shards * k * sizeof(unsigned long long);
  cudaError_t e = cudaFuncSetAttribute(topk_merge_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  topk_merge_packed_kernel<<<(unsigned)((nU + 3) / 4), 128, smem, stream>>>(reinterpret_cast<const unsigned long long*>(g), n_shards,
                                                                            nU, k, out_scores, out_items);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_select_topk(const SelectParams& p, cudaStream_t stream) {
  const int64_t grid = (p.nU + SEL_WARPS - 1) / SEL_WARPS;
  const bool vec = (p.d & 3) == 0 && p.d <= 256;  // wider rows take the scalar path (no fused metrics)
  const bool big = p.k > MAX_K_SMALL;             // k class: larger per-warp buffers, 2 CTAs per SM
  cudaError_t e;
#define DIVREC_LAUNCH_SELECT(NJ, DT, FULL, KB)                                                                              \
  do {                                                                                                                      \
    const size_t smem = (size_t)SEL_WARPS * SelCfg<KB>::SEL_WORDS * sizeof(uint32_t);                                       \
    e = cudaFuncSetAttribute(select_topk_kernel<NJ, DT, FULL, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                                                         \
    select_topk_kernel<NJ, DT, FULL, KB><<<(unsigned)grid, SEL_WARPS * 32, smem, stream>>>(p);                              \
  } while (0)
  if (big) {
    if (!vec) DIVREC_LAUNCH_SELECT(0, 0, false, 1);
    else if (p.dtype == 0 && p.d <= 128) DIVREC_LAUNCH_SELECT(4, 0, false, 1);
    else if (p.dtype == 0) DIVREC_LAUNCH_SELECT(8, 0, false, 1);
    else if (p.d <= 128) DIVREC_LAUNCH_SELECT(4, 1, false, 1);
    else DIVREC_LAUNCH_SELECT(8, 1, false, 1);
  }
  else if (!vec) DIVREC_LAUNCH_SELECT(0, 0, false, 0);
  else if (p.dtype == 0 && p.d == 128) DIVREC_LAUNCH_SELECT(4, 0, true, 0);   // the common full-row shapes get branch-free loads
  else if (p.dtype == 0 && p.d <= 128) DIVREC_LAUNCH_SELECT(4, 0, false, 0);
  else if (p.dtype == 0 && p.d == 256) DIVREC_LAUNCH_SELECT(8, 0, true, 0);
  else if (p.dtype == 0) DIVREC_LAUNCH_SELECT(8, 0, false, 0);
  else if (p.d == 128) DIVREC_LAUNCH_SELECT(4, 1, true, 0);
  else if (p.d <= 128) DIVREC_LAUNCH_SELECT(4, 1, false, 0);
  else if (p.d == 256) DIVREC_LAUNCH_SELECT(8, 1, true, 0);
  else DIVREC_LAUNCH_SELECT(8, 1, false, 0);
#undef DIVREC_LAUNCH_SELECT
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
