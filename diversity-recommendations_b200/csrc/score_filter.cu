// score_filter.cu -- full-catalogue scoring on tcgen05 tensor cores fused with a per-user
// threshold filter (the first half of score + mask + top-k).
//
// Replaces the reference's per-user loop: divrec/train/utils.py:59-76 (get_model_recommendations)
// calling divrec/models/matrix_factorization.py:26-28 on the candidates produced by
// divrec/datasets/base_datasets.py:136-171.  The reference scores L = I - |seen_u| items per
// user, fully sorts them and keeps k; here S = U V^T is produced tile by tile in tensor memory
// and never written to HBM: each epilogue thread owns one user row, compares the tile's scores
// with the row's running threshold tau (a lower bound on the row's k-th best unmasked score) and
// appends the few survivors to a per-row candidate list.  When a list fills up, the owning warp
// compacts it (removes seen items, selects the k best, raises tau).  select_topk.cu finishes.
//
// Block roles (384 threads, 1 CTA / SM, persistent over 128-user blocks):
//   warps 0-7  epilogue: warp w reads TMEM lanes 32*(w%4).. and columns 64*(w/4).. of every tile
//   warp  8    TMA producer (one lane): A block once per user block, B k-blocks through a ring
//   warp  9    MMA issuer (one lane): tcgen05.mma kind::f16, M=128 N=128 K=16, fp32 accum in TMEM
//   warp  10   TMEM allocator
#include "common.cuh"
#include "kernels.h"

namespace divrec {

namespace {

constexpr int BM = 128;            // users per block  (TMEM lanes)
constexpr int BN = 128;            // items per tile   (TMEM columns per accumulator stage)
constexpr int BK = 64;             // 16-bit elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;         // K per tcgen05.mma for 16-bit inputs
constexpr int SLAB_BYTES = BM * BK * 2;  // 16 KB: 128 rows x 128 B
constexpr int ACC_STAGES = 4;      // 4 x 128 fp32 columns = all 512 TMEM columns
constexpr int B_STAGES = 6;
constexpr int EPI_WARPS = 8;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int NUM_THREADS = 384;
constexpr int SLOTS = LIST_CAP / 32;  // candidate slots per lane during compaction
constexpr int COLS_PER_THREAD = BN / 2;

struct Smem {
  uint64_t b_full[B_STAGES];
  uint64_t b_empty[B_STAGES];
  uint64_t acc_full[ACC_STAGES];
  uint64_t acc_empty[ACC_STAGES];
  uint64_t a_full;
  uint64_t a_empty;
  uint32_t tmem_base;
  uint32_t pad;
  uint32_t tau_key[BM];  // per-row threshold as ordered key (max over both lists of the row)
};

__device__ __forceinline__ float max8(const uint32_t* v) {
  float a = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1]));
  float b = fmaxf(__uint_as_float(v[2]), __uint_as_float(v[3]));
  float c = fmaxf(__uint_as_float(v[4]), __uint_as_float(v[5]));
  float d = fmaxf(__uint_as_float(v[6]), __uint_as_float(v[7]));
  return fmaxf(fmaxf(a, b), fmaxf(c, d));
}

// Warp-cooperative compaction of one candidate list (all 32 lanes call with uniform arguments).
//   L[0..n): (score bits, local item id), ascending item id.  seen[cursor..s_end): the row's seen
//   items (global ids, ascending) not yet applied to this list.
// Removes seen items, keeps the k best by (score desc, id asc) preserving id order, returns the
// new length; T_out = ordered key of a valid lower bound of the row's k-th best unmasked score.
__device__ __forceinline__ int compact_list(uint2* __restrict__ L, int n, const int32_t* __restrict__ seen,
                                            int64_t& cursor, int64_t s_end, int64_t item_off,
                                            int32_t nI, int k, uint32_t& T_out) {
  const uint32_t lane = lane_id();
  const uint32_t lt = lanemask_lt();
  // 1. apply the seen-item mask to L (binary search of each new seen id in the id-sorted list)
  if (seen != nullptr && n > 0) {
    const int32_t last_id = (int32_t)L[n - 1].y;
    while (cursor < s_end) {
      const int64_t idx = cursor + lane;
      int64_t sid = INT64_MAX;
      if (idx < s_end) sid = (int64_t)seen[idx] - item_off;
      const bool active = sid <= (int64_t)last_id;
      if (active && sid >= 0 && sid < nI) {
        int lo = 0, hi = n;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          const int32_t y = (int32_t)L[mid].y;
          if (y < (int32_t)sid) lo = mid + 1; else hi = mid;
        }
        if (lo < n && (int32_t)L[lo].y == (int32_t)sid) L[lo].x = F32_NEG_INF_BITS;
      }
      const int nact = __popc(__ballot_sync(0xffffffffu, active));
      cursor += nact;
      if (nact < 32) break;
    }
  }
  __syncwarp();
  // 2. load the list into registers (slot-major = list order), build ordered keys; 0 = dropped
  uint32_t key[SLOTS];
  uint32_t id[SLOTS];
  int nv_local = 0;
  uint32_t kmin = 0xffffffffu, kmax = 0;
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int idx = s * 32 + lane;
    uint2 e = make_uint2(F32_NEG_INF_BITS, 0);
    if (idx < n) e = L[idx];
    uint32_t kk = f32_to_key(e.x);
    if (kk <= KEY_NEG_INF) kk = 0;  // masked (-inf) or out of range
    key[s] = kk;
    id[s] = e.y;
    if (kk) { nv_local++; kmin = min(kmin, kk); kmax = max(kmax, kk); }
  }
  const int nv = __reduce_add_sync(0xffffffffu, nv_local);
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  __syncwarp();
  // 3. k-th largest key by bisection (exits early when a pivot separates exactly k entries)
  uint32_t T = 0;          // keep everything valid
  int quota = 0x7fffffff;  // how many entries equal to T may stay
  bool raise = false;
  if (nv >= k) {
    uint32_t lo = kmin, hi = kmax + 1;
    int c_lo = nv, c_hi = 0;
    while (hi - lo > 1 && c_lo != k) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      int c = 0;
#pragma unroll
      for (int s = 0; s < SLOTS; ++s) c += (key[s] >= mid) ? 1 : 0;
      c = __reduce_add_sync(0xffffffffu, c);
      if (c >= k) { lo = mid; c_lo = c; } else { hi = mid; c_hi = c; }
    }
    T = lo;
    raise = true;
    if (c_lo != k) quota = k - c_hi;  // hi == lo + 1: T is the exact k-th key, ties trimmed by id
  }
  // 4. order-preserving write-back
  int outn = 0, tie_seen = 0;
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const bool valid = key[s] != 0;
    const bool gt = valid && key[s] > T;
    const bool eq = valid && key[s] == T;
    const uint32_t eqb = __ballot_sync(0xffffffffu, eq);
    const int my_rank = tie_seen + __popc(eqb & lt);
    const bool keep = gt || (eq && my_rank < quota);
    tie_seen += __popc(eqb);
    const uint32_t kb = __ballot_sync(0xffffffffu, keep);
    if (keep) L[outn + __popc(kb & lt)] = make_uint2(key_to_f32(key[s]), id[s]);
    outn += __popc(kb);
  }
  __syncwarp();
  T_out = raise ? T : KEY_NEG_INF;
  return outn;
}

template <bool DUMP>
__global__ void __launch_bounds__(NUM_THREADS, 1)
score_filter_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const FilterParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem: [A slabs: kblocks x 16 KB][B ring: B_STAGES x 16 KB][Smem control block]
  uint8_t* smem_base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_slabs = smem_base;
  uint8_t* b_ring = smem_base + (size_t)p.kblocks * SLAB_BYTES;
  Smem* sm = reinterpret_cast<Smem*>(b_ring + (size_t)B_STAGES * SLAB_BYTES);

  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();

  if (threadIdx.x == 0) {
    for (int i = 0; i < B_STAGES; ++i) { mbar_init(smem_u32(&sm->b_full[i]), 1); mbar_init(smem_u32(&sm->b_empty[i]), 1); }
    for (int i = 0; i < ACC_STAGES; ++i) { mbar_init(smem_u32(&sm->acc_full[i]), 1); mbar_init(smem_u32(&sm->acc_empty[i]), EPI_WARPS); }
    mbar_init(smem_u32(&sm->a_full), 1);
    mbar_init(smem_u32(&sm->a_empty), 1);
    fence_mbar_init();
  }
  if (warp == 8 && lane == 0) { tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_b); }
  if (warp == 10) { tmem_alloc(smem_u32(&sm->tmem_base), ACC_STAGES * BN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&sm->tmem_base);

  const int n_tiles = p.n_tiles;
  const int kblocks = p.kblocks;

  if (warp == 8) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, a_phase = 0;
      const uint64_t pol_keep = l2_policy_evict_last();   // item planes: re-read by every user block
      const uint64_t pol_stream = l2_policy_evict_first();  // user planes: read once
      for (int ub = blockIdx.x; ub < p.n_ublocks; ub += gridDim.x) {
        mbar_wait(smem_u32(&sm->a_empty), a_phase ^ 1);
        mbar_arrive_expect_tx(smem_u32(&sm->a_full), (uint32_t)kblocks * SLAB_BYTES);
        for (int kb = 0; kb < kblocks; ++kb)
          tma_load_2d_hint(smem_u32(a_slabs + (size_t)kb * SLAB_BYTES), &tmap_a, smem_u32(&sm->a_full),
                           kb * BK, ub * BM, pol_stream);
        a_phase ^= 1;
        for (int t = 0; t < n_tiles; ++t) {
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(smem_u32(&sm->b_empty[stage]), phase ^ 1);
            mbar_arrive_expect_tx(smem_u32(&sm->b_full[stage]), SLAB_BYTES);
            tma_load_2d_hint(smem_u32(b_ring + (size_t)stage * SLAB_BYTES), &tmap_b,
                             smem_u32(&sm->b_full[stage]), kb * BK, t * BN, pol_keep);
            if (++stage == B_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 9) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(p.fmt, BM, BN);
      uint32_t stage = 0, phase = 0, a_phase = 0, acc = 0, acc_phase = 0;
      for (int ub = blockIdx.x; ub < p.n_ublocks; ub += gridDim.x) {
        mbar_wait(smem_u32(&sm->a_full), a_phase);
        a_phase ^= 1;
        tc_fence_after();
        for (int t = 0; t < n_tiles; ++t) {
          mbar_wait(smem_u32(&sm->acc_empty[acc]), acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(smem_u32(&sm->b_full[stage]), phase);
            tc_fence_after();
            const uint64_t adesc = umma_desc_sw128(smem_u32(a_slabs + (size_t)kb * SLAB_BYTES));
            const uint64_t bdesc = umma_desc_sw128(smem_u32(b_ring + (size_t)stage * SLAB_BYTES));
#pragma unroll
            for (int k4 = 0; k4 < BK / UMMA_K; ++k4) {
              // advance 32 bytes (16 x 2 B) inside the 128-byte swizzle row: +2 in 16-byte units
              umma_f16(d_tmem, adesc + 2 * k4, bdesc + 2 * k4, idesc, (uint32_t)((kb | k4) != 0));
            }
            umma_commit(smem_u32(&sm->b_empty[stage]));
            if (++stage == B_STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit(smem_u32(&sm->acc_full[acc]));
          if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
        umma_commit(smem_u32(&sm->a_empty));
      }
    }
  } else if (warp < EPI_WARPS) {
    // ================= epilogue: threshold filter =================
    const int q = warp & 3;   // TMEM lane quarter
    const int h = warp >> 2;  // column half of every tile
    const int row_local = q * 32 + lane;
    uint32_t acc = 0, acc_phase = 0;
    for (int ub = blockIdx.x; ub < p.n_ublocks; ub += gridDim.x) {
      const int64_t row = (int64_t)ub * BM + row_local;
      const bool valid_row = row < p.nU;
      // all epilogue warps are done with the previous block before its thresholds are reset
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
      if (h == 0) sm->tau_key[row_local] = valid_row ? KEY_NEG_INF : 0xffffffffu;
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");

      uint2* L = p.lists + ((size_t)row * 2 + h) * LIST_CAP;
      int cnt = 0;
      int64_t cursor = 0, s_end = 0;
      if (valid_row && p.seen_rowptr != nullptr) { cursor = p.seen_rowptr[row]; s_end = p.seen_rowptr[row + 1]; }

      for (int t = 0; t < n_tiles; ++t) {
        // ---- make room: a thread appends at most COLS_PER_THREAD entries per tile ----
        uint32_t need = __ballot_sync(0xffffffffu, valid_row && cnt > LIST_CAP - COLS_PER_THREAD);
        while (need) {
          const int src = __ffs(need) - 1;
          need &= need - 1;
          const int n_src = __shfl_sync(0xffffffffu, cnt, src);
          int64_t cur_src = __shfl_sync(0xffffffffu, cursor, src);
          const int64_t end_src = __shfl_sync(0xffffffffu, s_end, src);
          uint2* Lsrc = p.lists + (((size_t)ub * BM + q * 32 + src) * 2 + h) * LIST_CAP;
          uint32_t T;
          const int n_new = compact_list(Lsrc, n_src, p.seen_items, cur_src, end_src, p.item_id_offset, p.nI, p.k, T);
          if ((int)lane == src) { cnt = n_new; cursor = cur_src; }
          if (lane == 0) atomicMax(&sm->tau_key[q * 32 + src], T);
        }
        __syncwarp();
        const float tau = __uint_as_float(key_to_f32(*reinterpret_cast<volatile uint32_t*>(&sm->tau_key[row_local])));

        // ---- TMEM -> registers ----
        mbar_wait(smem_u32(&sm->acc_full[acc]), acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + h * COLS_PER_THREAD;
        uint32_t v0[32], v1[32];
        tmem_ld_32x32b_x32(taddr, v0);
        tmem_ld_32x32b_x32(taddr + 32, v1);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sm->acc_empty[acc]));
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }

        const int item0 = t * BN + h * COLS_PER_THREAD;
        if (DUMP) {
          float dump_scale = 1.0f;
          if (p.scale_exps != nullptr) dump_scale = ldexpf(ldexpf(1.0f, -p.scale_exps[0]), -p.scale_exps[1]);
          if (valid_row) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (item0 + j < p.nI) p.dump[(size_t)row * p.nI + item0 + j] = __uint_as_float(v0[j]) * dump_scale;
              if (item0 + 32 + j < p.nI) p.dump[(size_t)row * p.nI + item0 + 32 + j] = __uint_as_float(v1[j]) * dump_scale;
            }
          }
        } else {
          // ---- filter: group maxima first, element tests only in groups that can contain a survivor ----
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (max8(&v0[8 * g]) >= tau) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float s = __uint_as_float(v0[8 * g + j]);
                const int item = item0 + 8 * g + j;
                if (s >= tau && item < p.nI) { L[cnt] = make_uint2(__float_as_uint(s + 0.0f), (uint32_t)item); ++cnt; }
              }
            }
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (max8(&v1[8 * g]) >= tau) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float s = __uint_as_float(v1[8 * g + j]);
                const int item = item0 + 32 + 8 * g + j;
                if (s >= tau && item < p.nI) { L[cnt] = make_uint2(__float_as_uint(s + 0.0f), (uint32_t)item); ++cnt; }
              }
            }
          }
        }
        __syncwarp();
      }
      if (!DUMP && valid_row) p.list_count[row * 2 + h] = cnt;
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 10) { tc_fence_after(); tmem_dealloc(tmem_base, ACC_STAGES * BN); }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* f = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(f);
  return fn;
}

// rows x K 16-bit row-major table -> tensor map with a {64 x 128-row} box, 128-byte swizzle
bool make_tmap(CUtensorMap* m, const void* base, int64_t rows, int K, int fmt) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {BK, BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, fmt == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace

size_t score_filter_smem_bytes(int kblocks) {
  return 1024 + (size_t)(kblocks + B_STAGES) * SLAB_BYTES + sizeof(Smem);
}

// A16 [nU, K] and B16 [nI, K] are K-major 16-bit planes (fmt 0 = fp16, 1 = bf16), K = 64 * kblocks.
cudaError_t launch_score_filter(const void* A16, const void* B16, FilterParams p, int sm_count, cudaStream_t stream,
                                const char** err) {
  const int K = p.kblocks * BK;
  CUtensorMap ta, tb;
  if (!make_tmap(&ta, A16, p.nU, K, p.fmt) || !make_tmap(&tb, B16, p.nI, K, p.fmt)) {
    *err = "cuTensorMapEncodeTiled failed";
    return cudaErrorInvalidValue;
  }
  p.n_ublocks = (int)((p.nU + BM - 1) / BM);
  p.n_tiles = (p.nI + BN - 1) / BN;
  const size_t smem = score_filter_smem_bytes(p.kblocks);
  const int grid = p.n_ublocks < sm_count ? p.n_ublocks : sm_count;
  cudaError_t e;
  if (p.dump) {
    e = cudaFuncSetAttribute(score_filter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    score_filter_kernel<true><<<grid, NUM_THREADS, smem, stream>>>(ta, tb, p);
  } else {
    e = cudaFuncSetAttribute(score_filter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    score_filter_kernel<false><<<grid, NUM_THREADS, smem, stream>>>(ta, tb, p);
  }
  count_launch();
  return cudaGetLastError();
}

}  // namespace divrec
