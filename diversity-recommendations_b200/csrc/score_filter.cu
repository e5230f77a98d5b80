// score_filter.cu -- full-catalogue scoring on tcgen05 tensor cores fused with a per-user
// threshold filter (the first half of score + mask + top-k).
//
// Replaces the reference's per-user loop: divrec/train/utils.py:59-76 (get_model_recommendations)
// calling divrec/models/matrix_factorization.py:26-28 on the candidates produced by
// divrec/datasets/base_datasets.py:136-171.  The reference scores L = I - |seen_u| items per user,
// fully sorts them and keeps k.  Here S = U V^T is produced tile by tile in tensor memory from
// 16-bit operand planes (one pass, fp32 accumulate) and never written to HBM.  The tensor-core
// score A is only a FILTER: it differs from the exact fp32 score E by at most delta_u (prep.cu), the
// candidates it lets through are re-scored exactly in fp32 by select_topk.cu, so the final top-k is
// the top-k of the exact scores.
//
// Per 128-user block:
//   1. hint phase: N_SAMPLE item tiles spread over the catalogue are scored; every epilogue thread
//      keeps the maxima of 32-column groups (64 maxima per user row), and the row threshold thr is
//      set to the r-th largest of them, r chosen from k and the row's seen-item count so that with
//      probability ~1 - 1e-3 at least k + seen_u items of the catalogue score above thr
//      (distribution-free order statistics);
//   2. main phase: all item tiles in ascending order; a thread owns one user row and 64 columns of
//      each tile; 8-column groups whose maximum reaches thr are appended whole (8 scores + first
//      item id) to the thread's candidate list in global memory -- no per-element work, no
//      threshold updates, no compaction.
// thr is a hint, not a bound: select_topk.cu verifies per row that enough candidates were captured
// (k unmasked candidates at least 2 delta above thr) and hands rows that fail, or whose list
// overflowed, to the exact SIMT kernel in rescue.cu.  Small catalogues (< HINT_MIN_TILES tiles) skip
// phase 1 and keep every group.
//
// Block roles (608 threads, 1 CTA / SM, persistent over user blocks):
//   warps 0-15 epilogue in two sets of 8: set s = w/8 takes the item tiles of parity s, so four warps per
//              scheduler partition hide each other's TMEM-load and store latency; inside a set warp w reads
//              TMEM lanes 32*(w%4).. and columns 64*((w/4)%2).. of its tiles.  A user row therefore owns four
//              candidate lists (column half x tile parity).
//   warp  16   TMA producer (one lane): A block once per user block, B k-blocks through a ring
//   warp  17   MMA issuer (one lane): tcgen05.mma kind::f16, M=128 N=128 K=16, fp32 accum in TMEM
//   warp  18   TMEM allocator
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace divrec {

namespace {

constexpr int BM = 128;            // users per block  (TMEM lanes)
constexpr int BN = 128;            // items per tile   (TMEM columns per accumulator stage)
constexpr int BK = 64;             // 16-bit elements per k-block = one 128-byte swizzle row
constexpr int SLAB_BYTES = BM * BK * 2;  // 16 KB: 128 rows x 128 B
constexpr int TMEM_COLS = 512;     // accumulator stages (128 fp32 columns each) from column 0, the A operand at the top
constexpr int EPI_WARPS = 8;       // warps that consume one tile (one set)
constexpr int EPI_SETS = 2;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int W_TMA = EPI_WARPS * EPI_SETS, W_MMA = W_TMA + 1, W_ALLOC = W_TMA + 2;
constexpr int NUM_THREADS = (W_ALLOC + 1) * 32;
constexpr int COLS_PER_THREAD = BN / 2;
constexpr int GROUPS_PER_TILE = COLS_PER_THREAD / GROUP;  // 8 groups of 8 columns per thread per tile
constexpr int HS = 2 * N_SAMPLE;   // 32-column group maxima per thread in the hint phase
constexpr int SAMPLE_ITEMS_PER_TPG = 156250;  // catalogues >= TIGHTEN_MIN_ITEMS sample 16 * ceil(nI / this) tiles (8 per hint maximum at 1.25M items: 1.3 % of the pass)

struct Smem {
  uint64_t b_full[16];
  uint64_t b_empty[16];
  uint64_t acc_full[4];
  uint64_t acc_empty[4];
  uint64_t a_full;
  uint64_t a_empty;
  uint32_t tmem_base;
  uint32_t pad;
  float sorted[HS][EPI_THREADS];  // per-thread descending group maxima, exchanged between the two halves of a row
};

__device__ __forceinline__ float max3(float a, float b, float c) {
  float m;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(m) : "f"(a), "f"(b), "f"(c));
  return m;
}
__device__ __forceinline__ float max8(const uint32_t* v) {
  const float a = max3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
  const float b = max3(__uint_as_float(v[3]), __uint_as_float(v[4]), __uint_as_float(v[5]));
  return max3(a, b, fmaxf(__uint_as_float(v[6]), __uint_as_float(v[7])));
}
__device__ __forceinline__ float max32(const uint32_t* v) {
  const float a = max3(max8(v), max8(v + 8), max8(v + 16));
  return fmaxf(a, max8(v + 24));
}

// 2 delta of a user row in the (scaled) units of the tensor-core scores; same formula as select_topk.cu
__device__ __forceinline__ float row_margin2(const FilterParams& p, int64_t row) {
  int e2 = 0;
  if (p.scale_exps != nullptr) e2 = p.scale_exps[1];
  if (p.row_exps != nullptr) e2 += p.row_exps[row];
  const float nu = p.norms_u[row], nvmax = __uint_as_float(*p.nvmax_bits);
  float delta = p.c_margin * nu * nvmax;
  if (p.amax_bits != nullptr) {
    const float av = __uint_as_float(p.amax_bits[1]);
    delta += 3.64e-12f * sqrtf((float)p.d) * (nu * av + nvmax * nu);
  }
  return ldexpf(2.0f * delta + REC_ID_SLACK * nu * nvmax, e2) * 1.0001f;
}

// first item id of a candidate record: the group index rides in the low mantissa bytes of its first scores (kernels.h)
__device__ __forceinline__ uint32_t rec_item0(const float4& a, int id_bytes) {
  uint32_t gi = (__float_as_uint(a.x) & 0xffu) | ((__float_as_uint(a.y) & 0xffu) << 8);
  if (id_bytes > 2) gi |= ((__float_as_uint(a.z) & 0xffu) << 16) | ((__float_as_uint(a.w) & 0xffu) << 24);
  return gi * GROUP;
}

// Warp-cooperative tightening of ONE nearly full candidate list (large catalogues: the sampled hint cannot
// resolve a 1e-5 quantile, so the list fills many times during a pass).  X = a lower bound of the need-th
// largest GROUP MAXIMUM held by the list (need = k + seen_u + 8 distinct items score >= X, so at least k unmasked ones do):
// the k-th best unmasked score of the row is >= X and nothing below X - 2 delta can reach the top-k.  The
// threshold is raised to X - 2 delta and the groups that fall below it are dropped, order preserved.
// Returns the new group count; thr_io is updated.  All 32 lanes call with the same arguments.
__device__ __noinline__ uint32_t tighten_list(float* Ls, uint32_t n, int need, float margin2, uint32_t nI, int id_bytes,
                                              float& thr_io) {
  const uint32_t lane = lane_id();
  const uint32_t lt = lanemask_lt();
  constexpr int MAXPER = LIST_CAP_MAX / 32;  // groups per lane at the largest list capacity
  // pass 1: group maxima only (each is a real item's score): MAXPER registers per lane, MAXPER compares per bisection step
  float gm[MAXPER];
  int n_el = 0;
  uint32_t kmin = 0xffffffffu, kmax = 0;
#pragma unroll
  for (int j = 0; j < MAXPER; ++j) {
    const uint32_t g = (uint32_t)j * 32 + lane;
    gm[j] = -INFINITY;
    if (g < n) {
      float4 a = *reinterpret_cast<const float4*>(Ls + (size_t)g * REC_FLOATS);
      float4 b = *reinterpret_cast<const float4*>(Ls + (size_t)g * REC_FLOATS + 4);
      const uint32_t it0 = rec_item0(a, id_bytes);
      if (it0 + 7 >= nI) {  // columns past the catalogue (zero-filled operand rows) are not items
        if (it0 + 0 >= nI) a.x = -INFINITY; if (it0 + 1 >= nI) a.y = -INFINITY;
        if (it0 + 2 >= nI) a.z = -INFINITY; if (it0 + 3 >= nI) a.w = -INFINITY;
        if (it0 + 4 >= nI) b.x = -INFINITY; if (it0 + 5 >= nI) b.y = -INFINITY;
        if (it0 + 6 >= nI) b.z = -INFINITY; if (it0 + 7 >= nI) b.w = -INFINITY;
      }
      gm[j] = fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)), fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
      if (gm[j] > -INFINITY) {
        const uint32_t kk = f32_to_key(__float_as_uint(gm[j] + 0.0f));
        ++n_el; kmin = min(kmin, kk); kmax = max(kmax, kk);
      }
    }
  }
  n_el = __reduce_add_sync(0xffffffffu, n_el);
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  if (n_el < need) return n;  // cannot certify a higher threshold: the caller declares the list overflowed
  // bisection for a key with count(group maxima >= key) in [need, need + need/8]
  uint32_t lo = kmin, hi = kmax + 1;
  int c_lo = n_el;
  const int c_ok = need + (need >> 3);
  while (hi - lo > 1 && c_lo > c_ok) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    const float fm = __uint_as_float(key_to_f32(mid));
    int c = 0;
#pragma unroll
    for (int j = 0; j < MAXPER; ++j) c += (gm[j] >= fm) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= need) { lo = mid; c_lo = c; } else { hi = mid; }
  }
  const float X = __uint_as_float(key_to_f32(lo));
  const float thr_new = fmaxf(thr_io, X - margin2);
  thr_io = thr_new;
  // pass 2: order-preserving compaction, 32 groups at a time (a chunk is read completely before anything is
  // written, and writes land at or before the chunk, so later chunks are still intact)
  uint32_t outn = 0;
#pragma unroll
  for (int j = 0; j < MAXPER; ++j) {
    if ((uint32_t)j * 32 >= n) break;
    const uint32_t g = (uint32_t)j * 32 + lane;
    const bool keep = g < n && gm[j] >= thr_new;
    float4 a, b;
    if (keep) {
      a = *reinterpret_cast<const float4*>(Ls + (size_t)g * REC_FLOATS);
      b = *reinterpret_cast<const float4*>(Ls + (size_t)g * REC_FLOATS + 4);
    }
    const uint32_t kb = __ballot_sync(0xffffffffu, keep);
    __syncwarp();
    if (keep) {
      const uint32_t pos = outn + __popc(kb & lt);
      *reinterpret_cast<float4*>(Ls + (size_t)pos * REC_FLOATS) = a;   // the record carries its own index
      *reinterpret_cast<float4*>(Ls + (size_t)pos * REC_FLOATS + 4) = b;
    }
    outn += __popc(kb);
    __syncwarp();
  }
  return outn;
}

// tile visited at schedule position s: N_SAMPLE strided sample tiles first (when hints are on),
// then every tile in ascending order
__device__ __forceinline__ int sched_tile(int s, int n_sample, int stride) {
  return s < n_sample ? s * stride : s - n_sample;
}

// TIGHTEN: large catalogues only (the out-of-line tightening call costs the hot loop ~15 % in register shuffling,
// and below ~500k items the sampled hint keeps the lists short anyway).
// The MMA-issuing thread's loop, one instantiation per k-block count (see score_filter2.cu: the thread shares its scheduler
// partition with a quarter of the epilogue warps, so its instruction count per tile is what has to be small): k-blocks at
// compile time, descriptors advanced in their low word only, all of a tile's stages awaited before its first MMA.
template <int KB>
__device__ __forceinline__ void mma_issue_loop1(Smem* sm, const FilterParams& p, uint32_t tmem_base, uint32_t b_ring_addr,
                                                int n_sched) {
  const uint32_t idesc = umma_idesc_f16(p.fmt, BM, BN);
  const uint64_t bdesc0 = umma_desc_sw128(b_ring_addr);
  const uint32_t d_hi = (uint32_t)(bdesc0 >> 32), d_lo0 = (uint32_t)bdesc0;
  const uint32_t bar_b_full = smem_u32(&sm->b_full[0]), bar_b_empty = smem_u32(&sm->b_empty[0]);
  const uint32_t bar_acc_full = smem_u32(&sm->acc_full[0]), bar_acc_empty = smem_u32(&sm->acc_empty[0]);
  const uint32_t bar_a_full = smem_u32(&sm->a_full), bar_a_empty = smem_u32(&sm->a_empty);
  const uint32_t b_stages = (uint32_t)p.b_stages, acc_stages = (uint32_t)p.acc_stages;
  const uint32_t a_op = tmem_base + TMEM_COLS - KB * (BK / 2);
  uint32_t stage = 0, phase = 0, a_phase = 0, acc = 0, acc_phase = 0;
  for (int ub = blockIdx.x; ub < p.n_ublocks; ub += gridDim.x) {
    mbar_wait(bar_a_full, a_phase);  // the block's A operand has been written to TMEM
    a_phase ^= 1;
    tc_fence_after();
    for (int s = 0; s < n_sched; ++s) {
      uint32_t st[KB];
      mbar_wait(bar_acc_empty + 8 * acc, acc_phase ^ 1);
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        st[kb] = stage;
        mbar_wait(bar_b_full + 8 * stage, phase);
        if (++stage == b_stages) { stage = 0; phase ^= 1; }
      }
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        const uint32_t lo = d_lo0 + st[kb] * (uint32_t)(SLAB_BYTES >> 4);
        // B: +32 bytes (16 x 2 B) inside the 128-byte swizzle row = +2 in 16-byte units; A: 16 elements = 8 TMEM columns
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
          umma_f16_ts(d_tmem, a_op + (uint32_t)(kb * (BK / 2) + 8 * k4), ((uint64_t)d_hi << 32) | (uint64_t)(lo + 2u * k4), idesc,
                      (uint32_t)((kb | k4) != 0));
        umma_commit(bar_b_empty + 8 * st[kb]);
      }
      umma_commit(bar_acc_full + 8 * acc);
      if (++acc == acc_stages) { acc = 0; acc_phase ^= 1; }
    }
    umma_commit(bar_a_empty);  // every MMA that reads this block's A has completed
  }
}

template <bool TIGHTEN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
score_filter_kernel(const __grid_constant__ CUtensorMap tmap_b, const FilterParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem: [B ring: b_stages x 16 KB][Smem control block]; the A operand of the user block lives in TMEM
  uint8_t* smem_base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* b_ring = smem_base;
  Smem* sm = reinterpret_cast<Smem*>(b_ring + (size_t)p.b_stages * SLAB_BYTES);

  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int b_stages = p.b_stages;

  if (threadIdx.x == 0) {
    for (int i = 0; i < b_stages; ++i) { mbar_init(smem_u32(&sm->b_full[i]), 1); mbar_init(smem_u32(&sm->b_empty[i]), 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(smem_u32(&sm->acc_full[i]), 1); mbar_init(smem_u32(&sm->acc_empty[i]), EPI_WARPS); }
    mbar_init(smem_u32(&sm->a_full), EPI_WARPS);  // the eight set-0 warps that write A into TMEM
    mbar_init(smem_u32(&sm->a_empty), 1);
    fence_mbar_init();
  }
  if (warp == W_TMA && lane == 0) tma_prefetch_desc(&tmap_b);
  if (warp == W_ALLOC) { tmem_alloc(smem_u32(&sm->tmem_base), TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&sm->tmem_base);

  const int TPG = TIGHTEN ? p.sample_tpg : 1;  // sample tiles folded into one hint maximum
  const int kblocks = p.kblocks;
  const int n_sample = p.n_sample;
  const int n_sched = n_sample + p.n_tiles;
  const uint32_t acc_stages = (uint32_t)p.acc_stages;            // (512 - K/2) / 128: 3 for K <= 256, 2 up to K = 512
  const uint32_t a_col = TMEM_COLS - (uint32_t)kblocks * (BK / 2);  // A: K/2 packed columns at the top of TMEM

  if (warp == W_TMA) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint64_t pol_keep = l2_policy_evict_last();     // item plane: re-read by every user block
      for (int ub = blockIdx.x; ub < p.n_ublocks; ub += gridDim.x) {
        for (int s = 0; s < n_sched; ++s) {
          const int t = sched_tile(s, n_sample, p.sample_stride);
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(smem_u32(&sm->b_empty[stage]), phase ^ 1);
            mbar_arrive_expect_tx(smem_u32(&sm->b_full[stage]), SLAB_BYTES);
            tma_load_2d_hint(smem_u32(b_ring + (size_t)stage * SLAB_BYTES), &tmap_b,
                             smem_u32(&sm->b_full[stage]), kb * BK, t * BN, pol_keep);
            if (++stage == (uint32_t)b_stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == W_MMA) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t ring = smem_u32(b_ring);
      switch (kblocks) {
        case 1: mma_issue_loop1<1>(sm, p, tmem_base, ring, n_sched); break;
        case 2: mma_issue_loop1<2>(sm, p, tmem_base, ring, n_sched); break;
        case 3: mma_issue_loop1<3>(sm, p, tmem_base, ring, n_sched); break;
        case 4: mma_issue_loop1<4>(sm, p, tmem_base, ring, n_sched); break;
        case 5: mma_issue_loop1<5>(sm, p, tmem_base, ring, n_sched); break;
        case 6: mma_issue_loop1<6>(sm, p, tmem_base, ring, n_sched); break;
        case 7: mma_issue_loop1<7>(sm, p, tmem_base, ring, n_sched); break;
        default: mma_issue_loop1<8>(sm, p, tmem_base, ring, n_sched); break;
      }
    }
  } else if (warp < EPI_WARPS * EPI_SETS) {
    // ================= epilogue =================
    const int q = warp & 3;          // TMEM lane quarter
    const int h = (warp >> 2) & 1;   // column half of a tile
    const int pset = warp >> 3;      // tile parity this warp consumes
    const int row_local = q * 32 + lane;
    const int et = (warp & 7) * 32 + lane;  // 0..255 inside the set
    uint32_t base_seq = 0;                   // tiles issued before this user block (all sets count every tile)
    uint32_t a_phase = 0;
    for (int ub = blockIdx.x; ub < p.n_ublocks; ub += gridDim.x, base_seq += (uint32_t)n_sched) {
      const int64_t row = (int64_t)ub * BM + row_local;
      const bool valid_row = row < p.nU;
      float thr = valid_row ? -INFINITY : INFINITY;

      // ---------------- phase 0: set 0 writes the block's A operand into TMEM (half of the K columns per thread) ----------------
      if (pset == 0) {
        mbar_wait(smem_u32(&sm->a_empty), a_phase ^ 1);  // the previous block's MMAs are done with the old A
        a_phase ^= 1;
        tc_fence_after();
        const int words = kblocks * (BK / 4);            // 32-bit words per thread: K/2 columns split over the two halves
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.a16) + (size_t)row * (kblocks * (BK / 2)) + (size_t)h * words;
        const uint32_t tdst = tmem_base + ((uint32_t)(q * 32) << 16) + a_col + (uint32_t)h * words;
        for (int w0 = 0; w0 < words; w0 += 16) {
          uint32_t r[16];
          if (valid_row) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const uint4 t4 = __ldg(reinterpret_cast<const uint4*>(src + w0 + j));
              r[j] = t4.x; r[j + 1] = t4.y; r[j + 2] = t4.z; r[j + 3] = t4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = 0u;
          }
          tmem_st_32x32b_x16(tdst + w0, r);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sm->a_full));
      }

      // ---------------- phase 1: threshold hint from sampled tiles (set 0 scores them, both sets use the result) ----------------
      if (n_sample > 0) {
        if (pset == 0) {
          // hs[] = maxima of 32-column groups, each followed through TPG consecutive sample tiles (large catalogues sample
          // 4x the tiles so that the hint resolves their much smaller quantiles; the number of maxima stays HS)
          float hs[HS];
#pragma unroll
          for (int sg = 0; sg < N_SAMPLE; ++sg) {
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll 1
            for (int tt = 0; tt < TPG; ++tt) {
              const uint32_t seq = base_seq + (uint32_t)(sg * TPG + tt), acc = seq % acc_stages, acc_phase = (seq / acc_stages) & 1;
              mbar_wait(smem_u32(&sm->acc_full[acc]), acc_phase);
              tc_fence_after();
              const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + h * COLS_PER_THREAD;
              uint32_t v0[32];
              tmem_ld_32x32b_x32(taddr, v0);
              tmem_wait_ld();
              m0 = fmaxf(m0, max32(v0));
              tmem_ld_32x32b_x32(taddr + 32, v0);
              tmem_wait_ld();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(smem_u32(&sm->acc_empty[acc]));
              m1 = fmaxf(m1, max32(v0));
            }
            hs[2 * sg] = m0;
            hs[2 * sg + 1] = m1;
          }
          // bitonic sort, descending, fully unrolled in registers
#pragma unroll
          for (int size = 2; size <= HS; size <<= 1) {
#pragma unroll
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
              for (int i = 0; i < HS; ++i) {
                const int j = i ^ stride;
                if (j > i) {
                  const bool desc = (i & size) == 0;
                  const float a = hs[i], b = hs[j];
                  const float mx = fmaxf(a, b), mn = fminf(a, b);
                  hs[i] = desc ? mx : mn;
                  hs[j] = desc ? mn : mx;
                }
              }
            }
          }
#pragma unroll
          for (int i = 0; i < HS; ++i) sm->sorted[i][et] = hs[i];
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS * EPI_SETS) : "memory");
        if (valid_row) {
          // r-th largest of the row's 2*HS group maxima (both column halves)
          int64_t seen_u = 0;
          if (p.seen_rowptr != nullptr) seen_u = p.seen_rowptr[row + 1] - p.seen_rowptr[row];
          const float f0 = ((float)p.k + (float)seen_u + 8.0f) / (float)p.nI;  // needed fraction of the catalogue
          int r = 2 * HS + 1;                                                   // "no hint"
          if (f0 < 0.25f) {
            const float pg = 1.0f - __expf((float)(32 * TPG) * log1pf(-f0));  // P(the maximum of a group's 32 TPG sampled scores is in that fraction)
            const float n = (float)(2 * HS);
            if (TIGHTEN && p.hint_eps > 0.0f) {
              // exact binomial tail (the normal approximation is far off when n pg < 1: "5 sigma" at n pg = 0.7 is a 1e-5
              // tail, and an uncertified row of a multi-million-item catalogue costs hundreds of ms in the exact kernel):
              // smallest r past the mode with P(Bin(n, pg) >= r) <= pmf(r) / (1 - ratio_r) <= eps
              const float qq = pg / (1.0f - pg);
              float pmf = __expf(n * log1pf(-pg));
              for (int j = 0; j <= 2 * HS; ++j) {
                const float ratio = (n - (float)j) / ((float)j + 1.0f) * qq;  // pmf(j + 1) / pmf(j), decreasing in j
                if ((float)j > n * pg && ratio < 0.5f && pmf <= p.hint_eps * (1.0f - ratio)) { r = j; break; }
                pmf *= ratio;
              }
            } else {
              r = (int)ceilf(n * pg + p.hint_sigma * sqrtf(n * pg * (1.0f - pg))) + p.hint_extra;
            }
          }
          if (r <= 2 * HS - 4) {
            const int ea = row_local, eb = row_local + 128;  // the row's two column halves in `sorted`
            int i = 0, j = 0;
            float pick = INFINITY;
            for (int c = 0; c < r; ++c) {
              const float a = i < HS ? sm->sorted[i][ea] : -INFINITY;
              const float b = j < HS ? sm->sorted[j][eb] : -INFINITY;
              if (a >= b) { pick = a; ++i; } else { pick = b; ++j; }
            }
            thr = pick;
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS * EPI_SETS) : "memory");  // sorted[] is reused by the next block
      }

      // ---------------- phase 2: fixed-threshold group filter over this set's tiles ----------------
      // block-relative 32-bit offsets (a user block's lists span 128 * 4 * cap groups) on uniform 64-bit bases
      float* const Ls_blk = p.lists + (size_t)ub * (BM * 4) * ((size_t)p.cap * REC_FLOATS);
      const uint32_t goff = (uint32_t)((row_local * 4 + h * 2 + pset) * p.cap);
      uint32_t cnt = 0;
      const int id_bytes = rec_id_bytes(p.nI);
      // TIGHTEN: room is checked on every second tile of this thread, so two tiles' worth of slots are kept free
      const uint32_t cap_lim = (uint32_t)(p.cap - (TIGHTEN ? 2 : 1) * GROUPS_PER_TILE);
      int t = pset;
      while (t < p.n_tiles) {
        if (TIGHTEN) {
          // rare (large catalogues): lists that cannot take the next tiles' groups are tightened by the whole warp.  This
          // sits OUTSIDE the hot tile loop (which only breaks out to get here), so the out-of-line call does not force the
          // loop's state into local memory; everything the slow path needs is recomputed.
          uint32_t full = __ballot_sync(0xffffffffu, valid_row && cnt > cap_lim && cnt <= (uint32_t)p.cap);
          while (full) {
            const int src = __ffs(full) - 1;
            full &= full - 1;
            const int64_t row_src = (int64_t)ub * BM + q * 32 + src;
            int64_t seen_src = 0;
            if (p.seen_rowptr != nullptr) seen_src = p.seen_rowptr[row_src + 1] - p.seen_rowptr[row_src];
            const int64_t need64 = (int64_t)p.k + seen_src + 8;
            const uint32_t n_src = __shfl_sync(0xffffffffu, cnt, src);
            float thr_src = __shfl_sync(0xffffffffu, thr, src);
            uint32_t n_new = n_src;
            if (need64 + (need64 >> 3) + 16 <= (int64_t)cap_lim) {
              const long long tt0 = p.stats != nullptr ? clock64() : 0;
              const uint32_t g_src = (uint32_t)(((q * 32 + src) * 4 + h * 2 + pset) * p.cap);
              n_new = tighten_list(Ls_blk + (size_t)g_src * REC_FLOATS, n_src, (int)need64, row_margin2(p, row_src),
                                   (uint32_t)p.nI, id_bytes, thr_src);
              if (p.stats != nullptr && lane == 0) {  // measurement aid: tightening events and their cycles per CTA
                atomicAdd(reinterpret_cast<unsigned long long*>(p.stats) + (size_t)blockIdx.x * 8 + 4, 1ull);
                atomicAdd(reinterpret_cast<unsigned long long*>(p.stats) + (size_t)blockIdx.x * 8 + 5, (unsigned long long)(clock64() - tt0));
              }
            }
            if ((int)lane == src) { cnt = n_new; thr = thr_src; }
          }
          if (cnt > cap_lim) { thr = INFINITY; cnt = (uint32_t)p.cap + 1; }  // still no room: stop, the row goes to the rescue kernel
        }
      for (; t < p.n_tiles; t += EPI_SETS) {
        if (!TIGHTEN) {
          if (cnt > cap_lim) { thr = INFINITY; cnt = (uint32_t)p.cap + 1; }  // overflow: stop, the row goes to the rescue kernel
        } else if ((t & 2) == 0 && __any_sync(0xffffffffu, valid_row && cnt > cap_lim && cnt <= (uint32_t)p.cap)) {
          break;  // a list of this warp is nearly full: tighten it (above), then resume with this tile
        }
        const uint32_t seq = base_seq + (uint32_t)(n_sample + t), acc = seq % acc_stages, acc_phase = (seq / acc_stages) & 1;
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

        if (p.debug_mode == 1) continue;   // measurement aid: MMA + TMA pipeline alone (no filtering)
        const float thr_st = p.debug_mode == 2 ? INFINITY : thr;  // measurement aid: filter arithmetic without the stores
        const uint32_t gi0 = (uint32_t)(t * (BN / GROUP) + h * GROUPS_PER_TILE);  // index of this thread's first group of the tile
#pragma unroll
        for (int g = 0; g < GROUPS_PER_TILE; ++g) {
          const uint32_t* v = g < 4 ? &v0[8 * g] : &v1[8 * (g - 4)];
          const float m = max8(v);
          float* dst = Ls_blk + (size_t)(goff + cnt) * REC_FLOATS;
          // the record carries its group index in the low mantissa bytes of its first scores (taken AFTER the maximum)
          const uint32_t gi = gi0 + (uint32_t)g;
          const uint32_t w0 = (v[0] & 0xffffff00u) | (gi & 0xffu), w1 = (v[1] & 0xffffff00u) | ((gi >> 8) & 0xffu);
          const uint32_t w2 = id_bytes > 2 ? ((v[2] & 0xffffff00u) | ((gi >> 16) & 0xffu)) : v[2];
          const uint32_t w3 = id_bytes > 2 ? ((v[3] & 0xffffff00u) | (gi >> 24)) : v[3];
          // branch-free append: ONE predicated 256-bit store (STG.E.ENL2.256)
          asm volatile(
              "{\n"
              ".reg .pred p;\n"
              "setp.ge.f32 p, %0, %1;\n"
              "@p st.global.v8.b32 [%2], {%3, %4, %5, %6, %7, %8, %9, %10};\n"
              "}\n"
              ::"f"(m), "f"(thr_st), "l"(dst), "r"(w0), "r"(w1), "r"(w2), "r"(w3), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
              : "memory");
          cnt += (m >= thr) ? 1u : 0u;
        }
      }
      }  // while (t < n_tiles)
      if (valid_row) {
        p.list_count[row * 4 + h * 2 + pset] = (int32_t)cnt;  // cap + 1 = overflowed
        p.thr_out[row * 4 + h * 2 + pset] = thr;               // per list: the hint, raised wherever the list was tightened
      }
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == W_ALLOC) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
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

bool make_tmap_b(CUtensorMap* m, const void* base, int64_t rows, int K, int fmt) { return make_tmap(m, base, rows, K, fmt); }

int score_filter_b_stages(int kblocks) {
  // B ring: as many 16 KB stages as fit (the A operand is in TMEM), at least two tiles' worth when possible
  (void)kblocks;
  int s = (200 * 1024 - (int)sizeof(Smem)) / SLAB_BYTES;
  if (s > 16) s = 16;
  return s;
}

size_t score_filter_smem_bytes(int kblocks) {
  return 1024 + (size_t)score_filter_b_stages(kblocks) * SLAB_BYTES + sizeof(Smem);
}

// A16 [nU, K] and B16 [nI, K] are K-major 16-bit planes (fmt 0 = fp16, 1 = bf16), K = 64 * kblocks.
cudaError_t launch_score_filter(const void* A16, const void* B16, FilterParams p, int sm_count, cudaStream_t stream,
                                const char** err) {
  const int K = p.kblocks * BK;
  CUtensorMap tb;
  if (!make_tmap(&tb, B16, p.nI, K, p.fmt)) {
    *err = "cuTensorMapEncodeTiled failed";
    return cudaErrorInvalidValue;
  }
  if (K > 512) { *err = "embedding dimension too large for the tensor-memory operand (d <= 512)"; return cudaErrorInvalidValue; }
  p.a16 = A16;
  p.n_ublocks = (int)((p.nU + BM - 1) / BM);
  p.n_tiles = (p.nI + BN - 1) / BN;
  p.b_stages = score_filter_b_stages(p.kblocks);
  p.acc_stages = (TMEM_COLS - K / 2) / BN;
  // Large catalogues: an uncertified row costs an exact pass over millions of items (measured: 522 ms for ONE row of a
  // 10M-item catalogue), so the hint's failure probability is bounded by the exact binomial tail (hint_eps per row;
  // the kernel's "sigma" rule is a normal approximation, fine for the common shape where n pg ~ 17 and a failed row is
  // cheap, but a 1e-5 tail at n pg < 1), and more tiles are sampled (below) so that such a safe hint is still tight.
  p.hint_eps = p.nI >= TIGHTEN_MIN_ITEMS ? 1.0e-9f : 0.0f;  // per-row probability that the sampled hint is too high
  p.hint_sigma = p.nI >= TIGHTEN_MIN_ITEMS ? 5.0f : 3.0f;
  p.hint_extra = p.nI >= TIGHTEN_MIN_ITEMS ? 2 : 1;
  const bool tighten = p.nI >= TIGHTEN_MIN_ITEMS && !debug_env("DIVREC_DEBUG_NO_TIGHTEN");
  if (p.use_hint && p.n_tiles >= HINT_MIN_TILES) {
    // large catalogues: the number of sample tiles grows with the catalogue (1.3 % of the pass), so that the hint -- the
    // r-th largest of 64 group maxima with a 1e-9 exact binomial tail -- sits where ~1200 groups per row pass it whatever
    // the size (1.25M-item shard with 16 tiles: ~3000 groups and 1395 tightenings per CTA, filter 52.1 ms; with the hint
    // at ~1000 groups: 67 tightenings, 39.6 ms -- each tightening stalls the CTA's accumulator ring for ~25k cycles)
    p.sample_tpg = 1;
    if (tighten) {
      p.sample_tpg = (int)((p.nI + SAMPLE_ITEMS_PER_TPG - 1) / SAMPLE_ITEMS_PER_TPG);
      if (p.sample_tpg > 64) p.sample_tpg = 64;
      while (p.sample_tpg > 1 && N_SAMPLE * p.sample_tpg > (p.n_tiles - 1) / 2) --p.sample_tpg;
    }
    p.n_sample = N_SAMPLE * p.sample_tpg;
    p.sample_stride = (p.n_tiles - 1) / p.n_sample;  // sample tiles are full tiles (never the ragged last one)
  } else {
    p.n_sample = 0;
    p.sample_stride = 1;
  }
  const size_t smem = score_filter_smem_bytes(p.kblocks);
  int grid = p.n_ublocks < sm_count ? p.n_ublocks : sm_count;
  if (const char* g = debug_env("DIVREC_DEBUG_FILTER_GRID")) { const int v = atoi(g); if (v > 0 && v < grid) grid = v; }  // measurement aid
  cudaError_t e;
  if (tighten) {
    e = cudaFuncSetAttribute(score_filter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    score_filter_kernel<true><<<grid, NUM_THREADS, smem, stream>>>(tb, p);
  } else {
    e = cudaFuncSetAttribute(score_filter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    score_filter_kernel<false><<<grid, NUM_THREADS, smem, stream>>>(tb, p);
  }
  count_launch();
  return cudaGetLastError();
}

}  // namespace divrec
