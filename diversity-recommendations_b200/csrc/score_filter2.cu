// score_filter2.cu -- the score + filter kernel for the common shape (d <= 128, catalogue below 500k items):
// TWO 128-user blocks per CTA share every item tile.
//
// Why: with the A operand in tensor memory the tensor pipe needs a new 32 KB item tile every ~590 cycles, i.e.
// ~55 B/cycle of shared-memory ingest per SM -- measured to be the limit of a single SM's L2->smem path (the wait
// for B operand data was 19 % of the issuing thread's time and did not change when the grid was shrunk to 74 or 37
// CTAs, so it is not chip-level L2 contention).  Scoring two user blocks against each tile halves the ingest per
// MMA and removes that wait.  Same arithmetic, same candidate-list format and same certification contract as
// score_filter.cu (see there for the algorithm); what changes is the schedule:
//   TMEM      A(block 0) | A(block 1) at the top (2 x K/2 <= 128 columns), three 128-column accumulator stages
//             rotating over (tile, block) pairs;
//   epilogue  set b (8 warps) owns user block b and consumes every tile of it; a user row owns two candidate
//             lists (column halves) of 2 x LIST_CAP groups, stored in the four-list layout as slots (h, 0)+(h, 1);
//   MMA       per tile: 8 k-steps for block 0, 8 for block 1 against the same shared-memory stages, which are
//             released after the second pass.
//
// Threshold ladder (round 2).  The sampled hint has to be safe (4 sigma on ~53 group maxima), so it sits where ~360
// elements per row pass it while ~118 are needed.  The hint phase therefore also leaves four tighter LEVELS (the same
// order statistic at 3, 2, 1, 0 sigma); during the main pass every thread counts, per level, the CLEAN 64-column blocks
// (no seen item: dirty-block bitmap of row_mask_kernel) whose maximum reaches it.  At checkpoints (every ~1/8 of the
// catalogue) the two threads of a row pool their counts and raise the threshold to the highest value -- interpolated
// between levels -- whose projected number of unmasked items above it, at a 2.5 sigma lower confidence bound, still
// covers k + margin.  A threshold only ever rises, and everything at or above the final one was captured at every
// earlier (lower) one, so the select kernel's certification contract is unchanged (tools/sim/ladder_sim.py: captured
// elements per row 363 -> ~185, emitted groups 345 -> ~225, no uncertified row in 12 000).
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace divrec {

namespace {

#ifndef DIVREC_Z_HINT
#define DIVREC_Z_HINT 3.5f   // safety of the sampled hint on the masked rule, standard deviations (measured at cfg2: 4.0 -> 3 uncertified
                             // rows per step, 3.5 -> 6, 3.0 -> 26; the rescue path takes up to 32 at a fixed ~0.07 ms)
#endif
#ifndef DIVREC_Z_CP
#define DIVREC_Z_CP 2.5f     // lower confidence bound of the ladder's checkpoint counts
#endif
constexpr int BM = 128, BN = 128, BK = 64;
constexpr int SLAB_BYTES = BM * BK * 2;
constexpr int TMEM_COLS = 512;
constexpr int ACC_STAGES = 3;
constexpr int NB = 2;              // user blocks per CTA
constexpr int EPI_WARPS = 8;       // per set
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int W_TMA = EPI_WARPS * NB, W_MMA = W_TMA + 1, W_ALLOC = W_TMA + 2;
constexpr int NUM_THREADS = (W_ALLOC + 1) * 32;
constexpr int COLS_PER_THREAD = BN / 2;
constexpr int GROUPS_PER_TILE = COLS_PER_THREAD / GROUP;
constexpr int HS = 2 * N_SAMPLE;

struct Smem2 {
  uint64_t b_full[16];
  uint64_t b_empty[16];
  uint64_t acc_full[4];
  uint64_t acc_empty[4];
  uint64_t a_full;
  uint64_t a_empty;
  uint32_t tmem_base;
  uint32_t pad;
  alignas(16) float sorted[NB][HS][EPI_THREADS];  // hint phase; afterwards rows 0..5 = ladder checkpoint exchange, rows 8..9 = ladder levels (float4 per row)
  int dropped[NB][EPI_THREADS];   // sample groups a thread left out of the hint because they contain seen items
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
__device__ __forceinline__ int sched_tile(int s, int n_sample, int stride) {
  return s < n_sample ? s * stride : s - n_sample;
}

// Ladder checkpoint (rare path, out of line so that it does not weigh on the hot loop's registers): the highest
// threshold -- a level, or a log-linear interpolation between two levels -- whose projected number of UNMASKED items
// above it still covers k + margin.  cj = clean 64-column blocks (of nc seen so far, both column halves of the row)
// whose maximum reached level j; P(block max >= t) = q  <=>  per-item tail p = -log(1 - q) / 64; the count is taken
// at a 2.5 sigma lower confidence bound of q, plus one sigma of the part of the catalogue still to come.
// Returns -inf when no level qualifies.  Both threads of a row evaluate this on the same numbers.
__device__ __noinline__ float ladder_decide(float n_un, float need, float4 l4, uint32_t s01, uint32_t s23, float nc, float frac) {
  // (a checkpoint used to cost ~9 500 cycles against ~2 000 for a tile -- 11 % of the kernel over the seven of a pass: two
  // dependent global loads of the row's seen count, IEEE divisions, log1pf; the count now comes from shared memory and the
  // arithmetic uses the fast intrinsics -- the result only steers a threshold whose safety is checked by the select kernel)
  const float req = need + __fsqrt_rn(need * (1.0f - frac));
  const float cj[4] = {(float)(s01 & 0xffffu), (float)(s01 >> 16), (float)(s23 & 0xffffu), (float)(s23 >> 16)};
  const float lj[4] = {l4.x, l4.y, l4.z, l4.w};
  const float inv_nc = __fdividef(1.0f, nc);
  float nlo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float qh = cj[j] * inv_nc;
    const float se = __fsqrt_rn(fmaxf(qh * (1.0f - qh), 1e-9f) * inv_nc);
    const float qlo = fminf(fmaxf(qh - DIVREC_Z_CP * se, 1e-6f), 0.999f);
    nlo[j] = -__logf(1.0f - qlo) * (1.0f / 64.0f) * n_un;  // items above level j, lower confidence bound
  }
  float cand = -INFINITY;
#pragma unroll
  for (int j = 3; j >= 0; --j) {
    if (cand == -INFINITY && lj[j] < 3.0e38f && nlo[j] >= req) {
      cand = lj[j];
      if (j < 3 && lj[j + 1] < 3.0e38f && lj[j + 1] > lj[j] && nlo[j + 1] < req && nlo[j + 1] > 0.0f) {
        const float wgt = __fdividef(__logf(nlo[j]) - __logf(req), fmaxf(__logf(nlo[j]) - __logf(nlo[j + 1]), 1e-6f));
        cand += (lj[j + 1] - lj[j]) * fminf(fmaxf(wgt, 0.0f), 1.0f);
      }
    }
  }
  return cand;
}

// The MMA-issuing thread's loop.  It shares its scheduler partition with a quarter of the epilogue warps and gets an issue
// slot every fifth cycle or so: at ~240 instructions per item tile (runtime k-block loop, 64-bit descriptor arithmetic,
// statistics) the THREAD, not the tensor pipe, set the kernel's pace -- a trace of its clock showed 1 300 cycles per tile
// with every operand long resident, against 1 024 cycles of MMA work (tools/filter_trace.py).  So: k-blocks at compile
// time, descriptors advanced in their low word only (the address field cannot carry out), all of a tile's stages awaited
// before the first MMA, nothing else in the loop.
template <int KB>
__device__ __forceinline__ void mma_issue_loop(Smem2* sm, const FilterParams& p, uint32_t tmem_base, uint32_t b_ring_addr,
                                               int n_sched, int n_dblocks) {
  const uint32_t idesc = umma_idesc_f16(p.fmt, BM, BN);
  const uint64_t bdesc0 = umma_desc_sw128(b_ring_addr);
  const uint32_t d_hi = (uint32_t)(bdesc0 >> 32), d_lo0 = (uint32_t)bdesc0;
  const uint32_t bar_b_full = smem_u32(&sm->b_full[0]), bar_b_empty = smem_u32(&sm->b_empty[0]);
  const uint32_t bar_acc_full = smem_u32(&sm->acc_full[0]), bar_acc_empty = smem_u32(&sm->acc_empty[0]);
  const uint32_t bar_a_full = smem_u32(&sm->a_full), bar_a_empty = smem_u32(&sm->a_empty);
  const uint32_t b_stages = (uint32_t)p.b_stages;
  constexpr uint32_t A_COLS = KB * (BK / 2);
  const uint32_t a0 = tmem_base + TMEM_COLS - A_COLS, a1 = tmem_base + TMEM_COLS - 2 * A_COLS;
  uint32_t stage = 0, phase = 0, a_phase = 0, acc = 0, acc_phase = 0;
  for (int db = blockIdx.x; db < n_dblocks; db += gridDim.x) {
    mbar_wait(bar_a_full, a_phase);
    a_phase ^= 1;
    tc_fence_after();
    for (int s = 0; s < n_sched; ++s) {
      uint32_t st[KB];
#ifdef DIVREC_DIAG_TRACE2
      if (p.stats != nullptr && blockIdx.x == 0) p.stats[148 * 8 + ((db - (int)blockIdx.x) / (int)gridDim.x) * 256 + s] = clock64();
#endif
      // ---- user block 0: the tile's stages arrive together, all awaited before the first MMA ----
      mbar_wait(bar_acc_empty + 8 * acc, acc_phase ^ 1);
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        st[kb] = stage;
        mbar_wait(bar_b_full + 8 * stage, phase);
        if (++stage == b_stages) { stage = 0; phase ^= 1; }
      }
      tc_fence_after();
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        const uint32_t lo = d_lo0 + st[kb] * (uint32_t)(SLAB_BYTES >> 4);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
          umma_f16_ts(tmem_base + acc * BN, a0 + (uint32_t)(kb * (BK / 2) + 8 * k4), ((uint64_t)d_hi << 32) | (uint64_t)(lo + 2u * k4),
                      idesc, (uint32_t)((kb | k4) != 0));
      }
      umma_commit(bar_acc_full + 8 * acc);
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      // ---- user block 1 against the same stages, which are released behind it ----
      mbar_wait(bar_acc_empty + 8 * acc, acc_phase ^ 1);
      tc_fence_after();
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        const uint32_t lo = d_lo0 + st[kb] * (uint32_t)(SLAB_BYTES >> 4);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
          umma_f16_ts(tmem_base + acc * BN, a1 + (uint32_t)(kb * (BK / 2) + 8 * k4), ((uint64_t)d_hi << 32) | (uint64_t)(lo + 2u * k4),
                      idesc, (uint32_t)((kb | k4) != 0));
        umma_commit(bar_b_empty + 8 * st[kb]);
      }
      umma_commit(bar_acc_full + 8 * acc);
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
    umma_commit(bar_a_empty);
  }
}

// Seen-item patterns of every row, one pass over the CSR (one warp per row, words assembled in shared memory):
//   words [0, HS):           out[(2 s + h) * stride_rows + row] has bit c set when item (sample tile s, column 64 h + c)
//                            is in the row's seen list (the hint phase leaves such sample groups out of its statistic);
//   words [HS, HS + 2 W):    out[(HS + h * W + w) * stride_rows + row] has bit b set when the 64-column block
//                            (tile 64 w + b, column half h) holds a seen item ("dirty": the ladder counts clean blocks only).
// 44 MB at cfg2 (W = 4).
__global__ void __launch_bounds__(256)
row_mask_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ items, int64_t nU, int64_t stride_rows,
                int32_t nI, int64_t item_off, int n_sample, int sample_stride, int W, unsigned long long* __restrict__ out) {
  extern __shared__ unsigned long long rm_smem[];
  const int w = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int nw = HS + 2 * W;
  unsigned long long* m = rm_smem + (size_t)w * nw;
  // bits are set through the 32-bit halves of the words: shared memory has a native 32-bit atomic OR, a 64-bit one is a
  // compare-and-swap loop (ncu: 61 % of this kernel's stalls, 0.2 ms at cfg2 -- as long as a sixth of the filter itself)
  uint32_t* m32 = reinterpret_cast<uint32_t*>(m);
  for (int64_t row = (int64_t)blockIdx.x * 8 + w; row < nU; row += (int64_t)gridDim.x * 8) {
    for (int i = lane; i < nw; i += 32) m[i] = 0ull;
    __syncwarp();
    const int64_t s0 = rowptr[row], s1 = rowptr[row + 1];
    for (int64_t i = s0 + lane; i < s1; i += 32) {
      const int64_t id = (int64_t)__ldg(items + i) - item_off;
      if (id >= 0 && id < nI) {
        const int tile = (int)(id >> 7);
        const int col = (int)(id & 127);
        // dirty-block word (HS + half * W + tile / 64), bit tile % 64
        atomicOr(&m32[2 * (HS + (col >> 6) * W + (tile >> 6)) + ((tile >> 5) & 1)], 1u << (tile & 31));
        const int s = tile / sample_stride;
        if (s < n_sample && s * sample_stride == tile)   // sample-tile word (2 s + half), bit col % 64
          atomicOr(&m32[2 * (2 * s + (col >> 6)) + ((col >> 5) & 1)], 1u << (col & 31));
      }
    }
    __syncwarp();
    for (int i = lane; i < nw; i += 32) out[(size_t)i * stride_rows + row] = m[i];
    __syncwarp();
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
score_filter2_kernel(const __grid_constant__ CUtensorMap tmap_b, const FilterParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem_base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* b_ring = smem_base;
  Smem2* sm = reinterpret_cast<Smem2*>(b_ring + (size_t)p.b_stages * SLAB_BYTES);

  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int b_stages = p.b_stages;

  if (threadIdx.x == 0) {
    for (int i = 0; i < b_stages; ++i) { mbar_init(smem_u32(&sm->b_full[i]), 1); mbar_init(smem_u32(&sm->b_empty[i]), 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(smem_u32(&sm->acc_full[i]), 1); mbar_init(smem_u32(&sm->acc_empty[i]), EPI_WARPS); }
    mbar_init(smem_u32(&sm->a_full), EPI_WARPS * NB);  // every epilogue warp writes part of the two A operands
    mbar_init(smem_u32(&sm->a_empty), 1);
    fence_mbar_init();
  }
  if (warp == W_TMA && lane == 0) tma_prefetch_desc(&tmap_b);
  if (warp == W_ALLOC) { tmem_alloc(smem_u32(&sm->tmem_base), TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&sm->tmem_base);

  const int kblocks = p.kblocks;  // 1 or 2
  const int n_sample = p.n_sample;
  const int n_sched = n_sample + p.n_tiles;
  const int n_dblocks = (p.n_ublocks + NB - 1) / NB;
  const uint32_t a_cols = (uint32_t)kblocks * (BK / 2);  // TMEM columns of one A operand

  if (warp == W_TMA) {
    // ================= TMA producer: one pass over the schedule per PAIR of user blocks =================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint64_t pol_keep = l2_policy_evict_last();
      for (int db = blockIdx.x; db < n_dblocks; db += gridDim.x) {
        for (int s = 0; s < n_sched; ++s) {
          const int t = sched_tile(s, n_sample, p.sample_stride);
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(smem_u32(&sm->b_empty[stage]), phase ^ 1);
            mbar_arrive_expect_tx(smem_u32(&sm->b_full[stage]), SLAB_BYTES);
            tma_load_2d_hint(smem_u32(b_ring + (size_t)stage * SLAB_BYTES), &tmap_b, smem_u32(&sm->b_full[stage]), kb * BK,
                             t * BN, pol_keep);
            if (++stage == (uint32_t)b_stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == W_MMA) {
    // ================= MMA issuer =================
    if (lane == 0) {
      if (kblocks == 2) mma_issue_loop<2>(sm, p, tmem_base, smem_u32(b_ring), n_sched, n_dblocks);
      else mma_issue_loop<1>(sm, p, tmem_base, smem_u32(b_ring), n_sched, n_dblocks);
    }
  } else if (warp < EPI_WARPS * NB) {
    // ================= epilogue: set `bset` owns user block `bset` of the pair =================
    const int q = warp & 3;
    const int h = (warp >> 2) & 1;
    const int bset = warp >> 3;
    const int row_local = q * 32 + lane;
    const int et = (warp & 7) * 32 + lane;
    uint32_t base_seq = 0;  // accumulator uses before this pair: NB per schedule position
    uint32_t a_phase = 0;
    for (int db = blockIdx.x; db < n_dblocks; db += gridDim.x, base_seq += (uint32_t)(NB * n_sched)) {
      const int ub = db * NB + bset;
      const int64_t row = (int64_t)ub * BM + row_local;
      const bool valid_row = ub < p.n_ublocks && row < p.nU;
      float thr = valid_row ? -INFINITY : INFINITY;

      // ---- phase 0: this set's A operand -> TMEM ----
      {
        mbar_wait(smem_u32(&sm->a_empty), a_phase ^ 1);
        a_phase ^= 1;
        tc_fence_after();
        const int words = kblocks * (BK / 4);
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.a16) + (size_t)row * (kblocks * (BK / 2)) + (size_t)h * words;
        const uint32_t tdst = tmem_base + ((uint32_t)(q * 32) << 16) + TMEM_COLS - (uint32_t)(bset + 1) * a_cols + (uint32_t)h * words;
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

      // ---- phase 1: threshold hint (each set samples for its own block) ----
      float4 lv = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);  // ladder levels (tighter than the hint); +inf = unused
      if (n_sample > 0) {
        // A sample group (32 columns) that contains one of the row's seen items is left out of the statistic: the
        // remaining maxima are maxima over UNMASKED items only, so the hint can aim at k unmasked items instead of
        // provisioning for every seen item outranking them (row_mask_kernel; null = no mask / not built).
        const unsigned long long* hm = (p.hint_mask != nullptr && valid_row) ? p.hint_mask + row : nullptr;
        int dropped = 0;
        float hs[HS];
#pragma unroll
        for (int s = 0; s < N_SAMPLE; ++s) {
          const unsigned long long mk = hm != nullptr ? __ldg(hm + (size_t)(2 * s + h) * (size_t)p.hint_mask_stride) : 0ull;
          const uint32_t seq = base_seq + (uint32_t)(NB * s + bset), acc = seq % ACC_STAGES, acc_phase = (seq / ACC_STAGES) & 1;
          mbar_wait(smem_u32(&sm->acc_full[acc]), acc_phase);
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + h * COLS_PER_THREAD;
          uint32_t v0[32];
          tmem_ld_32x32b_x32(taddr, v0);
          tmem_wait_ld();
          const bool d0 = (uint32_t)mk != 0u, d1 = (uint32_t)(mk >> 32) != 0u;
          hs[2 * s] = d0 ? -INFINITY : max32(v0);
          tmem_ld_32x32b_x32(taddr + 32, v0);
          tmem_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&sm->acc_empty[acc]));
          hs[2 * s + 1] = d1 ? -INFINITY : max32(v0);
          dropped += (d0 ? 1 : 0) + (d1 ? 1 : 0);
        }
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
        for (int i = 0; i < HS; ++i) sm->sorted[bset][i][et] = hs[i];
        sm->dropped[bset][et] = dropped;
        if (bset == 0) asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
        else asm volatile("bar.sync 2, %0;" ::"n"(EPI_THREADS) : "memory");
        if (valid_row) {
          int64_t seen_u = 0;
          if (p.seen_rowptr != nullptr) seen_u = p.seen_rowptr[row + 1] - p.seen_rowptr[row];
          const int n_eff = 2 * HS - (sm->dropped[bset][row_local] + sm->dropped[bset][row_local + 128]);
          // groups left: maxima over unmasked items only -> aim at k (+ max(16, k/4): the 2 delta certification band)
          // unmasked items above the hint; too few groups left (heavy users): the worst-case rule on all 64 entries
          // (the dropped ones read -inf, which can only lower the hint)
          const bool masked_rule = hm != nullptr && n_eff >= HS && (int64_t)p.nI - seen_u > 0;
          const float f0 = masked_rule ? ((float)p.k + (float)max(16, p.k >> 2)) / (float)((int64_t)p.nI - seen_u)
                                       : ((float)p.k + (float)seen_u + 8.0f) / (float)p.nI;
          const int n_groups = masked_rule ? n_eff : 2 * HS;
          int r = 2 * HS + 1;
          if (f0 < 0.25f) {
            const float pg = 1.0f - __expf(32.0f * log1pf(-f0));
            const float n = (float)n_groups;
            // 4 sigma on the masked rule: an uncertified row costs a rescue launch (~0.17 ms of latency per call even for a
            // handful of rows), which at 3 sigma (24 rows per 138k) hits every call of the chunked host pipeline
            r = (int)ceilf(n * pg + (masked_rule ? DIVREC_Z_HINT : 3.0f) * sqrtf(n * pg * (1.0f - pg))) + 1;
          }
          // ladder levels: the same order statistic at 3, 2, 1, 0 sigma for k + margin UNMASKED items (ranks r_lv[j] <= r);
          // which of them is safe is decided later from live counts, so their own noise costs tightness, never exactness
          int r_lv[4] = {0, 0, 0, 0};
          {
            const float f1 = ((float)p.k + (float)max(16, p.k >> 2)) / fmaxf(1.0f, (float)((int64_t)p.nI - seen_u));
            if (f1 < 0.25f) {
              const float pg = 1.0f - __expf(32.0f * log1pf(-f1));
              const float n = (float)n_groups, sd = sqrtf(n * pg * (1.0f - pg));
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                int rj = (int)ceilf(n * pg + (float)(3 - j) * sd) + 1;
                rj = max(1, rj);
                r_lv[j] = (rj <= n_groups - 4 && (r > n_groups - 4 || rj <= r)) ? rj : 0;
              }
            }
          }
          const int r_walk = r <= n_groups - 4 ? r : max(max(r_lv[0], r_lv[1]), max(r_lv[2], r_lv[3]));
          {
            const int ea = row_local, eb = row_local + 128;
            int i = 0, j = 0;
            float pick = INFINITY;
            for (int c = 0; c < r_walk; ++c) {
              const float a = i < HS ? sm->sorted[bset][i][ea] : -INFINITY;
              const float b = j < HS ? sm->sorted[bset][j][eb] : -INFINITY;
              if (a >= b) { pick = a; ++i; } else { pick = b; ++j; }
              if (c + 1 == r_lv[0]) lv.x = pick;
              if (c + 1 == r_lv[1]) lv.y = pick;
              if (c + 1 == r_lv[2]) lv.z = pick;
              if (c + 1 == r_lv[3]) lv.w = pick;
            }
            if (r <= n_groups - 4) thr = pick;
          }
        }
        if (bset == 0) asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
        else asm volatile("bar.sync 2, %0;" ::"n"(EPI_THREADS) : "memory");
      }
      // the row's ladder levels live in shared memory (rows 8, 9 of this set's `sorted` area, dead until the next user
      // block's hint phase; both threads of a row write the same values)
      float4* const lv_s = reinterpret_cast<float4*>(&sm->sorted[bset][8][0]);
      lv_s[row_local] = lv;
      // the row's unmasked item count for the ladder's projections (row 10 of the same area; both threads of a row write it)
      float* const nun_s = &sm->sorted[bset][10][0];
      {
        int64_t seen_u = 0;
        if (valid_row && p.seen_rowptr != nullptr) seen_u = p.seen_rowptr[row + 1] - p.seen_rowptr[row];
        nun_s[row_local] = fmaxf(1.0f, (float)((int64_t)p.nI - seen_u));
      }

      // ---- phase 2: group filter over every tile of this set's block ----
      // The epilogue is instruction-bound (16 warps share the SM's issue slots with nothing to hide behind), so an append
      // is one predicate, ONE 256-bit store (the group's scores carry its index in their low mantissa bytes) and a bump
      // of the thread's running record pointer.  The group count is the pointer's distance.
      const uint32_t cap2 = 2u * (uint32_t)p.cap;                 // slots (h,0)+(h,1): one list of 2 * cap records
      const uint32_t cap_lim = cap2 - GROUPS_PER_TILE;
      float* const rec0 = p.lists + ((size_t)ub * (BM * 4) * (size_t)p.cap + (size_t)((row_local * 4 + h * 2) * p.cap)) * REC_FLOATS;
      float* rec = rec0;
      bool overflow = false;
      // ladder state: per level, clean 64-column blocks of this thread whose maximum reached it (16-bit fields), and the
      // clean blocks seen; the dirty-block bitmap is read 32 tiles at a time
      uint32_t c01 = 0, c23 = 0, ncl = 0, dirty = 0;
      const uint32_t* const dirty_w = (p.hint_mask != nullptr && valid_row)
          ? reinterpret_cast<const uint32_t*>(p.hint_mask + (size_t)(HS + h * p.mask_words) * (size_t)p.hint_mask_stride + row)
          : nullptr;
      const bool ladder = p.ladder_stride > 0;
      const float inv_tiles = 1.0f / (float)p.n_tiles;
      int next_cp = ladder ? p.ladder_stride : 0x7fffffff;
      uint32_t* const exch = reinterpret_cast<uint32_t*>(&sm->sorted[bset][0][0]);  // checkpoint exchange: rows 0..5 of `sorted`
      for (int t = 0; t < p.n_tiles; ++t) {
        if (ladder && (t & 31) == 0)
          dirty = dirty_w != nullptr ? __ldg(dirty_w + (size_t)(t >> 6) * (size_t)p.hint_mask_stride * 2 + ((t >> 5) & 1)) : 0u;
        // this set's accumulator sequence advances by NB per tile: running stage / phase instead of a division per tile
        const uint32_t seq = base_seq + (uint32_t)(NB * (n_sample + t) + bset), acc = seq % ACC_STAGES, acc_phase = (seq / ACC_STAGES) & 1;
        mbar_wait(smem_u32(&sm->acc_full[acc]), acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + h * COLS_PER_THREAD;
        uint32_t v0[32], v1[32];
        if (p.debug_mode == 3) {  // measurement aid: half of the tensor-memory read traffic, no filtering
          tmem_ld_32x32b_x32(taddr, v0);
          tmem_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&sm->acc_empty[acc]));
          if ((v0[0] ^ v0[13] ^ v0[31]) == 0x7fc12345u) thr = 0.f;  // keep the load alive
          continue;
        }
        tmem_ld_32x32b_x32(taddr, v0);
        tmem_ld_32x32b_x32(taddr + 32, v1);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sm->acc_empty[acc]));

        if (p.debug_mode == 1) continue;
        // records appended so far, from the low words of the pointers (a list is far smaller than 4 GB)
        if ((((uint32_t)reinterpret_cast<uintptr_t>(rec) - (uint32_t)reinterpret_cast<uintptr_t>(rec0)) >> 5) > cap_lim) { thr = INFINITY; overflow = true; }  // the row goes to the rescue kernel
        // index of this thread's first group of the tile (< 2^16: this kernel runs below TIGHTEN_MIN_ITEMS items); a record
        // carries it in the low mantissa bytes of its first two scores: (v | 0xff) ^ x with x = 0xff ^ byte
        const uint32_t gi0 = (uint32_t)(t * (BN / GROUP) + h * GROUPS_PER_TILE);
        const uint32_t xlo = 0xffu ^ (gi0 & 0xffu), xhi = 0xffu ^ (gi0 >> 8);
        float mt0 = -INFINITY, mt1 = -INFINITY, mt2 = -INFINITY;  // maximum of the 64-column block
#pragma unroll
        for (int g = 0; g < GROUPS_PER_TILE; ++g) {
          uint32_t* v = g < 4 ? &v0[8 * g] : &v1[8 * (g - 4)];
          const float m = max8(v);
#ifndef DIVREC_DIAG_NOEMBED
          v[0] = (v[0] | 0xffu) ^ (xlo ^ (uint32_t)g);  // gi0 is a multiple of 8: its low three bits take g
          v[1] = (v[1] | 0xffu) ^ xhi;
#endif
#if defined(DIVREC_DIAG_NOSTORE)
          asm volatile(
              "{\n"
              ".reg .pred p;\n"
              "setp.ge.f32 p, %1, %2;\n"
              "@p add.u64 %0, %0, 32;\n"
              "}\n"
              : "+l"(rec)
              : "f"(m), "f"(thr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
              : "memory");
#else
          asm volatile(
              "{\n"
              ".reg .pred p;\n"
              "setp.ge.f32 p, %1, %2;\n"
              "@p st.global.v8.b32 [%0], {%3, %4, %5, %6, %7, %8, %9, %10};\n"  // the whole append: one 256-bit store (STG.E.ENL2.256)
              "@p add.u64 %0, %0, 32;\n"
              "}\n"
              : "+l"(rec)
              : "f"(m), "f"(thr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
              : "memory");
#endif
          if (g % 3 == 0) mt0 = fmaxf(mt0, m); else if (g % 3 == 1) mt1 = fmaxf(mt1, m); else mt2 = fmaxf(mt2, m);
        }
#ifdef DIVREC_EXP_LADDER_EVEN
        if (ladder && ((t & 1) == 0 || t + 1 == next_cp)) {
#else
        if (ladder) {
#endif
          // a dirty block counts for nothing: its maximum is replaced by -inf, so the level compares need no second condition
          const bool clean = ((dirty >> (t & 31)) & 1u) == 0u;
          const float mt = clean ? max3(mt0, mt1, mt2) : -INFINITY;
          const float4 l4 = lv_s[row_local];  // (an explicit ld.shared + running stage counters measured 5 % slower)
          ncl += clean ? 1u : 0u;
          c01 += (mt >= l4.x) ? 1u : 0u;
          c01 += (mt >= l4.y) ? 0x10000u : 0u;
          c23 += (mt >= l4.z) ? 1u : 0u;
          c23 += (mt >= l4.w) ? 0x10000u : 0u;
          if (t + 1 == next_cp) {
            next_cp += p.ladder_stride;
            if (t + 1 < p.n_tiles) {
              // ---- checkpoint: pool the row's two halves, raise the threshold where the projected count allows ----
              const uint32_t slot = (uint32_t)(((t + 1) / p.ladder_stride) & 1) * 3u;
              exch[(slot + 0) * EPI_THREADS + et] = c01;
              exch[(slot + 1) * EPI_THREADS + et] = c23;
              exch[(slot + 2) * EPI_THREADS + et] = ncl;
              if (bset == 0) asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
              else asm volatile("bar.sync 2, %0;" ::"n"(EPI_THREADS) : "memory");
              const int pe = et ^ 128;  // the row's other column half
              const uint32_t s01 = c01 + exch[(slot + 0) * EPI_THREADS + pe], s23 = c23 + exch[(slot + 1) * EPI_THREADS + pe];
              const float nc = (float)(ncl + exch[(slot + 2) * EPI_THREADS + pe]);
              if (valid_row && nc >= 16.0f)
                thr = fmaxf(thr, ladder_decide(nun_s[row_local], (float)p.k + (float)max(16, p.k >> 2), l4, s01, s23, nc, (float)(t + 1) * inv_tiles));
            }
          }
        }
      }
      const uint32_t cnt = overflow ? cap2 + 1u : (((uint32_t)reinterpret_cast<uintptr_t>(rec) - (uint32_t)reinterpret_cast<uintptr_t>(rec0)) >> 5);
      if (valid_row) {
        p.list_count[row * 4 + h * 2] = (int32_t)cnt;  // 2 * cap + 1 = overflowed
        p.list_count[row * 4 + h * 2 + 1] = 0;
        p.thr_out[row * 4 + h * 2] = thr;
        p.thr_out[row * 4 + h * 2 + 1] = thr;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_ALLOC) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

}  // namespace

bool make_tmap_b(CUtensorMap* m, const void* base, int64_t rows, int K, int fmt);  // score_filter.cu

// sample schedule of launch_score_filter2 for a catalogue of nI items (0 tiles: no hint phase)
static void filter2_sample_schedule(int32_t nI, int* n_sample, int* stride) {
  const int n_tiles = (nI + BN - 1) / BN;
  if (n_tiles >= HINT_MIN_TILES) { *n_sample = N_SAMPLE; *stride = (n_tiles - 1) / N_SAMPLE; }
  else { *n_sample = 0; *stride = 1; }
}

static int mask_words_for(int32_t nI) { return ((nI + BN - 1) / BN + 63) / 64; }

size_t hint_mask_bytes(int64_t rows_padded, int32_t nI) {
  return (size_t)rows_padded * (size_t)(HS + 2 * mask_words_for(nI)) * sizeof(unsigned long long);
}

cudaError_t launch_hint_mask(const int64_t* seen_rowptr, const int32_t* seen_items, int64_t nU, int64_t rows_padded, int32_t nI,
                             int64_t item_id_offset, unsigned long long* out, cudaStream_t stream) {
  int n_sample, stride;
  filter2_sample_schedule(nI, &n_sample, &stride);
  if (n_sample == 0) return cudaSuccess;
  int64_t g = (nU + 7) / 8;
  if (g > 148 * 8) g = 148 * 8;
  const int W = mask_words_for(nI);
  const size_t smem = (size_t)8 * (HS + 2 * W) * sizeof(unsigned long long);
  row_mask_kernel<<<(unsigned)g, 256, smem, stream>>>(seen_rowptr, seen_items, nU, rows_padded, nI, item_id_offset, n_sample, stride, W, out);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_score_filter2(const void* A16, const void* B16, FilterParams p, int sm_count, cudaStream_t stream,
                                 const char** err) {
  const int K = p.kblocks * BK;
  CUtensorMap tb;
  if (!make_tmap_b(&tb, B16, p.nI, K, p.fmt)) {
    *err = "cuTensorMapEncodeTiled failed";
    return cudaErrorInvalidValue;
  }
  p.a16 = A16;
  p.n_ublocks = (int)((p.nU + BM - 1) / BM);
  p.n_tiles = (p.nI + BN - 1) / BN;
  int s = (204 * 1024 - (int)sizeof(Smem2)) / SLAB_BYTES;
  if (s > 16) s = 16;
  p.b_stages = s;
  p.acc_stages = ACC_STAGES;
  p.hint_sigma = 3.0f;
  p.hint_extra = 1;
  filter2_sample_schedule(p.nI, &p.n_sample, &p.sample_stride);
  if (!p.use_hint) { p.n_sample = 0; p.sample_stride = 1; }
  p.mask_words = mask_words_for(p.nI);
  // ladder checkpoints every ~1/8 of the catalogue (off without a hint phase: catalogues under HINT_MIN_TILES tiles)
  p.ladder_stride = p.n_sample > 0 ? ((p.n_tiles + 7) / 8 > 8 ? (p.n_tiles + 7) / 8 : 8) : 0;
#ifdef DIVREC_DIAG_NOLADDER
  p.ladder_stride = 0;
#endif
  const size_t smem = 1024 + (size_t)p.b_stages * SLAB_BYTES + sizeof(Smem2);
  const int n_dblocks = (p.n_ublocks + NB - 1) / NB;
  int grid = n_dblocks < sm_count ? n_dblocks : sm_count;
  if (const char* g = debug_env("DIVREC_DEBUG_FILTER_GRID")) { const int v = atoi(g); if (v > 0 && v < grid) grid = v; }  // measurement aid
  cudaError_t e = cudaFuncSetAttribute(score_filter2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  score_filter2_kernel<<<grid, NUM_THREADS, smem, stream>>>(tb, p);
  count_launch();
  return cudaGetLastError();
}

}  // namespace divrec
