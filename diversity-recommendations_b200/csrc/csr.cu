// csr.cu -- seen-item mask builder: [nnz, 2] (user, item) interaction pairs -> CSR with rows sorted ascending and unique.
//
// Replaces the part of RankingDataset.__iter__ that builds each user's `frozen` item set by scanning the whole
// interaction table (divrec/datasets/base_datasets.py:139-149: `interactions[:, 0] == user_id` per user, O(U nnz),
// 78 % of the reference's evaluation time at the MovieLens-20M shape), and the same scan inside the relevance metrics
// (divrec/metrics/precision_at_k.py:18 etc.).  Sort-free bucketing: count per user, exclusive scan, scatter into the
// user's bucket, then each bucket is sorted and de-duplicated on its own (the reference's set semantics), a second
// scan of the unique counts gives the final row pointers, and the buckets are compacted.  Pairs whose user is
// outside [0, nU) or whose item is outside [0, nI) are dropped (the reference filters `frozen` per scored user).
#include "common.cuh"
#include "kernels.h"

namespace divrec {

namespace {

constexpr int SORT_T = 256;          // threads of the per-row sort CTA
constexpr int SORT_SMEM = 8192;      // elements a row may have to be sorted in shared memory

__global__ void __launch_bounds__(256)
csr_count_kernel(const int64_t* __restrict__ pairs, int64_t nnz, int64_t nU, int64_t nI, int32_t* __restrict__ cnt) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
    const longlong2 pr = __ldg(reinterpret_cast<const longlong2*>(pairs) + i);
    if ((uint64_t)pr.x < (uint64_t)nU && (uint64_t)pr.y < (uint64_t)nI) atomicAdd(cnt + pr.x, 1);
  }
}

// out[0] = 0, out[i + 1] = sum_{j <= i} in[j]; one CTA walks the array in chunks of 1024 (n is a user count: <= a few
// million, ~1 us per 1024 elements)
__global__ void __launch_bounds__(1024)
csr_scan_kernel(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
  __shared__ long long warp_sum_s[32];
  __shared__ long long carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { carry_s = 0; out[0] = 0; }
  __syncthreads();
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t i = base + tid;
    long long v = i < n ? (long long)in[i] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) warp_sum_s[warp] = v;
    __syncthreads();
    if (warp == 0) {
      long long w = warp_sum_s[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      warp_sum_s[lane] = w;
    }
    __syncthreads();
    const long long incl = v + (warp > 0 ? warp_sum_s[warp - 1] : 0) + carry_s;
    if (i < n) out[i + 1] = incl;
    __syncthreads();
    if (tid == 1023) carry_s = incl;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
csr_fill_kernel(const int64_t* __restrict__ pairs, int64_t nnz, int64_t nU, int64_t nI, const int64_t* __restrict__ rowptr,
                int32_t* __restrict__ cursor, int32_t* __restrict__ raw) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
    const longlong2 pr = __ldg(reinterpret_cast<const longlong2*>(pairs) + i);
    if ((uint64_t)pr.x < (uint64_t)nU && (uint64_t)pr.y < (uint64_t)nI) {
      const int32_t pos = atomicAdd(cursor + pr.x, 1);
      raw[rowptr[pr.x] + pos] = (int32_t)pr.y;
    }
  }
}

// ordered block compaction step: thread `tid` contributes `mine` elements; returns its exclusive offset and adds the
// chunk's total to *base_s (all threads of the CTA call this)
__device__ __forceinline__ int block_offset(int mine, int* warp_tot, int* base_s, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  int before = *base_s;
  for (int w = 0; w < warp; ++w) before += warp_tot[w];
  __syncthreads();
  if (tid == SORT_T - 1) *base_s = before + incl;
  __syncthreads();
  return before + incl - mine;
}

// One CTA per row: the row's bucket `raw[s0, s0 + n)` -> sorted ascending and de-duplicated in `out[s0, ...)`,
// ucnt[row] = number of distinct items.  Two ways, both in the CTA's 32 KB of shared memory:
//   bitonic   rows of at most SORT_SMEM items: sort, then drop equal neighbours;
//   bitmap    any length: the item-id space is walked in windows of 32 * SORT_SMEM ids; the row's items falling into the
//             window set bits, the set bits are read out in order (sorted and unique by construction).
// The cheaper one for the row's length and the catalogue size is taken.
__global__ void __launch_bounds__(SORT_T)
csr_sort_rows_kernel(const int64_t* __restrict__ rowptr, int64_t nU, int64_t nI, const int32_t* __restrict__ raw,
                     int32_t* __restrict__ out, int32_t* __restrict__ ucnt) {
  extern __shared__ int32_t srt[];
  __shared__ int warp_tot[SORT_T / 32];
  __shared__ int base_s;
  const int tid = threadIdx.x;
  uint32_t* bits = reinterpret_cast<uint32_t*>(srt);
  const int64_t window = (int64_t)32 * SORT_SMEM;
  const int64_t n_win = (nI + window - 1) / window;
  for (int64_t row = blockIdx.x; row < nU; row += gridDim.x) {
    const int64_t s0 = rowptr[row];
    const int n = (int)(rowptr[row + 1] - s0);
    if (n <= 1) {
      if (tid == 0) { ucnt[row] = n; if (n == 1) out[s0] = raw[s0]; }
      continue;
    }
    const int32_t* seg = raw + s0;
    int32_t* dst = out + s0;
    int np2 = 2, lg = 1;
    while (np2 < n) { np2 <<= 1; ++lg; }
    // block-wide steps: bitonic = lg (lg + 1) / 2 stages of np2 / (2 SORT_T) iterations; bitmap = per window, zero + read-out
    // of the window's words and one pass over the row
    const int64_t words_win = (((nI < window ? nI : window) + 31) / 32 + SORT_T - 1) / SORT_T;
    const int64_t cost_bitonic = (int64_t)(lg * (lg + 1) / 2) * ((np2 / 2 + SORT_T - 1) / SORT_T + 1);
    const int64_t cost_bitmap = n_win * (words_win * 5 + (n + SORT_T - 1) / SORT_T + 2);
    const bool bitonic = n <= SORT_SMEM && cost_bitonic <= cost_bitmap;
    __syncthreads();
    if (tid == 0) base_s = 0;
    if (bitonic) {
      for (int i = tid; i < np2; i += SORT_T) srt[i] = i < n ? seg[i] : 0x7fffffff;
      __syncthreads();
      for (int size = 2; size <= np2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
          for (int t = tid; t < np2 / 2; t += SORT_T) {
            const int i = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
            const int j = i | stride;
            const bool asc = (i & size) == 0;
            const int32_t x = srt[i], y = srt[j];
            if (asc ? (x > y) : (x < y)) { srt[i] = y; srt[j] = x; }
          }
          __syncthreads();
        }
      }
      for (int c0 = 0; c0 < n; c0 += SORT_T) {
        const int i = c0 + tid;
        const bool keep = i < n && (i == 0 || srt[i] != srt[i - 1]);
        const int off = block_offset(keep ? 1 : 0, warp_tot, &base_s, tid);
        if (keep) dst[off] = srt[i];
      }
    } else {
      for (int64_t wnd = 0; wnd < n_win; ++wnd) {
        const int64_t lo = wnd * window;
        const int n_words = (int)(((nI - lo < window ? nI - lo : window) + 31) / 32);
        for (int i = tid; i < n_words; i += SORT_T) bits[i] = 0u;
        __syncthreads();
        for (int i = tid; i < n; i += SORT_T) {
          const int64_t v = (int64_t)seg[i] - lo;
          if (v >= 0 && v < window) atomicOr(&bits[v >> 5], 1u << (v & 31));
        }
        __syncthreads();
        for (int w0 = 0; w0 < n_words; w0 += SORT_T) {
          const int wi = w0 + tid;
          uint32_t word = wi < n_words ? bits[wi] : 0u;
          int off = block_offset(__popc(word), warp_tot, &base_s, tid);
          while (word) {
            const int bit = __ffs(word) - 1;
            word &= word - 1;
            dst[off++] = (int32_t)(lo + (int64_t)wi * 32 + bit);
          }
        }
        __syncthreads();
      }
    }
    __syncthreads();
    if (tid == 0) ucnt[row] = base_s;
  }
}

// one warp per row: the de-duplicated front of the row's bucket -> its final place
__global__ void __launch_bounds__(256)
csr_compact_kernel(const int64_t* __restrict__ rowptr_raw, const int64_t* __restrict__ rowptr, int64_t nU,
                   const int32_t* __restrict__ raw, int32_t* __restrict__ items) {
  const uint32_t lane = lane_id();
  for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < nU; row += (int64_t)gridDim.x * 8) {
    const int64_t src = rowptr_raw[row], dst = rowptr[row];
    const int n = (int)(rowptr[row + 1] - dst);
    for (int i = lane; i < n; i += 32) items[dst + i] = raw[src + i];
  }
}

inline size_t up(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace

size_t build_csr_workspace_bytes(int64_t nnz, int64_t nU) {
  return up((size_t)nU * 4) * 3 + up((size_t)(nU + 1) * 8) + 2 * up((size_t)(nnz > 0 ? nnz : 1) * 4);
}

cudaError_t launch_build_csr(const int64_t* pairs, int64_t nnz, int64_t nU, int64_t nI, int64_t* rowptr, int32_t* items,
                             void* workspace, cudaStream_t stream) {
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  int32_t* cnt = reinterpret_cast<int32_t*>(ws); ws += up((size_t)nU * 4);
  int32_t* cursor = reinterpret_cast<int32_t*>(ws); ws += up((size_t)nU * 4);
  int32_t* ucnt = reinterpret_cast<int32_t*>(ws); ws += up((size_t)nU * 4);
  int64_t* rowptr_raw = reinterpret_cast<int64_t*>(ws); ws += up((size_t)(nU + 1) * 8);
  int32_t* raw = reinterpret_cast<int32_t*>(ws); ws += up((size_t)(nnz > 0 ? nnz : 1) * 4);
  int32_t* sorted = reinterpret_cast<int32_t*>(ws);
  cudaError_t e = cudaMemsetAsync(cnt, 0, up((size_t)nU * 4) * 2, stream);  // cnt + cursor
  if (e != cudaSuccess) return e;
  int64_t g = (nnz + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  if (nnz > 0) { csr_count_kernel<<<(unsigned)g, 256, 0, stream>>>(pairs, nnz, nU, nI, cnt); count_launch(); }
  csr_scan_kernel<<<1, 1024, 0, stream>>>(cnt, nU, rowptr_raw); count_launch();
  if (nnz > 0) { csr_fill_kernel<<<(unsigned)g, 256, 0, stream>>>(pairs, nnz, nU, nI, rowptr_raw, cursor, raw); count_launch(); }
  e = cudaFuncSetAttribute(csr_sort_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_SMEM * 4);
  if (e != cudaSuccess) return e;
  int64_t gs = nU < 148 * 32 ? (nU > 0 ? nU : 1) : 148 * 32;
  csr_sort_rows_kernel<<<(unsigned)gs, SORT_T, SORT_SMEM * 4, stream>>>(rowptr_raw, nU, nI, raw, sorted, ucnt); count_launch();
  csr_scan_kernel<<<1, 1024, 0, stream>>>(ucnt, nU, rowptr); count_launch();
  int64_t gc = (nU + 7) / 8;
  if (gc > 148 * 16) gc = 148 * 16;
  csr_compact_kernel<<<(unsigned)gc, 256, 0, stream>>>(rowptr_raw, rowptr, nU, sorted, items); count_launch();
  return cudaGetLastError();
}

}  // namespace divrec
