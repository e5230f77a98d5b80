// prep.cu -- operand preparation for the tensor-core filter.
//
// fp32 tables: X is scaled by a power of two so that max|X| lies in [2^13, 2^14) (no fp16 overflow, the
// whole fp16 exponent range below it available) and rounded to ONE fp16 plane.  The tensor-core score
//     A = sum_i fp16(x_i) fp16(y_i)     (fp32 accumulate, products exact)
// then differs from the exact dot product E of MatrixFactorization.forward
// (divrec/models/matrix_factorization.py:26-28) by
//     |A - E| <= (2^-10 + 2^-22) |x||y|  (operand rounding, Cauchy-Schwarz)  +  d 2^-23 |x||y|  (accumulation)
// so delta_u = c |u| max_i |v_i| with c = 1.1 * 2^-10 covers d <= 1024.  The row norms |u| and the maximum
// item norm needed for delta are produced here, in the same pass that converts the table.
// bf16 tables are used as stored (exact operands; delta only covers the accumulation, c = 2^-12).
#include "common.cuh"
#include "kernels.h"

namespace divrec {

namespace {

__global__ void absmax_kernel(const float4* __restrict__ X4, int64_t n4, const float* __restrict__ X, int64_t n,
                              uint32_t* __restrict__ amax_bits) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = X4[i];
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  // tail (n not a multiple of 4, or an unaligned table)
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(X[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f && m < INFINITY) atomicMax(amax_bits, __float_as_uint(m));
}

__global__ void scale_exp_kernel(const uint32_t* __restrict__ amax_bits, int32_t* __restrict__ scale_exp) {
  const float amax = __uint_as_float(*amax_bits);
  int e = 0;
  if (amax > 0.f) e = 13 - ilogbf(amax);
  e = max(-100, min(100, e));
  *scale_exp = e;
}

// one warp per row: scaled fp16 copy (zero padded to d_pad) + fp32 row norm (+ global max norm)
__global__ void __launch_bounds__(256)
convert_f32_kernel(const float* __restrict__ X, int64_t rows, int d, int d_pad, __half* __restrict__ out,
                   const int32_t* __restrict__ scale_exp, float* __restrict__ norms, float* __restrict__ inv_norms,
                   uint32_t* __restrict__ normmax_bits) {
  const uint32_t lane = lane_id();
  const int e = *scale_exp;
  const bool vec = ((d & 3) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  float wmax = 0.f;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    const float* x = X + r * d;
    __half* o = out + r * d_pad;
    float ss = 0.f;
    for (int c = lane * 4; c < d_pad; c += 128) {
      float v[4];
      if (vec && c + 3 < d) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(x + c));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (c + j < d) ? x[c + j] : 0.f;
      }
      __half hv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ss = fmaf(v[j], v[j], ss);
        hv[j] = __float2half_rn(ldexpf(v[j], e));
      }
      *reinterpret_cast<uint2*>(o + c) = *reinterpret_cast<const uint2*>(hv);
    }
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss) * 1.000001f;  // round up: the norm feeds an error BOUND
    if (lane == 0 && norms) norms[r] = nrm;
    if (lane == 0 && inv_norms) inv_norms[r] = 1.0f / sqrtf(ss);
    wmax = fmaxf(wmax, nrm);
  }
  if (normmax_bits && lane == 0 && wmax > 0.f && wmax < INFINITY) atomicMax(normmax_bits, __float_as_uint(wmax));
}

// Item plane for the common shape (d_pad == 128, 16-byte aligned rows, d % 4 == 0): the table-wide exponent is known
// (absmax pass), FOUR rows per warp, eight lanes per row, four independent 128-bit loads in flight per lane.
__global__ void __launch_bounds__(256)
convert_f32_128_kernel(const float* __restrict__ X, int64_t rows, int d, __half* __restrict__ out,
                       const int32_t* __restrict__ scale_exp, float* __restrict__ norms, float* __restrict__ inv_norms,
                       uint32_t* __restrict__ normmax_bits) {
  const uint32_t lane = lane_id();
  const int sub = lane & 7, grp = lane >> 3;
  const int e = *scale_exp;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float wmax = 0.f;
  for (int64_t r0 = warp0 * 4; r0 < rows; r0 += nwarps * 4) {
    const int64_t r = r0 + grp;
    const bool on = r < rows;
    const float* x = X + r * d;
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = (sub + 8 * j) * 4;
      v[j] = (on && c < d) ? __ldg(reinterpret_cast<const float4*>(x + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) { ss = fmaf(v[j].x, v[j].x, ss); ss = fmaf(v[j].y, v[j].y, ss); ss = fmaf(v[j].z, v[j].z, ss); ss = fmaf(v[j].w, v[j].w, ss); }
#pragma unroll
    for (int o2 = 1; o2 < 8; o2 <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o2);
    if (on) {
      __half* o = out + r * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = (sub + 8 * j) * 4;
        __half hv[4] = {__float2half_rn(ldexpf(v[j].x, e)), __float2half_rn(ldexpf(v[j].y, e)), __float2half_rn(ldexpf(v[j].z, e)),
                        __float2half_rn(ldexpf(v[j].w, e))};
        *reinterpret_cast<uint2*>(o + c) = *reinterpret_cast<const uint2*>(hv);
      }
      const float nrm = sqrtf(ss) * 1.000001f;  // round up: the norm feeds an error BOUND
      if (sub == 0) {
        if (norms) norms[r] = nrm;
        if (inv_norms) inv_norms[r] = 1.0f / sqrtf(ss);
      }
      wmax = fmaxf(wmax, nrm);
    }
  }
#pragma unroll
  for (int o2 = 16; o2 > 0; o2 >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o2));
  if (normmax_bits && lane == 0 && wmax > 0.f && wmax < INFINITY) atomicMax(normmax_bits, __float_as_uint(wmax));
}

// User rows: ONE pass, each row scaled by its OWN power of two (max |x| of the row into [2^13, 2^14)): no global absmax pass
// over the table, and small-norm rows keep their full fp16 precision.  row_exp[r] is the exponent used (the select kernel
// and the list-tightening margin read it back), norms[r] the fp32 row norm (rounded up: it feeds an error bound).
__global__ void __launch_bounds__(256)
convert_f32_rows_kernel(const float* __restrict__ X, int64_t rows, int d, int d_pad, __half* __restrict__ out,
                        int32_t* __restrict__ row_exp, float* __restrict__ norms) {
  const uint32_t lane = lane_id();
  const bool vec = ((d & 3) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    const float* x = X + r * d;
    __half* o = out + r * d_pad;
    float ss = 0.f, am = 0.f;
    if (d_pad <= 512) {
      // the whole row in registers (up to 4 float4 per lane): one read of the table
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = (int)lane * 4 + j * 128;
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < d_pad) {
          if (vec && c + 3 < d) v[j] = __ldg(reinterpret_cast<const float4*>(x + c));
          else { v[j].x = c < d ? x[c] : 0.f; v[j].y = c + 1 < d ? x[c + 1] : 0.f; v[j].z = c + 2 < d ? x[c + 2] : 0.f; v[j].w = c + 3 < d ? x[c + 3] : 0.f; }
        }
        ss = fmaf(v[j].x, v[j].x, ss); ss = fmaf(v[j].y, v[j].y, ss); ss = fmaf(v[j].z, v[j].z, ss); ss = fmaf(v[j].w, v[j].w, ss);
        am = fmaxf(am, fmaxf(fmaxf(fabsf(v[j].x), fabsf(v[j].y)), fmaxf(fabsf(v[j].z), fabsf(v[j].w))));
      }
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o2));
      int e = 0;
      if (am > 0.f && am < INFINITY) e = max(-100, min(100, 13 - ilogbf(am)));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = (int)lane * 4 + j * 128;
        if (c < d_pad) {
          __half hv[4] = {__float2half_rn(ldexpf(v[j].x, e)), __float2half_rn(ldexpf(v[j].y, e)), __float2half_rn(ldexpf(v[j].z, e)),
                          __float2half_rn(ldexpf(v[j].w, e))};
          *reinterpret_cast<uint2*>(o + c) = *reinterpret_cast<const uint2*>(hv);
        }
      }
      ss = warp_sum(ss);
      if (lane == 0) { row_exp[r] = e; norms[r] = sqrtf(ss) * 1.000001f; }
    } else {
      for (int c = lane; c < d; c += 32) { const float t = x[c]; ss = fmaf(t, t, ss); am = fmaxf(am, fabsf(t)); }
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o2));
      int e = 0;
      if (am > 0.f && am < INFINITY) e = max(-100, min(100, 13 - ilogbf(am)));
      for (int c = lane; c < d_pad; c += 32) o[c] = __float2half_rn(ldexpf(c < d ? x[c] : 0.f, e));
      ss = warp_sum(ss);
      if (lane == 0) { row_exp[r] = e; norms[r] = sqrtf(ss) * 1.000001f; }
    }
  }
}

// The same for the common shape (d_pad == 128, 16-byte aligned rows, d % 4 == 0): FOUR rows per warp, eight lanes per row,
// four independent 128-bit loads in flight per lane (the one-row-per-warp form is latency-bound: one load per lane and a
// dependent reduce / scale / store chain per row).
__global__ void __launch_bounds__(256)
convert_f32_rows128_kernel(const float* __restrict__ X, int64_t rows, int d, __half* __restrict__ out,
                           int32_t* __restrict__ row_exp, float* __restrict__ norms) {
  const uint32_t lane = lane_id();
  const int sub = lane & 7, grp = lane >> 3;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r0 = warp0 * 4; r0 < rows; r0 += nwarps * 4) {
    const int64_t r = r0 + grp;
    const bool on = r < rows;
    const float* x = X + r * d;
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = (sub + 8 * j) * 4;
      v[j] = (on && c < d) ? __ldg(reinterpret_cast<const float4*>(x + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float ss = 0.f, am = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ss = fmaf(v[j].x, v[j].x, ss); ss = fmaf(v[j].y, v[j].y, ss); ss = fmaf(v[j].z, v[j].z, ss); ss = fmaf(v[j].w, v[j].w, ss);
      am = fmaxf(am, fmaxf(fmaxf(fabsf(v[j].x), fabsf(v[j].y)), fmaxf(fabsf(v[j].z), fabsf(v[j].w))));
    }
#pragma unroll
    for (int o2 = 1; o2 < 8; o2 <<= 1) {
      am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o2));
      ss += __shfl_xor_sync(0xffffffffu, ss, o2);
    }
    int e = 0;
    if (am > 0.f && am < INFINITY) e = max(-100, min(100, 13 - ilogbf(am)));
    if (on) {
      __half* o = out + r * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = (sub + 8 * j) * 4;
        __half hv[4] = {__float2half_rn(ldexpf(v[j].x, e)), __float2half_rn(ldexpf(v[j].y, e)), __float2half_rn(ldexpf(v[j].z, e)),
                        __float2half_rn(ldexpf(v[j].w, e))};
        *reinterpret_cast<uint2*>(o + c) = *reinterpret_cast<const uint2*>(hv);
      }
      if (sub == 0) { row_exp[r] = e; norms[r] = sqrtf(ss) * 1.000001f; }
    }
  }
}

__global__ void __launch_bounds__(256)
prep_bf16_kernel(const __nv_bfloat16* __restrict__ X, int64_t rows, int d, int d_pad, __nv_bfloat16* __restrict__ out,
                 float* __restrict__ norms, float* __restrict__ inv_norms, uint32_t* __restrict__ normmax_bits) {
  const uint32_t lane = lane_id();
  float wmax = 0.f;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    const __nv_bfloat16* x = X + r * d;
    float ss = 0.f;
    for (int c = lane; c < d_pad; c += 32) {
      const __nv_bfloat16 b = c < d ? x[c] : __float2bfloat16(0.f);
      const float v = __bfloat162float(b);
      ss = fmaf(v, v, ss);
      if (out) out[r * d_pad + c] = b;
    }
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss) * 1.000001f;
    if (lane == 0 && norms) norms[r] = nrm;
    if (lane == 0 && inv_norms) inv_norms[r] = 1.0f / sqrtf(ss);
    wmax = fmaxf(wmax, nrm);
  }
  if (normmax_bits && lane == 0 && wmax > 0.f && wmax < INFINITY) atomicMax(normmax_bits, __float_as_uint(wmax));
}

inline unsigned grid_for(int64_t work, int threads) {
  int64_t g = (work + threads - 1) / threads;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace

cudaError_t launch_prep_f32(const float* X, int64_t rows, int d, int d_pad, __half* out, uint32_t* amax_bits,
                            int32_t* scale_exp, float* norms, float* inv_norms, uint32_t* normmax_bits, cudaStream_t stream) {
  const int64_t n = rows * d;
  const bool al16 = (reinterpret_cast<uintptr_t>(X) & 15) == 0;
  const int64_t n4 = al16 ? n / 4 : 0;
  absmax_kernel<<<grid_for(n4 > 0 ? n4 : n, 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(X), n4, X, n, amax_bits);
  count_launch();
  scale_exp_kernel<<<1, 1, 0, stream>>>(amax_bits, scale_exp);
  count_launch();
  if (d_pad == 128 && (d & 3) == 0 && al16)
    convert_f32_128_kernel<<<grid_for(rows * 8, 256), 256, 0, stream>>>(X, rows, d, out, scale_exp, norms, inv_norms, normmax_bits);
  else
    convert_f32_kernel<<<grid_for(rows * 32, 256), 256, 0, stream>>>(X, rows, d, d_pad, out, scale_exp, norms, inv_norms, normmax_bits);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_prep_f32_rows(const float* X, int64_t rows, int d, int d_pad, __half* out, int32_t* row_exp, float* norms,
                                 cudaStream_t stream) {
  if (d_pad == 128 && (d & 3) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0)
    convert_f32_rows128_kernel<<<grid_for(rows * 8, 256), 256, 0, stream>>>(X, rows, d, out, row_exp, norms);
  else
    convert_f32_rows_kernel<<<grid_for(rows * 32, 256), 256, 0, stream>>>(X, rows, d, d_pad, out, row_exp, norms);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_prep_bf16(const void* X, int64_t rows, int d, int d_pad, void* out, float* norms, float* inv_norms,
                             uint32_t* normmax_bits, cudaStream_t stream) {
  prep_bf16_kernel<<<grid_for(rows * 32, 256), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(X), rows, d, d_pad,
                                                                 reinterpret_cast<__nv_bfloat16*>(out), norms, inv_norms, normmax_bits);
  count_launch();
  return cudaGetLastError();
}

}  // namespace divrec
