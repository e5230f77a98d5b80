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
  convert_f32_kernel<<<grid_for(rows * 32, 256), 256, 0, stream>>>(X, rows, d, d_pad, out, scale_exp, norms, inv_norms, normmax_bits);
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
