// prep.cu -- operand preparation for the tensor-core scorer.
//
// fp32-parity mode: an fp32 table X is scaled by a power of two so that max|X| lies in
// [2^13, 2^14) and split into three fp16 planes
//     hi  = fp16(x)            hs = fp16(hi * 2^-11)            lo = fp16((x - hi) * 2^11)
// With user planes laid out [hi | hs | lo] and item planes [hi | lo | hs] along K, one fp16 GEMM
// over K = 3 * d_pad accumulates  hi*hi' + (hi*2^-11)(lo'*2^11) + (lo*2^11)(hi'*2^-11)
//                              =  hi*hi' + hi*lo' + lo*hi'   in fp32,
// i.e. the fp32 product x*x' up to the dropped lo*lo' term (<= 2^-24 relative) and the rounding
// of lo (<= 2^-24 relative): the arithmetic of MatrixFactorization.forward
// (divrec/models/matrix_factorization.py:26-28) at fp32 accuracy on fp16 tensor cores.
// Values that are exact in fp16 after scaling (e.g. the k/8 grid used by the bit-exact parity
// tier) give lo == 0 and an exact result.
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
  // tail (n not a multiple of 4)
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

// one thread per 4 consecutive elements of a row (d_pad % 4 == 0); writes three 8-byte stores
__global__ void split_kernel(const float* __restrict__ X, int64_t rows, int d, int d_pad, int role,
                             __half* __restrict__ out, const int32_t* __restrict__ scale_exp) {
  const int q_per_row = d_pad >> 2;
  const int64_t total = rows * q_per_row;
  const int e = *scale_exp;
  const int K3 = 3 * d_pad;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / q_per_row;
    const int c = (int)(t - r * q_per_row) * 4;
    float x[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = (c + j < d) ? X[r * d + c + j] : 0.f;
    __half hi[4], hs[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float xs = ldexpf(x[j], e);
      const __half h = __float2half_rn(xs);
      const float hf = __half2float(h);
      hi[j] = h;
      hs[j] = __float2half_rn(hf * 0.00048828125f);        // 2^-11
      lo[j] = __float2half_rn((xs - hf) * 2048.0f);        // 2^11
    }
    __half* o = out + r * K3 + c;
    const int p_hs = role == 0 ? 1 : 2;
    const int p_lo = role == 0 ? 2 : 1;
    *reinterpret_cast<uint2*>(o) = *reinterpret_cast<const uint2*>(hi);
    *reinterpret_cast<uint2*>(o + (size_t)p_hs * d_pad) = *reinterpret_cast<const uint2*>(hs);
    *reinterpret_cast<uint2*>(o + (size_t)p_lo * d_pad) = *reinterpret_cast<const uint2*>(lo);
  }
}

__global__ void pad_bf16_kernel(const __nv_bfloat16* __restrict__ X, int64_t rows, int d, int d_pad,
                                __nv_bfloat16* __restrict__ out) {
  const int64_t total = rows * d_pad;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / d_pad;
    const int c = (int)(t - r * d_pad);
    out[t] = c < d ? X[r * d + c] : __float2bfloat16(0.f);
  }
}

inline unsigned grid_for(int64_t work, int threads) {
  int64_t g = (work + threads - 1) / threads;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace

cudaError_t launch_split_f32(const float* X, int64_t rows, int d, int d_pad, int role, __half* out,
                             uint32_t* amax_bits, int32_t* scale_exp, cudaStream_t stream) {
  const int64_t n = rows * d;
  const bool al16 = (reinterpret_cast<uintptr_t>(X) & 15) == 0;
  const int64_t n4 = al16 ? n / 4 : 0;
  absmax_kernel<<<grid_for(n4 > 0 ? n4 : n, 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(X), n4, X, n, amax_bits);
  count_launch();
  scale_exp_kernel<<<1, 1, 0, stream>>>(amax_bits, scale_exp);
  count_launch();
  split_kernel<<<grid_for(rows * (d_pad / 4), 256), 256, 0, stream>>>(X, rows, d, d_pad, role, out, scale_exp);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_pad_bf16(const void* X, int64_t rows, int d, int d_pad, void* out, cudaStream_t stream) {
  pad_bf16_kernel<<<grid_for(rows * d_pad, 256), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(X), rows, d,
                                                                    d_pad, reinterpret_cast<__nv_bfloat16*>(out));
  count_launch();
  return cudaGetLastError();
}

}  // namespace divrec
