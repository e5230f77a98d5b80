// Probe: tcgen05.mma with the A operand in tensor memory (TS form), B from shared memory (128B swizzle, via TMA).
// Layout under test: A[m][k] (16-bit) -> TMEM lane m, 32-bit column k/2 (low half = even k).
// Also times SS vs TS issue rate.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ts_mma ts_mma.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../../diversity-recommendations_b200/csrc/common.cuh"
using namespace divrec;

__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}

constexpr int K = 64;  // one 128-byte swizzle k-block
#ifndef PROBE_DBASE
#define PROBE_DBASE 0
#endif
#ifndef PROBE_DSTRIDE
#define PROBE_DSTRIDE 128   // column stride between the two accumulators the timing modes alternate
#endif
#ifndef PROBE_N
#define PROBE_N 128   // MMA N under test for the timing modes (-DPROBE_N=96: the three-column-split epilogue's tile)
#endif
// mode 0: TS correctness (out = A B^T); mode 1: time `iters` TS MMAs; mode 2: time SS MMAs
__global__ void __launch_bounds__(192) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                             const __half* A, float* out, int mode, int iters, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sb = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = sb;            // 16 KB (SS timing only)
  uint8_t* sB = sb + 16384;    // 16 KB
  __shared__ uint64_t full, done, scratch;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  if (threadIdx.x == 0) { mbar_init(smem_u32(&full), 1); mbar_init(smem_u32(&done), 1); mbar_init(smem_u32(&scratch), 1); fence_mbar_init(); }
  if (warp == 4) { tmem_alloc(smem_u32(&tbase), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tbase;
  const uint32_t A_COL = 256;  // A lives in columns [256, 256 + K/2)
  if (warp == 5 && lane == 0) {
    mbar_arrive_expect_tx(smem_u32(&full), 2 * 16384);
    tma_load_2d(smem_u32(sA), &tmA, smem_u32(&full), 0, 0);
    tma_load_2d(smem_u32(sB), &tmB, smem_u32(&full), 0, 0);
  }
  if (warp < 4) {
    // each thread writes its row of A (K halves = K/2 packed words) into TMEM
    const int row = warp * 32 + lane;
    uint32_t r[32];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(A + (size_t)row * K);
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = src[j];
    tmem_st_x32(tb + ((uint32_t)(warp * 32) << 16) + A_COL, r);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 5 && lane == 0) {
    mbar_wait(smem_u32(&full), 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_f16(0, 128, mode == 0 ? 128 : PROBE_N);
    const uint64_t bdesc = umma_desc_sw128(smem_u32(sB));
    const uint64_t adesc = umma_desc_sw128(smem_u32(sA));
    const long long t0 = clock64();
    if (mode == 0) {
      for (int k4 = 0; k4 < K / 16; ++k4) umma_f16_ts(tb, tb + A_COL + 8 * k4, bdesc + 2 * k4, idesc, k4 != 0);
    } else if (mode <= 2) {
      for (int it = 0; it < iters; ++it)
        for (int k4 = 0; k4 < K / 16; ++k4) {
          if (mode == 1) umma_f16_ts(tb + (it & 1) * PROBE_DSTRIDE + PROBE_DBASE, tb + A_COL + 8 * k4, bdesc + 2 * k4, idesc, k4 != 0);
          else umma_f16(tb + (it & 1) * PROBE_DSTRIDE + PROBE_DBASE, adesc + 2 * k4, bdesc + 2 * k4, idesc, k4 != 0);
        }
    } else if (mode == 3) {  // kernel-like: 4 MMAs, then a wait on an already-completed barrier + fence + commit
      for (int it = 0; it < iters; ++it) {
        mbar_wait(smem_u32(&full), 0); tc_fence_after();
        for (int k4 = 0; k4 < 4; ++k4) umma_f16_ts(tb + (it & 1) * PROBE_DSTRIDE + PROBE_DBASE, tb + A_COL + 8 * k4, bdesc + 2 * k4, idesc, k4 != 0);
        umma_commit(smem_u32(&scratch));
      }
    } else if (mode == 4) {  // 8 MMAs per two waits
      for (int it = 0; it < iters / 2; ++it) {
        mbar_wait(smem_u32(&full), 0); tc_fence_after(); mbar_wait(smem_u32(&full), 0); tc_fence_after();
        for (int k4 = 0; k4 < 8; ++k4) umma_f16_ts(tb + (it & 1) * PROBE_DSTRIDE + PROBE_DBASE, tb + A_COL + 8 * (k4 & 3), bdesc + 2 * (k4 & 3), idesc, k4 != 0);
        umma_commit(smem_u32(&scratch)); umma_commit(smem_u32(&scratch));
      }
    } else if (mode == 5) {  // 8 MMAs per two waits, the waits issued before the last 3 MMAs of the previous group
      for (int it = 0; it < iters / 2; ++it) {
        for (int k4 = 0; k4 < 5; ++k4) umma_f16_ts(tb + (it & 1) * PROBE_DSTRIDE + PROBE_DBASE, tb + A_COL + 8 * (k4 & 3), bdesc + 2 * (k4 & 3), idesc, k4 != 0);
        mbar_wait(smem_u32(&full), 0); mbar_wait(smem_u32(&full), 0);
        for (int k4 = 5; k4 < 8; ++k4) umma_f16_ts(tb + (it & 1) * PROBE_DSTRIDE + PROBE_DBASE, tb + A_COL + 8 * (k4 & 3), bdesc + 2 * (k4 & 3), idesc, 1);
        umma_commit(smem_u32(&scratch)); umma_commit(smem_u32(&scratch)); tc_fence_after();
      }
    } else if (mode == 6) {  // issue timestamps of 24 back-to-back MMAs
      long long ts[25];
      ts[0] = clock64();
#pragma unroll
      for (int i = 0; i < 24; ++i) { umma_f16_ts(tb + (i & 1) * 128, tb + A_COL + 8 * (i & 3), bdesc + 2 * (i & 3), idesc, 1); ts[i + 1] = clock64(); }
      if (blockIdx.x == 0) for (int i = 0; i < 25; ++i) cyc[148 + i] = ts[i] - ts[0];
    }
    umma_commit(smem_u32(&done));
    mbar_wait(smem_u32(&done), 0);
    const long long t1 = clock64();
    if (cyc) cyc[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp < 4 && mode == 0) {
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_wait_ld();
      for (int j = 0; j < 32; ++j) out[row * 128 + c0 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  PFN_encodeTiled enc = (PFN_encodeTiled)f;
  std::vector<__half> hA(128 * K), hB(128 * K);
  for (int i = 0; i < 128 * K; ++i) { hA[i] = __float2half((rand() % 17 - 8) / 8.0f); hB[i] = __float2half((rand() % 17 - 8) / 8.0f); }
  __half *dA, *dB; float* dO; long long* dC;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, 128 * 128 * 4); cudaMalloc(&dC, 256 * 8);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tA, tB;
  cuuint64_t gdim[2] = {K, 128}; cuuint64_t gstr[1] = {K * 2}; cuuint32_t box[2] = {64, 128}; cuuint32_t es[2] = {1, 1};
  enc(&tA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dA, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  enc(&tB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int smem = 1024 + 32768;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 192, smem>>>(tA, tB, dA, dO, 0, 0, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  printf("TS correctness run: %s\n", cudaGetErrorString(e));
  std::vector<float> hO(128 * 128);
  cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0; double maxd = 0;
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) {
    double s = 0; for (int k = 0; k < K; ++k) s += (double)__half2float(hA[m * K + k]) * __half2float(hB[n * K + k]);
    const double dd = fabs(s - hO[m * 128 + n]); if (dd > 1e-6) ++bad; if (dd > maxd) maxd = dd;
  }
  printf("TS layout check: %d / 16384 mismatches, max diff %g\n", bad, maxd);
  const char* names[] = {"", "TS (A in TMEM)", "SS (A in smem)", "TS, 4 MMA per wait+fence+commit", "TS, 8 MMA per 2 waits", "TS, 8 MMA, waits hoisted before the last 3", "TS issue timestamps"};
  for (int mode = 1; mode <= 6; ++mode) {
    const int iters = 2000;
    probe<<<148, 192, smem>>>(tA, tB, dA, dO, mode, iters, dC);
    e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, dC, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    if (mode < 6) printf("%s: %.1f cycles per M128xN%dxK16 MMA (%s)\n", names[mode], c / (iters * 4.0), PROBE_N, cudaGetErrorString(e));
    else { long long t[25]; cudaMemcpy(t, dC + 148, sizeof(t), cudaMemcpyDeviceToHost); printf("issue timestamps:"); for (int i = 0; i < 25; ++i) printf(" %lld", t[i]); printf("\n"); }
  }
  return 0;
}
