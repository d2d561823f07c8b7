// Feasibility test for conv-tap reuse of ONE activation tile: a k3 Conv1d tap is a row shift of the same operand, so if a
// tcgen05 shared-memory descriptor may start r rows (r * 128 B) inside a 128B-swizzled tile that TMA wrote, one (128 + 2)-row
// tile per K chunk can feed all three taps instead of three separately staged tiles.
//   mode 0: start address += r * 128, base_offset field = 0      mode 1: start address += r * 128, base_offset = r
// Prints the max abs error of D_r = A[r : r + 128] B^T against the host for r = 0..7 in both modes.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rowshift rowshift.cu && ./rowshift
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../matcha_tts_b200/csrc/ptx.cuh"
using namespace mtts;

constexpr int AROWS = 136, N = 256, KC = 64;

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                 // 136 x 128 B = 17408 B (17 KB), padded to 18432
  uint8_t* sB = smem + 18432;         // 256 x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 18432 + 32768);
  uint64_t* mbar = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(mbar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<256>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tD = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, AROWS * 128 + N * 128);
    tma_load_2d(sA, &tmA, bar, 0, 0);
    tma_load_2d(sB, &tmB, bar, 0, 0);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  constexpr uint32_t idesc = umma_idesc_f16(128, N);
  uint32_t phase = 0;
  for (int mode = 0; mode < 2; ++mode) {
    for (int r = 0; r < 8; ++r) {
      if (warp == 0) {
        uint64_t da = umma_desc_sw128(smem_u32(sA) + r * 128);
        if (mode == 1) da |= (uint64_t)(r & 7) << 49;
        const uint64_t db = umma_desc_sw128(smem_u32(sB));
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_f16(tD, da + 2 * kk, db + 2 * kk, idesc, kk != 0);
          umma_commit(mbar);
        }
        __syncwarp();
      }
      mbar_wait(mbar, phase);
      phase ^= 1;
      tc_fence_after();
      float* o = out + ((size_t)(mode * 8 + r) * 128 + warp * 32 + lane) * N;
      for (int c = 0; c < N / 32; ++c) {
        float v[32];
        tmem_ld32(tD + (uint32_t(warp * 32) << 16) + c * 32, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) o[c * 32 + j] = v[j];
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  }
  if (warp == 0) tmem_dealloc<256>(tD);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no driver entry point\n"); return 1; }
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
  std::vector<__half> A(AROWS * KC), B(N * KC);
  srand(1);
  for (auto& x : A) x = __float2half((rand() % 2001 - 1000) / 1000.f);
  for (auto& x : B) x = __float2half((rand() % 2001 - 1000) / 1000.f);
  __half *dA, *dB;
  float* dO;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dO, sizeof(float) * 16 * 128 * N);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dO, 0, sizeof(float) * 16 * 128 * N);
  CUtensorMap tA, tB;
  cuuint32_t estr[2] = {1, 1};
  cuuint64_t strides[1] = {KC * 2};
  {
    cuuint64_t dims[2] = {KC, AROWS}; cuuint32_t box[2] = {KC, AROWS};
    if (enc(&tA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dA, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode A failed\n"); return 1; }
  }
  {
    cuuint64_t dims[2] = {KC, N}; cuuint32_t box[2] = {KC, N};
    if (enc(&tB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode B failed\n"); return 1; }
  }
  const int smem = 18432 + 32768 + 64;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<<<1, 128, smem>>>(tA, tB, dO);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> O(16 * 128 * N);
  cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
  for (int mode = 0; mode < 2; ++mode)
    for (int r = 0; r < 8; ++r) {
      double worst = 0;
      for (int i = 0; i < 128; ++i)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int kx = 0; kx < KC; ++kx) ref += (double)__half2float(A[(i + r) * KC + kx]) * (double)__half2float(B[n * KC + kx]);
          worst = fmax(worst, fabs(ref - O[((size_t)(mode * 8 + r) * 128 + i) * N + n]));
        }
      printf("mode %d (base_offset %s) row shift %d: max abs err %.3e %s\n", mode, mode ? "= r" : "= 0", r, worst, worst < 1e-3 ? "OK" : "WRONG");
    }
  return 0;
}
