// FP32 issue-rate microbenchmark: the same 8 multiply-adds per iteration as 8 scalar FFMA or as 4 packed FFMA2
// (fma.rn.f32x2, sm_100).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>

template <int PACKED>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
  float2 v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (PACKED) v[i] = __ffma2_rn(v[i], a2, b2);
        else { v[i].x = fmaf(v[i].x, a, b); v[i].y = fmaf(v[i].y, a, b); }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += v[i].x + v[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  const int blocks = 148 * 8, threads = 256, iters = 4096;
  float* out;
  cudaMalloc(&out, blocks * threads * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int packed = 0; packed < 2; ++packed) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (packed) k<1><<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
      else k<0><<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fma = (double)blocks * threads * iters * 64;   // scalar multiply-adds
      if (rep == 2) printf("%s: %.3f ms  %.1f TFLOP/s fp32  (%.2f multiply-adds per SM per clock at 1.965 GHz)\n", packed ? "FFMA2" : "FFMA ",
                           ms, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.965e9);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
