// Microbenchmark: per-SM TMA load throughput/latency for [128 rows x 64 fp16] pieces (16 KB, 128B swizzle)
// as a function of ring depth, grid size, row pitch and whether all CTAs read the same tensor (weights)
// or disjoint tensors (activations).   nvcc -arch=sm_100a -O3 -o tma_bench tma_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../matcha_tts_b200/csrc/ptx.cuh"
using namespace mtts;

constexpr int PIECE = 16384;  // bytes per 128 box rows
__device__ __forceinline__ void wait_v(uint64_t* bar, uint32_t parity, int mode) {
  if (mode == 0) { mbar_wait(bar, parity); return; }
  if (mode == 1) {  // non-blocking test_wait spin
    uint32_t ok = 0;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
    return;
  }
  uint32_t ok = 0;  // try_wait with a small suspend-time hint (ns)
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(20) : "memory");
  } while (!ok);
}

__global__ void __launch_bounds__(64, 1) bench_kernel(const __grid_constant__ CUtensorMap tm, int depth, int npieces,
                                                      int cols_total, int rows_per_cta, int shared_data,
                                                      long long* out, int box_rows, int split, int freerun, int wmode, long long* trace) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + 13 * PIECE);
  uint64_t* empty_bar = full_bar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const int kchunks = cols_total / 64;
  const int row_base = shared_data ? 0 : blockIdx.x * rows_per_cta;
  long long t0 = clock64();
  if (warp == 0 && lane == 0) {
    long long acc_wait = 0, acc_exp = 0, acc_tma = 0, acc_misc = 0, tprev = clock64();
    for (int i = 0; i < npieces; ++i) {
      const int slot = i % depth, use = i / depth;
      long long ta = clock64();
      acc_misc += ta - tprev;
      if (!freerun && wmode != 4) wait_v(&empty_bar[slot], (use & 1) ^ 1, wmode == 3 ? 0 : wmode);
      else if (use > 0) wait_v(&full_bar[slot], (use - 1) & 1, wmode);   // only wait for the previous load into this slot
      const int pbytes = box_rows * 128 * split;
      long long tb = clock64();
      acc_wait += tb - ta;
      mbar_arrive_expect_tx(&full_bar[slot], pbytes);
      long long tc = clock64();
      acc_exp += tc - tb;
      const int kc = i % kchunks, rb = (i / kchunks) * box_rows * split % rows_per_cta;
      if (trace && blockIdx.x == 0 && i < 32) trace[i] = clock64() - t0;
      for (int sp = 0; sp < split; ++sp)
        tma_load_2d(smem + slot * pbytes + sp * box_rows * 128, &tm, &full_bar[slot], kc * 64, row_base + rb + sp * box_rows);
      tprev = clock64();
      acc_tma += tprev - tc;
    }
    if (trace && blockIdx.x == 0) { trace[24] = acc_wait / npieces; trace[25] = acc_exp / npieces; trace[26] = acc_tma / npieces; trace[27] = acc_misc / npieces; }
  } else if (warp == 1 && lane == 0) {
    long long tfirst = 0;
    for (int i = 0; i < npieces; ++i) {
      const int slot = i % depth, use = i / depth;
      if (wmode == 3) {
        uint32_t ok = 0;
        while (true) {
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                       : "=r"(ok) : "r"(smem_u32(&full_bar[slot])), "r"(use & 1) : "memory");
          if (ok) break;
          __nanosleep(100);
        }
      } else if (wmode == 5) {
        if (i >= npieces - depth) wait_v(&full_bar[slot], use & 1, 0);
        else { long long tt = clock64(); while (clock64() - tt < 300) {} }
      } else if (!freerun || i >= npieces - depth) wait_v(&full_bar[slot], use & 1, wmode == 4 ? 0 : wmode);
      if (i == 0) tfirst = clock64();
      if (trace && blockIdx.x == 0 && i < 32) trace[32 + i] = clock64() - t0;
      if (!freerun) mbar_arrive(&empty_bar[slot]);
    }
    long long t1 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = tfirst - t0;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  const size_t bytes = 512ull << 20;
  void* buf;
  cudaMalloc(&buf, bytes);
  cudaMemset(buf, 1, bytes);
  long long* out;
  cudaMalloc(&out, 148 * 2 * 8);
  long long* trace;
  cudaMalloc(&trace, 64 * 8);
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 13 * PIECE + 512);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("# cols(pitch B)  shared grid depth | per-CTA GB/s   aggregate TB/s   first-piece latency us   (clock %d MHz assumed 1.85 GHz under load)\n", clk_khz / 1000);
  const int npieces = 512;
  printf("# box_rows split depth grid | bytes/stage  per-CTA GB/s  aggregate TB/s  first latency us\n");
  for (int wmode : {0, 3})
  for (int freerun : {0}) {
    for (int box_rows : {128}) {
      for (int split : {1}) {
        for (int depth : {4}) {
          for (int grid : {148}) {
            const int cols = 256, rows_per_cta = 1024, shared_data = 1;
            const int pbytes = box_rows * 128 * split;
            if (pbytes * depth > 13 * PIECE) continue;
            CUtensorMap tm;
            cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows_per_cta};
            cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
            cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
            cuuint32_t es[2] = {1, 1};
            CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
            for (int rep = 0; rep < 3; ++rep)
              bench_kernel<<<grid, 64, 13 * PIECE + 512>>>(tm, depth, npieces, cols, rows_per_cta, shared_data, out, box_rows, split, freerun, wmode, trace);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            std::vector<long long> h(grid * 2);
            cudaMemcpy(h.data(), out, grid * 16, cudaMemcpyDeviceToHost);
            double cyc = 0, first = 0;
            for (int i = 0; i < grid; ++i) { cyc += h[2 * i]; first += h[2 * i + 1]; }
            cyc /= grid; first /= grid;
            const double sec = cyc / 1.9e9;
            const double gbs = (double)npieces * pbytes / sec / 1e9;
            { long long ht[64]; cudaMemcpy(ht, trace, 64 * 8, cudaMemcpyDeviceToHost);
              printf("issue   :"); for (int i = 0; i < 20; ++i) printf(" %lld", ht[i]); printf("\n");
              printf("consumed:"); for (int i = 0; i < 20; ++i) printf(" %lld", ht[32 + i]); printf("\n");
              printf("producer per-iteration cycles: empty-wait %lld  arrive.expect_tx %lld  tma-issue(+trace) %lld  loop-misc %lld\n", ht[24], ht[25], ht[26], ht[27]); fflush(stdout); }
            printf("wmode=%d freerun=%d box_rows=%3d depth=%d grid=%3d | %7d B/stage %8.1f GB/s/CTA %8.2f TB/s agg  first %.2f us (%.0f cycles/stage)\n", wmode, freerun, box_rows, depth, grid, pbytes, gbs, gbs * grid / 1e3,
                   first / 1.9e3, cyc / npieces);
          }
        }
      }
    }
  }
  return 0;
}
