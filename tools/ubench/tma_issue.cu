// Microbenchmark 2: what limits TMA issue?  Free-running producer(s): each issuing thread loops
// { wait for the load previously sent to this slot; expect_tx; cp.async.bulk.tensor } over its own ring of slots.
// Variants: number of issuing warps, 2-D box rows, 3-D box (several K chunks per instruction).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include "../../matcha_tts_b200/csrc/ptx.cuh"
using namespace mtts;

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(desc), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// nissuers warps, each with `depth` slots of `bytes` bytes
__global__ void __launch_bounds__(256, 1) issue_kernel(const __grid_constant__ CUtensorMap tm, int nissuers, int depth, int bytes,
                                                       int box_rows, int nk, int niter, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 212992);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 32; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (warp < nissuers && lane == 0) {
    uint8_t* base = smem + warp * depth * bytes;
    uint64_t* bar = bars + warp * 8;
    long long t0 = clock64();
    int slot = 0;
    uint32_t phase = 0;
    for (int i = 0; i < niter; ++i) {
      if (i >= depth) mbar_wait(&bar[slot], phase ^ 1);   // previous load into this slot finished
      mbar_arrive_expect_tx(&bar[slot], bytes);
      const int r = ((i * 7 + warp * 3) & 7) * box_rows % 1024;
      if (nk == 1) tma_load_2d(base + slot * bytes, &tm, &bar[slot], (i & 3) * 64, r);
      else tma_load_3d(base + slot * bytes, &tm, &bar[slot], 0, r, 0);
      if (++slot == depth) { slot = 0; phase ^= 1; }
    }
    // drain
    for (int s = 0; s < depth; ++s) {
      const int last_i = niter - 1 - ((niter - 1 - s) % depth);  // not exact; just wait for current phase of every slot
      (void)last_i;
    }
    long long t1 = clock64();
    out[blockIdx.x * 8 + warp] = t1 - t0;
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
  void* buf;
  cudaMalloc(&buf, 64 << 20);
  cudaMemset(buf, 1, 64 << 20);
  long long* out;
  cudaMalloc(&out, 148 * 8 * 8);
  cudaFuncSetAttribute(issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 212992 + 512);
  struct Cfg { int nissuers, depth, box_rows, nk; };
  const Cfg cfgs[] = {{1, 4, 128, 1}, {2, 4, 128, 1}, {4, 2, 128, 1}, {1, 4, 256, 1}, {2, 2, 256, 1}, {1, 2, 128, 4}, {1, 3, 128, 4}, {2, 1, 128, 4}, {1, 2, 256, 2}, {1, 4, 64, 1}, {1, 2, 128, 2}, {1, 3, 128, 2}};
  const int niter = 400;
  for (const Cfg& c : cfgs) {
    const int bytes = c.box_rows * 128 * c.nk;
    CUtensorMap tm;
    CUresult r;
    if (c.nk == 1) {
      cuuint64_t dims[2] = {256, 1024 + 256};
      cuuint64_t strides[1] = {512};
      cuuint32_t box[2] = {64, (cuuint32_t)c.box_rows};
      cuuint32_t es[2] = {1, 1};
      r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {  // [k-chunk][row][64 cols] view of a [rows, 256] matrix
      cuuint64_t dims[3] = {64, 1024 + 256, 4};
      cuuint64_t strides[2] = {512, 128};
      cuuint32_t box[3] = {64, (cuuint32_t)c.box_rows, (cuuint32_t)c.nk};
      cuuint32_t es[3] = {1, 1, 1};
      r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { printf("encode failed %d (rows %d nk %d)\n", (int)r, c.box_rows, c.nk); continue; }
    for (int grid : {1, 148}) {
      for (int rep = 0; rep < 3; ++rep) issue_kernel<<<grid, 256, 212992 + 512>>>(tm, c.nissuers, c.depth, bytes, c.box_rows, c.nk, niter, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      std::vector<long long> h(grid * 8);
      cudaMemcpy(h.data(), out, grid * 64, cudaMemcpyDeviceToHost);
      double cyc = 0;
      for (int i = 0; i < grid; ++i) for (int w = 0; w < c.nissuers; ++w) cyc += h[i * 8 + w];
      cyc /= (grid * c.nissuers);
      const double per = cyc / niter;
      printf("issuers=%d depth=%d box=[%3d rows x 64] x%d  (%6d B/instr) grid=%3d : %6.0f cycles/instr/thread -> %6.1f B/clk/SM  (%5.1f TB/s agg @1.9GHz)\n",
             c.nissuers, c.depth, c.box_rows, c.nk, bytes, grid, per, bytes * c.nissuers / per, bytes * c.nissuers / per * 1.9e9 * grid / 1e12);
      fflush(stdout);
    }
  }
  return 0;
}
