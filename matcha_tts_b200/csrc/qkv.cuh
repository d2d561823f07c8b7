// The QKV projection of the transformer blocks (reference model.py:662-664: to_q / to_k / to_v, no bias; 256 -> 3 x 128)
// as its own tcgen05 kernel: q | k | v = a Wqkv^T for one 128-row tile per step of a persistent CTA.
//
// Through the generic gemm_tc_kernel<128, EPI_QKV> this GEMM ran at 410-550 TFLOP/s: K is only 256, so a (row tile, N
// tile) unit is 8 K16 steps long and the 64 KB activation tile was staged once for EACH of the three 128-wide N tiles.
// Here a CTA owns whole row tiles: the `a` tile is staged ONCE (4 K-chunk tiles, double-buffered across row tiles), the
// twelve 16 KB weight pieces stream through a ring, and the three 128-column accumulators rotate through FOUR tensor
// memory slots -- the q part of tile i+1 starts in the slot tile i did not use while the epilogue still drains tile i, and
// every later part finds the slot the epilogue emptied one part earlier -- so MMA issue and the drain overlap throughout.
//
// Warp roles (352 threads, one CTA per SM): warp 0 TMA producer of `a` (waits for the previous kernel), warp 1 TMEM
// allocator + MMA issuer, warp 2 TMA producer of the weight pieces (constants: no dependency wait), warps 3-10 epilogue
// (TMEM lane quarter = warp % 4, 64 of a slot's 128 columns each, swizzled staging -> coalesced 64-byte row stores).
#pragma once
#include <cuda.h>

#include "gemm_tc.cuh"
#include "ptx.cuh"

namespace mtts {

struct QkvParams {
  int M;        // rows of the level's flat row space
  __half* q;    // [rows, 128]
  __half* k;    // [rows, 128]
  __half* v;    // [rows, 128]
  int w_hint;
  int pdl_late;
  long long* tl;   // debug: [gridDim.x][128] clock64 stamps (tools/qkv_timeline.py), or null
};

constexpr int QKV_NST = 5;                                     // weight ring slots
constexpr int QKV_PIECE = 16384;                               // [128 N rows x 64 K]
constexpr int QKV_A_BYTES = 65536;                             // 4 K-chunk tiles of 128 rows x 128 B
constexpr int QKV_THREADS = 96 + 32 * GEMM_EPI_WARPS;          // 352
constexpr int QKV_OFF_RING = 2 * QKV_A_BYTES;
constexpr int QKV_OFF_STAGE = QKV_OFF_RING + QKV_NST * QKV_PIECE;
constexpr int QKV_OFF_BAR = QKV_OFF_STAGE + GEMM_EPI_WARPS * GEMM_STAGING_BYTES;
constexpr int QKV_SMEM = QKV_OFF_BAR + 256;
static_assert(QKV_SMEM <= 232448, "exceeds the 227 KB of shared memory one CTA can own");

__global__ void __launch_bounds__(QKV_THREADS, 1)
qkv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const QkvParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + QKV_OFF_BAR);
  uint64_t* a_full = bars;                         // [2]
  uint64_t* a_empty = bars + 2;                    // [2]
  uint64_t* w_full = bars + 4;                     // [QKV_NST]
  uint64_t* w_empty = bars + 4 + QKV_NST;          // [QKV_NST]
  uint64_t* s_full = bars + 4 + 2 * QKV_NST;       // [4] accumulator slot complete
  uint64_t* s_empty = bars + 8 + 2 * QKV_NST;      // [4] accumulator slot drained (8 epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12 + 2 * QKV_NST);
  static_assert((12 + 2 * QKV_NST + 1) * 8 <= 256, "barrier block");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (!p.pdl_late) pdl_launch_dependents();
  const int m_tiles = (p.M + 127) / 128;
  const int nt = ((int)blockIdx.x < m_tiles) ? (m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < QKV_NST; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], GEMM_EPI_WARPS); }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 2) {
    // ===================================== TMA producer: Wqkv pieces (constants: no dependency wait) =====
    uint32_t it = 0;
    const uint64_t pol = l2_policy_evict_last();
    for (int i = 0; i < nt; ++i) {
      for (int pc = 0; pc < 12; ++pc, ++it) {   // piece pc: part pc/4 (q, k, v), K chunk pc%4
        const uint32_t slot = it % QKV_NST, use = it / QKV_NST;
        mbar_wait(&w_empty[slot], (use & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&w_full[slot], QKV_PIECE);
          if (p.w_hint) tma_load_2d_hint(smem + QKV_OFF_RING + slot * QKV_PIECE, &tmW, &w_full[slot], (pc & 3) * 64, (pc >> 2) * 128, pol);
          else tma_load_2d(smem + QKV_OFF_RING + slot * QKV_PIECE, &tmW, &w_full[slot], (pc & 3) * 64, (pc >> 2) * 128);
        }
        __syncwarp();
      }
    }
  } else if (warp == 0) {
    // ===================================== TMA producer: the `a` tile, once per row tile ==========
    pdl_wait();
    for (int i = 0; i < nt; ++i) {
      const int buf = i & 1, r0 = ((int)blockIdx.x + i * (int)gridDim.x) * 128;
      mbar_wait(&a_empty[buf], ((i >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&a_full[buf], QKV_A_BYTES);
#pragma unroll
        for (int c = 0; c < 4; ++c) tma_load_2d(smem + buf * QKV_A_BYTES + c * 16384, &tmA, &a_full[buf], c * 64, r0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    constexpr uint32_t idesc = umma_idesc_f16(128, 128);
    const uint32_t ring = smem_u32(smem + QKV_OFF_RING);
    uint32_t it = 0;
    for (int i = 0; i < nt; ++i) {
      const uint32_t abuf = smem_u32(smem + (i & 1) * QKV_A_BYTES);
      mbar_wait(&a_full[i & 1], (i >> 1) & 1);
      if (p.tl && lane == 0 && i < 5) p.tl[(size_t)blockIdx.x * 128 + i * 16 + 0] = clock64();
      for (int part = 0; part < 3; ++part) {
        const uint32_t g = 3u * (uint32_t)i + part, slot4 = g & 3u;
        mbar_wait(&s_empty[slot4], ((g >> 2) & 1) ^ 1);   // the epilogue has drained the slot's previous accumulator
        tc_fence_after();
        if (p.tl && lane == 0 && i < 5) p.tl[(size_t)blockIdx.x * 128 + i * 16 + 1 + 2 * part] = clock64();
        for (int kc = 0; kc < 4; ++kc, ++it) {
          const uint32_t slot = it % QKV_NST, use = it / QKV_NST;
          mbar_wait(&w_full[slot], use & 1);
          tc_fence_after();
          const uint64_t da0 = umma_desc_sw128(abuf + kc * 16384), db0 = umma_desc_sw128(ring + slot * QKV_PIECE);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_f16(tmem_base + slot4 * 128, da0 + 2 * kk, db0 + 2 * kk, idesc, (kc | kk) != 0);
            umma_commit(&w_empty[slot]);
            if (kc == 3) umma_commit(&s_full[slot4]);
            if (kc == 3 && part == 2) umma_commit(&a_empty[i & 1]);
          }
          __syncwarp();
        }
        if (p.tl && lane == 0 && i < 5) p.tl[(size_t)blockIdx.x * 128 + i * 16 + 2 + 2 * part] = clock64();
      }
    }
  } else {
    // ===================================== epilogue =========================================
    const int ew = warp - 3;            // 0..7
    const int q4 = warp & 3;            // TMEM lane quarter
    const int ch = ew >> 2;             // column half of a slot: [ch*64, ch*64 + 64)
    const uint32_t st = smem_u32(smem + QKV_OFF_STAGE + ew * GEMM_STAGING_BYTES);
    for (int i = 0; i < nt; ++i) {
      const int r0 = ((int)blockIdx.x + i * (int)gridDim.x) * 128;
      const int rw0 = r0 + q4 * 32;
      const int rows_valid = min(32, p.M - rw0);
      for (int part = 0; part < 3; ++part) {
        const uint32_t g = 3u * (uint32_t)i + part, slot4 = g & 3u;
        if (lane == 0) {
          mbar_wait(&s_full[slot4], (g >> 2) & 1);
          if (p.pdl_late && i + 1 == nt && part == 2) pdl_launch_dependents();
        }
        __syncwarp();
        tc_fence_after();
        if (p.tl && ew == 0 && lane == 0 && i < 5) p.tl[(size_t)blockIdx.x * 128 + i * 16 + 8 + 2 * part] = clock64();
        const uint32_t taddr = tmem_base + (uint32_t(q4 * 32) << 16) + slot4 * 128 + ch * 64;
        __half* dst = (part == 0 ? p.q : (part == 1 ? p.k : p.v)) + (size_t)rw0 * 128 + ch * 64;
        float vbuf[2][32];
        tmem_ld32(taddr, vbuf[0]);
        tmem_ld_wait();
        tmem_ld32(taddr + 32, vbuf[1]);
        epi_store_h32(st, lane, vbuf[0], dst, 128, rows_valid);
        tmem_ld_wait();
        // both chunks are in registers: hand the slot back before the second store
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[slot4]);
        epi_store_h32(st, lane, vbuf[1], dst + 32, 128, rows_valid);
        if (p.tl && ew == 0 && lane == 0 && i < 5) p.tl[(size_t)blockIdx.x * 128 + i * 16 + 9 + 2 * part] = clock64();
      }
    }
    if (nt == 0 && p.pdl_late && lane == 0) pdl_launch_dependents();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

}  // namespace mtts
