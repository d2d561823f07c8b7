// The QKV projection of the transformer blocks (reference model.py:662-664: to_q / to_k / to_v, no bias; 256 -> 3 x 128)
// as its own tcgen05 kernel: q | k | v = a Wqkv^T for one 128-row tile per step of a persistent CTA.
//
// Through the generic gemm_tc_kernel (128-wide N tiles, removed since) this GEMM ran at 410-550 TFLOP/s: K is only 256, so a (row tile, N
// tile) unit is 8 K16 steps long and the 64 KB activation tile was staged once for EACH of the three 128-wide N tiles.
// Here a CTA owns whole row tiles: the `a` tile is staged ONCE (4 K-chunk tiles, double-buffered across row tiles), the
// twelve 16 KB weight pieces stream through a ring, and the three 128-column accumulators rotate through FOUR tensor
// memory slots -- the q part of tile i+1 starts in the slot tile i did not use while the epilogue still drains tile i, and
// every later part finds the slot the epilogue emptied one part earlier -- so MMA issue and the drain overlap throughout.
//
// CG = 2 (default): a CTA pair (cluster of two, tcgen05 cta_group::2) per 256 rows -- each CTA stages its own 128-row `a` tile and
// HALF of every weight piece (64 of the 128 output rows), so the 192 KB of Wqkv a row tile streams from L2 (what bounded
// this kernel: 53 GB/s per SM against the 119 GB/s its MMAs consume, profiles/r02z_qkv_timeline.txt) halve per SM, and the
// ring holds eight 8 KB half pieces.  The leader issues the MMAs; commits are multicast; the peer's loads and its epilogue's
// slot releases signal the leader's barriers.
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

constexpr int QKV_NST = 5;                                     // weight ring slots (CTA pairs: 8 slots of half pieces in the same 80 KB)
constexpr int QKV_NST2 = 8;
constexpr int QKV_PIECE = 16384;                               // [128 N rows x 64 K]
constexpr int QKV_A_BYTES = 65536;                             // 4 K-chunk tiles of 128 rows x 128 B
constexpr int QKV_THREADS = 96 + 32 * GEMM_EPI_WARPS;          // 352
constexpr int QKV_OFF_RING = 2 * QKV_A_BYTES;
constexpr int QKV_OFF_STAGE = QKV_OFF_RING + QKV_NST * QKV_PIECE;
constexpr int QKV_OFF_BAR = QKV_OFF_STAGE + GEMM_EPI_WARPS * GEMM_STAGING_BYTES;
constexpr int QKV_SMEM = QKV_OFF_BAR + 256;
static_assert(QKV_SMEM <= 232448, "exceeds the 227 KB of shared memory one CTA can own");

template <int CG>
__global__ void __launch_bounds__(QKV_THREADS, 1)
qkv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const QkvParams p) {
  constexpr int NST = (CG == 2) ? QKV_NST2 : QKV_NST;
  constexpr int PIECE = QKV_PIECE / CG;          // this CTA's share of a weight piece
  const uint32_t crank = (CG == 2) ? cluster_ctarank() : 0u;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + QKV_OFF_BAR);
  uint64_t* a_full = bars;                         // [2]
  uint64_t* a_empty = bars + 2;                    // [2]
  uint64_t* w_full = bars + 4;                     // [NST]
  uint64_t* w_empty = bars + 4 + NST;              // [NST]
  uint64_t* s_full = bars + 4 + 2 * NST;           // [4] accumulator slot complete
  uint64_t* s_empty = bars + 8 + 2 * NST;          // [4] accumulator slot drained (8 epilogue warps per CTA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12 + 2 * NST);
  static_assert((12 + 2 * NST + 1) * 8 <= 256, "barrier block");
  static_assert(NST * PIECE <= QKV_NST * QKV_PIECE, "the half-piece ring fits the ring region");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (!p.pdl_late) pdl_launch_dependents();
  const int m_tiles = (p.M + 127) / 128;
  // a unit = CG consecutive row tiles; unit u0 + i * nunits is this CTA's (pair's) i-th
  const int m_units = (m_tiles + CG - 1) / CG, u0 = (int)blockIdx.x / CG, nunits = (int)gridDim.x / CG;
  const int nt = (u0 < m_units) ? (m_units - u0 + nunits - 1) / nunits : 0;
  auto unit_r0 = [&](int i) -> int { return ((u0 + i * nunits) * CG + (int)crank) * 128; };   // first row of this CTA in its i-th unit

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NST; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], GEMM_EPI_WARPS * CG); }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1) {
    if constexpr (CG == 2) tmem_alloc_pair<512>(tmem_slot);
    else tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them remotely
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 2) {
    // ===================================== TMA producer: Wqkv pieces (constants: no dependency wait) =====
    uint32_t it = 0;
    const uint64_t pol = l2_policy_evict_last();
    for (int i = 0; i < nt; ++i) {
      for (int pc = 0; pc < 12; ++pc, ++it) {   // piece pc: part pc/4 (q, k, v), K chunk pc%4
        const uint32_t slot = it % NST, use = it / NST;
        mbar_wait(&w_empty[slot], (use & 1) ^ 1);
        if (elect_one()) {
          uint8_t* dst = smem + QKV_OFF_RING + slot * PIECE;
          if constexpr (CG == 2) {   // this CTA's 64 of the piece's 128 output rows; both halves are counted on the leader's barrier
            const uint32_t lbar = mapa_u32(smem_u32(&w_full[slot]), 0);
            if (crank == 0) mbar_arrive_expect_tx(&w_full[slot], 2 * PIECE);
            if (p.w_hint) tma_load_2d_pair_hint(dst, &tmW, lbar, (pc & 3) * 64, (pc >> 2) * 128 + (int)crank * 64, pol);
            else tma_load_2d_pair(dst, &tmW, lbar, (pc & 3) * 64, (pc >> 2) * 128 + (int)crank * 64);
          } else {
            mbar_arrive_expect_tx(&w_full[slot], PIECE);
            if (p.w_hint) tma_load_2d_hint(dst, &tmW, &w_full[slot], (pc & 3) * 64, (pc >> 2) * 128, pol);
            else tma_load_2d(dst, &tmW, &w_full[slot], (pc & 3) * 64, (pc >> 2) * 128);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 0) {
    // ===================================== TMA producer: the `a` tile, once per row tile ==========
    pdl_wait();
    for (int i = 0; i < nt; ++i) {
      const int buf = i & 1, r0 = unit_r0(i);
      mbar_wait(&a_empty[buf], ((i >> 1) & 1) ^ 1);
      if (elect_one()) {
        if constexpr (CG == 2) {
          const uint32_t lbar = mapa_u32(smem_u32(&a_full[buf]), 0);
          if (crank == 0) mbar_arrive_expect_tx(&a_full[buf], 2 * QKV_A_BYTES);
#pragma unroll
          for (int c = 0; c < 4; ++c) tma_load_2d_pair(smem + buf * QKV_A_BYTES + c * 16384, &tmA, lbar, c * 64, r0);
        } else {
          mbar_arrive_expect_tx(&a_full[buf], QKV_A_BYTES);
#pragma unroll
          for (int c = 0; c < 4; ++c) tma_load_2d(smem + buf * QKV_A_BYTES + c * 16384, &tmA, &a_full[buf], c * 64, r0);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1 && crank == 0) {
    // ===================================== MMA issuer (pair: the leader CTA only) ============
    constexpr uint32_t idesc = umma_idesc_f16(128 * CG, 128);
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
          const uint32_t slot = it % NST, use = it / NST;
          mbar_wait(&w_full[slot], use & 1);
          tc_fence_after();
          const uint64_t da0 = umma_desc_sw128(abuf + kc * 16384), db0 = umma_desc_sw128(ring + slot * PIECE);
          if (elect_one()) {
            if constexpr (CG == 2) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_f16_pair(tmem_base + slot4 * 128, da0 + 2 * kk, db0 + 2 * kk, idesc, (kc | kk) != 0);
              umma_commit_pair(&w_empty[slot]);
              if (kc == 3) umma_commit_pair(&s_full[slot4]);
              if (kc == 3 && part == 2) umma_commit_pair(&a_empty[i & 1]);
            } else {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_f16(tmem_base + slot4 * 128, da0 + 2 * kk, db0 + 2 * kk, idesc, (kc | kk) != 0);
              umma_commit(&w_empty[slot]);
              if (kc == 3) umma_commit(&s_full[slot4]);
              if (kc == 3 && part == 2) umma_commit(&a_empty[i & 1]);
            }
          }
          __syncwarp();
        }
        if (p.tl && lane == 0 && i < 5) p.tl[(size_t)blockIdx.x * 128 + i * 16 + 2 + 2 * part] = clock64();
      }
    }
  } else if (warp >= 3) {
    // ===================================== epilogue =========================================
    const int ew = warp - 3;            // 0..7
    const int q4 = warp & 3;            // TMEM lane quarter
    const int ch = ew >> 2;             // column half of a slot: [ch*64, ch*64 + 64)
    const uint32_t st = smem_u32(smem + QKV_OFF_STAGE + ew * GEMM_STAGING_BYTES);
    for (int i = 0; i < nt; ++i) {
      const int r0 = unit_r0(i);
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
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&s_empty[slot4]), 0));   // the leader's MMA warp waits for both CTAs
          else mbar_arrive(&s_empty[slot4]);
        }
        epi_store_h32(st, lane, vbuf[1], dst + 32, 128, rows_valid);
        if (p.tl && ew == 0 && lane == 0 && i < 5) p.tl[(size_t)blockIdx.x * 128 + i * 16 + 9 + 2 * part] = clock64();
      }
    }
    if (nt == 0 && p.pdl_late && lane == 0) pdl_launch_dependents();
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();   // the leader's MMAs read the peer's shared memory until the last commit has completed
  else __syncthreads();
  if (warp == 1) {
    if constexpr (CG == 2) tmem_dealloc_pair<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace mtts
