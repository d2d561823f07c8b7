// Implicit-GEMM on tcgen05 tensor cores for every contraction of the estimator:
// Conv1d k3 (stride 1 and 2), ConvTranspose1d k4 s2, 1x1 convs and all Linear layers.
//
//   D[r, n] = sum_seg  A_seg[r + shift_seg, col_seg : col_seg + 64*nchunks] . W[n, k_seg : ...]
//
// Activations are channels-last fp16 in a FLAT row space (utterance b, frame t) -> row b*Lp + t with
// zero guard rows between utterances, so a conv tap is just a row shift of the same 2-D TMA tensor
// map (out-of-range rows are zero-filled by TMA; guard rows give the per-utterance zero padding) and
// `torch.cat` along channels is a second tensor map in the K loop.  Weights are pre-packed
// [N, Ktot] K-major fp16.  Accumulation is fp32 in TMEM.
//
// Warp roles (352 threads, persistent over tiles, 1 CTA / SM):
//   warp 0   TMA producer of the activation (A) tiles, 128x64 per stage (128B swizzle, mbarrier tx)
//   warp 1   TMEM allocator + single-thread tcgen05.mma issuer (4 x K16 per stage), tcgen05.commit
//   warp 2   TMA producer of the weight (B) tiles, BNx64 per stage.  A cp.async.bulk.tensor costs its
//            issuing thread ~330 cycles whatever the box size (profiles/r01_tma_issue_microbench.txt), so
//            one thread issuing both tiles would cap a 512-cycle stage; the weight producer also never
//            waits for the previous kernel (weights are constants), so its first tiles land during PDL overlap.
//   warps 3-10 epilogue: two warps per TMEM lane quarter, each owning half of the tile's columns;
//             tcgen05.ld (one accumulator row per thread, 32 columns at a time), fused math,
//             swizzled smem staging -> 64-byte-per-row coalesced global stores; residual tiles are
//             prefetched into registers one chunk ahead.  Two accumulator stages in TMEM overlap
//             epilogue(i) with mma(i+1).
// Every launch uses programmatic dependent launch: the prologue (barriers, TMEM alloc, descriptor
// prefetch) runs while the previous kernel drains; griddepcontrol.wait precedes the first global access.
//
// Fused epilogues (reference file:line each one replaces is listed in DESIGN.md):
//   EPI_STATS  +bias, fp16 store, deterministic GroupNorm partial sums per (utterance, group); optionally a
//              second accumulator fed by extra K chunks of the same A tiles (the ResnetBlock1D 1x1 res_conv)
//   EPI_PLAIN  +bias (+residual) (*row mask), fp16 store
//   EPI_FINAL  final 1x1 projection * mask, Euler update of the fp32 state z (channels-first) and
//              refresh of the z channels of the first conv's operand buffer
#pragma once
#include <cuda.h>

#include "ptx.cuh"

namespace mtts {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_NCG = 2;                                   // epilogue column groups per TMEM lane quarter (4 = 16 warps was
                                                              // measured: no faster per tile, and it costs a pipeline stage)
constexpr int GEMM_EPI_WARPS = 4 * GEMM_NCG;                  // 8
constexpr int GEMM_THREADS = 96 + 32 * GEMM_EPI_WARPS;        // 352: A producer, MMA, B producer, 8 epilogue warps
constexpr int GEMM_MAX_SEGS = 9;
constexpr int GEMM_STAGING_BYTES = 32 * 64;                    // per epilogue warp: 32 rows x 32 fp16, swizzled

enum { EPI_STATS = 0, EPI_PLAIN = 1, EPI_FINAL = 5 };
__host__ __device__ constexpr bool epi_has_stats(int e) { return e == EPI_STATS; }

struct GemmSeg {
  int src;        // 0/1: which A tensor map
  int row_shift;  // tap offset in rows
  int col0;       // first column of the source
  int nchunks;    // number of 64-column K chunks
};

struct GemmParams {
  int M;        // output rows
  int n_tiles;  // N / BN
  int num_segs;
  GemmSeg seg[GEMM_MAX_SEGS];
  // common epilogue operands
  const float* bias;   // [N]
  __half* out;         // out[row*ldo + n]
  int ldo;
  const __half* resid;  // resid[row*ldr + n] or null
  int ldr;
  const float* rowmask;  // mask[row*mask_mul + n_tile*mask_nstep] or null
  int mask_mul, mask_nstep;
  const int* rowb;  // utterance id per row, -1 on guard rows
  int Lp;           // rows per utterance incl. guard rows
  // EPI_STATS
  float* stats_part;  // [B][S][16]
  int S;
  // EPI_STATS, optional second GEMM sharing the A tiles (ResnetBlock1D.res_conv next to block1's conv):
  // K chunks [res_chunk0, total) accumulate into a second TMEM accumulator -> res_out = acc1 + res_bias
  int res_chunk0;       // 0 = no second GEMM
  const float* res_bias;
  __half* res_out;      // [row*ldo + n]
  // EPI_FINAL
  float* zout;         // (B, n_valid, T) channels-first fp32
  const float* zbase;  // same layout or null
  float zscale;
  __half* x0;  // first-conv operand buffer rows [row*ldx0 + j] or null
  int ldx0;
  int T;
  int n_valid;
  // debug: per-CTA timeline [gridDim.x][16] (clock64 / globaltimer stamps), null in production
  long long* tl;
  long long* tl2;  // debug: per-CTA per-tile stamps [gridDim.x][64]: tile i -> [4i] first operands seen by the MMA warp, [4i+1] last MMA
                   // issued, [4i+2] accumulator seen by the epilogue, [4i+3] epilogue done (clock64; tools/gemm_tiles.py)
  int w_hint;  // 1: weight (B) tiles are loaded with the L2 evict_last policy
  int a_prefetch;  // 1: L2-prefetch the CTA's first activation tiles before the dependency wait
  int dbg;         // debug experiments (tools/gemm_repeat.py): 1 skip global stores, 2 skip smem staging + stores, 4 skip TMEM loads
  int pdl_late;    // 1: griddepcontrol.launch_dependents when the CTA's last accumulator is complete instead of at entry
  int tap3;        // n = 1 / 2 (256-wide single-CTA conv tiles only): the segments are the -1 / 0 / +1 taps of n sources (seg[t*n + s],
                   //    same columns per source), optionally followed by the res_conv segments (shift 0, one per source, K chunks from
                   //    res_chunk0 on).  tmA0 / tmA1 then have 130-row boxes and one (128 + 2)-row activation tile per K chunk feeds
                   //    all three taps through row-shifted shared-memory descriptors (tools/ubench/rowshift.cu); the res_conv pass
                   //    stages the tiles once more and reads them one row in
  int relu;        // EPI_PLAIN: 1 = ReLU after bias (+ residual), before the row mask (text encoder FFN / duration predictor convs)
  int tma_out;     // 1 (256-wide STATS / PLAIN tiles): the fp16 output tiles leave as 32 x 32 TMA boxes (tmOut; the res_conv half through
                   //    tmRes) straight from the per-warp staging tile, whose layout IS the 64-byte swizzle, instead of the
                   //    shared-memory transpose + 64-byte-per-row st.global: per-tile stamps (tools/conv_tiles.py) showed the epilogue
                   //    -- 2.6-3.2 us per 128 x 256 tile, 4.8 us with the res_conv half -- bounding the tile period once the operands
                   //    arrive fast enough (CTA pairs), and the unit GEMM runs 20 % faster with its st.global removed
  int m_major;     // 1: a CTA owns whole row tiles and walks their N tiles back to back (launch grid <= row tiles):
                   //    the epilogue of one N tile overlaps the main loop of the next even with one row tile per CTA
};

// KSUB = number of 64-column K chunks per pipeline stage.  KSUB = 2 needs the 3-D tensor maps ([nk][rows][64]
// views, box {64, 128, 2}): one 32 KB TMA instruction per operand and stage -- a cp.async.bulk.tensor costs its
// issuing thread ~330 cycles whatever the box size, which would cap the 256-cycle stage of a 128-wide N tile.
// CG = 2: CTA pair (cluster of two, tcgen05 cta_group::2).  The pair computes a 256-row x BN tile; each CTA stages its
// own 128 activation rows and HALF of the weight tile (BN/2 rows), so a stage is 32 KB instead of 48 KB: a third less
// L2 -> shared-memory traffic per FLOP and six stages instead of four in flight (the 4 x 48 KB ring cannot cover the
// TMA latency under load: 192 KB / ~1.5 us < the 185 GB/s one SM's MMA stream consumes).
template <int BN, int EPI = EPI_PLAIN, int KSUB = 1, int CG = 1>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2 * KSUB;
  static constexpr int B_BYTES = (BN / CG) * GEMM_BK * 2 * KSUB;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (CG == 2) ? 6 : ((BN == 256) ? 4 : (KSUB == 2 ? 3 : 4));
  // per-column epilogue parameters of ALL n-tiles, staged once per CTA: [bias | p1 | p2] x PAR_N
  static constexpr int PAR_N = (BN == 128 ? 512 : 1024);  // max N of one launch
  static constexpr int PAR_BYTES = PAR_N * 4;
  static constexpr int RED_BYTES = 0;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + GEMM_EPI_WARPS * GEMM_STAGING_BYTES + PAR_BYTES + RED_BYTES + 256;
  static_assert(TOTAL <= 232448, "exceeds the 227 KB of shared memory one CTA can own");
};

// ---- epilogue helpers -------------------------------------------------------------------------
// Per-warp staging tile: 32 rows x 64 B (32 fp16); 16-byte unit u of row r lives at
// r*64 + ((u ^ ((r >> 1) & 3)) << 4): conflict-free both for "thread = row" accesses and for the
// coalesced pattern (8 rows x 64 B per instruction: lane -> row it*8 + lane/4, unit lane%4).
__device__ __forceinline__ uint32_t epi_st_addr(uint32_t st, int row, int unit) {
  return st + row * 64 + ((unit ^ ((row >> 1) & 3)) << 4);
}
// store 32 fp32 values of "my" row as fp16, coalesced 64 B per row
__device__ __forceinline__ void epi_store_h32(uint32_t st, int lane, const float* v, __half* gtile, int ld,
                                              int rows_valid, int dbg = 0) {
  if (dbg & 2) return;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    sts128(epi_st_addr(st, lane, j), make_uint4(pack_h2(v[8 * j + 0], v[8 * j + 1]), pack_h2(v[8 * j + 2], v[8 * j + 3]),
                                                pack_h2(v[8 * j + 4], v[8 * j + 5]), pack_h2(v[8 * j + 6], v[8 * j + 7])));
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int row = it * 8 + (lane >> 2), unit = lane & 3;
    const uint4 u = lds128(epi_st_addr(st, row, unit));
    if (row < rows_valid && !(dbg & 1)) stg128(gtile + (size_t)row * ld + unit * 8, u);
  }
  __syncwarp();
}
// the same tile handed to TMA: the staging layout is the 64-byte swizzle of a {32 columns, 32 rows} box; rows past the
// tensor's end are clipped.  Lane 0 owns the bulk groups of its warp.
__device__ __forceinline__ void epi_staging_acquire(int lane) {   // the previous TMA store has read the staging tile
  if (lane == 0) tma_store_wait_read<0>();
  __syncwarp();
}
__device__ __forceinline__ void epi_store_h32_tma(uint32_t st, int lane, const float* v, const CUtensorMap* tm, int col, int row0,
                                                  int rows_valid) {
  epi_staging_acquire(lane);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    sts128(epi_st_addr(st, lane, j), make_uint4(pack_h2(v[8 * j + 0], v[8 * j + 1]), pack_h2(v[8 * j + 2], v[8 * j + 3]),
                                                pack_h2(v[8 * j + 4], v[8 * j + 5]), pack_h2(v[8 * j + 6], v[8 * j + 7])));
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0 && rows_valid > 0) {
    tma_store_2d(tm, st, col, row0);
    tma_store_commit();
  }
}
// issue the coalesced loads of a 32x32 fp16 residual tile (consumed later by epi_resid_add)
__device__ __forceinline__ void epi_resid_issue(uint4 (&rr)[4], int lane, const __half* gtile, int ld, int rows_valid) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int row = it * 8 + (lane >> 2), unit = lane & 3;
    rr[it] = make_uint4(0, 0, 0, 0);
    if (row < rows_valid) rr[it] = ldg128(gtile + (size_t)row * ld + unit * 8);
  }
}
__device__ __forceinline__ void epi_resid_add(uint32_t st, int lane, const uint4 (&rr)[4], float* v) {
#pragma unroll
  for (int it = 0; it < 4; ++it) sts128(epi_st_addr(st, it * 8 + (lane >> 2), lane & 3), rr[it]);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 u = lds128(epi_st_addr(st, lane, j));
    float2 f;
    f = unpack_h2(u.x); v[8 * j + 0] += f.x; v[8 * j + 1] += f.y;
    f = unpack_h2(u.y); v[8 * j + 2] += f.x; v[8 * j + 3] += f.y;
    f = unpack_h2(u.z); v[8 * j + 4] += f.x; v[8 * j + 5] += f.y;
    f = unpack_h2(u.w); v[8 * j + 6] += f.x; v[8 * j + 7] += f.y;
  }
  // the next staging writes of this thread target its own row only; cross-row reuse is ordered by
  // the __syncwarp()s inside epi_store_h32
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * GEMM_EPI_WARPS) : "memory"); }

template <int BN, int EPI, int KSUB = 1, int CG = 1>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
               const __grid_constant__ CUtensorMap tmRes, const GemmParams p) {
  using SM = GemmSmem<BN, EPI, KSUB, CG>;
  static_assert(KSUB == 1 || (KSUB == 2 && BN == 128), "two K chunks per stage only for the 128-wide N tile");
  static_assert(CG == 1 || (CG == 2 && BN == 256 && KSUB == 1 && (EPI == EPI_STATS || EPI == EPI_PLAIN)), "CTA pairs: 256-wide conv tiles");
  // CTA pair: rank 0 is the leader (issues the MMAs, owns the full / accumulator-empty barriers both CTAs signal)
  const uint32_t crank = (CG == 2) ? cluster_ctarank() : 0u;
  constexpr int PN = SM::PAR_N;
  constexpr int STAGES = SM::STAGES;
  // accumulator stage stride: the STATS variant may feed a second accumulator (res_conv) per tile; with 128-wide tiles
  // both accumulators of both stages fit (2 x 2 x 128 columns), with 256-wide tiles the dual mode has one stage
  constexpr uint32_t ACC_STRIDE = (epi_has_stats(EPI) && BN == 128) ? 2 * BN : BN;
  constexpr bool DUAL_DOUBLE = (epi_has_stats(EPI) && BN == 128);
  constexpr uint32_t TMEM_COLS = 2 * ACC_STRIDE;  // two accumulator stages
  static_assert(TMEM_COLS == 512 || TMEM_COLS == 256, "BN must be 128 or 256");

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // 128B-swizzled tiles need 1024-byte aligned bases
  uint8_t* staging = smem + STAGES * SM::STAGE_BYTES;
  float* s_par = reinterpret_cast<float*>(staging + GEMM_EPI_WARPS * GEMM_STAGING_BYTES);
  uint8_t* s_red = reinterpret_cast<uint8_t*>(s_par) + SM::PAR_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_red + SM::RED_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES | TAP_B_STAGES] <= 8
  uint64_t* empty_bar = bars + 8;               // [STAGES | TAP_B_STAGES]
  uint64_t* tfull_bar = bars + 16;              // [2]
  uint64_t* tempty_bar = bars + 18;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  // tap-sharing mode (p.tap3): the ring region is re-cut into an activation ring of TAP_A_STAGES x 17 KB (130 rows x 128 B)
  // and a weight ring of TAP_B_STAGES x 32 KB (CTA pair: x 16 KB, this CTA's half of the weight tile); full_bar / empty_bar
  // guard the weight ring, afull / aempty the activation ring.  A pair with tap sharing pulls 17 + 3 x 16 = 65 KB from L2
  // per K chunk and row tile, against 113 KB for the single CTA and 96 KB for the pair with one activation tile per tap.
  constexpr bool TAP3_OK = (BN == 256 && KSUB == 1 && (EPI == EPI_STATS || EPI == EPI_PLAIN));
  constexpr int TAP_A_BYTES = 17 * 1024, TAP_A_STAGES = 3, TAP_B_OFF = 52 * 1024, TAP_B_BYTES = (BN / CG) * GEMM_BK * 2;
  constexpr int TAP_B_STAGES = (CG == 2) ? 8 : STAGES;
  static_assert(!TAP3_OK || (TAP_A_STAGES * TAP_A_BYTES <= TAP_B_OFF && TAP_B_OFF + TAP_B_STAGES * TAP_B_BYTES <= STAGES * SM::STAGE_BYTES),
                "tap-sharing rings must fit the stage ring");
  static_assert(STAGES <= 8 && TAP_B_STAGES <= 8, "barrier block");
  uint64_t* afull = bars + 24;                  // [TAP_A_STAGES]
  uint64_t* aempty = bars + 27;                 // [TAP_A_STAGES]
  const bool tap3 = TAP3_OK && p.tap3 != 0;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // The next kernel may start its prologue (it still waits for our completion).  Early: at once -- best when one solve
  // is in flight.  Late (p.pdl_late): when this CTA's last accumulator is complete -- with several solves in flight the
  // dependents' CTAs would otherwise sit on SM slots (shared memory, TMEM) that another solve's ready kernels could use.
  if (!p.pdl_late) pdl_launch_dependents();
  long long* tl = p.tl ? p.tl + (size_t)blockIdx.x * 16 : nullptr;
  if (tl && threadIdx.x == 0) { tl[0] = clock64(); tl[8] = (long long)globaltimer_ns(); }

  int total_chunks = 0;
  for (int s = 0; s < p.num_segs; ++s) total_chunks += p.seg[s].nchunks;
  const int m_tiles = (p.M + GEMM_BM - 1) / GEMM_BM;
  // a "unit" is what one CTA (CG = 1) or one CTA pair (CG = 2) works on: CG consecutive row tiles x one N tile
  const int m_units = (m_tiles + CG - 1) / CG;
  const int total_tiles = m_units * p.n_tiles;
  const int unit0 = blockIdx.x / CG, nunits = gridDim.x / CG;
  // i-th tile of this CTA (or -1): round-robin over all tiles, or whole row tiles with their N tiles back to back
  auto cta_tile = [&](int i) -> int {
    if (CG == 2 || !p.m_major) { const int t = unit0 + i * nunits; return t < total_tiles ? t : -1; }
    const int mt = blockIdx.x + (i / p.n_tiles) * gridDim.x;
    return mt < m_tiles ? mt * p.n_tiles + (i % p.n_tiles) : -1;
  };
  // first row of this CTA inside unit tile `tile`
  auto tile_r0 = [&](int tile) -> int { return ((tile / p.n_tiles) * CG + (int)crank) * GEMM_BM; };

  if (threadIdx.x == 0) {
    // pair: the leader's full barrier counts the bytes of both CTAs' loads (its own two producers arrive and expect
    // twice their bytes; the peer's TMA instructions complete_tx on it -- the transaction count may run negative inside
    // a phase), its accumulator-empty barrier collects the epilogue warps of both CTAs; empty / accumulator-full
    // barriers are signalled in both CTAs by multicast commits
    for (int i = 0; i < (tap3 ? TAP_B_STAGES : STAGES); ++i) { mbar_init(&full_bar[i], tap3 ? 1 : 2); mbar_init(&empty_bar[i], 1); }
    if (tap3) for (int i = 0; i < TAP_A_STAGES; ++i) { mbar_init(&afull[i], 1); mbar_init(&aempty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], GEMM_EPI_WARPS * CG); }
    fence_mbar_init();
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    if (p.tma_out) { tma_prefetch_desc(&tmOut); tma_prefetch_desc(&tmRes); }
  }
  if (warp == 1) {
    if constexpr (CG == 2) tmem_alloc_pair<TMEM_COLS>(tmem_slot);
    else tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them remotely
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tl && threadIdx.x == 0) tl[1] = clock64();
  // per-column epilogue parameters (weights: independent of the previous kernel) for every n-tile.  Staged AFTER the block
  // barrier and closed by a barrier of the epilogue warps alone: their global-memory round trip (~0.6 us) no longer holds
  // back the producers' first TMA instructions -- with several solves in flight the predecessor has long finished and
  // every cycle of a CTA's prologue is a cycle its SM does nothing else
  if (warp >= 3) {
    const int ncols = min(p.n_tiles * BN, PN);
    for (int i = threadIdx.x - 96; i < ncols; i += 32 * GEMM_EPI_WARPS) {
      s_par[i] = p.bias ? p.bias[i] : 0.f;
      if constexpr (epi_has_stats(EPI)) if (p.res_chunk0 > 0 && i < 256) s_par[256 + i] = p.res_bias[i];  // the conv's own N is 256
    }
    epi_bar_sync();
  }

  if (warp == 2) {
    // ===================================== TMA producer: weights (no dependency wait) ==========
    int stage = 0;
    uint32_t phase = 0;
    const uint64_t pol = l2_policy_evict_last();
    if constexpr (TAP3_OK) if (tap3) {   // chunk-major: the three tap tiles of a K chunk follow each other; then the res_conv chunks
      const int nsrc = p.tap3;
      int CH = 0;
      for (int q = 0; q < nsrc; ++q) CH += p.seg[q].nchunks;
      const bool has_res = p.res_chunk0 > 0;
      auto put = [&](int kc, int n0) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sb = smem + TAP_B_OFF + stage * TAP_B_BYTES;
          if constexpr (CG == 2) {   // this CTA's half of the weight tile; both halves are counted on the leader's barrier
            const uint32_t lbar = mapa_u32(smem_u32(&full_bar[stage]), 0);
            if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * TAP_B_BYTES);
            if (p.w_hint) tma_load_2d_pair_hint(sb, &tmB, lbar, kc * GEMM_BK, n0 + (int)crank * (BN / 2), pol);
            else tma_load_2d_pair(sb, &tmB, lbar, kc * GEMM_BK, n0 + (int)crank * (BN / 2));
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], TAP_B_BYTES);
            if (p.w_hint) tma_load_2d_hint(sb, &tmB, &full_bar[stage], kc * GEMM_BK, n0, pol);
            else tma_load_2d(sb, &tmB, &full_bar[stage], kc * GEMM_BK, n0);
          }
        }
        __syncwarp();
        if (++stage == TAP_B_STAGES) { stage = 0; phase ^= 1; }
      };
      for (int ti = 0, tile; (tile = cta_tile(ti)) >= 0; ++ti) {
        const int n0 = (tile % p.n_tiles) * BN;
        for (int c = 0; c < CH; ++c)
          for (int t = (p.dbg & 8) ? 2 : 0; t < 3; ++t) put(t * CH + c, n0);
        if (has_res)
          for (int c = 0; c < CH; ++c) put(p.res_chunk0 + c, n0);
      }
    }
    if (!tap3)
    for (int ti = 0, tile; (tile = cta_tile(ti)) >= 0; ++ti) {
      const int n0 = (tile % p.n_tiles) * BN;
      for (int kc = 0; kc < total_chunks; kc += KSUB) {
        mbar_wait(&empty_bar[stage], phase ^ 1);   // converged warp; one elected lane issues
        if (elect_one()) {
          uint8_t* sb = smem + stage * SM::STAGE_BYTES + SM::A_BYTES;
          if constexpr (CG == 2) {   // this CTA's half of the weight tile; bytes are counted on the leader's barrier
            const uint32_t lbar = mapa_u32(smem_u32(&full_bar[stage]), 0);
            if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * SM::B_BYTES);
            if (p.w_hint) tma_load_2d_pair_hint(sb, &tmB, lbar, kc * GEMM_BK, n0 + (int)crank * (BN / 2), pol);
            else tma_load_2d_pair(sb, &tmB, lbar, kc * GEMM_BK, n0 + (int)crank * (BN / 2));
          } else {
          mbar_arrive_expect_tx(&full_bar[stage], SM::B_BYTES);
          if constexpr (KSUB == 1) {
            if (p.w_hint) tma_load_2d_hint(sb, &tmB, &full_bar[stage], kc * GEMM_BK, n0, pol);
            else tma_load_2d(sb, &tmB, &full_bar[stage], kc * GEMM_BK, n0);
          } else {
            if (p.w_hint) tma_load_3d_hint(sb, &tmB, &full_bar[stage], 0, n0, kc, pol);
            else tma_load_3d(sb, &tmB, &full_bar[stage], 0, n0, kc);
          }
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  }

  if (KSUB == 1 && warp == 0 && cta_tile(0) >= 0 && p.a_prefetch) {
    // warm the TLB / L2 path of this CTA's first activation tiles while the previous kernel drains
    if (elect_one()) {
      const int r0 = tile_r0(cta_tile(0));
      int kc = 0;
      for (int s = 0; s < p.num_segs && kc < STAGES; ++s) {
        const GemmSeg sg = p.seg[s];
        for (int c = 0; c < sg.nchunks && kc < STAGES; ++c, ++kc)
          tma_prefetch_2d(sg.src ? &tmA1 : &tmA0, sg.col0 + c * GEMM_BK, r0 + sg.row_shift);
      }
    }
    __syncwarp();
  }

  pdl_wait();  // everything below touches memory the previous kernel may still be using
  if (tl && threadIdx.x == 0) { tl[2] = clock64(); tl[9] = (long long)globaltimer_ns(); }

  if (warp == 0) {
    // ===================================== TMA producer: activations ===========================
    int stage = 0;
    uint32_t phase = 0;
    if constexpr (TAP3_OK) if (tap3) {   // one 130-row tile (rows r0 - 1 .. r0 + 128) per K chunk and pass
      const int nsrc = p.tap3;
      const int passes = p.res_chunk0 > 0 ? 2 : 1;
      for (int ti = 0, tile; (tile = cta_tile(ti)) >= 0; ++ti) {
        const int r0 = tile_r0(tile);
        for (int pass = 0; pass < passes; ++pass)
          for (int q = 0; q < nsrc; ++q) {
            const GemmSeg sg = p.seg[q];
            const CUtensorMap* tm = q ? &tmA1 : &tmA0;
            for (int c = 0; c < sg.nchunks; ++c) {
              mbar_wait(&aempty[stage], phase ^ 1);
              if (elect_one()) {
                if constexpr (CG == 2) {   // both CTAs' tiles are counted on the leader's barrier
                  if (crank == 0) mbar_arrive_expect_tx(&afull[stage], 2 * 130 * GEMM_BK * 2);
                  tma_load_2d_pair(smem + stage * TAP_A_BYTES, tm, mapa_u32(smem_u32(&afull[stage]), 0), sg.col0 + c * GEMM_BK, r0 - 1);
                } else {
                  mbar_arrive_expect_tx(&afull[stage], 130 * GEMM_BK * 2);
                  tma_load_2d(smem + stage * TAP_A_BYTES, tm, &afull[stage], sg.col0 + c * GEMM_BK, r0 - 1);
                }
              }
              __syncwarp();
              if (++stage == TAP_A_STAGES) { stage = 0; phase ^= 1; }
            }
          }
      }
    }
    if (!tap3)
    for (int ti = 0, tile; (tile = cta_tile(ti)) >= 0; ++ti) {
      const int r0 = tile_r0(tile);
      for (int s = 0; s < p.num_segs; ++s) {
        const GemmSeg sg = p.seg[s];
        const CUtensorMap* tm = sg.src ? &tmA1 : &tmA0;
        for (int c = 0; c < sg.nchunks; c += KSUB) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            if constexpr (CG == 2) {
              const uint32_t lbar = mapa_u32(smem_u32(&full_bar[stage]), 0);
              if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * SM::A_BYTES);
              tma_load_2d_pair(smem + stage * SM::STAGE_BYTES, tm, lbar, sg.col0 + c * GEMM_BK, r0 + sg.row_shift);
            } else {
            mbar_arrive_expect_tx(&full_bar[stage], SM::A_BYTES);
            if constexpr (KSUB == 1)
              tma_load_2d(smem + stage * SM::STAGE_BYTES, tm, &full_bar[stage], sg.col0 + c * GEMM_BK, r0 + sg.row_shift);
            else
              tma_load_3d(smem + stage * SM::STAGE_BYTES, tm, &full_bar[stage], 0, r0 + sg.row_shift, sg.col0 / GEMM_BK + c);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && crank == 0) {
    // ===================================== MMA issuer (pair: the leader CTA only) ============
    constexpr uint32_t idesc = umma_idesc_f16(GEMM_BM * CG, BN);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    if constexpr (TAP3_OK) if (tap3) {
      // tap-sharing: a K chunk of the activations is staged once (rows r0 - 1 .. r0 + 128); tap t reads it through a
      // descriptor that starts t rows (t * 128 B) into the 128B-swizzled tile -- the swizzle is a function of the
      // shared-memory address bits, so the shifted view addresses exactly the rows TMA wrote.  With a res_conv the
      // chunks are staged a second time and read one row in (shift 0) into the second accumulator half; the two halves
      // are handed over separately like in the tap-by-tap order (barrier pairs [0] / [1]).
      int CH = 0;
      for (int q = 0; q < p.tap3; ++q) CH += p.seg[q].nchunks;
      const bool has_res = p.res_chunk0 > 0;
      int sa_i = 0;
      uint32_t sa_ph = 0;
      for (int ti = 0, tile; (tile = cta_tile(ti)) >= 0; ++ti) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * ACC_STRIDE;
        for (int c = 0; c < CH; ++c) {
          mbar_wait(&afull[sa_i], sa_ph);
          if (p.tl2 && lane == 0 && c == 0 && ti < 16) p.tl2[(size_t)blockIdx.x * 64 + 4 * ti] = clock64();
          for (int t = (p.dbg & 8) ? 2 : 0; t < 3; ++t) {   // dbg 8: one tap only (timing experiment, wrong results)
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint64_t da = umma_desc_sw128(smem_u32(smem + sa_i * TAP_A_BYTES) + t * (GEMM_BK * 2));
            const uint64_t db = umma_desc_sw128(smem_u32(smem + TAP_B_OFF + stage * TAP_B_BYTES));
            if (elect_one()) {
              if constexpr (CG == 2) {
#pragma unroll
                for (int k = 0; k < GEMM_BK / 16; ++k) umma_f16_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, (c | (t != ((p.dbg & 8) ? 2 : 0)) | k) != 0);
                umma_commit_pair(&empty_bar[stage]);
                if (t == 2) umma_commit_pair(&aempty[sa_i]);
                if (t == 2 && c + 1 == CH) umma_commit_pair(&tfull_bar[has_res ? 0 : as]);
              } else {
#pragma unroll
                for (int k = 0; k < GEMM_BK / 16; ++k) umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (c | (t != ((p.dbg & 8) ? 2 : 0)) | k) != 0);
                umma_commit(&empty_bar[stage]);
                if (t == 2) umma_commit(&aempty[sa_i]);
                if (t == 2 && c + 1 == CH) umma_commit(&tfull_bar[has_res ? 0 : as]);
              }
            }
            __syncwarp();
            if (++stage == TAP_B_STAGES) { stage = 0; phase ^= 1; }
          }
          if (++sa_i == TAP_A_STAGES) { sa_i = 0; sa_ph ^= 1; }
        }
        if (has_res) {
          mbar_wait(&tempty_bar[1], aphase ^ 1);   // the previous tile's res half has been drained
          tc_fence_after();
          for (int c = 0; c < CH; ++c) {
            mbar_wait(&afull[sa_i], sa_ph);
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint64_t da = umma_desc_sw128(smem_u32(smem + sa_i * TAP_A_BYTES) + GEMM_BK * 2);   // shift 0 = one row in
            const uint64_t db = umma_desc_sw128(smem_u32(smem + TAP_B_OFF + stage * TAP_B_BYTES));
            if (elect_one()) {
              if constexpr (CG == 2) {
#pragma unroll
                for (int k = 0; k < GEMM_BK / 16; ++k) umma_f16_pair(d_tmem + BN, da + 2 * k, db + 2 * k, idesc, (c | k) != 0);
                umma_commit_pair(&empty_bar[stage]);
                umma_commit_pair(&aempty[sa_i]);
                if (c + 1 == CH) umma_commit_pair(&tfull_bar[1]);
              } else {
#pragma unroll
                for (int k = 0; k < GEMM_BK / 16; ++k) umma_f16(d_tmem + BN, da + 2 * k, db + 2 * k, idesc, (c | k) != 0);
                umma_commit(&empty_bar[stage]);
                umma_commit(&aempty[sa_i]);
                if (c + 1 == CH) umma_commit(&tfull_bar[1]);
              }
            }
            __syncwarp();
            if (++stage == TAP_B_STAGES) { stage = 0; phase ^= 1; }
            if (++sa_i == TAP_A_STAGES) { sa_i = 0; sa_ph ^= 1; }
          }
        }
        if (p.tl2 && lane == 0 && ti < 16) p.tl2[(size_t)blockIdx.x * 64 + 4 * ti + 1] = clock64();
        if (has_res) { aphase ^= 1; }   // dual accumulator: one TMEM stage
        else { as ^= 1; if (as == 0) aphase ^= 1; }
      }
    }
    if (!tap3)
    for (int ti = 0, tile; (tile = cta_tile(ti)) >= 0; ++ti) {
      // converged warp: waits by every lane, tcgen05 instructions by one elected lane (uniform operands)
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      tc_fence_after();
      const int rc0 = (epi_has_stats(EPI) && p.res_chunk0 > 0) ? p.res_chunk0 : total_chunks;  // dual: both TMEM halves
      // dual accumulators in one 256-wide TMEM stage: the two halves are handed over separately -- barrier pair [0] for
      // the conv accumulator (complete after chunk rc0-1, free once y / the statistics are out), pair [1] for the
      // res_conv accumulator (complete at the end, free once res is out) -- so the epilogue drains the conv half while
      // the res chunks run and the next tile's conv chunks run while the res half drains
      const bool dual_split = !DUAL_DOUBLE && rc0 < total_chunks;
      for (int kc = 0; kc < total_chunks; kc += KSUB) {
        const uint32_t d_tmem = tmem_base + as * ACC_STRIDE + ((kc >= rc0) ? BN : 0);
        const int kfirst = (kc >= rc0) ? rc0 : 0;
        if (dual_split && kc == rc0) { mbar_wait(&tempty_bar[1], aphase ^ 1); tc_fence_after(); }
        mbar_wait(&full_bar[stage], phase);
        if (tl && lane == 0 && kc == 0 && ti == 0) tl[3] = clock64();
        if (p.tl2 && lane == 0 && kc == 0 && ti < 16) p.tl2[(size_t)blockIdx.x * 64 + 4 * ti] = clock64();
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * SM::STAGE_BYTES);
        const uint64_t da = umma_desc_sw128(sa);
        const uint64_t db = umma_desc_sw128(sa + SM::A_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int sub = 0; sub < KSUB; ++sub)   // K sub-chunk tiles are consecutive in the stage; descriptor address in 16-B units
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k)
              if constexpr (CG == 2) umma_f16_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, ((kc - kfirst) | k) != 0);
              else
              umma_f16(d_tmem, da + sub * ((GEMM_BM * GEMM_BK * 2) >> 4) + 2 * k, db + sub * ((BN * GEMM_BK * 2) >> 4) + 2 * k, idesc,
                       ((kc - kfirst) | sub | k) != 0);
          if constexpr (CG == 2) {
            umma_commit_pair(&empty_bar[stage]);
            if (dual_split && kc + KSUB == rc0) umma_commit_pair(&tfull_bar[0]);
            if (kc + KSUB >= total_chunks) umma_commit_pair(&tfull_bar[dual_split ? 1 : as]);
          } else {
          umma_commit(&empty_bar[stage]);
          if (dual_split && kc + KSUB == rc0) umma_commit(&tfull_bar[0]);
          if (kc + KSUB >= total_chunks) umma_commit(&tfull_bar[dual_split ? 1 : as]);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (tl && lane == 0) tl[4] = clock64();
      if (p.tl2 && lane == 0 && ti < 16) p.tl2[(size_t)blockIdx.x * 64 + 4 * ti + 1] = clock64();
      if (!DUAL_DOUBLE && rc0 < total_chunks) { aphase ^= 1; }   // dual accumulator with 256-wide tiles: one TMEM stage
      else { as ^= 1; if (as == 0) aphase ^= 1; }
    }
  } else if (warp >= 3) {
    // ===================================== epilogue =========================================
    constexpr int CW = BN / GEMM_NCG;  // columns per epilogue warp
    constexpr int NCH = CW / 32;  // 32-column chunks per warp
    const int ew = warp - 3;
    const int q = warp & 3;       // TMEM lane quarter this warp may access
    const int hcol = ew >> 2;     // which column group of the tile
    const int cbase = hcol * CW;
    const uint32_t st = smem_u32(staging + ew * GEMM_STAGING_BYTES);
    const uint32_t spar = smem_u32(s_par);
    int as = 0;
    uint32_t aphase = 0;
    for (int ti = 0, tile; (tile = cta_tile(ti)) >= 0; ++ti) {
      const int n_tile = tile % p.n_tiles;
      const int r0 = tile_r0(tile);
      const int n0 = n_tile * BN;
      const int rw0 = r0 + q * 32;       // first row of this warp
      const int row = rw0 + lane;        // my row
      const int rows_valid = min(32, p.M - rw0);  // may be <= 0
      const bool row_ok = row < p.M;
      const bool last_tile = p.pdl_late && cta_tile(ti + 1) < 0;

      const uint32_t sp0 = spar + (n0 + cbase) * 4;  // bias of this warp's first column
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + as * ACC_STRIDE + cbase;

      if constexpr (EPI == EPI_STATS || EPI == EPI_PLAIN) {
        const bool tma_out = (BN == 256) && p.tma_out != 0;
        const bool has_res = (EPI == EPI_PLAIN && p.resid != nullptr);
        const __half* rbase = has_res ? p.resid + (size_t)rw0 * p.ldr + n0 + cbase : nullptr;
        uint4 rr[4];
        if (has_res) epi_resid_issue(rr, lane, rbase, p.ldr, rows_valid);
        float mrow = 1.f;
        if constexpr (EPI == EPI_PLAIN) if (p.rowmask && row_ok) mrow = p.rowmask[(size_t)row * p.mask_mul + (n0 >> 8) * p.mask_nstep];  // ConvT: one output phase per 256 columns
        int myb = -1;
        if constexpr (EPI == EPI_STATS) if (row_ok) myb = p.rowb[row];

        if (lane == 0) { mbar_wait(&tfull_bar[as], aphase); if (last_tile) pdl_launch_dependents(); }
        __syncwarp();
        tc_fence_after();
        if (tl && ew == 0 && lane == 0 && ti == 0) tl[5] = clock64();
        if (p.tl2 && ew == 0 && lane == 0 && ti < 16) p.tl2[(size_t)blockIdx.x * 64 + 4 * ti + 2] = clock64();

        float gs[(EPI == EPI_STATS) ? 2 * NCH : 1];
        __half* obase = p.out + (size_t)rw0 * p.ldo + n0 + cbase;
        float vbuf[2][32];  // accumulator chunk c+1 is fetched from TMEM while chunk c is processed
        if (!(p.dbg & 4)) tmem_ld32(taddr, vbuf[0]);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          float* v = vbuf[c & 1];
          tmem_ld_wait();
          if (tl && ew == 0 && lane == 0 && c == 0 && ti == 0) tl[11] = clock64();
          if (c + 1 < NCH && !(p.dbg & 4)) tmem_ld32(taddr + (c + 1) * 32, vbuf[(c + 1) & 1]);
          if (c + 1 == NCH && !(DUAL_DOUBLE && epi_has_stats(EPI) && p.res_chunk0 > 0)) {
            // this warp's part of the accumulator is in registers: hand the TMEM stage (dual mode: the conv half) back now, not
            // after the arithmetic, the statistics and the stores of the last chunk -- in dual mode the next tile's conv
            // chunks wait for exactly this arrive (tools/conv_tiles.py: tile period 5.8 us for 3.8 us of MMAs)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              uint64_t* eb = &tempty_bar[(!DUAL_DOUBLE && epi_has_stats(EPI) && p.res_chunk0 > 0) ? 0 : as];
              if constexpr (CG == 2) mbar_arrive_cluster(mapa_u32(smem_u32(eb), 0));
              else mbar_arrive(eb);
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = lds_f4(sp0 + (c * 32 + 4 * j) * 4);
            v[4 * j + 0] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
          }
          if (has_res) {
            if (tma_out) epi_staging_acquire(lane);
            epi_resid_add(st, lane, rr, v);
            if (c + 1 < NCH) epi_resid_issue(rr, lane, rbase + (c + 1) * 32, p.ldr, rows_valid);
          }
          if constexpr (EPI == EPI_STATS) {  // one 32-column chunk == one GroupNorm group
            float a0 = 0.f, b0 = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) { a0 += v[j]; b0 = fmaf(v[j], v[j], b0); }
            gs[2 * c] = a0; gs[2 * c + 1] = b0;
          }
          if constexpr (EPI == EPI_PLAIN) {
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = (mrow == 0.f) ? 0.f : v[j] * mrow;  // never NaN * 0 on guard rows
          }
          if (tma_out) epi_store_h32_tma(st, lane, v, &tmOut, n0 + cbase + c * 32, rw0, rows_valid);
          else epi_store_h32(st, lane, v, obase + c * 32, p.ldo, rows_valid, p.dbg);
          if (tl && ew == 0 && lane == 0 && ti == 0) tl[12 + (c != 0)] = clock64();
        }
        if constexpr (EPI == EPI_STATS) {
          // deterministic per-(utterance, group) partial sums of this warp's 32 rows x 4 groups:
          // slot = 32-row block index relative to the utterance's first block, 16 floats per slot
          // (group g -> [2g] sum, [2g+1] sum of squares); this warp owns groups hcol*4 .. hcol*4+3.
          const int wb = rw0 >> 5;
          const int b0 = __shfl_sync(0xffffffffu, myb, 0);
          if (__all_sync(0xffffffffu, myb == b0)) {
            if (b0 >= 0) {  // uniform
#pragma unroll
              for (int j = 0; j < 2 * NCH; ++j) {
                float x = gs[j];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
                gs[j] = x;
              }
              if (lane == 0) {
                float* dst = p.stats_part + ((size_t)b0 * p.S + (wb - ((b0 * p.Lp) >> 5))) * 16 + (((n0 + cbase) >> 5) << 1);
#pragma unroll
                for (int j = 0; j < 2 * NCH; ++j) dst[j] = gs[j];
              }
            }
          } else {  // utterance boundary inside the warp's rows: serial, fixed order
            if (tma_out) epi_staging_acquire(lane);
            const uint32_t sf = st;                 // [32][9] floats
            const uint32_t sb = st + 32 * 9 * 4;    // [32] ints
#pragma unroll
            for (int j = 0; j < 2 * NCH; ++j) sts_f32(sf + (lane * 9 + j) * 4, gs[j]);
            sts_u32(sb + lane * 4, (uint32_t)myb);
            __syncwarp();
            if (lane < 2 * NCH) {
              int cur = -1;
              float acc = 0.f;
              for (int i = 0; i < 32; ++i) {
                const int bi = (int)lds_u32(sb + i * 4);
                if (bi != cur) {
                  if (cur >= 0) p.stats_part[((size_t)cur * p.S + (wb - ((cur * p.Lp) >> 5))) * 16 + (((n0 + cbase) >> 5) << 1) + lane] = acc;
                  cur = bi;
                  acc = 0.f;
                }
                acc += lds_f32(sf + (i * 9 + lane) * 4);
              }
              if (cur >= 0) p.stats_part[((size_t)cur * p.S + (wb - ((cur * p.Lp) >> 5))) * 16 + (((n0 + cbase) >> 5) << 1) + lane] = acc;
            }
            __syncwarp();
          }
        }
        if (tl && ew == 0 && lane == 0 && ti == 0) tl[14] = clock64();
        if constexpr (EPI == EPI_STATS) {
          if (p.res_chunk0 > 0) {  // second accumulator: res = acc1 + res_bias (no statistics, no mask)
            if constexpr (!DUAL_DOUBLE) {   // the conv half went back above; wait for the res half (see the MMA warp)
              if (lane == 0) mbar_wait(&tfull_bar[1], aphase);
              __syncwarp();
              tc_fence_after();
            }
            __half* rob = p.res_out + (size_t)rw0 * p.ldo + n0 + cbase;
            tmem_ld32(taddr + BN, vbuf[0]);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
              float* v = vbuf[c & 1];
              tmem_ld_wait();
              if (c + 1 < NCH) tmem_ld32(taddr + BN + (c + 1) * 32, vbuf[(c + 1) & 1]);
              if (c + 1 == NCH) {   // the res half is in registers: hand it back
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                  if constexpr (CG == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[DUAL_DOUBLE ? as : 1]), 0));
                  else mbar_arrive(&tempty_bar[DUAL_DOUBLE ? as : 1]);
                }
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b4 = lds_f4(spar + (256 + n0 + cbase + c * 32 + 4 * j) * 4);
                v[4 * j + 0] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
              }
              if (tma_out) epi_store_h32_tma(st, lane, v, &tmRes, n0 + cbase + c * 32, rw0, rows_valid);
              else epi_store_h32(st, lane, v, rob + c * 32, p.ldo, rows_valid);
            }
          }
        }
      } else {  // EPI_FINAL
        const int b = row_ok ? p.rowb[row] : -1;
        const int t = row - b * p.Lp;
        const float m = (b >= 0) ? p.rowmask[row] : 0.f;
        // the state z of this row's channels is requested BEFORE the accumulator wait (it does not depend on the MMAs): one
        // memory round trip under the main loop instead of one per 32-channel chunk after it.  All state loads precede the
        // first store (zout may alias zbase; every element is read and written by the same thread).
        float zb[NCH][32];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int col0 = cbase + c * 32;
          const int nj = min(32, p.n_valid - col0);  // warp-uniform
          const size_t idx0 = ((size_t)max(b, 0) * p.n_valid + col0) * p.T + t;
#pragma unroll
          for (int j = 0; j < 32; ++j) zb[c][j] = (b >= 0 && p.zbase != nullptr && j < nj) ? p.zbase[idx0 + (size_t)j * p.T] : 0.f;
        }
        if (lane == 0) { mbar_wait(&tfull_bar[as], aphase); if (last_tile) pdl_launch_dependents(); }
        __syncwarp();
        tc_fence_after();
        if (tl && ew == 0 && lane == 0 && ti == 0) tl[5] = clock64();
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int col0 = cbase + c * 32;
          const int nj = min(32, p.n_valid - col0);  // warp-uniform
          if (nj > 0) {
            float v[32];
            tmem_ld32(taddr + c * 32, v);
            const size_t idx0 = ((size_t)max(b, 0) * p.n_valid + col0) * p.T + t;
            tmem_ld_wait();
            if (b >= 0) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float o = fmaf(p.zscale, (v[j] + lds_f32(sp0 + (c * 32 + j) * 4)) * m, zb[c][j]);
                v[j] = o;
              }
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nj) p.zout[idx0 + (size_t)j * p.T] = v[j];
              if (p.x0) {
                uint4* xd = reinterpret_cast<uint4*>(p.x0 + (size_t)row * p.ldx0 + col0);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (8 * j < nj)
                    xd[j] = make_uint4(pack_h2(v[8 * j] * m, v[8 * j + 1] * m), pack_h2(v[8 * j + 2] * m, v[8 * j + 3] * m),
                                       pack_h2(v[8 * j + 4] * m, v[8 * j + 5] * m), pack_h2(v[8 * j + 6] * m, v[8 * j + 7] * m));
              }
            }
          }
        }
      }

      if (tl && ew == 0 && lane == 0) tl[6] = clock64();
      if (p.tl2 && ew == 0 && lane == 0 && ti < 16) p.tl2[(size_t)blockIdx.x * 64 + 4 * ti + 3] = clock64();
      // release the accumulator stage back to the MMA warp (STATS / PLAIN did it as soon as the accumulator was in registers)
      const bool dual_split = !DUAL_DOUBLE && epi_has_stats(EPI) && p.res_chunk0 > 0;
      if constexpr (EPI != EPI_STATS && EPI != EPI_PLAIN) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          uint64_t* eb = &tempty_bar[as];
          if constexpr (CG == 2) mbar_arrive_cluster(mapa_u32(smem_u32(eb), 0));   // the leader's MMA warp waits for both CTAs
          else mbar_arrive(eb);
        }
      }
      if (dual_split) { aphase ^= 1; }   // dual accumulator, 256-wide: single stage
      else { as ^= 1; if (as == 0) aphase ^= 1; }
    }
    if constexpr (BN == 256 && (EPI == EPI_STATS || EPI == EPI_PLAIN))
      if (p.tma_out && lane == 0) tma_store_wait_read<0>();   // TMA has read the staging tiles before the CTA (and its shared memory) retires;
                                                              // the writes themselves are ordered by kernel completion
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();   // the leader's MMAs read the peer's shared memory until the last commit
  else __syncthreads();
  if (warp == 1) {
    if constexpr (CG == 2) tmem_dealloc_pair<TMEM_COLS>(tmem_base);
    else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
  if (tl && threadIdx.x == 0) { tl[7] = clock64(); tl[10] = (long long)globaltimer_ns(); }
}

}  // namespace mtts
