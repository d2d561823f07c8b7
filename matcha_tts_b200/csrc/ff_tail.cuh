// Fused "transformer tail" of BasicTransformerBlock (reference model.py:737-744, FeedForward :641-644,
// SnakeBeta :600-609) for one 128-row tile, everything after the attention product being row-local:
//
//   x_a = x_r + o Wo^T + b_o                      (attn1.to_out + residual, :703, :737)
//   c   = LayerNorm3(x_a)                         (:739)
//   s   = SnakeBeta(c W1^T + b1)                  (ff.net.0)
//   out = (x_a + s W2^T + b2) * mask              (ff.net.2 + residual :742; consumers mask the block output)
//
// One CTA per SM, persistent over row tiles.  Neither c (128x256) nor s (128x1024) ever leaves the SM:
//   TMEM  cols [0,256)   R : to_out accumulator -> x_a (fp32, written back by the epilogue) -> FF2 accumulates on top
//         cols [256,384) D1 : FF1 accumulator, one 128-wide chunk of the hidden dimension (single buffer: the epilogue
//                             moves it to registers at once, then FF1 of the next chunk may overwrite it)
//         cols [384,512) Cc : c = LayerNorm3(x_a) as packed fp16 pairs -- the A operand of FF1 is read from TENSOR
//                             MEMORY.  The tensor-memory A read costs ~64 cycles per K16 step whatever N is (measured:
//                             16 N=64 steps took 0.5 us, not 0.26), so FF1 uses N=128 steps (64 cycles each anyway)
//   SMEM  S  2x32 KB : s chunk j (128x128 fp16 as two 128B-swizzled 128x64 K tiles), two buffers (A operand of FF2)
//         ring 4x32 KB : TMA-fed operand pieces, ONE cp.async.bulk.tensor each -- an issue costs the thread
//                        ~330 cycles whatever the box size (profiles/r01_tma_issue_microbench.txt), so pieces
//                        are as large as a 512-cycle MMA group needs:
//                          o  tile    box {64, 128 rows, 2 K-chunks} of o  viewed as [2][rows][64]
//                          Wo K-chunk box {64, 256 rows}                 (B operand, N = 256)
//                          W1 chunk j, half h: box {64, 128 rows, 2 K-chunks} of W1 viewed as [4][1024][64] (N = 128, K = 128)
//                          W2 chunk j, half h: box {64, 256 rows}        (B operand, N = 256, K = 64)
// Warp roles: warp 0 TMA producer, warp 1 issues to_out + FF1, warp 2 issues FF2 (two issuers: the tcgen05 queue is
// short, so every barrier wait / commit of a single issuer idles the tensor pipe), warps 3-18 epilogue (four column
// groups per TMEM lane quarter).  The MMA stream is software-pipelined over the 8 hidden chunks
//   FF1_0 | FF1_1 FF2_0 | FF1_2 FF2_1 | ... | FF1_7 FF2_6 | FF2_7
// so that the tensor pipe works on FF1_{j+1} and FF2_{j-1} while the epilogue warps apply SnakeBeta to chunk j.
//
// CG = 2 (the default) is the CTA-pair variant (same maths; its own TMEM / shared-memory plan, see TAIL2_* below): a cluster of two CTAs works on two
// consecutive 128-row tiles with tcgen05 cta_group::2 MMAs (M = 256).  Each CTA stages only ITS HALF of every weight
// piece (Wo / W2: 128 of the 256 output rows, W1: 64 of the chunk's 128 hidden units; the tensor maps have half-size
// boxes), so the shared-memory traffic of the FF loop -- the bound of the single-CTA kernel --
// drops from ~320 KB to ~192 KB per hidden chunk.  The pair leader (cluster rank 0) issues every MMA; its ring "full"
// barriers count the bytes of both CTAs' TMA loads (the peer's loads complete_tx on the leader's barrier), ring "empty"
// and every MMA -> epilogue barrier is signalled in both CTAs by multicast commits, and every epilogue -> MMA barrier
// lives in the leader and collects the epilogue warps of both CTAs (the peer arrives through its shared::cluster address).
#pragma once
#include <cuda.h>

#include "gemm_tc.cuh"
#include "ptx.cuh"

namespace mtts {

struct TailParams {
  int M;                 // rows (flat row space of the level)
  const __half* xr;      // [rows, 256] resnet output (residual)
  const float* b_o;      // [256]
  const float* ln_g;     // [256]
  const float* ln_b;     // [256]
  const float* b1;       // [1024]
  const float* sn_a;     // [1024] exp(alpha)
  const float* sn_ib;    // [1024] 1/(exp(beta)+1e-9)
  const float* b2;       // [256]
  const float* rowmask;  // [rows]
  __half* out;           // [rows, 256]
  int w_hint;            // 1: weight pieces are loaded with the L2 evict_last policy
  int pdl_late;          // 1: griddepcontrol.launch_dependents at the last tile's final epilogue instead of at entry
  long long* tl;         // debug timeline [gridDim.x][128] clock64 stamps of the first tile (null in production)
};

constexpr int TAIL_NST = 4;
constexpr int TAIL_PIECE = 32768;
constexpr int TAIL_NSB = 2;                                           // s chunk buffers
constexpr int TAIL_SBYTES = 32768;                                    // one s chunk: 128 rows x 128 fp16
constexpr int TAIL_NJ = 8;                                            // hidden chunks of 128
constexpr int TAIL_NCG = 4;                                           // column groups -> 16 epilogue warps (4 per scheduler)
constexpr int TAIL_THREADS = 96 + 128 * TAIL_NCG;                      // producer, two MMA issuers, 16 epilogue warps
constexpr int TAIL_OFF_S = 0;                                         // 2 x 32 KB
constexpr int TAIL_OFF_RING = TAIL_OFF_S + TAIL_NSB * TAIL_SBYTES;    // 65536
constexpr int TAIL_OFF_PAR = TAIL_OFF_RING + TAIL_NST * TAIL_PIECE;   // 196608
constexpr int TAIL_PAR_FLOATS = 4 * 256 + 3 * 1024;                   // b_o ln_g ln_b b2 | b1 sn_a sn_ib
// epilogue staging (16 warps x 32 rows x 64 B = 32 KB) aliases the S buffers: it is only used at the start (E1)
// and at the end (E3) of a tile, when no FF2 MMA can be reading S (r_full / r_done imply all MMAs retired)
constexpr int TAIL_OFF_STAGING = TAIL_OFF_S;
constexpr int TAIL_OFF_RED = TAIL_OFF_PAR + TAIL_PAR_FLOATS * 4;      // LayerNorm partials [128][NCG] float2
constexpr int TAIL_OFF_BAR = TAIL_OFF_RED + 128 * TAIL_NCG * 8;
constexpr int TAIL_SMEM = TAIL_OFF_BAR + 256;
static_assert(TAIL_SMEM <= 232448, "exceeds the 227 KB of shared memory one CTA can own");
// CTA pairs (CG = 2) stage half of every weight piece, so their ring needs 4 x 16 KB, not 4 x 32 KB.  The 64 KB that frees
// hold c = LayerNorm3(x_a) in SHARED memory (four 128B-swizzled K tiles, the A operand of FF1 like any other): an MMA whose
// A operand comes from tensor memory costs ~40 cycles more than its 64-cycle floor (104 cycles per K16 step at N = 128) --
// and the 128 tensor-memory columns c occupied become a SECOND FF1 accumulator, so FF1 of chunk j + 1 no longer waits for
// the epilogue to pick up chunk j.  The `o` tile of the next row tile lands in the first half of the same region (it is
// dead once to_out has completed, and c is dead once the last FF1 has).
constexpr int TAIL2_OFF_C = TAIL_OFF_RING;                              // 64 KB: c (4 K tiles of 16 KB); o (2 K tiles) before it
constexpr int TAIL2_OFF_RING = TAIL2_OFF_C + 65536;                     // 4 x 16 KB
constexpr int TAIL2_PIECE = TAIL_PIECE / 2;
static_assert(TAIL2_OFF_RING + TAIL_NST * TAIL2_PIECE <= TAIL_OFF_PAR, "the pair layout fits the single-CTA ring region");

template <int CG>
__global__ void __launch_bounds__(TAIL_THREADS, 1)
ff_tail_kernel(const __grid_constant__ CUtensorMap tmO3, const __grid_constant__ CUtensorMap tmWo,
               const __grid_constant__ CUtensorMap tmW1_3, const __grid_constant__ CUtensorMap tmW2, const TailParams p) {
  static_assert(CG == 1 || CG == 2, "one CTA per row tile, or a CTA pair per two row tiles");
  // CG == 2: tmWo / tmW2 have 128-row boxes (this CTA's half of the 256 output rows), tmW1_3 box {64, 64 hidden units, 2 K-chunks}
  const uint32_t crank = (CG == 2) ? cluster_ctarank() : 0u;
  // a unit = CG consecutive row tiles, one per CTA of the pair
  const int unit0 = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, nunits = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int NCG = TAIL_NCG;
  constexpr int NEW = 4 * NCG;       // epilogue warps
  constexpr int CW1 = 256 / NCG;     // columns per warp in the 256-wide phases
  constexpr int NCH1 = CW1 / 32;
  constexpr int CW2 = 128 / NCG;     // columns per warp in a 128-wide FF1 chunk (= 32: one TMEM load)
  static_assert(CW2 == 32, "one 32-column TMEM load per thread and hidden chunk");

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  float* s_par = reinterpret_cast<float*>(smem + TAIL_OFF_PAR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TAIL_OFF_BAR);
  uint64_t* full_bar = bars;                        // [TAIL_NST]
  uint64_t* empty_bar = bars + TAIL_NST;            // [TAIL_NST]
  uint64_t* r_full = bars + 2 * TAIL_NST;           // to_out accumulator complete
  uint64_t* c_ready = r_full + 1;                   // c written to TMEM, x_a written back to TMEM
  uint64_t* d1_full = r_full + 2;                   // FF1 chunk accumulator complete ([1] unused)
  uint64_t* d1_empty = r_full + 4;                  // epilogue has read the FF1 chunk accumulator ([1] unused)
  uint64_t* s_ready = r_full + 6;                   // [2] s chunk written to smem
  uint64_t* s_empty = r_full + 9;                   // [2] FF2 has consumed the s chunk
  uint64_t* r_done = r_full + 12;                   // all FF2 MMAs of the tile complete
  uint64_t* r_empty = r_full + 13;                  // epilogue has read the final accumulator
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r_full + 14);
  uint64_t* o_full = r_full + 15;                   // CG == 2: the o tile has landed in the c region
  uint64_t* c_free = r_full + 16;                   // CG == 2: the last FF1 of the tile has read c (the next o tile may land)
  static_assert((2 * TAIL_NST + 17) * 8 <= 256, "barrier block");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (!p.pdl_late) pdl_launch_dependents();   // late: when the CTA's last tile reaches its final epilogue (see gemm_tc.cuh)
  const int m_tiles = (p.M + 127) / 128;
  const int m_units = (m_tiles + CG - 1) / CG;

  if (threadIdx.x == 0) {
    for (int i = 0; i < TAIL_NST; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(r_full, 1); mbar_init(c_ready, CG * NEW);
    for (int i = 0; i < 2; ++i) { mbar_init(&d1_full[i], 1); mbar_init(&d1_empty[i], CG * NEW); }
    for (int i = 0; i < TAIL_NSB; ++i) { mbar_init(&s_ready[i], CG * NEW); mbar_init(&s_empty[i], 1); }
    mbar_init(r_done, 1); mbar_init(r_empty, CG * NEW);
    mbar_init(o_full, 1); mbar_init(c_free, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmO3); tma_prefetch_desc(&tmWo); tma_prefetch_desc(&tmW1_3); tma_prefetch_desc(&tmW2);
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
  if (warp >= 3) {  // weights-only parameters: staged by the epilogue warps behind their own barrier (gemm_tc.cuh), so that the
                    // producer's first TMA instructions do not wait for this global-memory round trip
    for (int i = threadIdx.x - 96; i < 256; i += 32 * NEW) {
      s_par[i] = p.b_o[i]; s_par[256 + i] = p.ln_g[i]; s_par[512 + i] = p.ln_b[i]; s_par[768 + i] = p.b2[i];
    }
    for (int i = threadIdx.x - 96; i < 1024; i += 32 * NEW) {
      s_par[1024 + i] = p.b1[i]; s_par[2048 + i] = p.sn_a[i]; s_par[3072 + i] = p.sn_ib[i];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * NEW) : "memory");
  }
  const uint32_t tR = tmem_base;
  const uint32_t tD1 = tmem_base + 256;   // CG == 2: two FF1 accumulators, columns [256, 384) and [384, 512)
  const uint32_t tC = tmem_base + 384;    // CG == 1: c as packed fp16 pairs

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    // converged warp, one elected lane per instruction (keeps descriptors/coordinates in uniform registers)
    if (unit0 < m_units) {
      uint32_t it = 0;  // running ring item counter
      const uint64_t pol = l2_policy_evict_last();
      constexpr uint32_t WBYTES = TAIL_PIECE / CG;   // this CTA's share of a weight piece
      constexpr int R_OFF = (CG == 2) ? TAIL2_OFF_RING : TAIL_OFF_RING, R_PIECE = (CG == 2) ? TAIL2_PIECE : TAIL_PIECE;
      auto slot_acquire = [&]() -> uint32_t {
        const uint32_t slot = it % TAIL_NST, use = it / TAIL_NST;
        mbar_wait(&empty_bar[slot], (use & 1) ^ 1);
        ++it;
        return slot;
      };
      // CG == 2: every load signals the LEADER's full barrier; the leader's producer arrives once per item and expects the
      // bytes of both CTAs (the peer's complete_tx may come first: the transaction count runs negative inside the phase)
      auto put2 = [&](const CUtensorMap* tm, int c0) {  // weights: this CTA's 256 / CG output rows x 64 cols
        const uint32_t slot = slot_acquire();
        if (elect_one()) {
          uint8_t* dst = smem + R_OFF + slot * R_PIECE;
          if (crank == 0) mbar_arrive_expect_tx(&full_bar[slot], CG * WBYTES);
          if constexpr (CG == 2) {
            const uint32_t lbar = mapa_u32(smem_u32(&full_bar[slot]), 0);
            if (p.w_hint) tma_load_2d_pair_hint(dst, tm, lbar, c0, (int)crank * 128, pol);
            else tma_load_2d_pair(dst, tm, lbar, c0, (int)crank * 128);
          } else {
            if (p.w_hint) tma_load_2d_hint(dst, tm, &full_bar[slot], c0, 0, pol);
            else tma_load_2d(dst, tm, &full_bar[slot], c0, 0);
          }
        }
        __syncwarp();
      };
      auto put_w1 = [&](int j) {   // hidden units [128j + (128 / CG) rank, + 128 / CG): two pieces of K = 128 each
        for (int hh = 0; hh < 2; ++hh) {
          const uint32_t slot = slot_acquire();
          if (elect_one()) {
            uint8_t* dst = smem + R_OFF + slot * R_PIECE;
            if (crank == 0) mbar_arrive_expect_tx(&full_bar[slot], CG * WBYTES);
            if constexpr (CG == 2) {
              const uint32_t lbar = mapa_u32(smem_u32(&full_bar[slot]), 0);
              if (p.w_hint) tma_load_3d_pair_hint(dst, &tmW1_3, lbar, 0, j * 128 + (int)crank * 64, 2 * hh, pol);
              else tma_load_3d_pair(dst, &tmW1_3, lbar, 0, j * 128 + (int)crank * 64, 2 * hh);
            } else {
              if (p.w_hint) tma_load_3d_hint(dst, &tmW1_3, &full_bar[slot], 0, j * 128, 2 * hh, pol);
              else tma_load_3d(dst, &tmW1_3, &full_bar[slot], 0, j * 128, 2 * hh);
            }
          }
          __syncwarp();
        }
      };
      bool first = true;
      uint32_t n_tile_p = 0;
      for (int unit = unit0; unit < m_units; unit += nunits, ++n_tile_p) {
        const int tile = CG * unit + (int)crank;   // CG == 2: may be == m_tiles on the last unit (rows out of range load as zeros)
        put2(&tmWo, 0);    // Wo K-chunk 0
        put2(&tmWo, 64);   // Wo K-chunk 1
        if (first) { pdl_wait(); first = false; }  // o is the first operand produced by the previous kernel
        if constexpr (CG == 2) {   // the o tile goes to the first half of the c region, free once the previous tile's last FF1 has read c
          mbar_wait(c_free, (n_tile_p & 1) ^ 1);
          if (elect_one()) {
            if (crank == 0) mbar_arrive_expect_tx(o_full, CG * TAIL_PIECE);   // each CTA loads its own row tile of o
            tma_load_3d_pair(smem + TAIL2_OFF_C, &tmO3, mapa_u32(smem_u32(o_full), 0), 0, tile * 128, 0);
          }
          __syncwarp();
        } else {
          const uint32_t slot = slot_acquire();
          if (elect_one()) {
            mbar_arrive_expect_tx(&full_bar[slot], TAIL_PIECE);
            tma_load_3d(smem + TAIL_OFF_RING + slot * TAIL_PIECE, &tmO3, &full_bar[slot], 0, tile * 128, 0);
          }
          __syncwarp();
        }
        put_w1(0);
        put_w1(1);
        for (int j = 0; j < TAIL_NJ; ++j) {   // tensor-pipe order: FF1_{j+1} FF2_j FF1_{j+2} FF2_{j+1} ...
          put2(&tmW2, j * 128);
          put2(&tmW2, j * 128 + 64);
          if (j + 2 < TAIL_NJ) put_w1(j + 2);
        }
      }
    }
  } else if ((warp == 1 || warp == 2) && crank == 0) {
    // ===================================== MMA issuers (CG == 2: pair leader only) ============
    // Converged warps (addresses and descriptors stay in uniform registers); only the tcgen05 instructions are
    // issued by one elected lane.  Warp 1 owns the to_out + FF1 stream (accumulators R then D1), warp 2 the FF2
    // stream (accumulates into R); both walk the same ring-item numbering as the producer.
    constexpr uint32_t idesc256 = umma_idesc_f16(128 * CG, 256);   // CG == 2: M = 256, both CTAs' 128 rows
    constexpr uint32_t idesc128 = umma_idesc_f16(128 * CG, 128);
    auto mma_ss = [](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
      if constexpr (CG == 2) umma_f16_pair(d, a, b, idesc, acc); else umma_f16(d, a, b, idesc, acc);
    };
    auto mma_ts = [](uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
      if constexpr (CG == 2) umma_f16_ts_pair(d, a, b, idesc, acc); else umma_f16_ts(d, a, b, idesc, acc);
    };
    auto commit = [](uint64_t* bar) {   // CG == 2: arrives on the barrier at this offset in BOTH CTAs
      if constexpr (CG == 2) umma_commit_pair(bar); else umma_commit(bar);
    };
    const uint32_t ring = smem_u32(smem + ((CG == 2) ? TAIL2_OFF_RING : TAIL_OFF_RING));
    const uint32_t sbuf = smem_u32(smem + TAIL_OFF_S);
    const uint32_t cbuf = smem_u32(smem + TAIL2_OFF_C);   // CG == 2
    auto slot_wait = [&](uint32_t item) -> uint32_t {
      const uint32_t slot = item % TAIL_NST, use = item / TAIL_NST;
      mbar_wait(&full_bar[slot], use & 1);
      tc_fence_after();
      return ring + slot * ((CG == 2) ? TAIL2_PIECE : TAIL_PIECE);
    };
    // ring items of one tile (two per chunk operand): 0,1 Wo K-chunks | 2 o | W1_0 (3,4) W1_1 (5,6) | then for j = 0..7:
    // W2_j (2 items), W1_{j+2} (2 items, while it exists)
    // (CG == 2: the o tile is not a ring item -- it has its own region and barrier -- so everything after it moves down by one)
    constexpr uint32_t OI = (CG == 2) ? 0u : 1u;
    auto item_w1 = [](int k) -> uint32_t { return (k < 2 ? 3u + 2u * k : 4u * k + 1u) - (1u - OI); };          // first of two
    auto item_w2 = [](int j) -> uint32_t { return (j <= 5 ? 7u + 4u * j : 31u + 2u * (j - 6)) - (1u - OI); };  // first of two
    constexpr uint32_t ITEMS = 34 + OI;
    long long* tl = (p.tl != nullptr) ? p.tl + (size_t)unit0 * 128 : nullptr;
    uint32_t n_tile = 0;
    if (warp == 1) {
      uint32_t n_d1e = 0;
      for (int unit = unit0; unit < m_units; unit += nunits, ++n_tile) {
        if (n_tile != 0) tl = nullptr;
        const uint32_t base = n_tile * ITEMS;
        // ---- to_out: R = o Wo^T
        mbar_wait(r_empty, (n_tile & 1) ^ 1);
        tc_fence_after();
        if (tl && lane == 0) tl[0] = clock64();
        {
          uint32_t a;
          if constexpr (CG == 2) { mbar_wait(o_full, n_tile & 1); tc_fence_after(); a = cbuf; }
          else a = slot_wait(base + 2);
          for (int k = 0; k < 2; ++k) {
            const uint32_t b = slot_wait(base + k);
            const uint64_t da = umma_desc_sw128(a + k * 16384), db = umma_desc_sw128(b);
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) mma_ss(tR, da + 2 * kk, db + 2 * kk, idesc256, (k | kk) != 0);
              commit(&empty_bar[(base + k) % TAIL_NST]);
              if (k == 1) {
                if constexpr (CG == 1) commit(&empty_bar[(base + 2) % TAIL_NST]);
                commit(r_full);
              }
            }
            __syncwarp();
          }
        }
        if (tl && lane == 0) tl[1] = clock64();
        // ---- FF1 chunks (A = c from tensor memory)
        mbar_wait(c_ready, n_tile & 1);
        tc_fence_after();
        if (tl && lane == 0) tl[2] = clock64();
        for (int j = 0; j < TAIL_NJ; ++j) {
          // CG == 2: two FF1 accumulators (chunk parity); CG == 1: one.  n_d1e counts the chunks issued so far.
          const uint32_t db_i = (CG == 2) ? (n_d1e & 1u) : 0u, use = (CG == 2) ? (n_d1e >> 1) : n_d1e;
          mbar_wait(&d1_empty[db_i], (use & 1) ^ 1);
          ++n_d1e;
          tc_fence_after();
          if (tl && lane == 0 && j < 8) tl[4 + 4 * j] = clock64();
          for (int hh = 0; hh < 2; ++hh) {   // K halves: one ring piece [2 K-chunks][128 hidden units][64]
            const uint32_t item = base + item_w1(j) + hh;
            const uint32_t w = slot_wait(item);
            const uint64_t db0 = umma_desc_sw128(w);
            if (elect_one()) {
#pragma unroll
              for (int sub = 0; sub < 2; ++sub)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {  // B: this CTA's 128 / CG hidden units x 64 per K-chunk (16 / CG KB)
                  if constexpr (CG == 2)          // A: c from shared memory, K tile 2 hh + sub
                    mma_ss(tD1 + db_i * 128, umma_desc_sw128(cbuf + (2 * hh + sub) * 16384) + 2 * kk, db0 + sub * ((16384 / CG) >> 4) + 2 * kk,
                           idesc128, (hh | sub | kk) != 0);
                  else                            // A: 8 TMEM columns per K16 step
                    mma_ts(tD1, tC + (2 * hh + sub) * 32 + kk * 8, db0 + sub * ((16384 / CG) >> 4) + 2 * kk, idesc128, (hh | sub | kk) != 0);
                }
              commit(&empty_bar[item % TAIL_NST]);
              if (hh == 1) {
                commit(&d1_full[db_i]);
                if (CG == 2 && j == TAIL_NJ - 1) commit(c_free);   // c has been read: the next tile's o may land
              }
            }
            __syncwarp();
          }
          if (tl && lane == 0 && j < 8) tl[5 + 4 * j] = clock64();
        }
      }
    } else {
      uint32_t n_sr = 0;   // FF2 chunks consumed so far (buffer = n % 3, use = n / 3)
      for (int unit = unit0; unit < m_units; unit += nunits, ++n_tile) {
        if (n_tile != 0) tl = nullptr;
        const uint32_t base = n_tile * ITEMS;
        mbar_wait(c_ready, n_tile & 1);   // x_a is in R
        tc_fence_after();
        for (int j = 0; j < TAIL_NJ; ++j) {
          const uint32_t b = n_sr % TAIL_NSB;
          mbar_wait(&s_ready[b], (n_sr / TAIL_NSB) & 1);
          ++n_sr;
          tc_fence_after();
          if (tl && lane == 0 && j < 8) tl[6 + 4 * j] = clock64();
          for (int hh = 0; hh < 2; ++hh) {   // K halves: s K tile hh x W2 piece hh
            const uint32_t item = base + item_w2(j) + hh;
            const uint32_t w = slot_wait(item);
            const uint64_t da = umma_desc_sw128(sbuf + b * TAIL_SBYTES + hh * 16384), db = umma_desc_sw128(w);
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) mma_ss(tR, da + 2 * kk, db + 2 * kk, idesc256, 1u);  // on top of x_a
              commit(&empty_bar[item % TAIL_NST]);
              if (hh == 1) {
                commit(&s_empty[b]);
                if (j == TAIL_NJ - 1) commit(r_done);
              }
            }
            __syncwarp();
          }
          if (tl && lane == 0 && j < 8) tl[7 + 4 * j] = clock64();
        }
      }
    }
  } else if (warp >= 3) {
    // ===================================== epilogue =========================================
    // CG == 2: the epilogue -> MMA barriers live in the pair leader: arrive through the shared::cluster address
    auto arrive_leader = [&](uint64_t* bar) {
      if constexpr (CG == 2) mbar_arrive_cluster(mapa_u32(smem_u32(bar), 0)); else mbar_arrive(bar);
    };
    const int ew = warp - 3;
    const int q = warp & 3;     // TMEM lane quarter
    const int cg = ew >> 2;     // column group
    const uint32_t st = smem_u32(smem + TAIL_OFF_STAGING + ew * GEMM_STAGING_BYTES);
    const uint32_t spar = smem_u32(s_par);
    const uint32_t red = smem_u32(smem + TAIL_OFF_RED);
    const uint32_t sbuf = smem_u32(smem + TAIL_OFF_S);
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const int trow = q * 32 + lane;  // row inside the tile
    uint32_t n_tile = 0, n_d1f[2] = {0, 0}, n_s = 0;   // n_s: s chunks produced so far (buffer = n % 3, use = n / 3)
    long long* tl = (p.tl != nullptr && ew == 0 && lane == 0 && crank == 0) ? p.tl + (size_t)unit0 * 128 + 64 : nullptr;
    pdl_wait();  // x_r / out belong to the dependency chain
    for (int unit = unit0; unit < m_units; unit += nunits, ++n_tile) {
      if (n_tile != 0) tl = nullptr;
      const int tile = CG * unit + (int)crank;
      const int rw0 = tile * 128 + q * 32;
      const int row = rw0 + lane;
      const int rows_valid = min(32, p.M - rw0);
      // ------------------------------------------------ E1: x_a, LayerNorm3 -> c
      {
        const __half* rbase = p.xr + (size_t)rw0 * 256 + cg * CW1;
        uint4 rr[4];
        epi_resid_issue(rr, lane, rbase, 256, rows_valid);
        if (lane == 0) mbar_wait(r_full, n_tile & 1);
        __syncwarp();
        tc_fence_after();
        if (tl) tl[0] = clock64();
        const uint32_t ta = tR + lane_off + cg * CW1;
        float lsum = 0.f, lsq = 0.f;
#pragma unroll
        for (int c = 0; c < NCH1; ++c) {
          float v[32];
          tmem_ld32(ta + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = lds_f4(spar + (cg * CW1 + c * 32 + 4 * j) * 4);
            v[4 * j + 0] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
          }
          epi_resid_add(st, lane, rr, v);
          __syncwarp();
          if (c + 1 < NCH1) epi_resid_issue(rr, lane, rbase + (c + 1) * 32, 256, rows_valid);
#pragma unroll
          for (int j = 0; j < 32; ++j) { lsum += v[j]; lsq = fmaf(v[j], v[j], lsq); }
          tmem_st32(ta + c * 32, v);
        }
        sts_f32(red + (trow * NCG + cg) * 8, lsum);
        sts_f32(red + (trow * NCG + cg) * 8 + 4, lsq);
        tmem_st_wait();
        asm volatile("bar.sync 1, %0;" ::"n"(32 * NEW) : "memory");
        float tsum = 0.f, tsq = 0.f;
#pragma unroll
        for (int g = 0; g < NCG; ++g) {  // fixed order: every warp of a row sees identical statistics
          tsum += lds_f32(red + (trow * NCG + g) * 8);
          tsq += lds_f32(red + (trow * NCG + g) * 8 + 4);
        }
        const float mean = tsum * (1.f / 256.f);
        const float rstd = rsqrtf(fmaxf(tsq * (1.f / 256.f) - mean * mean, 0.f) + 1e-5f);
#pragma unroll
        for (int c = 0; c < NCH1; ++c) {
          float v[32];
          tmem_ld32(ta + c * 32, v);
          tmem_ld_wait();
          const int col = cg * CW1 + c * 32;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 g4 = lds_f4(spar + (256 + col + 4 * j) * 4);
            const float4 b4 = lds_f4(spar + (512 + col + 4 * j) * 4);
            v[4 * j + 0] = fmaf((v[4 * j + 0] - mean) * rstd, g4.x, b4.x);
            v[4 * j + 1] = fmaf((v[4 * j + 1] - mean) * rstd, g4.y, b4.y);
            v[4 * j + 2] = fmaf((v[4 * j + 2] - mean) * rstd, g4.z, b4.z);
            v[4 * j + 3] = fmaf((v[4 * j + 3] - mean) * rstd, g4.w, b4.w);
          }
          uint32_t pk[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) pk[u] = pack_h2(v[2 * u], v[2 * u + 1]);
          if constexpr (CG == 2) {   // c into shared memory: K tile col / 64, 16-byte units (col % 64) / 8 .. + 4 of the row (128B swizzle)
            const uint32_t crow = smem_u32(smem + TAIL2_OFF_C) + (col >> 6) * 16384 + trow * 128;
#pragma unroll
            for (int u = 0; u < 4; ++u)
              sts128(crow + (((((col & 63) >> 3) + u) ^ (trow & 7)) << 4), make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]));
          } else {
            tmem_st16(tC + lane_off + (col >> 1), pk);  // c as fp16 pairs: A operand of FF1, read from TMEM
          }
        }
        if constexpr (CG == 2) fence_proxy_async_smem();
        else tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(c_ready);
        if (tl) tl[1] = clock64();
      }
      // ------------------------------------------------ E2: SnakeBeta on the 8 hidden chunks
#pragma unroll 1
      for (int j = 0; j < TAIL_NJ; ++j) {
        const uint32_t sb = n_s % TAIL_NSB;  // s buffer
        const uint32_t db_i = (CG == 2) ? (n_d1f[0] & 1u) : 0u, d1use = (CG == 2) ? (n_d1f[0] >> 1) : n_d1f[0];   // n_d1f[0]: chunks so far
        if (lane == 0) {
          mbar_wait(&s_empty[sb], ((n_s / TAIL_NSB) & 1) ^ 1);  // FF2_{j-2} no longer reads S[sb]
          mbar_wait(&d1_full[db_i], d1use & 1);
        }
        ++n_s;
        ++n_d1f[0];
        __syncwarp();
        tc_fence_after();
        if (tl && j < 8) tl[4 + 2 * j] = clock64();
        float v[32];
        tmem_ld32(tD1 + db_i * 128 + lane_off + cg * 32, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(&d1_empty[db_i]);  // the accumulator chunk is in registers: FF1_{j+1 / j+2} may overwrite it
        const uint32_t pb = spar + (1024 + j * 128 + cg * 32) * 4;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const float4 b4 = lds_f4(pb + jj * 16);
          const float4 a4 = lds_f4(pb + 4096 + jj * 16);
          const float4 i4 = lds_f4(pb + 8192 + jj * 16);
          float x, sn;
          x = v[4 * jj + 0] + b4.x; sn = fast_sin(x * a4.x); v[4 * jj + 0] = fmaf(sn * sn, i4.x, x);
          x = v[4 * jj + 1] + b4.y; sn = fast_sin(x * a4.y); v[4 * jj + 1] = fmaf(sn * sn, i4.y, x);
          x = v[4 * jj + 2] + b4.z; sn = fast_sin(x * a4.z); v[4 * jj + 2] = fmaf(sn * sn, i4.z, x);
          x = v[4 * jj + 3] + b4.w; sn = fast_sin(x * a4.w); v[4 * jj + 3] = fmaf(sn * sn, i4.w, x);
        }
        // chunk columns cg*32 .. +32 -> K tile cg/2, 16-byte units (cg%2)*4 .. +4 of the row (128B swizzle)
        const uint32_t srow = sbuf + sb * TAIL_SBYTES + (cg >> 1) * 16384 + trow * 128;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          sts128(srow + ((((cg & 1) * 4 + u) ^ (trow & 7)) << 4),
                 make_uint4(pack_h2_sat(v[8 * u], v[8 * u + 1]), pack_h2_sat(v[8 * u + 2], v[8 * u + 3]),
                            pack_h2_sat(v[8 * u + 4], v[8 * u + 5]), pack_h2_sat(v[8 * u + 6], v[8 * u + 7])));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) arrive_leader(&s_ready[sb]);
        if (tl && j < 8) tl[5 + 2 * j] = clock64();
      }
      // ------------------------------------------------ E3: out = (x_a + FF + b2) * mask
      {
        float mrow = 0.f;
        if (row < p.M) mrow = p.rowmask[row];
        if (lane == 0) {
          mbar_wait(r_done, n_tile & 1);
          if (p.pdl_late && unit + nunits >= m_units) pdl_launch_dependents();
        }
        __syncwarp();
        tc_fence_after();
        if (tl) tl[2] = clock64();
        const uint32_t ta = tR + lane_off + cg * CW1;
        __half* obase = p.out + (size_t)rw0 * 256 + cg * CW1;
        float vbuf[2][32];
        tmem_ld32(ta, vbuf[0]);
#pragma unroll
        for (int c = 0; c < NCH1; ++c) {
          float* v = vbuf[c & 1];
          tmem_ld_wait();
          if (c + 1 < NCH1) tmem_ld32(ta + (c + 1) * 32, vbuf[(c + 1) & 1]);
          if (c == NCH1 - 1) {   // the whole accumulator is in registers: the next tile's to_out may overwrite R while this one is stored
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_leader(r_empty);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = lds_f4(spar + (768 + cg * CW1 + c * 32 + 4 * j) * 4);
            v[4 * j + 0] = (mrow == 0.f) ? 0.f : (v[4 * j + 0] + b4.x) * mrow;
            v[4 * j + 1] = (mrow == 0.f) ? 0.f : (v[4 * j + 1] + b4.y) * mrow;
            v[4 * j + 2] = (mrow == 0.f) ? 0.f : (v[4 * j + 2] + b4.z) * mrow;
            v[4 * j + 3] = (mrow == 0.f) ? 0.f : (v[4 * j + 3] + b4.w) * mrow;
          }
          epi_store_h32(st, lane, v, obase + c * 32, 256, rows_valid);
        }
        if (tl) tl[3] = clock64();
      }
    }
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
