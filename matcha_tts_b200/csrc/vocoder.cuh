// Kernels of the HiFi-GAN generator (reference hifigan/models.py:148-195, ResBlock1 :14-98, config v1): every Conv1d and
// ConvTranspose1d of the vocoder as ONE implicit-GEMM kernel on tcgen05 tensor cores, plus the two row-local ends (mel packing,
// conv_post + tanh).
//
//   D[(b, t), n] = sum_{tap, ci}  X[b, t + (tap - taps/2) * dil, ci] . W[n, tap, ci]          (zero outside 0 <= t < L)
//
// Activations are channels-last fp16 [B][L][C] -- NO guard rows: the operand tensor map is 3-D (channel, frame, utterance) and
// TMA zero-fills the frames before 0 and after L - 1, which is the convolution's zero padding.  A ConvTranspose1d with kernel
// 2u, stride u, padding u/2 (all four `ups`) is the same kernel: a 3-tap convolution over the INPUT frames whose N = u * Cout
// columns are the u output phases, so that the [B][L_in][u * Cout] result is, without any copy, the [B][u * L_in][Cout]
// upsampled signal (the phase / tap pairs that do not exist carry zero weights).
//
// One (128 + (taps - 1) * dil)-row activation tile per K chunk feeds ALL taps: tap t is the same tile read through a
// shared-memory descriptor that starts t * dil rows further in (the swizzle is a function of the address bits, so the shifted
// view addresses exactly the rows TMA wrote -- tools/ubench/rowshift.cu, already used by the estimator's k3 convs).  The
// dilated k = 11 convs therefore read their input once, not eleven times.  Weights are packed chunk-major
// ([N][chunk][tap][CK]) and arrive as 3-D boxes of `wb` consecutive k-tiles (<= 32 KB per TMA instruction: an instruction
// costs its issuing thread ~330 cycles whatever it moves).
//
// CK = 64 channels per K chunk with the 128-byte swizzle; CK = 32 with the 64-byte swizzle for the 32-channel level.
// BN = N tile = accumulator columns: 256 / 128 / 64 / 32 (the level's channel count, or 256 for the wide ups / conv_pre).
//
// Warp roles (persistent over (utterance, row tile, N tile)): warp 0 TMA producer of the activation tiles, warp 1 TMEM
// allocator + MMA issuer, warp 2 TMA producer of the weight boxes (constants: no dependency wait), then the epilogue warps
// (8 for BN = 256, one CTA per SM; 4 for the narrower tiles, 2-3 CTAs per SM) with two accumulator stages in TMEM.
//
// Fused epilogue (models.py:84-91, :183-193):  v = (acc + bias [+ res] [+ acc_in]) * scale;  out_raw = v;
// out_act = leaky_relu(v, slope) -- the activation every consumer conv applies to its input is applied once, by the
// producer, next to the raw residual stream.
#pragma once
#include <cuda.h>

#include "gemm_tc.cuh"
#include "ptx.cuh"

namespace mtts {

struct VocConvParams {
  int B, L;              // utterances, rows per utterance (input rows == GEMM rows)
  int tiles_per_utt;     // ceil(L / 128)
  int n_tiles;           // N / BN
  int n_chunks;          // Cin / CK
  int taps, dil;         // tap t reads row t0 + r + (t - taps / 2) * dil
  int wb;                // k-tiles per weight box
  int w_resident;        // 1: one N tile and the whole filter fits the weight ring: loaded once per CTA, never released
  const float* bias;     // [N]
  const __half* res;     // [B * L][ld] or null
  float res_unslope;     // 0: `res` holds the residual itself.  u > 1: it holds leaky_relu(x, 1 / u) and x = min(a, u * a) is recovered
                         //    from it -- the activated copy is all the stream keeps (fp16(x / u) * u carries the same 2^-11 relative
                         //    rounding a raw fp16 copy would), which saves one store and one buffer per conv pair
  const __half* acc_in;  // [B * L][ld] or null
  __half* out_raw;       // [B * L][ld] or null
  __half* out_act;       // [B * L][ld] or null
  int ld;                // = N
  float slope, scale;
  int w_hint, pdl_late;
};

constexpr int VOC_MAX_HALO_ROWS = 178;   // 128 + (11 - 1) * 5
constexpr int VOC_PAR_N = 2048;

// Per-variant resources.  The 256-wide tiles own an SM (two 256-column accumulators = all of TMEM, 4 x 32 KB weight ring).
// The narrower tiles are latency-bound per CTA -- an M128 x N32..128 x K16 MMA is 16-66 cycles of tensor work but a tap of
// the issue loop, the TMA issue and the epilogue of a tile each cost hundreds -- so they run 2 (N = 128, 64) or 3 (N = 32)
// CTAs per SM with 4 epilogue warps each, shallower rings, and (when the whole filter fits the ring: N = 32, and k = 3 at
// N = 64) the weights loaded ONCE per CTA instead of once per tile.
template <int CK, int BN>
struct VocSmem {
  static constexpr int CTAS = BN == 256 ? 1 : (BN == 32 ? 3 : 2);          // CTAs per SM
  static constexpr int EPI_WARPS = BN == 256 ? 8 : 4;
  static constexpr int THREADS = 96 + 32 * EPI_WARPS;
  static constexpr int A_STAGE = (VOC_MAX_HALO_ROWS * CK * 2 + 1023) / 1024 * 1024;
  static constexpr int NA = BN == 256 ? 3 : 2;
  // narrow tiles: the epilogue moves its tiles by TMA (the residual load, requested one chunk ahead, and the store, as 32 x 32
  // boxes through two 2 KB buffers per warp in the 64-byte swizzle) instead of register transposes -- it is the critical
  // path there.  (Every launch has ONE output since the residual stream is kept in its activated form only.)
  static constexpr bool TMA_EPI = BN <= 128;
  static constexpr int W_TILE = BN * CK * 2;
  static constexpr int WB_MAX = BN == 32 ? 4 : 1;   // k-tiles per weight box
  static constexpr int W_STAGE = WB_MAX * W_TILE;
  static constexpr int NW = BN == 128 ? 3 : (BN == 32 ? 3 : 4);
  static constexpr int PAR_N = BN == 256 ? VOC_PAR_N : BN;
  static constexpr int OFF_W = NA * A_STAGE;
  static constexpr int OFF_STAGE = OFF_W + NW * W_STAGE;
  static constexpr int EPI_BOXES = BN == 128 ? 2 : 3;   // a third box (the resblocks' running sum, also prefetched) where it fits
  static constexpr int EPI_STAGING = TMA_EPI ? EPI_BOXES * 2048 : GEMM_STAGING_BYTES;   // per epilogue warp
  static constexpr int OFF_PAR = OFF_STAGE + EPI_WARPS * EPI_STAGING;
  static constexpr int OFF_BAR = OFF_PAR + PAR_N * 4;
  static constexpr int TOTAL = OFF_BAR + 256;
  static_assert(OFF_STAGE % 1024 == 0, "epilogue boxes sit on swizzle-atom boundaries");
  static_assert(W_TILE % 1024 == 0 && WB_MAX >= 1, "weight k-tiles are whole swizzle atoms");
  static_assert(CTAS * (TOTAL + 1024) <= 233472, "CTAS of these must fit the 228 KB of shared memory of one SM");
};

// K-major operand tile stored as rows of 64 bytes (32 x 16-bit) with the 64-byte swizzle (CU_TENSOR_MAP_SWIZZLE_64B):
// 8-row groups are 512 B apart, layout type 4
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}

// epi_resid_add with the residual optionally recovered from its activated copy: x = min(a, un * a) (VocConvParams::res_unslope)
__device__ __forceinline__ void voc_resid_add(uint32_t st, int lane, const uint4 (&rr)[4], float* v, float un) {
#pragma unroll
  for (int it = 0; it < 4; ++it) sts128(epi_st_addr(st, it * 8 + (lane >> 2), lane & 3), rr[it]);
  __syncwarp();
  const float m = (un != 0.f) ? un : 1.f;   // min(a, 1 * a) = a
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 u = lds128(epi_st_addr(st, lane, j));
    float2 f;
    f = unpack_h2(u.x); v[8 * j + 0] += fminf(f.x, f.x * m); v[8 * j + 1] += fminf(f.y, f.y * m);
    f = unpack_h2(u.y); v[8 * j + 2] += fminf(f.x, f.x * m); v[8 * j + 3] += fminf(f.y, f.y * m);
    f = unpack_h2(u.z); v[8 * j + 4] += fminf(f.x, f.x * m); v[8 * j + 5] += fminf(f.y, f.y * m);
    f = unpack_h2(u.w); v[8 * j + 6] += fminf(f.x, f.x * m); v[8 * j + 7] += fminf(f.y, f.y * m);
  }
}

template <int CK, int BN>
__global__ void __launch_bounds__((VocSmem<CK, BN>::THREADS), (VocSmem<CK, BN>::CTAS))
voc_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmRes,
                const __grid_constant__ CUtensorMap tmAcc, const __grid_constant__ CUtensorMap tmRaw, const __grid_constant__ CUtensorMap tmAct,
                const VocConvParams p) {
  using SM = VocSmem<CK, BN>;
  static_assert(CK == 64 || CK == 32, "K chunk = one swizzle row");
  constexpr int NA = SM::NA, NW = SM::NW, EW = SM::EPI_WARPS;
  constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  float* s_par = reinterpret_cast<float*>(smem + SM::OFF_PAR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::OFF_BAR);
  uint64_t* afull = bars;                     // [NA]
  uint64_t* aempty = bars + NA;               // [NA]
  uint64_t* wfull = bars + 2 * NA;            // [NW]
  uint64_t* wempty = bars + 2 * NA + NW;      // [NW]
  uint64_t* tfull = bars + 2 * NA + 2 * NW;   // [2]
  uint64_t* tempty = tfull + 2;               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint64_t* ebar = tempty + 3;                // [2 * EW] (TMA epilogue): residual / running-sum box landed, per warp
  static_assert((2 * NA + 2 * NW + 5 + (SM::TMA_EPI ? 2 * EW : 0)) * 8 <= 256, "barrier block");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (!p.pdl_late) pdl_launch_dependents();

  const int total_tiles = p.B * p.tiles_per_utt * p.n_tiles;
  const int KT = p.n_chunks * p.taps;
  const int halo_lo = (p.taps / 2) * p.dil;
  const int box_rows = 128 + (p.taps - 1) * p.dil;
  auto cta_tile = [&](int i) -> int { const int t = (int)blockIdx.x + i * (int)gridDim.x; return t < total_tiles ? t : -1; };

  if (threadIdx.x == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(&afull[i], 1); mbar_init(&aempty[i], 1); }
    for (int i = 0; i < NW; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], EW); }
    if (SM::TMA_EPI) for (int i = 0; i < 2 * EW; ++i) mbar_init(&ebar[i], 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  if (warp >= 3) {
    const int ncols = min(p.n_tiles * BN, SM::PAR_N);
    for (int i = threadIdx.x - 96; i < ncols; i += SM::THREADS - 96) s_par[i] = p.bias ? p.bias[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 2) {
    // ===================================== TMA producer: weight boxes (constants: no dependency wait) =====
    int stage = 0;
    uint32_t phase = 0;
    const uint64_t pol = l2_policy_evict_last();
    const int ngroups = (KT + p.wb - 1) / p.wb;
    for (int ti = 0, tile; (tile = cta_tile(ti)) >= 0; ++ti) {
      if (p.w_resident && ti > 0) break;   // the filter stays where the first tile put it
      const int n0 = (tile % p.n_tiles) * BN;
      for (int g = 0; g < ngroups; ++g) {
        mbar_wait(&wempty[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sb = smem + SM::OFF_W + stage * SM::W_STAGE;
          mbar_arrive_expect_tx(&wfull[stage], (uint32_t)p.wb * SM::W_TILE);   // the full box, also where it overhangs KT
          if (p.w_hint) tma_load_3d_hint(sb, &tmW, &wfull[stage], 0, n0, g * p.wb, pol);
          else tma_load_3d(sb, &tmW, &wfull[stage], 0, n0, g * p.wb);
        }
        __syncwarp();
        if (++stage == NW) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 0) {
    // ===================================== TMA producer: activation tiles (one per K chunk, all taps) =====
    pdl_wait();
    int stage = 0;
    uint32_t phase = 0;
    for (int ti = 0, tile; (tile = cta_tile(ti)) >= 0; ++ti) {
      const int m = tile / p.n_tiles, b = m / p.tiles_per_utt, t0 = (m % p.tiles_per_utt) * 128;
      for (int c = 0; c < p.n_chunks; ++c) {
        mbar_wait(&aempty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&afull[stage], (uint32_t)box_rows * CK * 2);
          tma_load_3d(smem + stage * SM::A_STAGE, &tmA, &afull[stage], c * CK, t0 - halo_lo, b);
        }
        __syncwarp();
        if (++stage == NA) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    // The whole warp runs the loop converged (waits, address arithmetic); the tcgen05 instructions sit under elect.sync, which
    // the compiler turns into uniform-datapath issue (a plain `lane == 0` test makes it loop over the "active" lanes instead).
    // A tap of a narrow tile is 2-4 short MMAs: the loop around them has to cost less than they do.
    constexpr uint32_t idesc = umma_idesc_f16(128, BN);
    int sa = 0, sw = 0, as = 0;
    uint32_t pa = 0, pw = 0, aphase = 0;
    const uint32_t a_base = smem_u32(smem), w_base = smem_u32(smem + SM::OFF_W);
    const uint32_t tap_step = (uint32_t)p.dil * (CK * 2);
    const bool resident = p.w_resident != 0;
    for (int ti = 0; cta_tile(ti) >= 0; ++ti) {
      mbar_wait(&tempty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      const bool w_wait = !(resident && ti > 0);   // a resident filter is complete after the first tile
      int kt = 0, kw = 0;   // k-tile of the conv, k-tile inside the current weight box
      if (resident) sw = 0;
      for (int c = 0; c < p.n_chunks; ++c) {
        mbar_wait(&afull[sa], pa);
        tc_fence_after();
        uint32_t a_addr = a_base + sa * SM::A_STAGE;
        for (int t = 0; t < p.taps; ++t, ++kt, a_addr += tap_step) {
          if (kw == 0 && w_wait) { mbar_wait(&wfull[sw], pw); tc_fence_after(); }
          const uint32_t w_addr = w_base + sw * SM::W_STAGE + kw * SM::W_TILE;
          const uint64_t da = (CK == 64) ? umma_desc_sw128(a_addr) : umma_desc_sw64(a_addr);
          const uint64_t db = (CK == 64) ? umma_desc_sw128(w_addr) : umma_desc_sw64(w_addr);
          const bool last_of_box = (kw + 1 == p.wb) || (kt + 1 == KT);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < CK / 16; ++k) umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kt | k) != 0);
            if (last_of_box && !resident) umma_commit(&wempty[sw]);
          }
          if (last_of_box) { kw = 0; if (++sw == NW) { sw = 0; pw ^= 1; } }
          else ++kw;
        }
        if (elect_one()) umma_commit(&aempty[sa]);
        if (++sa == NA) { sa = 0; pa ^= 1; }
      }
      if (elect_one()) umma_commit(&tfull[as]);
      __syncwarp();
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  } else if (warp - 3 < EW) {
    // ===================================== epilogue =========================================
    constexpr int NCG = EW / 4;          // column groups per TMEM lane quarter
    constexpr int CW = BN / NCG;         // columns per epilogue warp
    constexpr int NCH = CW / 32;         // 32-column chunks per warp
    static_assert(CW % 32 == 0, "epilogue works on 32-column chunks");
    const int ew = warp - 3;
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int cbase = (ew >> 2) * CW;
    const uint32_t spar = smem_u32(s_par);
    int as = 0;
    uint32_t aphase = 0;
    pdl_wait();   // residual reads, and stores into buffers the previous kernel may still be reading
    if constexpr (SM::TMA_EPI) {
      // Per warp two 32-row x 64-byte boxes in the 64-byte swizzle (16-byte unit u of row r at r * 64 + ((u ^ ((r >> 1) & 3)) << 4):
      // what TMA reads / writes, and conflict-free for "thread = row" accesses): [0] the residual tile, requested for the NEXT chunk
      // as soon as this chunk has read it, [1] the output tile, whose store overlaps the next chunk's TMEM load and arithmetic.
      // The running sum of the resblocks (two launches per level) goes through a third box, prefetched like the residual, where
      // shared memory allows (N <= 64); at N = 128 it is read straight from global memory, one 64-byte row per lane.
      constexpr bool ACC_BOX = SM::EPI_BOXES == 3;
      const uint32_t sb = smem_u32(smem + SM::OFF_STAGE + ew * SM::EPI_STAGING);
      const uint32_t b_res = sb, b_out = sb + 2048, b_acc = sb + 4096;
      uint64_t* rbar = &ebar[2 * ew];
      uint64_t* abar = &ebar[2 * ew + 1];
      const bool has_res = p.res != nullptr, has_acc = p.acc_in != nullptr;
      const float un = p.res_unslope != 0.f ? p.res_unslope : 1.f;   // min(a, 1 * a) = a
      __half* const outp = p.out_act ? p.out_act : p.out_raw;
      const CUtensorMap* const tm_out = p.out_act ? &tmAct : &tmRaw;
      uint32_t lph = 0;
      // coordinates of flattened chunk e = ti * NCH + ch of this warp: column, row inside the utterance, utterance; false past the end
      auto chunk_at = [&](int e, int& col, int& row, int& bb) -> bool {
        const int tile = cta_tile(e / NCH);
        if (tile < 0) return false;
        const int m = tile / p.n_tiles;
        bb = m / p.tiles_per_utt;
        row = (m % p.tiles_per_utt) * 128 + q * 32;
        col = (tile % p.n_tiles) * BN + cbase + (e % NCH) * 32;
        return true;
      };
      auto request = [&](int e) {   // lane 0
        int col, row, bb;
        if (!chunk_at(e, col, row, bb)) return;
        if (has_res) {
          mbar_arrive_expect_tx(rbar, 2048);
          tma_load_3d(reinterpret_cast<void*>(smem + SM::OFF_STAGE + ew * SM::EPI_STAGING), &tmRes, rbar, col, row, bb);
        }
        if (ACC_BOX && has_acc) {
          mbar_arrive_expect_tx(abar, 2048);
          tma_load_3d(reinterpret_cast<void*>(smem + SM::OFF_STAGE + ew * SM::EPI_STAGING + 4096), &tmAcc, abar, col, row, bb);
        }
      };
      if ((has_res || (ACC_BOX && has_acc)) && lane == 0) request(0);
      int e = 0;
      for (int ti = 0, tile; (tile = cta_tile(ti)) >= 0; ++ti) {
        const bool last_tile = p.pdl_late && cta_tile(ti + 1) < 0;
        if (lane == 0) mbar_wait(&tfull[as], aphase);
        __syncwarp();
        tc_fence_after();
        if (last_tile && lane == 0) pdl_launch_dependents();
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + as * BN + cbase;
#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch, ++e) {
          int col, row, bb;
          chunk_at(e, col, row, bb);
          uint4 ar[ACC_BOX ? 1 : 4];
          if constexpr (!ACC_BOX) {
            if (has_acc) {
              const bool ok = row + lane < p.L;
              const __half* ap = p.acc_in + ((size_t)bb * p.L + (ok ? row + lane : 0)) * p.ld + col;
#pragma unroll
              for (int j = 0; j < 4; ++j) ar[j] = ok ? ldg128(ap + 8 * j) : make_uint4(0, 0, 0, 0);
            }
          }
          float v[32];
          tmem_ld32(taddr + ch * 32, v);
          tmem_ld_wait();
          if (ch + 1 == NCH) {   // the accumulator is in registers: hand the TMEM stage back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bv = lds_f4(spar + (col + 4 * j) * 4);
            v[4 * j + 0] += bv.x; v[4 * j + 1] += bv.y; v[4 * j + 2] += bv.z; v[4 * j + 3] += bv.w;
          }
          if (has_res) {
            mbar_wait(rbar, lph);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 u = lds128(b_res + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4));
              float2 f;
              f = unpack_h2(u.x); v[8 * j + 0] += fminf(f.x, f.x * un); v[8 * j + 1] += fminf(f.y, f.y * un);
              f = unpack_h2(u.y); v[8 * j + 2] += fminf(f.x, f.x * un); v[8 * j + 3] += fminf(f.y, f.y * un);
              f = unpack_h2(u.z); v[8 * j + 4] += fminf(f.x, f.x * un); v[8 * j + 5] += fminf(f.y, f.y * un);
              f = unpack_h2(u.w); v[8 * j + 6] += fminf(f.x, f.x * un); v[8 * j + 7] += fminf(f.y, f.y * un);
            }
          }
          if (ACC_BOX && has_acc) {
            mbar_wait(abar, lph);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 u = lds128(b_acc + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4));
              float2 f;
              f = unpack_h2(u.x); v[8 * j + 0] += f.x; v[8 * j + 1] += f.y;
              f = unpack_h2(u.y); v[8 * j + 2] += f.x; v[8 * j + 3] += f.y;
              f = unpack_h2(u.z); v[8 * j + 4] += f.x; v[8 * j + 5] += f.y;
              f = unpack_h2(u.w); v[8 * j + 6] += f.x; v[8 * j + 7] += f.y;
            }
          }
          if (has_res || (ACC_BOX && has_acc)) {
            lph ^= 1;
            __syncwarp();                       // every lane has read its rows: the boxes may be refilled
            if (lane == 0) request(e + 1);
          }
          if constexpr (!ACC_BOX) {
            if (has_acc) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float2 f;
                f = unpack_h2(ar[j].x); v[8 * j + 0] += f.x; v[8 * j + 1] += f.y;
                f = unpack_h2(ar[j].y); v[8 * j + 2] += f.x; v[8 * j + 3] += f.y;
                f = unpack_h2(ar[j].z); v[8 * j + 4] += f.x; v[8 * j + 5] += f.y;
                f = unpack_h2(ar[j].w); v[8 * j + 6] += f.x; v[8 * j + 7] += f.y;
              }
            }
          }
          if (p.scale != 1.f) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= p.scale;
          }
          if (p.out_act) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], p.slope * v[j]);   // 0 <= slope < 1
          }
          if (lane == 0) tma_store_wait_read<0>();   // the previous chunk's store has read the box
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128(b_out + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4),
                   make_uint4(pack_h2(v[8 * j + 0], v[8 * j + 1]), pack_h2(v[8 * j + 2], v[8 * j + 3]), pack_h2(v[8 * j + 4], v[8 * j + 5]),
                              pack_h2(v[8 * j + 6], v[8 * j + 7])));
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && row < p.L && outp) {
            tma_store_3d(tm_out, b_out, col, row, bb);
            tma_store_commit();
          }
        }
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
      if (lane == 0) tma_store_wait<0>();
    } else {
    const uint32_t st = smem_u32(smem + SM::OFF_STAGE + ew * SM::EPI_STAGING);
    for (int ti = 0, tile; (tile = cta_tile(ti)) >= 0; ++ti) {
      const int m = tile / p.n_tiles, b = m / p.tiles_per_utt, t0 = (m % p.tiles_per_utt) * 128;
      const int n0 = (tile % p.n_tiles) * BN + cbase;
      const int tw0 = t0 + q * 32;
      const int rows_valid = min(32, p.L - tw0);   // may be <= 0
      const size_t g0 = ((size_t)b * p.L + tw0) * p.ld + n0;
      const bool last_tile = p.pdl_late && cta_tile(ti + 1) < 0;
      uint4 rr[4], ra[4];
      if (p.res) epi_resid_issue(rr, lane, p.res + g0, p.ld, rows_valid);
      if (p.acc_in) epi_resid_issue(ra, lane, p.acc_in + g0, p.ld, rows_valid);
      if (lane == 0) mbar_wait(&tfull[as], aphase);
      __syncwarp();
      tc_fence_after();
      if (last_tile && lane == 0) pdl_launch_dependents();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + as * BN + cbase;
#pragma unroll 1
      for (int ch = 0; ch < NCH; ++ch) {
        float v[32];
        tmem_ld32(taddr + ch * 32, v);
        tmem_ld_wait();
        if (ch + 1 == NCH) {   // the accumulator is in registers: hand the TMEM stage back before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[as]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = lds_f4(spar + (n0 + ch * 32 + 4 * j) * 4);
          v[4 * j + 0] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
        }
        if (p.res) {
          voc_resid_add(st, lane, rr, v, p.res_unslope);
          __syncwarp();
          if (ch + 1 < NCH) epi_resid_issue(rr, lane, p.res + g0 + (ch + 1) * 32, p.ld, rows_valid);
        }
        if (p.acc_in) {
          epi_resid_add(st, lane, ra, v);
          __syncwarp();
          if (ch + 1 < NCH) epi_resid_issue(ra, lane, p.acc_in + g0 + (ch + 1) * 32, p.ld, rows_valid);
        }
        if (p.scale != 1.f) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= p.scale;
        }
        if (p.out_raw) epi_store_h32(st, lane, v, p.out_raw + g0 + ch * 32, p.ld, rows_valid);
        if (p.out_act) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], p.slope * v[j]);   // 0 <= slope < 1
          epi_store_h32(st, lane, v, p.out_act + g0 + ch * 32, p.ld, rows_valid);
        }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    }
    if (p.pdl_late && cta_tile(0) < 0 && lane == 0) pdl_launch_dependents();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// mel (B, n_mels, T) fp32 -> channels-last fp16 [B][T][ldc] (columns >= n_mels stay zero: the buffer is zero-filled once)
__global__ void voc_pack_mel_kernel(const float* __restrict__ mel, __half* __restrict__ out, int B, int n_mels, int T, int ldc) {
  pdl_launch_dependents();
  pdl_wait();
  const int groups = n_mels / 8;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)B * groups * T) return;
  const int t = (int)(i % T), cg = (int)((i / T) % groups), b = (int)(i / ((long)T * groups));
  const float* src = mel + ((size_t)b * n_mels + cg * 8) * T + t;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = src[(size_t)j * T];
  stg128(out + ((size_t)b * T + t) * ldc + cg * 8, make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7])));
}

// conv_post (C -> 1, k = 7, padding 3) + tanh (models.py:192-193) on the already-activated last level [B][L][C] (C = 32):
// 256 output samples per block; the 262 input rows are staged with an 80-byte pitch (conflict-free 16-byte row reads)
constexpr int VOC_POST_C = 32;
constexpr int VOC_POST_PITCH = 80;
__global__ void __launch_bounds__(256) voc_post_kernel(const __half* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                       float* __restrict__ wav, int L) {
  __shared__ __align__(16) uint8_t rows[262 * VOC_POST_PITCH];
  __shared__ float wt[7 * VOC_POST_C];
  pdl_launch_dependents();
  const int b = blockIdx.y, t0 = blockIdx.x * 256;
  if (threadIdx.x < 7 * VOC_POST_C) {   // w is (1, C, 7): wt[tap][c]
    const int tap = threadIdx.x / VOC_POST_C, c = threadIdx.x % VOC_POST_C;
    wt[threadIdx.x] = w[c * 7 + tap];
  }
  const float bb = bias[0];
  pdl_wait();
  for (int i = threadIdx.x; i < 262 * 4; i += 256) {
    const int r = i >> 2, u = i & 3, t = t0 - 3 + r;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (t >= 0 && t < L) v = ldg128(x + ((size_t)b * L + t) * VOC_POST_C + u * 8);
    *reinterpret_cast<uint4*>(rows + r * VOC_POST_PITCH + u * 16) = v;
  }
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t >= L) return;
  float acc = bb;
#pragma unroll
  for (int tap = 0; tap < 7; ++tap) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint4 v = *reinterpret_cast<const uint4*>(rows + (threadIdx.x + tap) * VOC_POST_PITCH + u * 16);
      const float* ww = wt + tap * VOC_POST_C + u * 8;
      float2 f;
      f = unpack_h2(v.x); acc = fmaf(f.x, ww[0], acc); acc = fmaf(f.y, ww[1], acc);
      f = unpack_h2(v.y); acc = fmaf(f.x, ww[2], acc); acc = fmaf(f.y, ww[3], acc);
      f = unpack_h2(v.z); acc = fmaf(f.x, ww[4], acc); acc = fmaf(f.y, ww[5], acc);
      f = unpack_h2(v.w); acc = fmaf(f.x, ww[6], acc); acc = fmaf(f.y, ww[7], acc);
    }
  }
  wav[(size_t)b * L + t] = tanhf(acc);
}

}  // namespace mtts
