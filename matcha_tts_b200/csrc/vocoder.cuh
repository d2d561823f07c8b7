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
// Warp roles (352 threads, persistent over (utterance, row tile, N tile), one CTA per SM): warp 0 TMA producer of the
// activation tiles, warp 1 TMEM allocator + MMA issuer, warp 2 TMA producer of the weight boxes (constants: no dependency
// wait), warps 3-10 epilogue (4 warps for BN = 32) with two accumulator stages in TMEM.
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
  const float* bias;     // [N]
  const __half* res;     // [B * L][ld] or null
  const __half* acc_in;  // [B * L][ld] or null
  __half* out_raw;       // [B * L][ld] or null
  __half* out_act;       // [B * L][ld] or null
  int ld;                // = N
  float slope, scale;
  int w_hint, pdl_late;
};

constexpr int VOC_THREADS = 352;
constexpr int VOC_MAX_HALO_ROWS = 178;   // 128 + (11 - 1) * 5
constexpr int VOC_PAR_N = 2048;

template <int CK, int BN>
struct VocSmem {
  static constexpr int A_STAGE = (VOC_MAX_HALO_ROWS * CK * 2 + 1023) / 1024 * 1024;
  static constexpr int NA = 3;
  static constexpr int W_TILE = BN * CK * 2;
  static constexpr int WB_MAX = (32768 / W_TILE) < 16 ? (32768 / W_TILE) : 16;
  static constexpr int W_STAGE = WB_MAX * W_TILE;
  static constexpr int NW = 4;
  static constexpr int EPI_WARPS = BN >= 64 ? 8 : 4;
  static constexpr int OFF_W = NA * A_STAGE;
  static constexpr int OFF_STAGE = OFF_W + NW * W_STAGE;
  static constexpr int OFF_PAR = OFF_STAGE + 8 * GEMM_STAGING_BYTES;
  static constexpr int OFF_BAR = OFF_PAR + VOC_PAR_N * 4;
  static constexpr int TOTAL = OFF_BAR + 256;
  static_assert(W_TILE % 1024 == 0 && WB_MAX >= 1, "weight k-tiles are whole swizzle atoms");
  static_assert(TOTAL <= 232448, "exceeds the 227 KB of shared memory one CTA can own");
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

template <int CK, int BN>
__global__ void __launch_bounds__(VOC_THREADS, 1)
voc_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const VocConvParams p) {
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
  static_assert((2 * NA + 2 * NW + 5) * 8 <= 256, "barrier block");

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
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  if (warp >= 3) {
    const int ncols = min(p.n_tiles * BN, VOC_PAR_N);
    for (int i = threadIdx.x - 96; i < ncols; i += VOC_THREADS - 96) s_par[i] = p.bias ? p.bias[i] : 0.f;
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
    constexpr uint32_t idesc = umma_idesc_f16(128, BN);
    int sa = 0, sw = 0, as = 0;
    uint32_t pa = 0, pw = 0, aphase = 0;
    const uint32_t a_base = smem_u32(smem), w_base = smem_u32(smem + SM::OFF_W);
    for (int ti = 0; cta_tile(ti) >= 0; ++ti) {
      mbar_wait(&tempty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      int kt = 0, kw = 0;   // k-tile of the conv, k-tile inside the current weight box
      for (int c = 0; c < p.n_chunks; ++c) {
        mbar_wait(&afull[sa], pa);
        for (int t = 0; t < p.taps; ++t, ++kt) {
          if (kw == 0) mbar_wait(&wfull[sw], pw);
          tc_fence_after();
          const uint32_t a_addr = a_base + sa * SM::A_STAGE + t * p.dil * (CK * 2);
          const uint32_t w_addr = w_base + sw * SM::W_STAGE + kw * SM::W_TILE;
          const uint64_t da = (CK == 64) ? umma_desc_sw128(a_addr) : umma_desc_sw64(a_addr);
          const uint64_t db = (CK == 64) ? umma_desc_sw128(w_addr) : umma_desc_sw64(w_addr);
          const bool last_of_box = (kw + 1 == p.wb) || (kt + 1 == KT);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < CK / 16; ++k) umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kt | k) != 0);
            if (last_of_box) umma_commit(&wempty[sw]);
            if (t + 1 == p.taps) umma_commit(&aempty[sa]);
            if (kt + 1 == KT) umma_commit(&tfull[as]);
          }
          __syncwarp();
          if (last_of_box) { kw = 0; if (++sw == NW) { sw = 0; pw ^= 1; } }
          else ++kw;
        }
        if (++sa == NA) { sa = 0; pa ^= 1; }
      }
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
    const uint32_t st = smem_u32(smem + SM::OFF_STAGE + ew * GEMM_STAGING_BYTES);
    const uint32_t spar = smem_u32(s_par);
    int as = 0;
    uint32_t aphase = 0;
    pdl_wait();   // residual reads, and stores into buffers the previous kernel may still be reading
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
          epi_resid_add(st, lane, rr, v);
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
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f) + p.slope * fminf(v[j], 0.f);
          epi_store_h32(st, lane, v, p.out_act + g0 + ch * 32, p.ld, rows_valid);
        }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
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
