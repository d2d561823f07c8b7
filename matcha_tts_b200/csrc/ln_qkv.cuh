// Fused tail of ResnetBlock1D + head of BasicTransformerBlock for one 128-row tile (reference model.py:773-775,
// :786-789, :735, :662-664): everything between the second conv of the resnet and the attention product is
// row-local once the GroupNorm statistics are known, so one kernel does
//
//   x_r = Mish(GroupNorm(y)) * m + res                      (Block1D #2 + residual; written to global, NOT masked)
//   a   = LayerNorm1(x_r)                                    (never leaves the SM: A operand of the QKV GEMM)
//   q | k | v = a Wqkv^T                                     (q pre-scaled; v stored transposed per head)
//
// instead of a GroupNorm-apply launch followed by a GEMM launch (each launch costs ~6 us of dependent-launch
// latency whatever it does: profiles/r01_gemm_repeat.txt).
// Warp roles (608 threads, one CTA per SM, persistent over row tiles): warp 0 TMA producer of the weight pieces
// ([128 rows x 128 K] = 32 KB, one 3-D box instruction each, six per tile through a 4-slot ring), warp 1 MMA issuer,
// warp 2 idle, warps 3-18 transform the tile (8 rows per warp, 8 channels per lane, LayerNorm by warp shuffles, `a`
// written as four 128B-swizzled 128x64 K-chunk tiles) and afterwards run the epilogue (TMEM lane quarter = warp % 4,
// 96 accumulator columns per warp).
#pragma once
#include <cuda.h>

#include "gemm_tc.cuh"
#include "ptx.cuh"

namespace mtts {

struct LnQkvParams {
  int M;                    // rows of the level's flat row space
  int L, Lp, S;             // frames per utterance, rows per utterance, GroupNorm partial slots per utterance
  const __half* y;          // [rows, 256] raw conv output
  const __half* res;        // [rows, 256] res_conv output
  const float* stats_part;  // [B][S][16]
  const float* gamma;       // GroupNorm affine [256]
  const float* beta;
  const float* ln_g;        // LayerNorm1 affine [256]
  const float* ln_b;
  const float* rowmask;     // [rows]
  const int* rowb;          // [rows] utterance id, -1 on guard rows
  __half* xr;               // [rows, 256]
  __half* q;                // [rows, 128]
  __half* k;                // [rows, 128]
  __half* vt;               // [(b*2+h)*64 + d][Lpad]
  __half* v;                // [rows, 128] row-major (attention2.cuh); when set, vt is not written
  int Lpad;
  int w_hint;
};

constexpr int LQ_NST = 4;
constexpr int LQ_PIECE = 32768;
constexpr int LQ_THREADS = 96 + 512;
constexpr int LQ_OFF_A = 0;                                 // 4 x 16 KB; reused as epilogue staging (16 x 2 KB)
constexpr int LQ_OFF_RING = 65536;
constexpr int LQ_OFF_BAR = LQ_OFF_RING + LQ_NST * LQ_PIECE; // 196608
constexpr int LQ_SMEM = LQ_OFF_BAR + 256;

__global__ void __launch_bounds__(LQ_THREADS, 1)
ln_qkv_kernel(const __grid_constant__ CUtensorMap tmW3, const LnQkvParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LQ_OFF_BAR);
  uint64_t* full_bar = bars;                // [4]
  uint64_t* empty_bar = bars + LQ_NST;      // [4]
  uint64_t* a_ready = bars + 2 * LQ_NST;    // a tile written (16 warps)
  uint64_t* d_full = a_ready + 1;           // accumulator complete
  uint64_t* d_empty = a_ready + 2;          // epilogue has read the accumulator and released the A/staging smem
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  const int m_tiles = (p.M + 127) / 128;

  if (threadIdx.x == 0) {
    for (int i = 0; i < LQ_NST; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(a_ready, 16); mbar_init(d_full, 1); mbar_init(d_empty, 16);
    fence_mbar_init();
    tma_prefetch_desc(&tmW3);
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer: Wqkv pieces (constants: no dependency wait) =====
    if ((int)blockIdx.x < m_tiles) {
      uint32_t it = 0;
      const uint64_t pol = l2_policy_evict_last();
      for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
        for (int pc = 0; pc < 6; ++pc, ++it) {   // piece pc: N tile pc/2 (q, k, v), K pair pc%2
          const uint32_t slot = it % LQ_NST, use = it / LQ_NST;
          mbar_wait(&empty_bar[slot], (use & 1) ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&full_bar[slot], LQ_PIECE);
            if (p.w_hint) tma_load_3d_hint(smem + LQ_OFF_RING + slot * LQ_PIECE, &tmW3, &full_bar[slot], 0, (pc >> 1) * 128, (pc & 1) * 2, pol);
            else tma_load_3d(smem + LQ_OFF_RING + slot * LQ_PIECE, &tmW3, &full_bar[slot], 0, (pc >> 1) * 128, (pc & 1) * 2);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    constexpr uint32_t idesc = umma_idesc_f16(128, 128);
    const uint32_t abuf = smem_u32(smem + LQ_OFF_A);
    const uint32_t ring = smem_u32(smem + LQ_OFF_RING);
    uint32_t it = 0, n_tile = 0;
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++n_tile) {
      mbar_wait(a_ready, n_tile & 1);
      tc_fence_after();
      for (int pc = 0; pc < 6; ++pc, ++it) {
        const uint32_t slot = it % LQ_NST, use = it / LQ_NST;
        mbar_wait(&full_bar[slot], use & 1);
        tc_fence_after();
        const uint64_t da0 = umma_desc_sw128(abuf + (pc & 1) * 2 * 16384), db0 = umma_desc_sw128(ring + slot * LQ_PIECE);
        if (elect_one()) {
#pragma unroll
          for (int sub = 0; sub < 2; ++sub)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)   // descriptor address field in 16-byte units
              umma_f16(tmem_base + (pc >> 1) * 128, da0 + sub * (16384 >> 4) + 2 * kk, db0 + sub * (16384 >> 4) + 2 * kk, idesc,
                       ((pc & 1) | sub | kk) != 0);
          umma_commit(&empty_bar[slot]);
          if (pc == 5) umma_commit(d_full);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 3) {
    // ===================================== transform + epilogue warps ========================
    const int ew = warp - 3;            // 0..15
    const int q4 = warp & 3;            // TMEM lane quarter (epilogue)
    const int cg = ew >> 2;             // column group (epilogue): accumulator columns [cg*96, cg*96+96)
    const int c0 = lane * 8, g = lane >> 2;
    const uint32_t abuf = smem_u32(smem + LQ_OFF_A);
    const uint32_t st = abuf + ew * GEMM_STAGING_BYTES;     // epilogue staging aliases the A tile (MMAs are done by then)
    float gam[8], bet[8], lg[8], lb[8];
    {
      const float4 g0 = *reinterpret_cast<const float4*>(p.gamma + c0), g1 = *reinterpret_cast<const float4*>(p.gamma + c0 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(p.beta + c0), b1 = *reinterpret_cast<const float4*>(p.beta + c0 + 4);
      const float4 l0 = *reinterpret_cast<const float4*>(p.ln_g + c0), l1 = *reinterpret_cast<const float4*>(p.ln_g + c0 + 4);
      const float4 m0 = *reinterpret_cast<const float4*>(p.ln_b + c0), m1 = *reinterpret_cast<const float4*>(p.ln_b + c0 + 4);
      gam[0] = g0.x; gam[1] = g0.y; gam[2] = g0.z; gam[3] = g0.w; gam[4] = g1.x; gam[5] = g1.y; gam[6] = g1.z; gam[7] = g1.w;
      bet[0] = b0.x; bet[1] = b0.y; bet[2] = b0.z; bet[3] = b0.w; bet[4] = b1.x; bet[5] = b1.y; bet[6] = b1.z; bet[7] = b1.w;
      lg[0] = l0.x; lg[1] = l0.y; lg[2] = l0.z; lg[3] = l0.w; lg[4] = l1.x; lg[5] = l1.y; lg[6] = l1.z; lg[7] = l1.w;
      lb[0] = m0.x; lb[1] = m0.y; lb[2] = m0.z; lb[3] = m0.w; lb[4] = m1.x; lb[5] = m1.y; lb[6] = m1.z; lb[7] = m1.w;
    }
    pdl_wait();
    uint32_t n_tile = 0;
    int cur_b = -2;
    float ga[8], be[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { ga[j] = 0.f; be[j] = 0.f; }
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++n_tile) {
      const int r0 = tile * 128;
      // ------------------------------------------------ transform: rows r0 + ew*8 .. +8, two batches of 4
      if (lane == 0) mbar_wait(d_empty, (n_tile & 1) ^ 1);   // previous tile's epilogue is done with the A/staging smem
      __syncwarp();
#pragma unroll
      for (int batch = 0; batch < 2; ++batch) {
        uint4 yv[4], rv[4];
        float m[4];
        int bb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = r0 + ew * 8 + batch * 4 + i;
          yv[i] = make_uint4(0, 0, 0, 0);
          rv[i] = make_uint4(0, 0, 0, 0);
          m[i] = 0.f;
          bb[i] = -1;
          if (row < p.M) {
            bb[i] = p.rowb[row];
            if (bb[i] >= 0) {
              yv[i] = ldg128(p.y + (size_t)row * 256 + c0);
              rv[i] = ldg128(p.res + (size_t)row * 256 + c0);
              m[i] = p.rowmask[row];
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int trow = ew * 8 + batch * 4 + i;   // row inside the tile
          const int row = r0 + trow;
          uint4 o = make_uint4(0, 0, 0, 0), o2 = make_uint4(0, 0, 0, 0);
          if (bb[i] >= 0) {                          // warp-uniform
            if (bb[i] != cur_b) {                    // new utterance: GroupNorm statistics from the conv's partial sums
              cur_b = bb[i];
              const int first = (cur_b * p.Lp) >> 5, last = (cur_b * p.Lp + p.L - 1) >> 5;
              double s = 0.0, ss = 0.0;
              for (int sl = lane & 3; sl <= last - first; sl += 4) {
                const float2 pp = *reinterpret_cast<const float2*>(p.stats_part + ((size_t)cur_b * p.S + sl) * 16 + 2 * g);
                s += (double)pp.x;
                ss += (double)pp.y;
              }
              s += __shfl_xor_sync(0xffffffffu, s, 1);  ss += __shfl_xor_sync(0xffffffffu, ss, 1);
              s += __shfl_xor_sync(0xffffffffu, s, 2);  ss += __shfl_xor_sync(0xffffffffu, ss, 2);
              const double n = 32.0 * (double)p.L;
              const double mean = s / n;
              double var = ss / n - mean * mean;
              if (var < 0.0) var = 0.0;
              const float meanf = (float)mean, rstd = (float)(1.0 / sqrt(var + 1e-5));
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                ga[j] = gam[j] * rstd;
                be[j] = bet[j] - meanf * ga[j];
              }
            }
            float v[8], rr[8];
            float2 f;
            f = unpack_h2(yv[i].x); v[0] = f.x; v[1] = f.y;
            f = unpack_h2(yv[i].y); v[2] = f.x; v[3] = f.y;
            f = unpack_h2(yv[i].z); v[4] = f.x; v[5] = f.y;
            f = unpack_h2(yv[i].w); v[6] = f.x; v[7] = f.y;
            f = unpack_h2(rv[i].x); rr[0] = f.x; rr[1] = f.y;
            f = unpack_h2(rv[i].y); rr[2] = f.x; rr[3] = f.y;
            f = unpack_h2(rv[i].z); rr[4] = f.x; rr[5] = f.y;
            f = unpack_h2(rv[i].w); rr[6] = f.x; rr[7] = f.y;
            float s = 0.f, ss = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[j] = mish_f(fmaf(v[j], ga[j], be[j])) * m[i] + rr[j];
              s += v[j];
              ss = fmaf(v[j], v[j], ss);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
              s += __shfl_xor_sync(0xffffffffu, s, off);
              ss += __shfl_xor_sync(0xffffffffu, ss, off);
            }
            const float lmean = s * (1.f / 256.f);
            const float lrstd = rsqrtf(fmaxf(ss * (1.f / 256.f) - lmean * lmean, 0.f) + 1e-5f);
            float a[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fmaf((v[j] - lmean) * lrstd, lg[j], lb[j]);
            o = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
            o2 = make_uint4(pack_h2(a[0], a[1]), pack_h2(a[2], a[3]), pack_h2(a[4], a[5]), pack_h2(a[6], a[7]));
          }
          if (row < p.M) stg128(p.xr + (size_t)row * 256 + c0, o);   // guard rows: zeros
          // a: channel c0..c0+7 -> K chunk c0/64, 16-byte unit (c0%64)/8 of the row, 128B swizzle
          sts128(abuf + (lane >> 3) * 16384 + trow * 128 + (((lane & 7) ^ (trow & 7)) << 4), o2);
        }
      }
      fence_proxy_async_smem();   // a is read by the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
      // ------------------------------------------------ epilogue: q | k row-major, v transposed per head
      const int rw0 = r0 + q4 * 32;
      const int row = rw0 + lane;
      const int rows_valid = min(32, p.M - rw0);
      const int b = (row < p.M) ? p.rowb[row] : -1;
      const int t = row - b * p.Lp;
      if (lane == 0) mbar_wait(d_full, n_tile & 1);
      __syncwarp();
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q4 * 32) << 16) + cg * 96;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int col = cg * 96 + c * 32;   // accumulator column: [0,128) q, [128,256) k, [256,384) v
        float v[32];
        tmem_ld32(taddr + c * 32, v);
        tmem_ld_wait();
        if (col < 256 || p.v != nullptr) {
          __half* dst = (col < 128 ? p.q : (col < 256 ? p.k : p.v)) + (size_t)rw0 * 128 + (col & 127);
          epi_store_h32(st, lane, v, dst, 128, rows_valid);
        } else if (b >= 0) {
          const int cc = col - 256;         // head cc/64, dim cc%64; lanes = consecutive frames
          __half* dst = p.vt + ((size_t)(b * 2 + (cc >> 6)) * 64 + (cc & 63)) * p.Lpad + t;
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[(size_t)j * p.Lpad] = __float2half_rn(v[j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

}  // namespace mtts
