// Bandwidth-bound kernels around the GEMMs: mask/row-map preparation, the first conv's operand
// buffer, GroupNorm-apply + Mish (+time embedding, +residual, +LayerNorm), the data-independent
// time-embedding path, and weight packing.  All activations are channels-last fp16 in the flat
// row space described in gemm_tc.cuh; every thread block handles rows of ONE utterance so the
// GroupNorm statistics (which span all frames of an utterance, padded ones included -- reference
// model.py:769) are finalised once per block from the deterministic partial sums.
#pragma once
#include "ptx.cuh"

namespace mtts {

// ---------------------------------------------------------------------------------------------
// mask (B,T) float -> flat per-row masks / utterance ids for both U-Net levels, masked-key counts
// level T : row b*LpT+t, guard rows t in [T, LpT);  level H = ceil(T/2): row b*LpH+m, guard row m = H,
// mask_H[b,m] = mask[b, 2m]  (reference model.py:1003  mask_down[:, :, ::2])
// ---------------------------------------------------------------------------------------------
__global__ void mask_prep_kernel(const float* __restrict__ mask, int T, int H, int LpT, int LpH, float* __restrict__ mT,
                                 float* __restrict__ mH, int* __restrict__ rowbT, int* __restrict__ rowbH,
                                 int* __restrict__ npadT, int* __restrict__ npadH) {
  const int b = blockIdx.x;
  __shared__ int cT, cH;
  if (threadIdx.x == 0) { cT = 0; cH = 0; }
  __syncthreads();
  int nT = 0, nH = 0;
  for (int t = threadIdx.x; t < LpT; t += blockDim.x) {
    float m = (t < T) ? mask[(size_t)b * T + t] : 0.f;
    mT[(size_t)b * LpT + t] = m;
    rowbT[(size_t)b * LpT + t] = (t < T) ? b : -1;
    if (t < T && m == 0.f) ++nT;
  }
  for (int t = threadIdx.x; t < LpH; t += blockDim.x) {
    float m = (t < H) ? mask[(size_t)b * T + 2 * t] : 0.f;
    mH[(size_t)b * LpH + t] = m;
    rowbH[(size_t)b * LpH + t] = (t < H) ? b : -1;
    if (t < H && m == 0.f) ++nH;
  }
  atomicAdd(&cT, nT);
  atomicAdd(&cH, nH);
  __syncthreads();
  if (threadIdx.x == 0) { npadT[b] = cT; npadH[b] = cH; }
}

// ---------------------------------------------------------------------------------------------
// X0[row][c] = fp16( cat[z, mu, spks][b, c, t] * mask[b,t] ), zero pad channels and guard rows
// (reference model.py:975-979 cat + the x*mask of Block1D / res_conv, :774, :789)
// ---------------------------------------------------------------------------------------------
__global__ void prep_x0_kernel(const float* __restrict__ z, const float* __restrict__ mu,
                               const float* __restrict__ spks, const float* __restrict__ mT, int T, int LpT, int nf,
                               int nspk, int cinp, __half* __restrict__ x0, int z_only) {
  extern __shared__ float tile[];  // [cinp][33]
  const int b = blockIdx.y, t0 = blockIdx.x * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int t = t0 + lane;
  const float m = (t < T) ? mT[(size_t)b * LpT + t] : 0.f;
  const int cend = z_only ? nf : cinp;
  for (int c = warp; c < cend; c += nw) {
    float v = 0.f;
    if (t < T) {
      if (c < nf) v = z[((size_t)b * nf + c) * T + t];
      else if (c < 2 * nf) v = mu[((size_t)b * nf + (c - nf)) * T + t];
      else if (c < 2 * nf + nspk) v = spks[(size_t)b * nspk + (c - 2 * nf)];
    }
    tile[c * 33 + lane] = v * m;
  }
  __syncthreads();
  const int npairs = cend / 2;
  for (int i = threadIdx.x; i < 32 * npairs; i += blockDim.x) {
    const int tl = i / npairs, cp = i % npairs;
    if (t0 + tl < LpT) {
      __half2 hv = __floats2half2_rn(tile[(2 * cp) * 33 + tl], tile[(2 * cp + 1) * 33 + tl]);
      *reinterpret_cast<__half2*>(x0 + ((size_t)b * LpT + t0 + tl) * cinp + 2 * cp) = hv;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// GroupNorm(8, 256) apply + Mish + mask, fused with what follows it in the reference:
//   MODE 0 (Block1D #1 of a ResnetBlock1D, and final_block):
//        h = (Mish(GN(y)) * m + temb[c]) * m                      model.py:773-775, :786-787
//   MODE 1 (Block1D #2):  x_r = Mish(GN(y)) * m + res  (NOT masked, :788-789),  a = LayerNorm1(x_r) :735
// y is the raw conv output (+bias) in fp16; statistics come from the conv epilogue's partial sums.
// grid = (ceil(Lp/16), B), block = 128 (4 warps x 4 rows, every load issued up front, 8 channels per lane)
// ---------------------------------------------------------------------------------------------
struct GnParams {
  const __half* y;
  const float* stats_part;  // [B][S][16] (sum, sumsq) per group
  int S, L, Lp;
  const float* gamma;
  const float* beta;
  const float* rowmask;
  const float* temb;  // [n_t][256] or null
  int t_off, t_stride, t_ld;  // temb row = t_off + b*t_stride, row pitch t_ld floats
  __half* out;        // MODE 0: h ; MODE 1: x_r
  const __half* res;  // MODE 1
  const float* ln_g;
  const float* ln_b;
  __half* out2;  // MODE 1: a
};

constexpr int GN_THREADS = 128;       // 4 warps
constexpr int GN_RPW = 4;             // rows per warp, all in flight at once (2 and 8 measured slower: profiles/r01i_variants.txt)
constexpr int GN_ROWS = 4 * GN_RPW;   // rows per block
template <int MODE>
__global__ void __launch_bounds__(GN_THREADS) gn_apply_kernel(const GnParams p) {
  pdl_launch_dependents();
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = lane * 8, g = lane >> 2;
  // weights do not depend on the previous kernel: fetch them while it drains
  float ga[8], be[8], te[8], lg[8], lb[8];
  {
    const float4 g0 = *reinterpret_cast<const float4*>(p.gamma + c0), g1 = *reinterpret_cast<const float4*>(p.gamma + c0 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(p.beta + c0), b1 = *reinterpret_cast<const float4*>(p.beta + c0 + 4);
    ga[0] = g0.x; ga[1] = g0.y; ga[2] = g0.z; ga[3] = g0.w; ga[4] = g1.x; ga[5] = g1.y; ga[6] = g1.z; ga[7] = g1.w;
    be[0] = b0.x; be[1] = b0.y; be[2] = b0.z; be[3] = b0.w; be[4] = b1.x; be[5] = b1.y; be[6] = b1.z; be[7] = b1.w;
    if (MODE == 1) {
      const float4 l0 = *reinterpret_cast<const float4*>(p.ln_g + c0), l1 = *reinterpret_cast<const float4*>(p.ln_g + c0 + 4);
      const float4 m0 = *reinterpret_cast<const float4*>(p.ln_b + c0), m1 = *reinterpret_cast<const float4*>(p.ln_b + c0 + 4);
      lg[0] = l0.x; lg[1] = l0.y; lg[2] = l0.z; lg[3] = l0.w; lg[4] = l1.x; lg[5] = l1.y; lg[6] = l1.z; lg[7] = l1.w;
      lb[0] = m0.x; lb[1] = m0.y; lb[2] = m0.z; lb[3] = m0.w; lb[4] = m1.x; lb[5] = m1.y; lb[6] = m1.z; lb[7] = m1.w;
    }
  }
  pdl_wait();
  // 1) issue every global load of this block up front (rows, time embedding, statistics partials):
  //    one memory round trip instead of three dependent ones
  constexpr int RPW = GN_RPW;
  const int tw0 = (blockIdx.x * 4 + warp) * RPW;
  uint4 yv[RPW], rv[RPW];
  float m[RPW];
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int t = tw0 + i;
    const size_t row = (size_t)b * p.Lp + t;
    yv[i] = make_uint4(0, 0, 0, 0);
    rv[i] = make_uint4(0, 0, 0, 0);
    m[i] = 0.f;
    if (t < p.L) {
      yv[i] = ldg128(p.y + row * 256 + c0);
      if (MODE == 1) rv[i] = ldg128(p.res + row * 256 + c0);
      m[i] = p.rowmask[row];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) te[j] = 0.f;
  if (MODE == 0 && p.temb) {
    const float* tp = p.temb + (size_t)(p.t_off + b * p.t_stride) * p.t_ld + c0;
    const float4 t0 = *reinterpret_cast<const float4*>(tp), t1 = *reinterpret_cast<const float4*>(tp + 4);
    te[0] = t0.x; te[1] = t0.y; te[2] = t0.z; te[3] = t0.w; te[4] = t1.x; te[5] = t1.y; te[6] = t1.z; te[7] = t1.w;
  }
  // 2) finalise the GroupNorm statistics of utterance b: the 4 lanes that share a group sum its per-32-row partials
  //    in a fixed order (deterministic) and combine by shuffles -- per warp, no shared memory, no block barrier.
  //    fp32 is enough: the partials are fp32 sums already and |mean| is O(std) for these conv outputs.
  float mean, rstd;
  {
    const int first = (b * p.Lp) >> 5, last = (b * p.Lp + p.L - 1) >> 5;
    float s = 0.f, ss = 0.f;
    for (int sl = lane & 3; sl <= last - first; sl += 4) {
      const float2 pp = *reinterpret_cast<const float2*>(p.stats_part + ((size_t)b * p.S + sl) * 16 + 2 * g);
      s += pp.x;
      ss += pp.y;
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);  ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);  ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    const float inv_n = 1.f / (32.f * (float)p.L);
    mean = s * inv_n;
    rstd = rsqrtf(fmaxf(fmaf(-mean, mean, ss * inv_n), 0.f) + 1e-5f);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ga[j] *= rstd;
    be[j] = be[j] - mean * ga[j];
  }
  // 3) normalise, Mish, mask (+temb | +res, LayerNorm), store
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int t = tw0 + i;
    if (t >= p.Lp) continue;   // warp-uniform
    const size_t row = (size_t)b * p.Lp + t;
    uint4 o = make_uint4(0, 0, 0, 0), o2 = make_uint4(0, 0, 0, 0);
    if (t < p.L) {
      float v[8];
      float2 f;
      f = unpack_h2(yv[i].x); v[0] = f.x; v[1] = f.y;
      f = unpack_h2(yv[i].y); v[2] = f.x; v[3] = f.y;
      f = unpack_h2(yv[i].z); v[4] = f.x; v[5] = f.y;
      f = unpack_h2(yv[i].w); v[6] = f.x; v[7] = f.y;
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (mish_f(fmaf(v[j], ga[j], be[j])) * m[i] + te[j]) * m[i];
      } else {
        float r[8];
        f = unpack_h2(rv[i].x); r[0] = f.x; r[1] = f.y;
        f = unpack_h2(rv[i].y); r[2] = f.x; r[3] = f.y;
        f = unpack_h2(rv[i].z); r[4] = f.x; r[5] = f.y;
        f = unpack_h2(rv[i].w); r[6] = f.x; r[7] = f.y;
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j] = mish_f(fmaf(v[j], ga[j], be[j])) * m[i] + r[j];
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
        o2 = make_uint4(pack_h2(a[0], a[1]), pack_h2(a[2], a[3]), pack_h2(a[4], a[5]), pack_h2(a[6], a[7]));
      }
      o = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
    }
    stg128(p.out + row * 256 + c0, o);  // guard rows are written as zeros
    if (MODE == 1) stg128(p.out2 + row * 256 + c0, o2);
  }
}

// ---------------------------------------------------------------------------------------------
// The same pass with the rows staged through shared memory by ONE bulk copy per block (default).
// gn_apply_kernel above keeps its operands in registers while they are in flight: 4 rows x 512 B per warp at ~18 warps
// per SM (110 registers) = 36 KB per SM, against the ~45 KB per SM that 6.5 TB/s x ~1 us of loaded HBM latency needs --
// it ran at 3.6-4.3 TB/s (profiles/r02v_ncu_stage01_summary.txt).  Here a block of GN2_ROWS rows of one utterance issues
// cp.async.bulk for its whole 32 KB slab of y (and of res, MODE 1) the moment the previous kernel has completed, finalises
// the statistics and fetches its per-row operands while the bytes fly, and then streams the rows out of shared memory:
// 32-64 KB in flight per block and several blocks per SM, no registers tied up.  The arithmetic is gn_apply_kernel's,
// operation for operation (same bits).
// grid = (ceil(Lp / GN2_ROWS), B), block = 256 (8 warps x GN2_ROWS / 8 rows, 8 channels per lane).  Which of the two kernels a
// launch gets is decided in mtts_api.cu::launch_gn (this one where throughput counts, the register-staged one for one small
// solve at a time: its chain is shorter -- no barrier, no TMA round trip; 32-row blocks measured between the two)
// ---------------------------------------------------------------------------------------------
constexpr int GN2_THREADS = 256;
template <int MODE, int GN2_ROWS>
constexpr int gn2_smem_bytes() { return 128 + GN2_ROWS * 512 * (MODE == 1 ? 2 : 1); }

template <int MODE, int GN2_ROWS>
__global__ void __launch_bounds__(GN2_THREADS) gn_apply2_kernel(const GnParams p) {
  extern __shared__ __align__(128) uint8_t gsm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(gsm);
  uint8_t* sy = gsm + 128;
  uint8_t* sr = sy + GN2_ROWS * 512;
  pdl_launch_dependents();
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = lane * 8, g = lane >> 2;
  const int t0 = blockIdx.x * GN2_ROWS;
  const int nload = min(GN2_ROWS, p.L - t0);   // rows of this block that exist in y (the rest, up to Lp, are guard rows)
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  // weights do not depend on the previous kernel: fetch them while it drains
  float ga[8], be[8], te[8], lg[8], lb[8];
  {
    const float4 g0 = *reinterpret_cast<const float4*>(p.gamma + c0), g1 = *reinterpret_cast<const float4*>(p.gamma + c0 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(p.beta + c0), b1 = *reinterpret_cast<const float4*>(p.beta + c0 + 4);
    ga[0] = g0.x; ga[1] = g0.y; ga[2] = g0.z; ga[3] = g0.w; ga[4] = g1.x; ga[5] = g1.y; ga[6] = g1.z; ga[7] = g1.w;
    be[0] = b0.x; be[1] = b0.y; be[2] = b0.z; be[3] = b0.w; be[4] = b1.x; be[5] = b1.y; be[6] = b1.z; be[7] = b1.w;
    if (MODE == 1) {
      const float4 l0 = *reinterpret_cast<const float4*>(p.ln_g + c0), l1 = *reinterpret_cast<const float4*>(p.ln_g + c0 + 4);
      const float4 m0 = *reinterpret_cast<const float4*>(p.ln_b + c0), m1 = *reinterpret_cast<const float4*>(p.ln_b + c0 + 4);
      lg[0] = l0.x; lg[1] = l0.y; lg[2] = l0.z; lg[3] = l0.w; lg[4] = l1.x; lg[5] = l1.y; lg[6] = l1.z; lg[7] = l1.w;
      lb[0] = m0.x; lb[1] = m0.y; lb[2] = m0.z; lb[3] = m0.w; lb[4] = m1.x; lb[5] = m1.y; lb[6] = m1.z; lb[7] = m1.w;
    }
  }
  __syncthreads();   // the barrier is initialised before anyone polls it
  pdl_wait();
  if (threadIdx.x == 0 && nload > 0) {
    const size_t row0 = (size_t)b * p.Lp + t0;
    const uint32_t bytes = (uint32_t)nload * 512u;
    mbar_arrive_expect_tx(bar, bytes * (MODE == 1 ? 2u : 1u));
    bulk_load_1d(sy, p.y + row0 * 256, bytes, bar);
    if (MODE == 1) bulk_load_1d(sr, p.res + row0 * 256, bytes, bar);
  }
  // per-row masks, time embedding and statistics partials travel while the slab does
  constexpr int RPW = GN2_ROWS / 8;
  const int tw0 = t0 + warp * RPW;
  float m[RPW];
#pragma unroll
  for (int i = 0; i < RPW; ++i) m[i] = (tw0 + i < p.L) ? p.rowmask[(size_t)b * p.Lp + tw0 + i] : 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) te[j] = 0.f;
  if (MODE == 0 && p.temb) {
    const float* tp = p.temb + (size_t)(p.t_off + b * p.t_stride) * p.t_ld + c0;
    const float4 t0v = *reinterpret_cast<const float4*>(tp), t1v = *reinterpret_cast<const float4*>(tp + 4);
    te[0] = t0v.x; te[1] = t0v.y; te[2] = t0v.z; te[3] = t0v.w; te[4] = t1v.x; te[5] = t1v.y; te[6] = t1v.z; te[7] = t1v.w;
  }
  float mean, rstd;
  {
    const int first = (b * p.Lp) >> 5, last = (b * p.Lp + p.L - 1) >> 5;
    float s = 0.f, ss = 0.f;
    for (int sl = lane & 3; sl <= last - first; sl += 4) {
      const float2 pp = *reinterpret_cast<const float2*>(p.stats_part + ((size_t)b * p.S + sl) * 16 + 2 * g);
      s += pp.x;
      ss += pp.y;
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);  ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);  ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    const float inv_n = 1.f / (32.f * (float)p.L);
    mean = s * inv_n;
    rstd = rsqrtf(fmaxf(fmaf(-mean, mean, ss * inv_n), 0.f) + 1e-5f);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ga[j] *= rstd;
    be[j] = be[j] - mean * ga[j];
  }
  if (nload > 0) mbar_wait(bar, 0);
  const uint32_t sya = smem_u32(sy) + c0 * 2, sra = smem_u32(sr) + c0 * 2;
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int t = tw0 + i;
    if (t >= p.Lp) break;   // warp-uniform
    const size_t row = (size_t)b * p.Lp + t;
    uint4 o = make_uint4(0, 0, 0, 0), o2 = make_uint4(0, 0, 0, 0);
    if (t < p.L) {
      const uint4 yv = lds128(sya + (t - t0) * 512);
      float v[8];
      float2 f;
      f = unpack_h2(yv.x); v[0] = f.x; v[1] = f.y;
      f = unpack_h2(yv.y); v[2] = f.x; v[3] = f.y;
      f = unpack_h2(yv.z); v[4] = f.x; v[5] = f.y;
      f = unpack_h2(yv.w); v[6] = f.x; v[7] = f.y;
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (mish_f(fmaf(v[j], ga[j], be[j])) * m[i] + te[j]) * m[i];
      } else {
        const uint4 rv = lds128(sra + (t - t0) * 512);
        float r[8];
        f = unpack_h2(rv.x); r[0] = f.x; r[1] = f.y;
        f = unpack_h2(rv.y); r[2] = f.x; r[3] = f.y;
        f = unpack_h2(rv.z); r[4] = f.x; r[5] = f.y;
        f = unpack_h2(rv.w); r[6] = f.x; r[7] = f.y;
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j] = mish_f(fmaf(v[j], ga[j], be[j])) * m[i] + r[j];
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
        o2 = make_uint4(pack_h2(a[0], a[1]), pack_h2(a[2], a[3]), pack_h2(a[4], a[5]), pack_h2(a[6], a[7]));
      }
      o = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
    }
    stg128(p.out + row * 256 + c0, o);  // guard rows are written as zeros
    if (MODE == 1) stg128(p.out2 + row * 256 + c0, o2);
  }
}

// ---------------------------------------------------------------------------------------------
// time path (data independent, reference model.py:753-762, :828-832, :780)
// ---------------------------------------------------------------------------------------------
// t values of the fixed-step solver: t_i = i/n in double, rounded to fp32 like the reference's
// torch.tensor([i / n_timesteps]) (:1091); midpoint adds fp32(dt)*0.5 in fp32 (:1101).
__global__ void solver_times_kernel(float* __restrict__ tv, int n, int midpoint) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float t = (float)((double)i / (double)n);
  if (!midpoint) {
    tv[i] = t;
  } else {
    const float dt = (float)(1.0 / (double)n);
    tv[2 * i] = t;
    tv[2 * i + 1] = t + dt * 0.5f;
  }
}
// e[i][j] = sin(1000 t_i w_j), e[i][half+j] = cos(1000 t_i w_j); w_j supplied by the host exactly
// as the reference computes it.
__global__ void sinus_emb_kernel(const float* __restrict__ tv, const float* __restrict__ freqs, int n_t, int half,
                                 float* __restrict__ e) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_t * half) return;
  const int i = idx / half, j = idx % half;
  const float arg = (1000.f * tv[i]) * freqs[j];
  e[(size_t)i * 2 * half + j] = sinf(arg);
  e[(size_t)i * 2 * half + half + j] = cosf(arg);
}
// out[i][n] = act( b[n] + sum_k W[n][k] in[i][k] ), one warp per output feature, fp32.
// act: 0 none, 1 SiLU, 2 Mish
__global__ void __launch_bounds__(256) small_linear_kernel(const float* __restrict__ in, const float* __restrict__ W,
                                                            const float* __restrict__ bias, float* __restrict__ out,
                                                            int n_t, int K, int N, int act) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  float w[32];  // K <= 1024
#pragma unroll
  for (int j = 0; j < 32; ++j) w[j] = (lane + 32 * j < K) ? W[(size_t)n * K + lane + 32 * j] : 0.f;
  for (int i = 0; i < n_t; ++i) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (lane + 32 * j < K) acc = fmaf(w[j], in[(size_t)i * K + lane + 32 * j], acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) {
      float v = acc + bias[n];
      if (act == 1) v = v / (1.f + expf(-v));
      else if (act == 2) v = v * tanhf(v > 20.f ? v : log1pf(expf(v)));
      out[(size_t)i * N + n] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// weight packing (once, at load time)
// ---------------------------------------------------------------------------------------------
// dst[(n + n_off)*ldd + k_off + c] = fp16(scale * src[n*sn + c*sc + off])
__global__ void pack2d_kernel(const float* __restrict__ src, __half* __restrict__ dst, int N, int C, long sn, long sc,
                              long off, int n_off, int ldd, int k_off, float scale) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)N * C) return;
  const int n = (int)(i / C), c = (int)(i % C);
  dst[(size_t)(n + n_off) * ldd + k_off + c] = __float2half_rn(scale * src[n * sn + c * sc + off]);
}
// mode 0: scale * x; 1: exp(x); 2: 1/(exp(x)+1e-9)
__global__ void packf_kernel(const float* __restrict__ src, float* __restrict__ dst, int n, int mode, float scale) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = src[i];
  if (mode == 1) v = expf(v);
  else if (mode == 2) v = 1.0f / (expf(v) + 1e-9f);
  else v = scale * v;
  dst[i] = v;
}

}  // namespace mtts
