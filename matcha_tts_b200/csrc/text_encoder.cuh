// Kernels of the native TextEncoder + duration predictor (the step before the hot path, SURVEY.md section 8f row 1;
// reference model.py:148-535) around the tcgen05 implicit GEMMs of gemm_tc.cuh, which run every convolution / linear
// layer of it (k5 prenet convs, 1x1 q|k|v / output projections, k3 FFN and duration-predictor convs).
//
// Layout: tokens in a flat row space like the decoder's frames -- utterance b, token t -> row b*Lx + t, Lx = T_x + 2 zero
// guard rows (a k5 conv reads rows t-2 .. t+2) -- channels-last fp16, 256 columns per row (192 or 256 used, rest zero).
// Every stored activation is zero on padded tokens and guard rows: the reference multiplies by x_mask in front of every
// conv and masks its outputs, and padded tokens never reach a valid one (attention gives their keys exp(-1e4 - max) = 0;
// oracle test_padding_is_inert_and_outputs_are_masked), so masking early changes padded positions only.
#pragma once
#include "ptx.cuh"

namespace mtts {

constexpr int TE_LD = 256;        // columns per activation row
constexpr int TE_GUARD = 2;       // zero rows after every utterance

// ---------------------------------------------------------------------------------------------
// x = emb(tokens) * sqrt(C) (model.py:517), masked; row mask; x_mask output (B, 1, T_x) as floats (model.py:519)
// one warp per row (8 channels per lane), 8 rows per block
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) te_embed_kernel(const long long* __restrict__ tokens, const long long* __restrict__ lengths,
                                                       const float* __restrict__ emb, int n_vocab, int C, float scale, int B, int Tx,
                                                       __half* __restrict__ x, float* __restrict__ rowmask, float* __restrict__ x_mask) {
  pdl_launch_dependents();
  pdl_wait();
  const int Lx = Tx + TE_GUARD;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B * Lx) return;
  const int b = row / Lx, t = row - b * Lx;
  const bool valid = t < Tx && (long long)t < lengths[b];
  uint4 o = make_uint4(0, 0, 0, 0);
  const int c0 = lane * 8;
  if (valid && c0 < C) {
    long long tok = tokens[(size_t)b * Tx + t];
    tok = tok < 0 ? 0 : (tok >= n_vocab ? n_vocab - 1 : tok);
    const float* e = emb + (size_t)tok * C + c0;
    const float4 a = *reinterpret_cast<const float4*>(e), c = *reinterpret_cast<const float4*>(e + 4);
    o = make_uint4(pack_h2(a.x * scale, a.y * scale), pack_h2(a.z * scale, a.w * scale), pack_h2(c.x * scale, c.y * scale),
                   pack_h2(c.z * scale, c.w * scale));
  }
  stg128(x + (size_t)row * TE_LD + c0, o);
  if (lane == 0) {
    rowmask[row] = valid ? 1.f : 0.f;
    if (t < Tx) x_mask[(size_t)b * Tx + t] = valid ? 1.f : 0.f;
  }
}

// speaker embedding broadcast along time into columns [C, C + nspk) of the valid rows (model.py:523-524; the encoder masks
// its input first, :429, so padded rows stay zero)
__global__ void te_spk_kernel(const float* __restrict__ spks, const float* __restrict__ rowmask, int C, int nspk, int B, int Lx,
                              __half* __restrict__ x) {
  pdl_launch_dependents();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * Lx * nspk) return;
  const int row = idx / nspk, j = idx - row * nspk, b = row / Lx;
  x[(size_t)row * TE_LD + C + j] = __float2half_rn(rowmask[row] != 0.f ? spks[(size_t)b * nspk + j] : 0.f);
}

// ---------------------------------------------------------------------------------------------
// channel LayerNorm of the reference (model.py:148-166: mean / biased variance over the C channels of one token, eps 1e-4,
// gamma / beta), optional ReLU after it (prenet, :203-205), row mask; optionally the duration predictor's last layer on top
// (model.py:233-234: logw = (proj(h * m) + bias) * m, a C -> 1 projection) from the fp32 normalised values.
// one warp per row, 8 channels per lane
// ---------------------------------------------------------------------------------------------
struct TeLnParams {
  const __half* in;      // [rows, 256]
  __half* out;           // [rows, 256] or null
  const float* gamma;
  const float* beta;
  const float* rowmask;  // [rows]
  int rows, C, relu;
  const float* proj_w;   // [C] or null
  const float* proj_b;   // [1]
  float* logw;           // (B, 1, Tx)
  int Lx, Tx;
};
__global__ void __launch_bounds__(256) te_ln_kernel(const TeLnParams p) {
  pdl_launch_dependents();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int c0 = lane * 8;
  const bool act = c0 < p.C;
  float g[8], be[8], pw[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { g[j] = 0.f; be[j] = 0.f; pw[j] = 0.f; }
  if (act) {
    const float4 g0 = *reinterpret_cast<const float4*>(p.gamma + c0), g1 = *reinterpret_cast<const float4*>(p.gamma + c0 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(p.beta + c0), b1 = *reinterpret_cast<const float4*>(p.beta + c0 + 4);
    g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
    be[0] = b0.x; be[1] = b0.y; be[2] = b0.z; be[3] = b0.w; be[4] = b1.x; be[5] = b1.y; be[6] = b1.z; be[7] = b1.w;
    if (p.proj_w) {
      const float4 w0 = *reinterpret_cast<const float4*>(p.proj_w + c0), w1 = *reinterpret_cast<const float4*>(p.proj_w + c0 + 4);
      pw[0] = w0.x; pw[1] = w0.y; pw[2] = w0.z; pw[3] = w0.w; pw[4] = w1.x; pw[5] = w1.y; pw[6] = w1.z; pw[7] = w1.w;
    }
  }
  pdl_wait();
  if (row >= p.rows) return;
  const float m = p.rowmask[row];
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
  if (act && m != 0.f) {
    const uint4 u = ldg128(p.in + (size_t)row * TE_LD + c0);
    float2 f;
    f = unpack_h2(u.x); v[0] = f.x; v[1] = f.y;
    f = unpack_h2(u.y); v[2] = f.x; v[3] = f.y;
    f = unpack_h2(u.z); v[4] = f.x; v[5] = f.y;
    f = unpack_h2(u.w); v[6] = f.x; v[7] = f.y;
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += v[j];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  const float mean = s / (float)p.C;
  float q = 0.f;
  if (act) {
#pragma unroll
    for (int j = 0; j < 8; ++j) q = fmaf(v[j] - mean, v[j] - mean, q);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
  const float rstd = rsqrtf(q / (float)p.C + 1e-4f);
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float y = fmaf((v[j] - mean) * rstd, g[j], be[j]);
    if (p.relu) y = fmaxf(y, 0.f);
    y = (act && m != 0.f) ? y : 0.f;
    v[j] = y;
    dot = fmaf(y, pw[j], dot);
  }
  if (p.out) stg128(p.out + (size_t)row * TE_LD + c0, make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7])));
  if (p.proj_w) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
    const int b = row / p.Lx, t = row - b * p.Lx;
    if (lane == 0 && t < p.Tx) p.logw[(size_t)b * p.Tx + t] = (m != 0.f) ? dot + p.proj_b[0] : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------
// rotary tables (model.py:257-272): angle[t][j] = t * theta[j], theta[j] = base^(-2j/d) given by the host; cos / sin in fp32
// ---------------------------------------------------------------------------------------------
__global__ void te_rope_table_kernel(const float* __restrict__ theta, int T, int half, float* __restrict__ cs, float* __restrict__ sn) {
  pdl_launch_dependents();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * half) return;
  const int t = idx / half, j = idx - t * half;
  const float ang = (float)t * theta[j];
  cs[idx] = cosf(ang);
  sn[idx] = sinf(ang);
}

// ---------------------------------------------------------------------------------------------
// multi-head self-attention with rotary position embedding (model.py:335-365): q, k, v from the fused projection
// [rows][768] (q | k | v at columns 0 / 256 / 512, head h at h*c inside each; q already scaled by c^-1/2), RoPE on the first
// c/2 features of q and k, scores masked_fill(-1e4) on padded keys (their probabilities underflow to exactly 0, so they are
// skipped), softmax in fp32, P V -> o [rows][256].  3 % of the encoder's FLOPs: CUDA cores, one CTA per (16 queries, head,
// utterance), keys / values staged as fp16 tiles of 128 keys with an online softmax across tiles.
// ---------------------------------------------------------------------------------------------
constexpr int TE_AQ = 16;        // queries per CTA (4 warps x 4)
constexpr int TE_AK = 128;       // keys per tile
struct TeAttnParams {
  const __half* qkv;            // [rows, 768]
  __half* o;                    // [rows, 256]
  const long long* lengths;     // [B]
  const float* cs;              // [Tx][c/4]
  const float* sn;
  int Tx, Lx, c;                // c = head width (96 or 128)
};
__global__ void __launch_bounds__(128) te_attn_kernel(const TeAttnParams p) {
  extern __shared__ __align__(16) uint8_t te_smem[];
  pdl_launch_dependents();
  const int c = p.c, cp = c + 2, half = c / 4;          // cp: padded row (odd word stride: conflict-free per-lane rows)
  __half* sK = reinterpret_cast<__half*>(te_smem);      // [128][cp]
  __half* sV = sK + TE_AK * cp;                         // [128][cp]
  float* sQ = reinterpret_cast<float*>(sV + TE_AK * cp);   // [16][c]
  float* sP = sQ + TE_AQ * c;                           // [4 warps][128]
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * TE_AQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_wait();
  const int len = (int)min((long long)p.Tx, p.lengths[b]);
  const size_t rowbase = (size_t)b * p.Lx;
  // queries of this CTA (RoPE applied), fp32 in shared memory
  for (int i = threadIdx.x; i < TE_AQ * c; i += 128) {
    const int qi = i / c, d = i - qi * c, t = q0 + qi;
    float v = 0.f;
    if (t < len) {
      const __half* qr = p.qkv + (rowbase + t) * 768 + h * c;
      v = __half2float(qr[d]);
      if (d < 2 * half) {
        const int j = d < half ? d : d - half;
        const float cs = p.cs[t * half + j], sn = p.sn[t * half + j];
        const float other = __half2float(qr[d < half ? d + half : d - half]);
        v = d < half ? v * cs - other * sn : v * cs + other * sn;
      }
    }
    sQ[i] = v;
  }
  float m_run[4], l_run[4], acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { m_run[i] = -INFINITY; l_run[i] = 0.f; acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
  const int nd = c / 32;   // output features per lane (3 or 4)
  for (int k0 = 0; k0 < len; k0 += TE_AK) {
    const int nk = min(TE_AK, len - k0);
    __syncthreads();   // previous tile consumed (and sQ written)
    for (int i = threadIdx.x; i < TE_AK * (c / 2); i += 128) {
      const int kk = i / (c / 2), d2 = (i - kk * (c / 2)) * 2;
      float2 kv = make_float2(0.f, 0.f), vv = make_float2(0.f, 0.f);
      if (kk < nk) {
        const int t = k0 + kk;
        const __half* kr = p.qkv + (rowbase + t) * 768 + 256 + h * c;
        kv = __half22float2(*reinterpret_cast<const __half2*>(kr + d2));
        vv = __half22float2(*reinterpret_cast<const __half2*>(kr + 256 + d2));
        if (d2 < 2 * half) {   // d2 and d2+1 are on the same side of the rotation (half is even)
          const int j = d2 < half ? d2 : d2 - half;
          const float2 ot = __half22float2(*reinterpret_cast<const __half2*>(kr + (d2 < half ? d2 + half : d2 - half)));
          const float c0 = p.cs[t * half + j], s0 = p.sn[t * half + j], c1 = p.cs[t * half + j + 1], s1 = p.sn[t * half + j + 1];
          kv = d2 < half ? make_float2(kv.x * c0 - ot.x * s0, kv.y * c1 - ot.y * s1) : make_float2(kv.x * c0 + ot.x * s0, kv.y * c1 + ot.y * s1);
        }
      }
      *reinterpret_cast<__half2*>(sK + kk * cp + d2) = __floats2half2_rn(kv.x, kv.y);
      *reinterpret_cast<__half2*>(sV + kk * cp + d2) = __floats2half2_rn(vv.x, vv.y);
    }
    __syncthreads();
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) {   // unrolled: the running maxima / sums / accumulators stay in registers
      const int ql = warp * 4 + qi;
      if (q0 + ql >= len) continue;   // warp-uniform
      const float* qv = sQ + ql * c;
      float s[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int kk = lane + 32 * r;
        float a = 0.f;
        const __half2* kr = reinterpret_cast<const __half2*>(sK + kk * cp);
        for (int d2 = 0; d2 < c / 2; ++d2) {
          const float2 kf = __half22float2(kr[d2]);
          a = fmaf(qv[2 * d2], kf.x, a);
          a = fmaf(qv[2 * d2 + 1], kf.y, a);
        }
        s[r] = kk < nk ? a : -INFINITY;
      }
      float mx = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3]));
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m_run[qi], mx);
      const float alpha = __expf(m_run[qi] - m_new);
      float ps = 0.f;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float e = __expf(s[r] - m_new);
        sP[warp * TE_AK + lane + 32 * r] = e;
        ps += e;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
      l_run[qi] = l_run[qi] * alpha + ps;
      m_run[qi] = m_new;
      __syncwarp();
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      for (int kk = 0; kk < nk; ++kk) {
        const float pk = sP[warp * TE_AK + kk];
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (r < nd) o[r] = fmaf(pk, __half2float(sV[kk * cp + lane + 32 * r]), o[r]);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[qi][r] = acc[qi][r] * alpha + o[r];
      __syncwarp();
    }
  }
  // o rows of this CTA's queries; padded queries (and the columns beyond 2c) are zero
#pragma unroll
  for (int qi = 0; qi < 4; ++qi) {
    const int t = q0 + warp * 4 + qi;
    if (t >= p.Tx) continue;
    const float inv = (t < len && l_run[qi] > 0.f) ? 1.f / l_run[qi] : 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (r < nd) p.o[(rowbase + t) * TE_LD + h * c + lane + 32 * r] = __float2half_rn(acc[qi][r] * inv);
  }
}
__host__ inline size_t te_attn_smem(int c) { return (size_t)2 * TE_AK * (c + 2) * 2 + (size_t)TE_AQ * c * 4 + 4 * TE_AK * 4; }

// mu (B, n_feats, Tx) fp32 from the masked fp16 projection rows [rows][256]
__global__ void te_mu_out_kernel(const __half* __restrict__ mu16, int B, int Tx, int Lx, int nf, float* __restrict__ mu) {
  pdl_launch_dependents();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * nf * Tx) return;
  const int t = idx % Tx, f = (idx / Tx) % nf, b = idx / (Tx * nf);
  mu[idx] = __half2float(mu16[((size_t)b * Lx + t) * TE_LD + f]);
}

}  // namespace mtts
