// Masked self-attention of BasicTransformerBlock (reference model.py:670-705), 2 heads x 64.
//
// One CTA per (128-query tile, head, utterance).  Q/K tiles arrive by TMA (128B swizzle), S = Q K^T
// and O_j = P_j V_j run on tcgen05 with fp32 accumulators in TMEM; the online softmax keeps one
// query row per thread (tcgen05.ld 32x32b), writes the un-normalised probabilities as fp16 into a
// swizzled smem tile that is the A operand of the PV MMA, and rescales a register accumulator.
// q is pre-scaled by head_dim^-0.5 (folded into to_q's packed weight, exact power of two).
//
// Reference quirk reproduced (model.py:697): masked keys are filled with -finfo.min = +3.4e38, so
// an utterance with >= 1 masked key gives EVERY query the uniform mean of V over its MASKED keys;
// an utterance without masked keys gets ordinary softmax attention.  The first case is evaluated
// in closed form (a column mean of V) instead of through the L x L product.
#pragma once
#include <cuda.h>

#include "ptx.cuh"

namespace mtts {

constexpr int ATT_THREADS = 128;
// Q 16K + K 2x16K + V^T 2x16K + P 32K + barriers; two CTAs fit one SM (2 x (112.1 KB + 1 KB) <= 228 KB)
constexpr int ATT_SMEM = 16384 + 2 * 16384 + 2 * 16384 + 32768 + 128;

struct AttnParams {
  int L;      // frames per utterance at this level
  int Lp;     // rows per utterance in the flat row space (L + guard)
  int Lpad;   // row pitch of V^T
  const float* rowmask;  // flat per-row mask (0 on guard rows)
  const int* npad;       // [B] number of frames with mask == 0
  const __half* vt;      // [(b*2+h)*64 + d][Lpad]
  __half* out;           // [rows][128]
  int pdl_late;          // 1: release the dependent launch after the key/value loop instead of at entry
};

__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmVT, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // 128B-swizzled tiles need 1024-byte aligned bases
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 16384;   // 2 buffers
  uint8_t* sV = sK + 32768;   // 2 buffers x (2 boxes of 8 KB)
  uint8_t* sP = sV + 32768;   // 2 K-chunks of 16 KB
  float* s_mean = reinterpret_cast<float*>(sP);  // quirk path only (P is unused there)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 32768);
  uint64_t* bar_kv = bars;      // [2]
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_o = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  if (!p.pdl_late) pdl_launch_dependents();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int rowbase = b * p.Lp;
  const int my_t = q0 + tid;  // query frame handled by this thread

  // barriers / TMEM are set up while the previous kernel drains (both paths; the quirk path frees TMEM again)
  if (tid == 0) {
    mbar_init(&bar_kv[0], 1);
    mbar_init(&bar_kv[1], 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmVT);
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  // ---------------- quirk path: utterance has masked keys -> uniform mean of V over them ----------
  const int npad = p.npad[b];
  if (npad > 0) {
    const __half* vt = p.vt + (size_t)(b * 2 + h) * 64 * p.Lpad;
    for (int d = warp * 16; d < warp * 16 + 16; ++d) {
      float acc = 0.f;
      for (int t = lane; t < p.L; t += 32)
        if (p.rowmask[rowbase + t] == 0.f) acc += __half2float(vt[(size_t)d * p.Lpad + t]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) s_mean[d] = acc / (float)npad;
    }
    __syncthreads();
    if (p.pdl_late) pdl_launch_dependents();
    if (my_t < p.L) {
      uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t)(rowbase + my_t) * 128 + h * 64);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        dst[j] = make_uint4(pack_h2(s_mean[8 * j], s_mean[8 * j + 1]), pack_h2(s_mean[8 * j + 2], s_mean[8 * j + 3]),
                            pack_h2(s_mean[8 * j + 4], s_mean[8 * j + 5]), pack_h2(s_mean[8 * j + 6], s_mean[8 * j + 7]));
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem_base);
    return;
  }

  // ---------------- full path: flash-style softmax(Q K^T) V on tcgen05 ----------------------------
  const uint32_t tS = tmem_base;         // 128 columns
  const uint32_t tO = tmem_base + 128;   // 64 columns
  const uint32_t lane_off = uint32_t(warp * 32) << 16;

  const int nkv = (p.L + 127) / 128;
  auto issue_kv = [&](int j) {
    const int buf = j & 1;
    uint32_t bytes = 16384 + 16384 + (j == 0 ? 16384 : 0);
    mbar_arrive_expect_tx(&bar_kv[buf], bytes);
    if (j == 0) tma_load_2d(sQ, &tmQ, &bar_kv[buf], h * 64, rowbase + q0);
    tma_load_2d(sK + buf * 16384, &tmK, &bar_kv[buf], h * 64, rowbase + j * 128);
    tma_load_2d(sV + buf * 16384, &tmVT, &bar_kv[buf], j * 128, (b * 2 + h) * 64);
    tma_load_2d(sV + buf * 16384 + 8192, &tmVT, &bar_kv[buf], j * 128 + 64, (b * 2 + h) * 64);
  };
  if (warp == 0) {  // converged warp, one elected lane issues (uniform operands)
    if (elect_one()) {
      issue_kv(0);
      if (nkv > 1) issue_kv(1);
    }
    __syncwarp();
  }

  constexpr uint32_t idesc_s = umma_idesc_f16(128, 128);
  constexpr uint32_t idesc_o = umma_idesc_f16(128, 64);
  constexpr float LOG2E = 1.4426950408889634f;
  float m_run = -INFINITY, l_run = 0.f;
  float acc[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) acc[j] = 0.f;

  for (int j = 0; j < nkv; ++j) {
    const int buf = j & 1;
    if (warp == 0) {
      mbar_wait(&bar_kv[buf], (j >> 1) & 1);
      tc_fence_after();
      const uint64_t dq = umma_desc_sw128(smem_u32(sQ));
      const uint64_t dk = umma_desc_sw128(smem_u32(sK + buf * 16384));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tS, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        umma_commit(bar_s);
      }
      __syncwarp();
    }
    mbar_wait(bar_s, j & 1);
    tc_fence_after();

    // ---- online softmax over this tile's 128 keys (keys >= L are excluded) ----
    const int kvalid = min(128, p.L - j * 128);
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      float s[32];
      tmem_ld32(tS + lane_off + c * 32, s);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (c * 32 + i < kvalid) ? s[i] : -INFINITY);
    }
    const float m_new = fmaxf(m_run, mx);
    const float alpha = exp2f((m_run - m_new) * LOG2E);
    const float mb = m_new * LOG2E;
    float rsum = 0.f;
    const int r = tid;  // tile row of this thread
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      float s[32];
      tmem_ld32(tS + lane_off + c * 32, s);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float e = (c * 32 + i < kvalid) ? exp2f(fmaf(s[i], LOG2E, -mb)) : 0.f;
        rsum += e;
        s[i] = e;
      }
      uint8_t* prow = sP + (c >> 1) * 16384 + r * 128;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int unit = (c & 1) * 4 + u;  // 16-byte unit inside the 128-byte row
        *reinterpret_cast<uint4*>(prow + ((unit ^ (r & 7)) << 4)) =
            make_uint4(pack_h2(s[8 * u], s[8 * u + 1]), pack_h2(s[8 * u + 2], s[8 * u + 3]),
                       pack_h2(s[8 * u + 4], s[8 * u + 5]), pack_h2(s[8 * u + 6], s[8 * u + 7]));
      }
    }
    l_run = l_run * alpha + rsum;
    m_run = m_new;
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] *= alpha;

    fence_proxy_async_smem();  // P written with generic-proxy stores, read by the tensor core
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      const uint64_t dp0 = umma_desc_sw128(smem_u32(sP)), dv0 = umma_desc_sw128(smem_u32(sV + buf * 16384));
      if (elect_one()) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k)   // descriptor address field counts 16-byte units
            umma_f16(tO, dp0 + c * (16384 >> 4) + 2 * k, dv0 + c * (8192 >> 4) + 2 * k, idesc_o, (c | k) != 0);
        umma_commit(bar_o);
      }
      __syncwarp();
    }
    mbar_wait(bar_o, j & 1);
    tc_fence_after();
    {
      float o[32];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tmem_ld32(tO + lane_off + c * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[c * 32 + i] += o[i];
      }
    }
    if (warp == 0 && j + 2 < nkv) {  // both MMAs that read buffer `buf` have completed
      if (elect_one()) issue_kv(j + 2);
      __syncwarp();
    }
    tc_fence_before();
    __syncwarp();
  }

  if (p.pdl_late) pdl_launch_dependents();
  if (my_t < p.L) {
    const float inv = 1.f / l_run;
    uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t)(rowbase + my_t) * 128 + h * 64);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      dst[j] = make_uint4(pack_h2(acc[8 * j] * inv, acc[8 * j + 1] * inv), pack_h2(acc[8 * j + 2] * inv, acc[8 * j + 3] * inv),
                          pack_h2(acc[8 * j + 4] * inv, acc[8 * j + 5] * inv), pack_h2(acc[8 * j + 6] * inv, acc[8 * j + 7] * inv));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem_base);
}

}  // namespace mtts
