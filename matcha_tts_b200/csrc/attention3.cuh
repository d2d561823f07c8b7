// Masked self-attention of BasicTransformerBlock (reference model.py:670-705), 2 heads x 64.
//
// Work item = (128-query tile, head, utterance); persistent CTAs, two per SM, walk the items (item = blockIdx.x + k *
// gridDim.x): tensor memory, barriers and descriptors are set up once per CTA, and the Q / K tile of the NEXT item is
// requested as soon as the last Q K^T of the current one has completed (its V tile as soon as the last P V has), so the
// loads of an item overlap the softmax / P V / normalisation of its predecessor (measured: 57.9 -> 55.7 us at level T,
// B = 256 -- the second CTA of the SM already hid most of it; profiles/r02r_launch_table_B256_T344.txt).
//   * key/value tiles are up to 192 keys wide (KT = the utterance's frames split evenly, a multiple of 16): T/2-level
//     utterances (172 frames at T=344) need ONE tile -- plain softmax, no online rescaling -- and level-T ones two;
//   * V stays row-major [rows][64] like K: the P V product reads it as an MN-major B operand (instruction-descriptor
//     bit 16), 16 key rows per K16 step;
//   * K and V are single-buffered: the next K tile is fetched as soon as S = Q K^T has completed, the next V tile as
//     soon as O_j = P_j V_j has, so the loads overlap the softmax of the current tile (112 KB -> two CTAs per SM).
// S and O_j accumulate in TMEM (192 + 64 columns); the un-normalised probabilities go as fp16 into 128B-swizzled smem
// tiles (A operand of P V) and a register accumulator is rescaled between tiles.  q is pre-scaled by head_dim^-0.5
// (folded into to_q's packed weight).  The softmax is spread over TWO threads per query row (one thread per row
// executed 14 instructions per score at 1.8 IPC: issue-bound, profiles/r01i_ncu_attention3_first.txt):
//   * 256 threads: warps w and w+4 share TMEM lane quarter w%4; warp half 0 owns the S columns [0, c0), half 1 the
//     columns [c0, KT) (c0 = the multiple of 32 nearest to KT/2 from above), and the O columns [0,32) / [32,64);
//   * the row maximum is combined through 512 bytes of shared memory (half 0 writes, half 1 merges and writes back),
//     the row sums stay per-thread partials until the end (combined once, through the then idle P buffer);
//   * warps whose 32 query rows all lie beyond the utterance (the second query tile at level T/2 holds 44 of 128 rows)
//     skip the softmax passes and only keep the barriers.
//
// Reference quirk reproduced (model.py:697): masked keys are filled with -finfo.min = +3.4e38, so an utterance with
// >= 1 masked key gives EVERY query the uniform mean of V over its MASKED keys (closed form below); an utterance
// without masked keys gets ordinary softmax attention.
#pragma once
#include <cuda.h>

#include "ptx.cuh"

namespace mtts {

constexpr int ATT2_KT_MAX = 192;
// Q 16 KB | K 24 KB | V 24 KB | P 3 x 16 KB | barriers: two CTAs fit one SM (2 x (112.1 KB + 1 KB) <= 228 KB)
constexpr int ATT2_OFF_K = 16384;
constexpr int ATT2_OFF_V = ATT2_OFF_K + ATT2_KT_MAX * 128;
constexpr int ATT2_OFF_P = ATT2_OFF_V + ATT2_KT_MAX * 128;
constexpr int ATT2_OFF_BAR = ATT2_OFF_P + 3 * 16384;

struct Attn2Params {
  int L;      // frames per utterance at this level
  int Lp;     // rows per utterance in the flat row space (L + guard)
  int KT;     // keys per tile (multiple of 16, <= 192); the K / V tensor maps have KT-row boxes
  int nkv;    // key tiles per utterance: ceil(L / KT)
  const float* rowmask;  // flat per-row mask (0 on guard rows)
  const int* npad;       // [B] number of frames with mask == 0
  const __half* v;       // [rows][128] (quirk path)
  __half* out;           // [rows][128]
  int pdl_late;          // 1: release the dependent launch after the key/value loop instead of at entry
  int B;                 // utterances (items = B * 2 heads * ceil(L / 128) query tiles)
};

// instruction descriptor with an MN-major B operand (V: keys x dims, dims contiguous)
__host__ __device__ constexpr uint32_t umma_idesc_f16_bmn(uint32_t M, uint32_t N) {
  return umma_idesc_f16(M, N) | (1u << 16);
}

constexpr int ATT3_THREADS = 256;
constexpr int ATT3_OFF_X = ATT2_OFF_BAR + 64;          // [128] floats: row-maximum exchange
constexpr int ATT3_SMEM = ATT3_OFF_X + 512;
static_assert(2 * (ATT3_SMEM + 1024) <= 233472, "two attention CTAs per SM");


// one chunk of N S columns held in registers.  MASK = false: every column is a valid key (no per-element predicates);
// MASK = true: columns >= lim are excluded
template <int N, bool MASK>
__device__ __forceinline__ float att3_chunk_max(const float* s, int lim, float mx) {
  if constexpr (MASK) {
#pragma unroll
    for (int i = 0; i < N; ++i) mx = fmaxf(mx, (i < lim) ? s[i] : -INFINITY);
  } else {
#pragma unroll
    for (int i = 0; i < N; i += 2) mx = fmaxf(mx, fmaxf(s[i], s[i + 1]));
  }
  return mx;
}
// s <- exp2(s * log2e - mb) (0 on excluded columns), returns the chunk's sum; the N/8 16-byte units go to the P row
template <int N, bool MASK>
__device__ __forceinline__ float att3_chunk_exp(float* s, int lim, float mb, uint8_t* prow, int unit0, int rx) {
  constexpr float LOG2E = 1.4426950408889634f;
  float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
  for (int i = 0; i < N; i += 2) {
    float e0 = exp2f(fmaf(s[i], LOG2E, -mb)), e1 = exp2f(fmaf(s[i + 1], LOG2E, -mb));
    if constexpr (MASK) { e0 = (i < lim) ? e0 : 0.f; e1 = (i + 1 < lim) ? e1 : 0.f; }
    rs0 += e0; rs1 += e1;
    s[i] = e0; s[i + 1] = e1;
  }
#pragma unroll
  for (int u = 0; u < N / 8; ++u)
    *reinterpret_cast<uint4*>(prow + (((unit0 + u) ^ rx) << 4)) =
        make_uint4(pack_h2(s[8 * u], s[8 * u + 1]), pack_h2(s[8 * u + 2], s[8 * u + 3]),
                   pack_h2(s[8 * u + 4], s[8 * u + 5]), pack_h2(s[8 * u + 6], s[8 * u + 7]));
  return rs0 + rs1;
}

__global__ void __launch_bounds__(ATT3_THREADS, 2)
attention3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const Attn2Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // 128B-swizzled tiles need 1024-byte aligned bases
  uint8_t* sQ = smem;
  uint8_t* sK = smem + ATT2_OFF_K;
  uint8_t* sV = smem + ATT2_OFF_V;
  uint8_t* sP = smem + ATT2_OFF_P;   // ceil(KT/64) K-chunks of 16 KB
  float* s_mean = reinterpret_cast<float*>(sP);  // quirk path only (P is unused there): [4][64]
  float* s_l = reinterpret_cast<float*>(sP);     // row-sum exchange after the last tile: [2][128]
  float* s_mx = reinterpret_cast<float*>(smem + ATT3_OFF_X);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT2_OFF_BAR);
  uint64_t* bar_k = bars;       // Q (first tile) + K tile landed
  uint64_t* bar_v = bars + 1;   // V tile landed
  uint64_t* bar_s = bars + 2;   // S complete
  uint64_t* bar_o = bars + 3;   // O_j complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  if (!p.pdl_late) pdl_launch_dependents();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qd = warp & 3;     // TMEM lane quarter
  const int hf = warp >> 2;    // column half
  const int r = qd * 32 + lane;  // tile row of this thread

  if (tid == 0) {
    mbar_init(bar_k, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const uint32_t tS = tmem_base;                  // KT columns (<= 192)
  const uint32_t tO = tmem_base + ATT2_KT_MAX;    // 64 columns
  const uint32_t lane_off = uint32_t(qd * 32) << 16;
  const int KT = p.KT, nkv = p.nkv;
  const uint32_t kv_bytes = (uint32_t)KT * 128u;
  const int nqt = (p.L + 127) >> 7;
  const int nitems = p.B * 2 * nqt;
  // item -> (query tile fastest, then head, then utterance): a CTA's neighbours in the grid read the same K / V from L2
  auto item_b = [&](int it) { return it / (2 * nqt); };
  auto issue_k = [&](int it, int j) {   // one elected lane of warp 0; tile 0 brings the item's Q along
    const int b = item_b(it), rem = it - b * 2 * nqt, h = rem / nqt, q0 = (rem - h * nqt) * 128;
    mbar_arrive_expect_tx(bar_k, kv_bytes + (j == 0 ? 16384u : 0u));
    if (j == 0) tma_load_2d(sQ, &tmQ, bar_k, h * 64, b * p.Lp + q0);
    tma_load_2d(sK, &tmK, bar_k, h * 64, b * p.Lp + j * KT);
  };
  auto issue_v = [&](int it, int j) {
    const int b = item_b(it), rem = it - b * 2 * nqt, h = rem / nqt;
    mbar_arrive_expect_tx(bar_v, kv_bytes);
    tma_load_2d(sV, &tmV, bar_v, h * 64, b * p.Lp + j * KT);
  };

  const uint32_t idesc_s = umma_idesc_f16(128, (uint32_t)KT);
  constexpr uint32_t idesc_o = umma_idesc_f16_bmn(128, 64);
  constexpr float LOG2E = 1.4426950408889634f;
  // this thread's S columns [cb, cb + nc): nc is a multiple of 16, cb a multiple of 32
  const int c0 = min(KT, ((KT + 63) >> 6) << 5);
  const int cb = hf ? c0 : 0;
  const int nc = hf ? KT - c0 : c0;
  const int n32 = nc >> 5, rem16 = (nc >> 4) & 1;

  uint32_t n_use = 0;        // key / value tiles processed so far: parity of the four barriers
  bool pre = false;          // the first tiles of the current item were requested while the previous item ran
#pragma unroll 1
  for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
    const int b = item_b(it), rem = it - b * 2 * nqt, h = rem / nqt, q0 = (rem - h * nqt) * 128;
    const int rowbase = b * p.Lp;
    const int my_t = q0 + r;     // query frame handled by this thread (together with thread tid ^ 128)
    const int nxt = it + (int)gridDim.x;
    const bool last_item = nxt >= nitems;
    // written by the solve's prologue, several launches back
    const int npad = p.npad[b];
    const bool nxt_full = !last_item && p.npad[item_b(nxt)] == 0;

    // ---------------- quirk path: utterance has masked keys -> uniform mean of V over them ----------
    if (npad > 0) {
      const int d = tid & 63, part = tid >> 6;
      const __half* vp = p.v + (size_t)rowbase * 128 + h * 64 + d;
      float acc = 0.f;
      for (int t = part; t < p.L; t += 4)
        if (p.rowmask[rowbase + t] == 0.f) acc += __half2float(vp[(size_t)t * 128]);
      __syncthreads();   // the previous item's readers of the P region are done
      s_mean[part * 64 + d] = acc;
      __syncthreads();
      if (tid < 64) s_mean[tid] = ((s_mean[tid] + s_mean[64 + tid]) + (s_mean[128 + tid] + s_mean[192 + tid])) / (float)npad;
      __syncthreads();
      if (last_item && p.pdl_late) pdl_launch_dependents();
      if (my_t < p.L) {
        uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t)(rowbase + my_t) * 128 + h * 64 + hf * 32);
        const float* sm = s_mean + hf * 32;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst[j] = make_uint4(pack_h2(sm[8 * j], sm[8 * j + 1]), pack_h2(sm[8 * j + 2], sm[8 * j + 3]),
                              pack_h2(sm[8 * j + 4], sm[8 * j + 5]), pack_h2(sm[8 * j + 6], sm[8 * j + 7]));
      }
      __syncthreads();   // s_mean is read before the next item reuses the region
      pre = false;
      continue;
    }

    // ---------------- full path: softmax(Q K^T) V on tcgen05 ------------------------------------------
    if (!pre && warp == 0) {  // converged warp, one elected lane issues (uniform operands)
      if (elect_one()) {
        issue_k(it, 0);
        issue_v(it, 0);
      }
      __syncwarp();
    }
    const bool wvalid = q0 + qd * 32 < p.L;   // warp-uniform: at least one query row of this warp exists
    float m_run = -INFINITY, l_run = 0.f;     // l_run: partial row sum over this thread's columns
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.f;

    for (int j = 0; j < nkv; ++j, ++n_use) {
      const uint32_t ph = n_use & 1;
      if (warp == 0) {
        mbar_wait(bar_k, ph);
        tc_fence_after();
        const uint64_t dq = umma_desc_sw128(smem_u32(sQ));
        const uint64_t dk = umma_desc_sw128(smem_u32(sK));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tS, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
          umma_commit(bar_s);
        }
        __syncwarp();
      }
      mbar_wait(bar_s, ph);
      tc_fence_after();
      if (warp == 0) {  // K (and after the last tile Q) is free: fetch what comes next while this tile goes through the softmax
        if (elect_one()) {
          if (j + 1 < nkv) issue_k(it, j + 1);
          else if (nxt_full) issue_k(nxt, 0);
        }
        __syncwarp();
      }

      // ---- softmax over this tile's keys (keys >= L are excluded) ----
      const int kvalid = min(KT, p.L - j * KT) - cb;   // valid keys among this thread's columns (may be <= 0)
      const uint32_t ts = tS + lane_off + cb;
      float mx = -INFINITY;
      if (wvalid) {
#pragma unroll 1
        for (int c = 0; c < n32; ++c) {
          float s[32];
          tmem_ld32(ts + c * 32, s);
          tmem_ld_wait();
          const int lim = kvalid - c * 32;
          if (lim >= 32) mx = att3_chunk_max<32, false>(s, lim, mx);
          else mx = att3_chunk_max<32, true>(s, lim, mx);
        }
        if (rem16) {
          float s[16];
          tmem_ld16(ts + n32 * 32, s);
          tmem_ld_wait();
          const int lim = kvalid - n32 * 32;
          if (lim >= 16) mx = att3_chunk_max<16, false>(s, lim, mx);
          else mx = att3_chunk_max<16, true>(s, lim, mx);
        }
      }
      // row maximum over both column halves (half 0 always holds a valid key: its columns start at key 0 of the tile)
      if (hf == 0) s_mx[r] = mx;
      __syncthreads();
      if (hf == 1) { mx = fmaxf(mx, s_mx[r]); s_mx[r] = mx; }
      __syncthreads();
      if (hf == 0) mx = s_mx[r];
      const float m_new = fmaxf(m_run, mx);
      const float alpha = exp2f((m_run - m_new) * LOG2E);
      const float mb = m_new * LOG2E;
      float rsum = 0.f;
      if (wvalid) {
        const int rx = r & 7;
#pragma unroll 1
        for (int c = 0; c < n32; ++c) {
          float s[32];
          tmem_ld32(ts + c * 32, s);
          tmem_ld_wait();
          const int gc = (cb >> 5) + c;  // 32-column chunk of the tile -> K-chunk gc/2 of P, 16-byte units (gc%2)*4.. of the 128-byte row
          uint8_t* prow = sP + (gc >> 1) * 16384 + r * 128;
          const int lim = kvalid - c * 32;
          if (lim >= 32) rsum += att3_chunk_exp<32, false>(s, lim, mb, prow, (gc & 1) * 4, rx);
          else rsum += att3_chunk_exp<32, true>(s, lim, mb, prow, (gc & 1) * 4, rx);
        }
        if (rem16) {
          float s[16];
          tmem_ld16(ts + n32 * 32, s);
          tmem_ld_wait();
          const int gc = (cb >> 5) + n32;
          uint8_t* prow = sP + (gc >> 1) * 16384 + r * 128;
          const int lim = kvalid - n32 * 32;
          if (lim >= 16) rsum += att3_chunk_exp<16, false>(s, lim, mb, prow, (gc & 1) * 4, rx);
          else rsum += att3_chunk_exp<16, true>(s, lim, mb, prow, (gc & 1) * 4, rx);
        }
      }
      l_run = l_run * alpha + rsum;
      m_run = m_new;
      if (j > 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] *= alpha;
      }

      fence_proxy_async_smem();  // P written with generic-proxy stores, read by the tensor core
      tc_fence_before();
      __syncthreads();
      if (warp == 0) {
        mbar_wait(bar_v, ph);
        tc_fence_after();
        const uint64_t dp0 = umma_desc_sw128(smem_u32(sP)), dv0 = umma_desc_sw128(smem_u32(sV));
        const int ksteps = KT >> 4;
        if (elect_one()) {
          for (int k = 0; k < ksteps; ++k)   // P: K-major, 16 keys = 32 B inside the 128-byte row of K-chunk k/4;
                                             // V: MN-major, 16 keys = 16 rows of 128 B = 2048 B (descriptor address in 16-byte units)
            umma_f16(tO, dp0 + (k >> 2) * (16384 >> 4) + 2 * (k & 3), dv0 + k * (2048 >> 4), idesc_o, k != 0);
          umma_commit(bar_o);
        }
        __syncwarp();
      }
      mbar_wait(bar_o, ph);
      tc_fence_after();
      if (warp == 0) {  // V (and P) are free
        if (elect_one()) {
          if (j + 1 < nkv) issue_v(it, j + 1);
          else if (nxt_full) issue_v(nxt, 0);
        }
        __syncwarp();
      }
      if (wvalid) {
        float o[32];
        tmem_ld32(tO + lane_off + hf * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] += o[i];
      }
      tc_fence_before();
      __syncwarp();
    }
    pre = nxt_full;

    if (last_item && p.pdl_late) pdl_launch_dependents();
    s_l[hf * 128 + r] = l_run;   // P is idle: every P V product has completed
    __syncthreads();
    if (my_t < p.L) {
      const float inv = 1.f / (s_l[r] + s_l[128 + r]);   // fixed order: both threads of the row use the same sum
      uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t)(rowbase + my_t) * 128 + h * 64 + hf * 32);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        dst[j] = make_uint4(pack_h2(acc[8 * j] * inv, acc[8 * j + 1] * inv), pack_h2(acc[8 * j + 2] * inv, acc[8 * j + 3] * inv),
                            pack_h2(acc[8 * j + 4] * inv, acc[8 * j + 5] * inv), pack_h2(acc[8 * j + 6] * inv, acc[8 * j + 7] * inv));
    }
    // the next item's softmax writes P only after two more block barriers (row-maximum exchange): s_l has been read by then
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem_base);
}

}  // namespace mtts
