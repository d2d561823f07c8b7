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
// Warp roles (192 threads, persistent over tiles, 1 CTA / SM):
//   warp 0   TMA producer   (A tile 128x64 + B tile BNx64 per stage, 128B swizzle, mbarrier tx)
//   warp 1   TMEM allocator + single-thread tcgen05.mma issuer (4 x K16 per stage), tcgen05.commit
//   warps 2-5 epilogue: tcgen05.ld (one accumulator row per thread), fused math, smem-staged
//             coalesced stores.  Two accumulator stages in TMEM overlap epilogue(i) with mma(i+1).
//
// Fused epilogues (reference file:line each one replaces is listed in DESIGN.md):
//   EPI_STATS  +bias, fp16 store, deterministic GroupNorm partial sums per (utterance, group)
//   EPI_PLAIN  +bias (+residual) (*row mask), fp16 store
//   EPI_LN     +bias +residual -> fp16 store, then LayerNorm(256) of the same row -> second fp16 store
//   EPI_SNAKE  +bias, SnakeBeta, fp16 store
//   EPI_QKV    q | k row-major, v transposed per (utterance, head) for the attention kernel
//   EPI_FINAL  final 1x1 projection * mask, Euler update of the fp32 state z (channels-first) and
//              refresh of the z channels of the first conv's operand buffer
#pragma once
#include <cuda.h>

#include "ptx.cuh"

namespace mtts {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_MAX_SEGS = 9;
constexpr int GEMM_STAGE_PITCH = 144;                          // bytes per staged row (128 + 16 pad)
constexpr int GEMM_STAGING_BYTES = 32 * GEMM_STAGE_PITCH;      // per epilogue warp

enum { EPI_STATS = 0, EPI_PLAIN = 1, EPI_LN = 2, EPI_SNAKE = 3, EPI_QKV = 4, EPI_FINAL = 5 };

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
  // EPI_LN
  const float* ln_g;
  const float* ln_b;
  __half* out2;
  // EPI_SNAKE
  const float* sn_a;   // exp(alpha) [N]
  const float* sn_ib;  // 1/(exp(beta)+1e-9) [N]
  // EPI_QKV
  __half* q;
  __half* k;
  __half* vt;  // [(b*2+h)*64 + d][Lpad]
  int Lpad;
  // EPI_FINAL
  float* zout;         // (B, n_valid, T) channels-first fp32
  const float* zbase;  // same layout or null
  float zscale;
  __half* x0;  // first-conv operand buffer rows [row*ldx0 + j] or null
  int ldx0;
  int T;
  int n_valid;
};

template <int BN>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int PAR_BYTES = 3 * BN * 4;
  static constexpr int TOTAL = 1024 /*align slack*/ + STAGES * STAGE_BYTES + 4 * GEMM_STAGING_BYTES + PAR_BYTES + 256;
};

// ---- epilogue helpers -------------------------------------------------------------------------
// store 64 fp32 values of "my" row as fp16 through the warp's staging buffer, coalesced 128 B / row
__device__ __forceinline__ void epi_store_h64(uint8_t* st, int lane, const float* v, __half* gtile, int ld,
                                              int rows_valid) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint4 u;
    u.x = pack_h2(v[8 * j + 0], v[8 * j + 1]);
    u.y = pack_h2(v[8 * j + 2], v[8 * j + 3]);
    u.z = pack_h2(v[8 * j + 4], v[8 * j + 5]);
    u.w = pack_h2(v[8 * j + 6], v[8 * j + 7]);
    *reinterpret_cast<uint4*>(st + lane * GEMM_STAGE_PITCH + j * 16) = u;
  }
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    int row = it * 4 + (lane >> 3), c = lane & 7;
    uint4 u = *reinterpret_cast<const uint4*>(st + row * GEMM_STAGE_PITCH + c * 16);
    if (row < rows_valid) *reinterpret_cast<uint4*>(gtile + (size_t)row * ld + c * 8) = u;
  }
  __syncwarp();
}
// add 64 fp16 residual values of "my" row (coalesced global read through the staging buffer)
__device__ __forceinline__ void epi_add_resid_h64(uint8_t* st, int lane, float* v, const __half* gtile, int ld,
                                                  int rows_valid) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    int row = it * 4 + (lane >> 3), c = lane & 7;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (row < rows_valid) u = *reinterpret_cast<const uint4*>(gtile + (size_t)row * ld + c * 8);
    *reinterpret_cast<uint4*>(st + row * GEMM_STAGE_PITCH + c * 16) = u;
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint4 u = *reinterpret_cast<const uint4*>(st + lane * GEMM_STAGE_PITCH + j * 16);
    float2 f;
    f = unpack_h2(u.x); v[8 * j + 0] += f.x; v[8 * j + 1] += f.y;
    f = unpack_h2(u.y); v[8 * j + 2] += f.x; v[8 * j + 3] += f.y;
    f = unpack_h2(u.z); v[8 * j + 4] += f.x; v[8 * j + 5] += f.y;
    f = unpack_h2(u.w); v[8 * j + 6] += f.x; v[8 * j + 7] += f.y;
  }
  __syncwarp();
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using SM = GemmSmem<BN>;
  constexpr int STAGES = SM::STAGES;
  constexpr uint32_t TMEM_COLS = 2 * BN;  // two accumulator stages (512 or 256 columns)
  static_assert(TMEM_COLS == 512 || TMEM_COLS == 256, "BN must be 128 or 256");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + STAGES * SM::STAGE_BYTES;
  float* s_par = reinterpret_cast<float*>(staging + 4 * GEMM_STAGING_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_par) + SM::PAR_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;      // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  int total_chunks = 0;
  for (int s = 0; s < p.num_segs; ++s) total_chunks += p.seg[s].nchunks;
  const int m_tiles = (p.M + GEMM_BM - 1) / GEMM_BM;
  const int total_tiles = m_tiles * p.n_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    fence_mbar_init();
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int r0 = (tile / p.n_tiles) * GEMM_BM;
      const int n0 = (tile % p.n_tiles) * BN;
      int kc = 0;
      for (int s = 0; s < p.num_segs; ++s) {
        const GemmSeg sg = p.seg[s];
        const CUtensorMap* tm = sg.src ? &tmA1 : &tmA0;
        for (int c = 0; c < sg.nchunks; ++c, ++kc) {
          if (lane == 0) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * SM::STAGE_BYTES;
            mbar_arrive_expect_tx(&full_bar[stage], SM::STAGE_BYTES);
            tma_load_2d(sa, tm, &full_bar[stage], sg.col0 + c * GEMM_BK, r0 + sg.row_shift);
            tma_load_2d(sa + SM::A_BYTES, &tmB, &full_bar[stage], kc * GEMM_BK, n0);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    constexpr uint32_t idesc = umma_idesc_f16(GEMM_BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      if (lane == 0) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kc = 0; kc < total_chunks; ++kc) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * SM::STAGE_BYTES);
          const uint64_t da = umma_desc_sw128(sa);
          const uint64_t db = umma_desc_sw128(sa + SM::A_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k)
            umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kc | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[as]);
      } else {
        for (int kc = 0; kc < total_chunks; ++kc)
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      __syncwarp();
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  } else {
    // ===================================== epilogue =========================================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    uint8_t* st = staging + (warp - 2) * GEMM_STAGING_BYTES;
    const int et = threadIdx.x - 64;  // 0..127
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles;
      const int r0 = (tile / p.n_tiles) * GEMM_BM;
      const int n0 = n_tile * BN;
      const int rw0 = r0 + q * 32;       // first row of this warp
      const int row = rw0 + lane;        // my row
      const int rows_valid = min(32, p.M - rw0);  // may be <= 0
      const bool row_ok = row < p.M;

      // per-tile column parameters -> smem (all 4 epilogue warps)
      epi_bar_sync();
      for (int i = et; i < BN; i += 128) {
        s_par[i] = p.bias ? p.bias[n0 + i] : 0.f;
        if constexpr (EPI == EPI_LN) { s_par[BN + i] = p.ln_g[n0 + i]; s_par[2 * BN + i] = p.ln_b[n0 + i]; }
        if constexpr (EPI == EPI_SNAKE) { s_par[BN + i] = p.sn_a[n0 + i]; s_par[2 * BN + i] = p.sn_ib[n0 + i]; }
      }
      epi_bar_sync();

      if (lane == 0) mbar_wait(&tfull_bar[as], aphase);
      __syncwarp();
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + as * BN;

      if constexpr (EPI == EPI_STATS || EPI == EPI_PLAIN || EPI == EPI_LN || EPI == EPI_SNAKE) {
        float mrow = 1.f;
        if constexpr (EPI == EPI_PLAIN) if (p.rowmask && row_ok) mrow = p.rowmask[(size_t)row * p.mask_mul + n_tile * p.mask_nstep];
        int myb = -1;
        if constexpr (EPI == EPI_STATS) if (row_ok) myb = p.rowb[row];
        float gs[(EPI == EPI_STATS) ? 16 : 1];
        float lsum = 0.f, lsq = 0.f;
#pragma unroll 1
        for (int u = 0; u < BN / 64; ++u) {
          float v[64];
          tmem_ld32(taddr + u * 64, v);
          tmem_ld32(taddr + u * 64 + 32, v + 32);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] += s_par[u * 64 + j];
          if ((EPI == EPI_PLAIN && p.resid != nullptr) || EPI == EPI_LN)
            epi_add_resid_h64(st, lane, v, p.resid + (size_t)rw0 * p.ldr + n0 + u * 64, p.ldr, rows_valid);
          if constexpr (EPI == EPI_SNAKE) {
#pragma unroll
            for (int j = 0; j < 64; ++j) {
              float s = sinf(v[j] * s_par[BN + u * 64 + j]);
              v[j] = fmaf(s * s, s_par[2 * BN + u * 64 + j], v[j]);
            }
          }
          if constexpr (EPI == EPI_STATS) {
            float a0 = 0.f, b0 = 0.f, a1 = 0.f, b1 = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) { a0 += v[j]; b0 = fmaf(v[j], v[j], b0); }
#pragma unroll
            for (int j = 32; j < 64; ++j) { a1 += v[j]; b1 = fmaf(v[j], v[j], b1); }
#pragma unroll
            for (int uu = 0; uu < BN / 64; ++uu)
              if (uu == u) { gs[4 * uu + 0] = a0; gs[4 * uu + 1] = b0; gs[4 * uu + 2] = a1; gs[4 * uu + 3] = b1; }
          }
          if constexpr (EPI == EPI_LN) {
#pragma unroll
            for (int j = 0; j < 64; ++j) { lsum += v[j]; lsq = fmaf(v[j], v[j], lsq); }
            tmem_st32(taddr + u * 64, v);
            tmem_st32(taddr + u * 64 + 32, v + 32);
          }
          if constexpr (EPI == EPI_PLAIN) {
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] *= mrow;
          }
          epi_store_h64(st, lane, v, p.out + (size_t)rw0 * p.ldo + n0 + u * 64, p.ldo, rows_valid);
        }
        if constexpr (EPI == EPI_STATS) {
          // deterministic per-(utterance, group) partial sums of this warp's 32 rows
          float* sf = reinterpret_cast<float*>(st);
          int* sb = reinterpret_cast<int*>(st + 32 * 17 * 4);
#pragma unroll
          for (int j = 0; j < 16; ++j) sf[lane * 17 + j] = gs[j];
          sb[lane] = myb;
          __syncwarp();
          if (lane < 16) {
            int cur = -1;
            float acc = 0.f;
            const int wb = rw0 >> 5;
            for (int i = 0; i < 32; ++i) {
              int bi = sb[i];
              if (bi != cur) {
                if (cur >= 0) p.stats_part[((size_t)cur * p.S + (wb - ((cur * p.Lp) >> 5))) * 16 + lane] = acc;
                cur = bi;
                acc = 0.f;
              }
              acc += sf[i * 17 + lane];
            }
            if (cur >= 0) p.stats_part[((size_t)cur * p.S + (wb - ((cur * p.Lp) >> 5))) * 16 + lane] = acc;
          }
          __syncwarp();
        }
        if constexpr (EPI == EPI_LN) {
          tmem_st_wait();
          const float mean = lsum * (1.f / BN);
          const float var = fmaxf(lsq * (1.f / BN) - mean * mean, 0.f);
          const float rstd = rsqrtf(var + 1e-5f);
#pragma unroll 1
          for (int u = 0; u < BN / 64; ++u) {
            float v[64];
            tmem_ld32(taddr + u * 64, v);
            tmem_ld32(taddr + u * 64 + 32, v + 32);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 64; ++j)
              v[j] = fmaf((v[j] - mean) * rstd, s_par[BN + u * 64 + j], s_par[2 * BN + u * 64 + j]);
            epi_store_h64(st, lane, v, p.out2 + (size_t)rw0 * p.ldo + n0 + u * 64, p.ldo, rows_valid);
          }
        }
      } else if constexpr (EPI == EPI_QKV) {
        if (n_tile < 2) {
          __half* dst = n_tile == 0 ? p.q : p.k;
#pragma unroll 1
          for (int u = 0; u < BN / 64; ++u) {
            float v[64];
            tmem_ld32(taddr + u * 64, v);
            tmem_ld32(taddr + u * 64 + 32, v + 32);
            tmem_ld_wait();
            epi_store_h64(st, lane, v, dst + (size_t)rw0 * BN + u * 64, BN, rows_valid);
          }
        } else {
          const int b = row_ok ? p.rowb[row] : -1;
          const int t = row - b * p.Lp;
#pragma unroll 1
          for (int u = 0; u < BN / 32; ++u) {
            float v[32];
            tmem_ld32(taddr + u * 32, v);
            tmem_ld_wait();
            if (b >= 0) {
              // column c = u*32 + j -> head c/64, dim c%64 ; lanes = consecutive frames -> coalesced
              __half* dst = p.vt + ((size_t)(b * 2 + (u >> 1)) * 64 + (u & 1) * 32) * p.Lpad + t;
#pragma unroll
              for (int j = 0; j < 32; ++j) dst[(size_t)j * p.Lpad] = __float2half_rn(v[j]);
            }
          }
        }
      } else {  // EPI_FINAL
        const int b = row_ok ? p.rowb[row] : -1;
        const int t = row - b * p.Lp;
        const float m = (b >= 0) ? p.rowmask[row] : 0.f;
#pragma unroll 1
        for (int u = 0; u < 3; ++u) {  // 80 valid columns = 32 + 32 + 16
          float v[32];
          tmem_ld32(taddr + u * 32, v);
          tmem_ld_wait();
          if (b >= 0) {
            const int nj = (u == 2) ? (p.n_valid - 64) : 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j < nj) {
                const size_t idx = ((size_t)b * p.n_valid + u * 32 + j) * p.T + t;
                float o = (v[j] + s_par[u * 32 + j]) * m;
                o = p.zbase ? fmaf(p.zscale, o, p.zbase[idx]) : o;
                p.zout[idx] = o;
                v[j] = o * m;
              }
            }
            if (p.x0) {
              uint4* xd = reinterpret_cast<uint4*>(p.x0 + (size_t)row * p.ldx0 + u * 32);
              const int nv = nj / 8;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (j < nv)
                  xd[j] = make_uint4(pack_h2(v[8 * j], v[8 * j + 1]), pack_h2(v[8 * j + 2], v[8 * j + 3]),
                                     pack_h2(v[8 * j + 4], v[8 * j + 5]), pack_h2(v[8 * j + 6], v[8 * j + 7]));
            }
          }
        }
      }

      // release the accumulator stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace mtts
