// Second Block1D's GroupNorm-apply of a ResnetBlock1D + residual + LayerNorm1 + the QKV projection of the transformer
// block in ONE launch per stage (reference model.py:773-775, :788-789, :735, :662-664).  Everything between the second
// conv of the resnet and the attention product is row-local once the GroupNorm statistics are known:
//
//   x_r = Mish(GroupNorm(y)) * m + res                      (written to global for the tail kernel, NOT masked)
//   a   = LayerNorm1(x_r)                                    (never leaves the SM: the A operand of the QKV GEMM)
//   q | k | v = a Wqkv^T                                     (q pre-scaled by head_dim^-1/2 at pack time)
//
// It replaces a stand-alone GroupNorm-apply pass (2 KB of HBM traffic per row for zero FLOPs: 41.6 us at level T, B=256)
// followed by a GEMM launch that re-staged the `a` tile once per 128-wide N tile (31.3 us): the fused launch moves
// y + res in, x_r + q + k + v out -- 1.15 KB / row less -- and stages `a` once for all three N tiles.
//
// One CTA per SM, persistent over 128-row tiles, 320 threads:
//   warp 0     TMA producer of the weight pieces ([128 N rows x 64 K] = 16 KB, twelve per tile through a 4-slot ring);
//              never waits for the previous kernel (weights are constants)
//   warp 1     TMEM allocator + tcgen05.mma issuer: per tile 12 pieces x 4 K16 steps into three 128-column accumulators
//   warps 2-9  workers: transform 16 rows each (one row per warp instruction group: 8 channels per lane, GroupNorm
//              scale/shift folded per utterance, Mish, mask, residual, LayerNorm by warp shuffles -- the arithmetic of
//              gn_apply_kernel<1>, instruction for instruction, so both paths give the same bits) into the 128B-swizzled
//              K-chunk tiles of `a`, then drain the accumulators (TMEM lane quarter = warp % 4, 192 columns per warp).
// `a` is double-buffered (2 x 64 KB): the MMAs of tile i run while the workers transform tile i+1; the loads of a row
// batch are requested two batches (>= 2 us) before they are used, the first two batches of tile i+2 before the epilogue
// of tile i.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "gemm_tc.cuh"
#include "ptx.cuh"

namespace mtts {

struct GnbQkvParams {
  int M;                    // rows of the level's flat row space
  int L, Lp, S;             // frames per utterance, rows per utterance, GroupNorm partial slots per utterance
  const __half* y;          // [rows, 256] raw conv output (+bias)
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
  __half* v;                // [rows, 128]
  int w_hint;
  int pdl_late;
  long long* tl;            // debug: [gridDim.x][128] clock64 stamps of worker warp 0 / the MMA warp (tools/gq_timeline.py), or null
};

constexpr int GQ_NST = 4;
constexpr int GQ_PIECE = 16384;                                // [128 N rows x 64 K] of Wqkv
constexpr int GQ_WORKERS = 8;
constexpr int GQ_THREADS = 64 + 32 * GQ_WORKERS;               // 320
constexpr int GQ_OFF_A = 0;                                    // 2 x (4 x 16 KB K-chunk tiles of `a`); the first 16 KB of a buffer double as
                                                               // the epilogue staging of the tile whose MMAs have completed (8 x 2 KB)
constexpr int GQ_A_BYTES = 65536;
constexpr int GQ_OFF_RING = 2 * GQ_A_BYTES;
constexpr int GQ_OFF_PAR = GQ_OFF_RING + GQ_NST * GQ_PIECE;    // gamma | beta | ln_g | ln_b, 256 floats each
constexpr int GQ_OFF_SET2 = GQ_OFF_PAR + 4 * 256 * 4;          // per worker lane: folded GroupNorm scale / shift of the NEXT utterance
                                                               // ([warp][4 quads][lane][4 floats]: conflict-free 16-byte accesses)
constexpr int GQ_OFF_BAR = GQ_OFF_SET2 + GQ_WORKERS * 32 * 16 * 4;
constexpr int GQ_SMEM = GQ_OFF_BAR + 256;
static_assert(GQ_SMEM <= 232448, "exceeds the 227 KB of shared memory one CTA can own");

// 10 warps = 3 on one scheduler partition: 168 registers per thread is the most a 320-thread CTA can own (16 K per partition)
__global__ void __launch_bounds__(GQ_THREADS, 1)
gnb_qkv_kernel(const __grid_constant__ CUtensorMap tmW, const GnbQkvParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  float* s_par = reinterpret_cast<float*>(smem + GQ_OFF_PAR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GQ_OFF_BAR);
  uint64_t* full_bar = bars;                // [GQ_NST]
  uint64_t* empty_bar = bars + GQ_NST;      // [GQ_NST]
  uint64_t* a_ready = bars + 2 * GQ_NST;    // `a` tile written (one arrive per worker warp)
  uint64_t* d_full = a_ready + 1;           // accumulators complete (and that tile's `a` buffer no longer read)
  uint64_t* d_empty = a_ready + 2;          // a worker warp has drained its part of the accumulators
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (!p.pdl_late) pdl_launch_dependents();
  const int m_tiles = (p.M + 127) / 128;
  const int nt = ((int)blockIdx.x < m_tiles) ? (m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;   // tiles of this CTA

  if (threadIdx.x == 0) {
    for (int i = 0; i < GQ_NST; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(a_ready, GQ_WORKERS); mbar_init(d_full, 1); mbar_init(d_empty, GQ_WORKERS);
    fence_mbar_init();
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  if (warp >= 2) {   // affine parameters (weights: independent of the previous kernel)
    for (int i = threadIdx.x - 64; i < 256; i += 32 * GQ_WORKERS) {
      s_par[i] = p.gamma[i]; s_par[256 + i] = p.beta[i]; s_par[512 + i] = p.ln_g[i]; s_par[768 + i] = p.ln_b[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer: Wqkv pieces (constants: no dependency wait) =====
    uint32_t it = 0;
    const uint64_t pol = l2_policy_evict_last();
    for (int i = 0; i < nt; ++i) {
      for (int pc = 0; pc < 12; ++pc, ++it) {   // piece pc: N tile pc/4 (q, k, v), K chunk pc%4
        const uint32_t slot = it % GQ_NST, use = it / GQ_NST;
        mbar_wait_sleep(&empty_bar[slot], (use & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[slot], GQ_PIECE);
          if (p.w_hint) tma_load_2d_hint(smem + GQ_OFF_RING + slot * GQ_PIECE, &tmW, &full_bar[slot], (pc & 3) * 64, (pc >> 2) * 128, pol);
          else tma_load_2d(smem + GQ_OFF_RING + slot * GQ_PIECE, &tmW, &full_bar[slot], (pc & 3) * 64, (pc >> 2) * 128);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    constexpr uint32_t idesc = umma_idesc_f16(128, 128);
    const uint32_t ring = smem_u32(smem + GQ_OFF_RING);
    uint32_t it = 0;
    for (int i = 0; i < nt; ++i) {
      const uint32_t abuf = smem_u32(smem + GQ_OFF_A + (i & 1) * GQ_A_BYTES);
      mbar_wait_sleep(a_ready, i & 1);
      if (i > 0) mbar_wait_sleep(d_empty, (i - 1) & 1);   // the workers have drained the accumulators of tile i-1
      tc_fence_after();
      if (p.tl && lane == 0 && i < 6) p.tl[(size_t)blockIdx.x * 128 + i * 12 + 8] = clock64();
      for (int pc = 0; pc < 12; ++pc, ++it) {
        const uint32_t slot = it % GQ_NST, use = it / GQ_NST;
        mbar_wait_sleep(&full_bar[slot], use & 1);
        tc_fence_after();
        const uint64_t da0 = umma_desc_sw128(abuf + (pc & 3) * 16384), db0 = umma_desc_sw128(ring + slot * GQ_PIECE);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)   // descriptor address field in 16-byte units
            umma_f16(tmem_base + (pc >> 2) * 128, da0 + 2 * kk, db0 + 2 * kk, idesc, ((pc & 3) | kk) != 0);
          umma_commit(&empty_bar[slot]);
          if (pc == 11) umma_commit(d_full);
        }
        __syncwarp();
      }
      if (p.tl && lane == 0 && i < 6) p.tl[(size_t)blockIdx.x * 128 + i * 12 + 9] = clock64();
    }
  } else {
    // ===================================== workers: transform + epilogue ======================
    const int ew = warp - 2;            // 0..7
    const int q4 = warp & 3;            // TMEM lane quarter (epilogue)
    const int cg = ew >> 2;             // column half (epilogue): accumulator columns [cg*192, cg*192+192)
    const int c0 = lane * 8, g = lane >> 2;
    const uint32_t spar = smem_u32(s_par);
    const uint32_t set2 = smem_u32(smem + GQ_OFF_SET2) + (ew * 4 * 32 + lane) * 16;   // quad q at + q * 512
    // LayerNorm1's affine parameters are re-read from shared memory where they are used (16 registers the three row
    // batches in flight need more)
#define GQ_LOAD_LN(lg, lb)                                                                                   \
    float lg[8], lb[8];                                                                                      \
    {                                                                                                        \
      const float4 l0 = lds_f4(spar + (512 + c0) * 4), l1 = lds_f4(spar + (512 + c0 + 4) * 4);               \
      const float4 m0 = lds_f4(spar + (768 + c0) * 4), m1 = lds_f4(spar + (768 + c0 + 4) * 4);               \
      lg[0] = l0.x; lg[1] = l0.y; lg[2] = l0.z; lg[3] = l0.w; lg[4] = l1.x; lg[5] = l1.y; lg[6] = l1.z; lg[7] = l1.w; \
      lb[0] = m0.x; lb[1] = m0.y; lb[2] = m0.z; lb[3] = m0.w; lb[4] = m1.x; lb[5] = m1.y; lb[6] = m1.z; lb[7] = m1.w; \
    }
    pdl_wait();
    int cur_b = -2;
    float ga[8], be[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { ga[j] = 0.f; be[j] = 0.f; }

    // loads of one batch of 4 rows: y, res (16 B per lane), mask and utterance id (lane i < 4 loads row i's, broadcast later)
    struct Batch { uint4 yv[4], rv[4]; float m; int b; };
    auto issue = [&](Batch& bt, int row0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = row0 + i;
        bt.yv[i] = make_uint4(0, 0, 0, 0);
        bt.rv[i] = make_uint4(0, 0, 0, 0);
        if (row < p.M) {   // guard rows are read too (valid memory, ignored below): no dependent load on the utterance id
          bt.yv[i] = ldg128(p.y + (size_t)row * 256 + c0);
          bt.rv[i] = ldg128(p.res + (size_t)row * 256 + c0);
        }
      }
      bt.m = 0.f;
      bt.b = -1;
      if (lane < 4 && row0 + lane < p.M) { bt.m = p.rowmask[row0 + lane]; bt.b = p.rowb[row0 + lane]; }
    };
    // GroupNorm statistics of an utterance: the loads of the conv's partial sums (up to 4 slots per lane in flight) ...
    struct Stats { float2 pp[4]; int b, nsl; };
    auto stats_issue = [&](Stats& st_, int b) {
      st_.b = b;
      const int first = (b * p.Lp) >> 5, last = (b * p.Lp + p.L - 1) >> 5;
      st_.nsl = last - first + 1;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int sl = (lane & 3) + 4 * k;
        st_.pp[k] = make_float2(0.f, 0.f);
        if (sl < st_.nsl) st_.pp[k] = *reinterpret_cast<const float2*>(p.stats_part + ((size_t)b * p.S + sl) * 16 + 2 * g);
      }
    };
    // ... and their reduction into this lane's 8 folded scale / shift pairs (the arithmetic of gn_apply_kernel: fp32,
    // the 4 lanes of a group sum slots lane%4, lane%4 + 4, ... in order, then combine by shuffles)
    auto stats_fold = [&](const Stats& st_, float (&ga)[8], float (&be)[8]) {
      float s = 0.f, ss = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if ((lane & 3) + 4 * k < st_.nsl) { s += st_.pp[k].x; ss += st_.pp[k].y; }
      for (int sl = (lane & 3) + 16; sl < st_.nsl; sl += 4) {   // utterances longer than 16 slots (T > 480): the rest, not prefetched
        const float2 pp = *reinterpret_cast<const float2*>(p.stats_part + ((size_t)st_.b * p.S + sl) * 16 + 2 * g);
        s += pp.x;
        ss += pp.y;
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);  ss += __shfl_xor_sync(0xffffffffu, ss, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);  ss += __shfl_xor_sync(0xffffffffu, ss, 2);
      const float inv_n = 1.f / (32.f * (float)p.L);
      const float mean = s * inv_n;
      const float rstd = rsqrtf(fmaxf(fmaf(-mean, mean, ss * inv_n), 0.f) + 1e-5f);
      const float4 g0 = lds_f4(spar + c0 * 4), g1 = lds_f4(spar + (c0 + 4) * 4);
      const float4 b0 = lds_f4(spar + (256 + c0) * 4), b1 = lds_f4(spar + (256 + c0 + 4) * 4);
      ga[0] = g0.x; ga[1] = g0.y; ga[2] = g0.z; ga[3] = g0.w; ga[4] = g1.x; ga[5] = g1.y; ga[6] = g1.z; ga[7] = g1.w;
      be[0] = b0.x; be[1] = b0.y; be[2] = b0.z; be[3] = b0.w; be[4] = b1.x; be[5] = b1.y; be[6] = b1.z; be[7] = b1.w;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ga[j] *= rstd;
        be[j] = be[j] - mean * ga[j];
      }
    };
    auto stats_finish = [&](const Stats& st_) { cur_b = st_.b; stats_fold(st_, ga, be); };
    // One batch of 4 rows.  Common case (the four rows belong to the current utterance): one straight-line block, the
    // four rows' dependent chains (Mish, the LayerNorm shuffle reduction) interleaved.  Returns false (nothing stored) for a
    // batch with guard rows or an utterance boundary: those go through slow_rows(), once, after the tile's other batches.
    // The operands are unpacked FIRST and the next batch is requested into the same registers right after: the hardware
    // tracks loads in flight with a handful of scoreboards per warp, so a wait for this batch also waits for any younger
    // load that shares its scoreboard -- with two or three batches in flight the "prefetched" data arrived no earlier than
    // the batch requested last (profiles/r02g_gq_timeline.txt: 2.8 us per batch with a request in it, 1.3 us without).
    // One batch outstanding at every wait makes the wait exact; the request then overlaps this batch's arithmetic.
    auto fast_batch = [&](Batch& bt, int row0, uint32_t abuf, int trow0, int next_row0, bool has2) -> bool {
      int bi[4];
      float mi[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { bi[i] = __shfl_sync(0xffffffffu, bt.b, i); mi[i] = __shfl_sync(0xffffffffu, bt.m, i); }
      const bool uniform = bi[0] == cur_b && bi[1] == cur_b && bi[2] == cur_b && bi[3] == cur_b;   // warp-uniform
      float v[4][8], rr[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 f;
        f = unpack_h2(bt.yv[i].x); v[i][0] = f.x; v[i][1] = f.y;
        f = unpack_h2(bt.yv[i].y); v[i][2] = f.x; v[i][3] = f.y;
        f = unpack_h2(bt.yv[i].z); v[i][4] = f.x; v[i][5] = f.y;
        f = unpack_h2(bt.yv[i].w); v[i][6] = f.x; v[i][7] = f.y;
        f = unpack_h2(bt.rv[i].x); rr[i][0] = f.x; rr[i][1] = f.y;
        f = unpack_h2(bt.rv[i].y); rr[i][2] = f.x; rr[i][3] = f.y;
        f = unpack_h2(bt.rv[i].z); rr[i][4] = f.x; rr[i][5] = f.y;
        f = unpack_h2(bt.rv[i].w); rr[i][6] = f.x; rr[i][7] = f.y;
      }
      if (next_row0 >= 0) issue(bt, next_row0);
      // a batch that also holds guard rows and / or rows of the NEXT utterance (whose scale / shift set2 holds): same
      // straight-line code with a per-row choice of the parameter set and zeros on guard rows
      bool mixable = true;
#pragma unroll
      for (int i = 0; i < 4; ++i) mixable = mixable && (bi[i] < 0 || bi[i] == cur_b || (has2 && bi[i] == cur_b + 1));
      if (!uniform && !mixable) return false;
      auto compute = [&](auto mixed_tag) {
        constexpr bool MIXED = decltype(mixed_tag)::value;
        float ga2[8], be2[8];
        if constexpr (MIXED) {
          const float4 a0 = lds_f4(set2 + 0 * 512), a1 = lds_f4(set2 + 1 * 512), e0 = lds_f4(set2 + 2 * 512), e1 = lds_f4(set2 + 3 * 512);
          ga2[0] = a0.x; ga2[1] = a0.y; ga2[2] = a0.z; ga2[3] = a0.w; ga2[4] = a1.x; ga2[5] = a1.y; ga2[6] = a1.z; ga2[7] = a1.w;
          be2[0] = e0.x; be2[1] = e0.y; be2[2] = e0.z; be2[3] = e0.w; be2[4] = e1.x; be2[5] = e1.y; be2[6] = e1.z; be2[7] = e1.w;
        }
        float s[4], ss[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          s[i] = 0.f; ss[i] = 0.f;
          const bool second = MIXED && bi[i] == cur_b + 1;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float gj = (MIXED && second) ? ga2[j] : ga[j], bj = (MIXED && second) ? be2[j] : be[j];
            v[i][j] = mish_f(fmaf(v[i][j], gj, bj)) * mi[i] + rr[i][j];
            s[i] += v[i][j];
            ss[i] = fmaf(v[i][j], v[i][j], ss[i]);
          }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            s[i] += __shfl_xor_sync(0xffffffffu, s[i], off);
            ss[i] += __shfl_xor_sync(0xffffffffu, ss[i], off);
          }
        }
        GQ_LOAD_LN(lg, lb)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = row0 + i, trow = trow0 + i;
          const float lmean = s[i] * (1.f / 256.f);
          const float lrstd = rsqrtf(fmaxf(ss[i] * (1.f / 256.f) - lmean * lmean, 0.f) + 1e-5f);
          float a[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] = fmaf((v[i][j] - lmean) * lrstd, lg[j], lb[j]);
          uint4 o = make_uint4(pack_h2(v[i][0], v[i][1]), pack_h2(v[i][2], v[i][3]), pack_h2(v[i][4], v[i][5]), pack_h2(v[i][6], v[i][7]));
          uint4 o2 = make_uint4(pack_h2(a[0], a[1]), pack_h2(a[2], a[3]), pack_h2(a[4], a[5]), pack_h2(a[6], a[7]));
          if (MIXED && bi[i] < 0) { o = make_uint4(0, 0, 0, 0); o2 = make_uint4(0, 0, 0, 0); }   // guard rows: zeros
          if (!MIXED || row < p.M) stg128(p.xr + (size_t)row * 256 + c0, o);
          // a: channel c0..c0+7 -> K chunk c0/64, 16-byte unit (c0%64)/8 of the row, 128B swizzle
          sts128(abuf + (lane >> 3) * 16384 + trow * 128 + (((lane & 7) ^ (trow & 7)) << 4), o2);
        }
      };
      if (uniform) compute(std::false_type{});
      else compute(std::true_type{});
      return true;
    };
    // rows with guard rows / an utterance boundary among them (1-2 % of the batches): row by row, re-loading the operands
    auto slow_rows = [&](int row0, uint32_t abuf, int trow0) {
#pragma unroll 1
      for (int i = 0; i < 4; ++i) {
        const int row = row0 + i, trow = trow0 + i;
        uint4 o = make_uint4(0, 0, 0, 0), o2 = make_uint4(0, 0, 0, 0);
        const int bb = row < p.M ? p.rowb[row] : -1;
        if (bb >= 0) {                               // warp-uniform
          if (bb != cur_b) { Stats st_; stats_issue(st_, bb); stats_finish(st_); }
          const float mm = p.rowmask[row];
          const uint4 yy = ldg128(p.y + (size_t)row * 256 + c0), rq = ldg128(p.res + (size_t)row * 256 + c0);
          float v[8], rr[8];
          float2 f;
          f = unpack_h2(yy.x); v[0] = f.x; v[1] = f.y;
          f = unpack_h2(yy.y); v[2] = f.x; v[3] = f.y;
          f = unpack_h2(yy.z); v[4] = f.x; v[5] = f.y;
          f = unpack_h2(yy.w); v[6] = f.x; v[7] = f.y;
          f = unpack_h2(rq.x); rr[0] = f.x; rr[1] = f.y;
          f = unpack_h2(rq.y); rr[2] = f.x; rr[3] = f.y;
          f = unpack_h2(rq.z); rr[4] = f.x; rr[5] = f.y;
          f = unpack_h2(rq.w); rr[6] = f.x; rr[7] = f.y;
          float s = 0.f, ss = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[j] = mish_f(fmaf(v[j], ga[j], be[j])) * mm + rr[j];
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
          GQ_LOAD_LN(lg, lb)
          float a[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] = fmaf((v[j] - lmean) * lrstd, lg[j], lb[j]);
          o = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
          o2 = make_uint4(pack_h2(a[0], a[1]), pack_h2(a[2], a[3]), pack_h2(a[4], a[5]), pack_h2(a[6], a[7]));
        }
        if (row < p.M) stg128(p.xr + (size_t)row * 256 + c0, o);   // guard rows: zeros
        sts128(abuf + (lane >> 3) * 16384 + trow * 128 + (((lane & 7) ^ (trow & 7)) << 4), o2);
      }
    };
    // accumulators of tile i -> q | k | v, 192 columns per warp; staging = this warp's own rows of chunk 0 of the tile's
    // `a` buffer (its MMAs are complete; the only other writer of that region is this warp's transform two tiles later)
    auto epilogue = [&](int i) {
      const int r0 = ((int)blockIdx.x + i * (int)gridDim.x) * 128;
      const int rw0 = r0 + q4 * 32;
      const int rows_valid = min(32, p.M - rw0);
      const uint32_t st = smem_u32(smem + GQ_OFF_A + (i & 1) * GQ_A_BYTES) + ew * GEMM_STAGING_BYTES;
      if (lane == 0) { mbar_wait(d_full, i & 1); if (i + 1 == nt && p.pdl_late) pdl_launch_dependents(); }
      __syncwarp();
      tc_fence_after();
      if (p.tl && ew == 0 && lane == 0 && i < 6) p.tl[(size_t)blockIdx.x * 128 + i * 12 + 10] = clock64();
      if (p.tl && lane == 0 && i == 1) p.tl[(size_t)blockIdx.x * 128 + 72 + ew * 6 + 2] = clock64();
      const uint32_t taddr = tmem_base + (uint32_t(q4 * 32) << 16) + cg * 192;
      // one 32-column chunk in registers at a time, not unrolled: the two row batches prefetched for the next transform
      // stay in registers across the epilogue (a double-buffered chunk made the compiler spill them -- an STL of a value
      // still in flight waits for its load, which serialised the prefetch: profiles/r02f_ncu_gnbqkv.txt)
#pragma unroll 1
      for (int c = 0; c < 6; ++c) {
        const int col = cg * 192 + c * 32;   // accumulator column: [0,128) q, [128,256) k, [256,384) v
        float v[32];
        tmem_ld32(taddr + c * 32, v);
        tmem_ld_wait();
        __half* dst = (col < 128 ? p.q : (col < 256 ? p.k : p.v)) + (size_t)rw0 * 128 + (col & 127);
        epi_store_h32(st, lane, v, dst, 128, rows_valid);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d_empty);
      if (p.tl && ew == 0 && lane == 0 && i < 6) p.tl[(size_t)blockIdx.x * 128 + i * 12 + 11] = clock64();
      if (p.tl && lane == 0 && i == 1) p.tl[(size_t)blockIdx.x * 128 + 72 + ew * 6 + 3] = clock64();
    };

    // Pipeline per worker warp:  T(0) | T(1) E(0) | T(2) E(1) | ... | E(nt-1).  The MMAs of tile i run while the workers
    // transform tile i+1 into the other `a` buffer; the first two row batches (and the statistics partials) of tile i+2 are
    // requested before E(i) and land while it runs.  ONE instance of every lambda body in the instruction stream except
    // the 4-row fast path (x4, ~600 instructions each): with everything inlined per batch the loop body was 21 k
    // instructions = 340 KB, more than the instruction cache holds, and the kernel ran at the speed of its instruction
    // fetches (profiles/r02c_gq_timeline.txt).
    Batch X;
    auto wrow_of = [&](int i) { return ((int)blockIdx.x + i * (int)gridDim.x) * 128 + ew * 16; };
    // the utterance of this warp's first row in the CTA's tile i (if that is a guard row: of the rows before it), or -1
    auto first_utt = [&](int i) -> int {
      const int wrow0 = wrow_of(i);
      return (i < nt && wrow0 < p.M) ? wrow0 / p.Lp : -1;
    };
    if (nt > 0) {
      issue(X, wrow_of(0));
      const int b = first_utt(0);
      if (b >= 0) { Stats s0; stats_issue(s0, b); stats_finish(s0); }
    }
#pragma unroll 1
    for (int i = -1; i < nt; ++i) {
      if (i + 1 < nt) {
        const int j = i + 1;
        const int wrow0 = wrow_of(j);
        const int trow0 = ew * 16;
        const uint32_t abuf = smem_u32(smem + GQ_OFF_A + (j & 1) * GQ_A_BYTES);
        long long* tl = (p.tl && ew == 0 && lane == 0 && j < 6) ? p.tl + (size_t)blockIdx.x * 128 + j * 12 : nullptr;
        if (tl) tl[0] = clock64();
        if (p.tl && lane == 0 && j == 2) p.tl[(size_t)blockIdx.x * 128 + 72 + ew * 6 + 0] = clock64();
        // the partial sums of the NEXT tile's utterance are requested now and folded into ga / be after this tile's last
        // row: the statistics never cost a worker a memory round trip of their own
        Stats sn;
        const int b_next = first_utt(j + 1);
        const bool new_utt = b_next >= 0 && b_next != cur_b;
        if (new_utt) stats_issue(sn, b_next);
        // this warp's 16 rows run into the next utterance (5 % of the warp-tiles at T = 344): its scale / shift go to set2
        const bool has2 = wrow0 < p.M && cur_b >= 0 && (wrow0 + 15) / p.Lp > cur_b && (cur_b + 1) * p.Lp < p.M;
        if (has2) {
          Stats s2;
          float g2[8], e2[8];
          stats_issue(s2, cur_b + 1);
          stats_fold(s2, g2, e2);
          sts_f4(set2 + 0 * 512, make_float4(g2[0], g2[1], g2[2], g2[3])); sts_f4(set2 + 1 * 512, make_float4(g2[4], g2[5], g2[6], g2[7]));
          sts_f4(set2 + 2 * 512, make_float4(e2[0], e2[1], e2[2], e2[3])); sts_f4(set2 + 3 * 512, make_float4(e2[4], e2[5], e2[6], e2[7]));
        }
        unsigned slow = 0;
        // the four batches of the tile; the first one was requested before the previous epilogue, the first batch of the
        // next tile is requested by the last one (and lands while the epilogue runs)
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
          const int nxt = k < 3 ? wrow0 + 4 * (k + 1) : (j + 1 < nt ? wrow_of(j + 1) : -1);
          if (!fast_batch(X, wrow0 + 4 * k, abuf, trow0 + 4 * k, nxt, has2)) slow |= 1u << k;
          if (tl) tl[1 + k] = clock64();
        }
#pragma unroll 1
        for (int k = 0; k < 4; ++k)
          if (slow & (1u << k)) slow_rows(wrow0 + 4 * k, abuf, trow0 + 4 * k);
        if (tl) tl[5] = clock64();
        fence_proxy_async_smem();   // a is read by the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready);
        if (tl) tl[6] = clock64();
        if (p.tl && lane == 0 && j == 2) p.tl[(size_t)blockIdx.x * 128 + 72 + ew * 6 + 1] = clock64();
        if (new_utt) stats_finish(sn);
      }
      if (i >= 0) epilogue(i);
    }
    if (nt == 0 && p.pdl_late) pdl_launch_dependents();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

#undef GQ_LOAD_LN

}  // namespace mtts
