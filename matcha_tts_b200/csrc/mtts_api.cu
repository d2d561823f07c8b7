// C-ABI implementation (include/mtts.h): weight packing, workspace planning, TMA tensor maps,
// and the launch sequence of one estimator call / the whole ODE solve.
// Host logic only enqueues kernels on the caller's stream; no device allocation, no sync.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <utility>
#include <cstdlib>
#include <vector>

#include "../../include/mtts.h"
#include "attention3.cuh"
#include "elementwise.cuh"
#include "ff_tail.cuh"
#include "gemm_tc.cuh"
#include "qkv.cuh"

using namespace mtts;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return fail(MTTS_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));             \
  } while (0)

// Every C entry point that touches CUDA runs with the handle's device current and restores the caller's device on
// return: kernels, side streams, events and cudaFuncSetAttribute all bind to the CURRENT device, and the caller (e.g.
// torch with tensors on cuda:1 while cuda:0 is current) must not see its current device change.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); ok = false; return; }
    if (cur != dev) {
      if (cudaSetDevice(dev) != cudaSuccess) { cudaGetLastError(); ok = false; return; }
      prev = cur;
    }
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define DEVICE_GUARD(h)                                                                        \
  DeviceGuard _dg((h)->device);                                                                \
  if (!_dg.ok) return fail(MTTS_ECUDA, "cannot make device " + std::to_string((h)->device) + " current")

// ------------------------------------------------------------------------------------------------
// tensor maps (driver entry point fetched at run time: the .so has no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int init_encode() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
    return fail(MTTS_ECUDA, "cuTensorMapEncodeTiled entry point not available (needs an NVIDIA driver)");
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return 0;
}
// fp16 2-D row-major [rows, cols] with row pitch `pitch` elements; box = (64 cols, box_rows), 128B swizzle
static int make_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[200];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu pitch=%llu box_rows=%u", (int)r,
             (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)pitch, box_rows);
    return fail(MTTS_ECUDA, buf);
  }
  return 0;
}

// fp16 2-D row-major output [rows, cols]: box = (32 cols, 32 rows) in the 64-byte swizzle = one epilogue warp's staging tile
static int make_out_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[200];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(out) failed (%d) rows=%llu cols=%llu pitch=%llu", (int)r,
             (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)pitch);
    return fail(MTTS_ECUDA, buf);
  }
  return 0;
}

// fp16 [rows, nk*64] row-major viewed as [nk][rows][64]: one box = nk K-chunks of box_rows x 64, landing in
// shared memory as nk consecutive 128B-swizzled tiles (a single TMA instruction for a whole operand piece)
static int make_map3(CUtensorMap* m, const void* base, uint64_t rows, uint32_t nk, uint64_t pitch, uint32_t box_rows,
                     uint32_t box_nk = 0) {
  if (box_nk == 0) box_nk = nk;
  cuuint64_t dims[3] = {64, rows, nk};
  cuuint64_t strides[2] = {pitch * 2, 128};
  cuuint32_t box[3] = {64, box_rows, box_nk};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[200];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(3d) failed (%d) rows=%llu nk=%u pitch=%llu box_rows=%u", (int)r,
             (unsigned long long)rows, nk, (unsigned long long)pitch, box_rows);
    return fail(MTTS_ECUDA, buf);
  }
  return 0;
}

// An operand with both views: d2 = 2-D [rows, cols] (box 64 x box_rows), d3 = [cols/64][rows][64] with box
// {64, 128, 2} (two K chunks per TMA instruction; used by the 128-wide-N GEMM variant), when cols % 128 == 0.
struct TMap {
  CUtensorMap d2, d3;
  CUtensorMap d2h;   // 2-D with 128-row boxes: a CTA pair's half of a 256-row weight tile (cta_group::2 GEMMs)
  CUtensorMap d2t;   // activations: 2-D with 130-row boxes, one tile for the three taps of a k3 conv (GemmParams::tap3)
};
static int make_tmap(TMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t box_rows) {
  if (make_map(&m->d2, base, rows, cols, pitch, box_rows)) return MTTS_ECUDA;
  m->d2h = m->d2;   // half boxes: a CTA pair's share of a weight tile (256-row conv tiles, 128-row QKV pieces)
  if ((box_rows == 256 || box_rows == 128) && make_map(&m->d2h, base, rows, cols, pitch, box_rows / 2)) return MTTS_ECUDA;
  m->d2t = m->d2;
  if (box_rows == 128 && make_map(&m->d2t, base, rows, cols, pitch, 130)) return MTTS_ECUDA;
  m->d3 = m->d2;
  if (cols % 128 == 0 && make_map3(&m->d3, base, rows, (uint32_t)(cols / 64), pitch, 128, 2)) return MTTS_ECUDA;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct PackOp {
  int kind;  // 0: fp32 -> fp16 strided 2-D pack, 1: fp32 copy (mode 0/1/2)
  size_t dst;  // byte offset in the arena
  int N, C;
  long sn, sc, off;
  int n_off, ldd, k_off;
  float scale;
  int mode;
};
struct WEntry {
  std::string name;
  int64_t numel;
  std::vector<PackOp> ops;
  bool loaded = false;
};

struct StageW {
  int nsrc;
  int src_cols[2];  // padded columns of each input source
  size_t c1, c2, qkv, wo, ff1, ff2;  // fp16 weights (byte offsets); c1 also holds res_conv
  size_t c1_b, gn1_g, gn1_b, c2_b, gn2_g, gn2_b, res_b, ln1_g, ln1_b, o_b, ln3_g, ln3_b, ff1_b, sn_a, sn_ib, ff2_b;
  TMap m_c1, m_c2, m_qkv;
  CUtensorMap m_wo, m_ff1, m_ff2;
  CUtensorMap t_ff1;  // W1 as [4][1024][64], box {64, 128, 2}: half (K = 128) of one 128-wide hidden chunk of the fused tail (ff_tail.cuh)
  CUtensorMap m_wo_h, m_ff2_h, t_ff1_h;  // CTA-pair tail (ff_tail_kernel<2>): each CTA stages half of every weight piece (128-row / 64-unit boxes)
};

struct WsLayout {
  int B, T, H, LpT, LpH, rowsT, rowsH, S, cinp, nt_max;
  size_t maskT, maskH, rowbT, rowbH, npadT, npadH, tvals, te_e, te_h1, te_h2, te6, part;
  size_t x0, y, res, h1, xr, a, q, k, o, v;
  size_t skip0, xD0, skip1, xD1, xM0, xM1, xU0s, xU0, xU1s, xF, zmid;
  size_t total;
};

struct LevelMaps {
  TMap h1, a;
  CUtensorMap q;
  CUtensorMap k2, v2;  // k / v [rows][128] with KT-row boxes
  int KT, nkv;         // keys per tile (multiple of 16, <= 192) and tiles per utterance
  CUtensorMap o3;  // o as [2][rows][64], box {64, 128, 2} (fused tail)
};
struct Plan {
  WsLayout w;
  char* ws;
  int te_n = -1, te_solver = -1;  // the solver's time-embedding table in ws.te6 is valid for (n_timesteps, solver)
  LevelMaps lv[2];
  TMap x0, skip0, skip0_pair, xD0, skip1, xD1, xM0, xM1, xU0s, xU0, xU1s, xF;
};

struct GraphKey {
  const void *z, *mu, *mask, *spks, *ws;
  int B, T, n, solver;
  bool operator<(const GraphKey& o) const {
    return std::tie(z, mu, mask, spks, ws, B, T, n, solver) <
           std::tie(o.z, o.mu, o.mask, o.spks, o.ws, o.B, o.T, o.n, o.solver);
  }
};

struct MttsHandle {
  MttsConfig cfg;
  int device, num_sms;
  int cinp, nspk;
  char* arena = nullptr;
  size_t arena_bytes = 0;
  std::vector<WEntry> entries;
  StageW st[6];
  // non-stage weights
  size_t freqs, tw1, tb1, tw2, tb2, mlpW, mlpB;
  size_t w_down0, w_down1, w_up0, w_up1, w_fin, w_proj;
  size_t b_down0, b_down1, b_up0, b_up1, b_fin, gnf_g, gnf_b, b_proj;
  TMap m_down0, m_down1, m_up0, m_up1, m_fin, m_proj;
  bool maps_ready = false;
  std::map<std::tuple<const void*, int, int>, Plan> plans;
  std::map<GraphKey, std::pair<cudaGraphExec_t, int>> graphs;
  std::map<const void*, int> ws_nsub;   // chain partition last laid over each solver workspace
  int launch_count = 0, launch_limit = -1;
  bool a_prefetch = true;  // early L2 prefetch of the first activation tiles
  bool w_hint = true;      // weights are loaded with the L2 evict_last hint
  bool tma_out = true;      // 256-wide conv / linear tiles leave the epilogue as 32 x 32 TMA boxes (MTTS_NO_TMA_OUT=1: shared-memory transpose +
                            // st.global, same bits)
  int lanes = 1;            // mtts_set_lanes: solves the caller keeps in flight on as many handles / streams (MTTS_LANES in the environment)
  int dbg = 0;              // MTTS_SOLVE_DBG: GemmParams::dbg for every conv launch of the solve (timing experiments only)
  int gn_mode = 0;          // GroupNorm-apply pass: 0 = by launch size / concurrency (launch_gn), 1 = always the register-staged
                            // gn_apply_kernel (MTTS_GN_REGS=1), 2 = always the bulk-staged gn_apply2_kernel (MTTS_GN_BULK=1); same bits
  bool tap3 = true;         // single-source k3 convs stage one 130-row activation tile per K chunk for all three taps (MTTS_NO_TAP3=1: one tile per tap)
  bool qkv_pairs = true;    // qkv_kernel<2>: the QKV projection on CTA pairs (MTTS_QKV_PAIRS=0: one CTA per row tile)
  bool tail_pairs = true;   // ff_tail_kernel<2> (cta_group::2, two row tiles per CTA pair, half of every weight piece per CTA); MTTS_TAIL_PAIRS=0:
                            // one CTA per row tile.  On by default since the lanes own their SMs (mtts_set_lanes): see cta_pairs
  bool use_pdl = true;  // MTTS_NO_PDL=1 in the environment disables programmatic dependent launch
  int pair_min_chunks = 9;   // MTTS_PAIR_MIN_CHUNKS: shortest K (in 64-column chunks) that goes to the CTA-pair GEMM (9 = the first conv)
  bool pair_tap3 = true;  // CTA pairs use tap sharing as well (MTTS_PAIR_TAP3=0: one activation tile per tap, the round-1 pair kernel)
  bool cta_pairs = true;  // 256-wide conv GEMMs with K >= pair_min_chunks on CTA pairs (cta_group::2) with tap sharing; MTTS_PAIRS=0: single CTAs.
                          // Alone the pair kernels were always faster (main loop 2.5-2.9 us per 256 rows against 3.5-3.8 us per 128), but
                          // with several solves sharing ALL SMs their cluster launches lost that again (-1 %: both SMs of a TPC must be
                          // free at once).  Since every lane runs on its own share of the SMs (mtts_set_lanes) they pay: +4 % with four
                          // solves in flight, +1.3 % for one solve at a time, and less power per FLOP (sustained +3.7 %) --
                          // profiles/r05q_pairs_under_lanes.txt
  bool pdl_late = true;  // GEMM / tail / attention CTAs release their dependents (griddepcontrol.launch_dependents) when their last
                         // accumulator is complete, not at kernel entry: dependents released at entry sit
                         // on SM slots (shared memory, TMEM) that ready kernels of another chain / solve could use (+4.5% with three
                         // solves in flight, +3% with one)
  // utterance sub-batches ("chains") of one solve run on forked streams so that their kernels overlap:
  // every kernel of a chain is small (tens of tiles) and latency-bound on its own
  int nsub_override = 0;  // MTTS_NSUB in the environment; 0 = heuristic
  std::vector<cudaStream_t> side;
  std::vector<cudaEvent_t> ev_join;
  cudaEvent_t ev_fork = nullptr;
  // optional per-launch device timing (CUDA events on the launching stream)
  bool profiling = false;
  cudaStream_t prof_stream = nullptr;
  std::vector<cudaEvent_t> prof_events;  // pairs (start, stop)
  std::vector<int> prof_kind;            // MTTS_KIND_*
  std::vector<double> prof_flops;        // algorithmic FLOPs of the launch
  // optional in-kernel timeline of the GEMM launches (debug): [max_launches][148][16] int64
  long long* tl_buf = nullptr;
  long long* tail_tl = nullptr;  // ff_tail_kernel timeline (last launch wins), [148][128] int64
  int tl_max = 0, tl_count = 0;
  long long* tl2_buf = nullptr;  // per-tile stamps of the GEMM launches (last launch wins), [148][64] int64
};

static const char* kStageNames[6] = {"down_blocks.0", "down_blocks.1", "mid_blocks.0",
                                     "mid_blocks.1",  "up_blocks.0",   "up_blocks.1"};

// ------------------------------------------------------------------------------------------------
// arena layout + weight table
// ------------------------------------------------------------------------------------------------
static PackOp op_h(size_t dst, int N, int C, long sn, long sc, long off, int n_off, int ldd, int k_off,
                   float scale = 1.f) {
  PackOp o{};
  o.kind = 0; o.dst = dst; o.N = N; o.C = C; o.sn = sn; o.sc = sc; o.off = off;
  o.n_off = n_off; o.ldd = ldd; o.k_off = k_off; o.scale = scale;
  return o;
}
static PackOp op_f(size_t dst, int n, int mode = 0, float scale = 1.f) {
  PackOp o{};
  o.kind = 1; o.dst = dst; o.N = n; o.mode = mode; o.scale = scale;
  return o;
}

static void build_tables(MttsHandle* h) {
  const int C = h->cfg.channels, Cin = h->cfg.in_channels, TD = 4 * C, FD = 4 * C, AD = h->cfg.heads * h->cfg.head_dim;
  size_t cur = 0;
  auto alloc = [&](size_t bytes) { size_t o = cur; cur = align_up(cur + bytes, 256); return o; };
  auto add = [&](const std::string& name, int64_t numel, std::vector<PackOp> ops) {
    WEntry e; e.name = name; e.numel = numel; e.ops = std::move(ops);
    h->entries.push_back(std::move(e));
  };
  // conv k3 weight (N, Ci, 3) -> [N, 3*CiTot] tap-major
  auto conv3_ops = [&](size_t dst, int N, int Ci, int CiTot, int ld = 0) {
    std::vector<PackOp> v;
    if (ld == 0) ld = 3 * CiTot;
    for (int t = 0; t < 3; ++t) v.push_back(op_h(dst, N, Ci, (long)Ci * 3, 3, t, 0, ld, t * CiTot));
    return v;
  };

  h->freqs = alloc(sizeof(float) * (Cin / 2));
  add("@time_freqs", Cin / 2, {op_f(h->freqs, Cin / 2)});
  h->tw1 = alloc(sizeof(float) * TD * Cin); h->tb1 = alloc(sizeof(float) * TD);
  h->tw2 = alloc(sizeof(float) * TD * TD);  h->tb2 = alloc(sizeof(float) * TD);
  h->mlpW = alloc(sizeof(float) * 6 * C * TD); h->mlpB = alloc(sizeof(float) * 6 * C);
  add("time_mlp.linear_1.weight", (int64_t)TD * Cin, {op_f(h->tw1, TD * Cin)});
  add("time_mlp.linear_1.bias", TD, {op_f(h->tb1, TD)});
  add("time_mlp.linear_2.weight", (int64_t)TD * TD, {op_f(h->tw2, TD * TD)});
  add("time_mlp.linear_2.bias", TD, {op_f(h->tb2, TD)});

  for (int s = 0; s < 6; ++s) {
    StageW& w = h->st[s];
    int ci_real, ci_tot;
    if (s == 0) { w.nsrc = 1; w.src_cols[0] = h->cinp; w.src_cols[1] = 0; ci_real = Cin; ci_tot = h->cinp; }
    else if (s < 4) { w.nsrc = 1; w.src_cols[0] = C; w.src_cols[1] = 0; ci_real = C; ci_tot = C; }
    else { w.nsrc = 2; w.src_cols[0] = C; w.src_cols[1] = C; ci_real = 2 * C; ci_tot = 2 * C; }
    w.c1 = alloc(2ull * C * 4 * ci_tot); w.c2 = alloc(2ull * C * 3 * C);  // c1 = [conv taps (3*ci) | res_conv (ci)] along K
    w.qkv = alloc(2ull * 3 * AD * C); w.wo = alloc(2ull * C * AD);
    w.ff1 = alloc(2ull * FD * C); w.ff2 = alloc(2ull * C * FD);
    size_t* f256[] = {&w.c1_b, &w.gn1_g, &w.gn1_b, &w.c2_b, &w.gn2_g, &w.gn2_b, &w.res_b, &w.ln1_g,
                      &w.ln1_b, &w.o_b,  &w.ln3_g, &w.ln3_b, &w.ff2_b};
    for (size_t* p : f256) *p = alloc(sizeof(float) * C);
    w.ff1_b = alloc(sizeof(float) * FD); w.sn_a = alloc(sizeof(float) * FD); w.sn_ib = alloc(sizeof(float) * FD);

    const std::string r = std::string(kStageNames[s]) + ".0", t = std::string(kStageNames[s]) + ".1.0";
    add(r + ".mlp.1.weight", (int64_t)C * TD, {op_f(h->mlpW + sizeof(float) * (size_t)s * C * TD, C * TD)});
    add(r + ".mlp.1.bias", C, {op_f(h->mlpB + sizeof(float) * (size_t)s * C, C)});
    add(r + ".block1.block.0.weight", (int64_t)C * ci_real * 3, conv3_ops(w.c1, C, ci_real, ci_tot, 4 * ci_tot));
    add(r + ".block1.block.0.bias", C, {op_f(w.c1_b, C)});
    add(r + ".block1.block.1.weight", C, {op_f(w.gn1_g, C)});
    add(r + ".block1.block.1.bias", C, {op_f(w.gn1_b, C)});
    add(r + ".block2.block.0.weight", (int64_t)C * C * 3, conv3_ops(w.c2, C, C, C));
    add(r + ".block2.block.0.bias", C, {op_f(w.c2_b, C)});
    add(r + ".block2.block.1.weight", C, {op_f(w.gn2_g, C)});
    add(r + ".block2.block.1.bias", C, {op_f(w.gn2_b, C)});
    add(r + ".res_conv.weight", (int64_t)C * ci_real, {op_h(w.c1, C, ci_real, ci_real, 1, 0, 0, 4 * ci_tot, 3 * ci_tot)});
    add(r + ".res_conv.bias", C, {op_f(w.res_b, C)});
    add(t + ".norm1.weight", C, {op_f(w.ln1_g, C)});
    add(t + ".norm1.bias", C, {op_f(w.ln1_b, C)});
    // softmax scale head_dim^-0.5 folded into to_q (exact: power of two)
    const float qs = 1.0f / sqrtf((float)h->cfg.head_dim);
    add(t + ".attn1.to_q.weight", (int64_t)AD * C, {op_h(w.qkv, AD, C, C, 1, 0, 0, C, 0, qs)});
    add(t + ".attn1.to_k.weight", (int64_t)AD * C, {op_h(w.qkv, AD, C, C, 1, 0, AD, C, 0)});
    add(t + ".attn1.to_v.weight", (int64_t)AD * C, {op_h(w.qkv, AD, C, C, 1, 0, 2 * AD, C, 0)});
    add(t + ".attn1.to_out.0.weight", (int64_t)C * AD, {op_h(w.wo, C, AD, AD, 1, 0, 0, AD, 0)});
    add(t + ".attn1.to_out.0.bias", C, {op_f(w.o_b, C)});
    add(t + ".norm3.weight", C, {op_f(w.ln3_g, C)});
    add(t + ".norm3.bias", C, {op_f(w.ln3_b, C)});
    add(t + ".ff.net.0.alpha", FD, {op_f(w.sn_a, FD, 1)});
    add(t + ".ff.net.0.beta", FD, {op_f(w.sn_ib, FD, 2)});
    add(t + ".ff.net.0.proj.weight", (int64_t)FD * C, {op_h(w.ff1, FD, C, C, 1, 0, 0, C, 0)});
    add(t + ".ff.net.0.proj.bias", FD, {op_f(w.ff1_b, FD)});
    add(t + ".ff.net.2.weight", (int64_t)C * FD, {op_h(w.ff2, C, FD, FD, 1, 0, 0, FD, 0)});
    add(t + ".ff.net.2.bias", C, {op_f(w.ff2_b, C)});
  }
  h->w_down0 = alloc(2ull * C * 3 * C); h->w_down1 = alloc(2ull * C * 3 * C);
  h->w_up0 = alloc(2ull * 2 * C * 3 * C); h->w_up1 = alloc(2ull * C * 3 * C);
  h->w_fin = alloc(2ull * C * 3 * C); h->w_proj = alloc(2ull * 128 * C);
  h->b_down0 = alloc(sizeof(float) * C); h->b_down1 = alloc(sizeof(float) * C);
  h->b_up0 = alloc(sizeof(float) * 2 * C); h->b_up1 = alloc(sizeof(float) * C);
  h->b_fin = alloc(sizeof(float) * C); h->gnf_g = alloc(sizeof(float) * C); h->gnf_b = alloc(sizeof(float) * C);
  h->b_proj = alloc(sizeof(float) * 128);
  add("down_blocks.0.2.conv.weight", (int64_t)C * C * 3, conv3_ops(h->w_down0, C, C, C));
  add("down_blocks.0.2.conv.bias", C, {op_f(h->b_down0, C)});
  add("down_blocks.1.2.weight", (int64_t)C * C * 3, conv3_ops(h->w_down1, C, C, C));
  add("down_blocks.1.2.bias", C, {op_f(h->b_down1, C)});
  // ConvTranspose1d k4 s2 p1, weight (in, out, k):  out[2r] = W1^T x[r] + W3^T x[r-1];
  // out[2r+1] = W0^T x[r+1] + W2^T x[r].  Packed [512, 3*C]: K blocks = taps (r-1, r, r+1),
  // rows [0,C) = even outputs, rows [C,2C) = odd outputs; unused blocks stay zero.
  add("up_blocks.0.2.conv.weight", (int64_t)C * C * 4,
      {op_h(h->w_up0, C, C, 4, (long)C * 4, 3, 0, 3 * C, 0), op_h(h->w_up0, C, C, 4, (long)C * 4, 1, 0, 3 * C, C),
       op_h(h->w_up0, C, C, 4, (long)C * 4, 2, C, 3 * C, C), op_h(h->w_up0, C, C, 4, (long)C * 4, 0, C, 3 * C, 2 * C)});
  add("up_blocks.0.2.conv.bias", C, {op_f(h->b_up0, C), op_f(h->b_up0 + sizeof(float) * C, C)});
  add("up_blocks.1.2.weight", (int64_t)C * C * 3, conv3_ops(h->w_up1, C, C, C));
  add("up_blocks.1.2.bias", C, {op_f(h->b_up1, C)});
  add("final_block.block.0.weight", (int64_t)C * C * 3, conv3_ops(h->w_fin, C, C, C));
  add("final_block.block.0.bias", C, {op_f(h->b_fin, C)});
  add("final_block.block.1.weight", C, {op_f(h->gnf_g, C)});
  add("final_block.block.1.bias", C, {op_f(h->gnf_b, C)});
  const int NO = h->cfg.out_channels;
  add("final_proj.weight", (int64_t)NO * C, {op_h(h->w_proj, NO, C, C, 1, 0, 0, C, 0)});
  add("final_proj.bias", NO, {op_f(h->b_proj, NO)});
  h->arena_bytes = cur;
}

static int build_weight_maps(MttsHandle* h) {
  if (int e = init_encode()) return e;
  const int C = h->cfg.channels;
  char* a = h->arena;
  for (int s = 0; s < 6; ++s) {
    StageW& w = h->st[s];
    const int ci = w.src_cols[0] + w.src_cols[1];
    if (make_tmap(&w.m_c1, a + w.c1, C, 4 * ci, 4 * ci, 256)) return MTTS_ECUDA;
    if (make_tmap(&w.m_c2, a + w.c2, C, 3 * C, 3 * C, 256)) return MTTS_ECUDA;
    if (make_tmap(&w.m_qkv, a + w.qkv, 384, C, C, 128)) return MTTS_ECUDA;
    if (make_map(&w.m_wo, a + w.wo, C, 128, 128, 256)) return MTTS_ECUDA;
    if (make_map(&w.m_ff1, a + w.ff1, 4 * C, C, C, 256)) return MTTS_ECUDA;
    if (make_map(&w.m_ff2, a + w.ff2, C, 4 * C, 4 * C, 256)) return MTTS_ECUDA;
    if (make_map3(&w.t_ff1, a + w.ff1, 4 * C, 4, C, 128, 2)) return MTTS_ECUDA;
    if (make_map(&w.m_wo_h, a + w.wo, C, 128, 128, 128)) return MTTS_ECUDA;
    if (make_map(&w.m_ff2_h, a + w.ff2, C, 4 * C, 4 * C, 128)) return MTTS_ECUDA;
    if (make_map3(&w.t_ff1_h, a + w.ff1, 4 * C, 4, C, 64, 2)) return MTTS_ECUDA;
  }
  if (make_tmap(&h->m_down0, a + h->w_down0, C, 3 * C, 3 * C, 256)) return MTTS_ECUDA;
  if (make_tmap(&h->m_down1, a + h->w_down1, C, 3 * C, 3 * C, 256)) return MTTS_ECUDA;
  if (make_tmap(&h->m_up0, a + h->w_up0, 2 * C, 3 * C, 3 * C, 256)) return MTTS_ECUDA;
  if (make_tmap(&h->m_up1, a + h->w_up1, C, 3 * C, 3 * C, 256)) return MTTS_ECUDA;
  if (make_tmap(&h->m_fin, a + h->w_fin, C, 3 * C, 3 * C, 256)) return MTTS_ECUDA;
  if (make_tmap(&h->m_proj, a + h->w_proj, 128, C, C, 128)) return MTTS_ECUDA;
  h->maps_ready = true;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// workspace layout
// ------------------------------------------------------------------------------------------------
static const int kMaxTimes = 2048;  // rows of the time-embedding table (>= B and >= 2*n_timesteps)

static bool ws_layout(const MttsHandle* h, int B, int T, WsLayout* w) {
  if (B < 1 || T < 1 || B > kMaxTimes) return false;
  memset(w, 0, sizeof *w);
  const int C = h->cfg.channels, NF = h->cfg.out_channels;
  // Level T/2 holds H = ceil(T/2) frames (stride-2 conv, k3 p1: reference model.py:797; mask[:, :, ::2] :1003).  The stride-2
  // conv and the ConvTranspose read / write level T through a row-pair view, so rows per utterance are even at level T:
  // two guard rows for even T, three for odd T -- the ConvTranspose's surplus frame 2H-1 = T of an odd T lands on a
  // guard row and is masked to zero, which is the reference's F.interpolate(nearest) crop (model.py:1027-1028).
  w->B = B; w->T = T; w->H = (T + 1) / 2; w->LpT = T + 2 + (T & 1); w->LpH = w->LpT / 2;
  w->rowsT = B * w->LpT; w->rowsH = B * w->LpH;
  w->S = (T + 31) / 32 + 1;
  w->cinp = h->cinp; w->nt_max = kMaxTimes;
  size_t cur = 0;
  auto alloc = [&](size_t bytes) { size_t o = cur; cur = align_up(cur + bytes, 1024); return o; };
  const size_t rT = w->rowsT, rH = w->rowsH;
  w->maskT = alloc(4 * rT); w->maskH = alloc(4 * rH); w->rowbT = alloc(4 * rT); w->rowbH = alloc(4 * rH);
  w->npadT = alloc(4 * B); w->npadH = alloc(4 * B);
  w->tvals = alloc(4 * kMaxTimes); w->te_e = alloc(4ull * kMaxTimes * h->cfg.in_channels);
  w->te_h1 = alloc(4ull * kMaxTimes * 4 * C); w->te_h2 = alloc(4ull * kMaxTimes * 4 * C);
  w->te6 = alloc(4ull * kMaxTimes * 6 * C);
  w->part = alloc(4ull * B * w->S * 16);
  w->x0 = alloc(2 * rT * w->cinp);
  w->y = alloc(2 * rT * C); w->res = alloc(2 * rT * C); w->h1 = alloc(2 * rT * C); w->xr = alloc(2 * rT * C);
  w->a = alloc(2 * rT * C);
  w->q = alloc(2 * rT * 128); w->k = alloc(2 * rT * 128); w->o = alloc(2 * rT * 128);
  w->v = alloc(2 * rT * 128);
  w->skip0 = alloc(2 * rT * C); w->xD0 = alloc(2 * rH * C); w->skip1 = alloc(2 * rH * C); w->xD1 = alloc(2 * rH * C);
  w->xM0 = alloc(2 * rH * C); w->xM1 = alloc(2 * rH * C); w->xU0s = alloc(2 * rH * C);
  w->xU0 = alloc(2 * rT * C); w->xU1s = alloc(2 * rT * C); w->xF = alloc(2 * rT * C);
  w->zmid = alloc(4ull * B * NF * T);
  w->total = cur;
  return true;
}

static int get_plan(MttsHandle* h, void* ws, size_t ws_bytes, int B, int T, cudaStream_t stream, Plan** out) {
  auto key = std::make_tuple((const void*)ws, B, T);
  auto it = h->plans.find(key);
  if (it != h->plans.end()) { *out = &it->second; return 0; }
  Plan P;
  if (!ws_layout(h, B, T, &P.w)) return fail(MTTS_EINVAL, "unsupported shape: need 1 <= B <= 2048 and T >= 1");
  if (ws_bytes < P.w.total) return fail(MTTS_ENOMEM, "workspace too small (see mtts_workspace_bytes)");
  if ((reinterpret_cast<uintptr_t>(ws) & 1023) != 0) return fail(MTTS_EINVAL, "workspace must be 1024-byte aligned");
  if (int e = init_encode()) return e;
  P.ws = static_cast<char*>(ws);
  const WsLayout& w = P.w;
  const int C = h->cfg.channels;
  char* b = P.ws;
  for (int lv = 0; lv < 2; ++lv) {
    const uint64_t rows = lv ? w.rowsH : w.rowsT;
    LevelMaps& m = P.lv[lv];
    if (make_tmap(&m.h1, b + w.h1, rows, C, C, 128)) return MTTS_ECUDA;
    if (make_tmap(&m.a, b + w.a, rows, C, C, 128)) return MTTS_ECUDA;
    if (make_map3(&m.o3, b + w.o, rows, 2, 128, 128)) return MTTS_ECUDA;
    if (make_map(&m.q, b + w.q, rows, 128, 128, 128)) return MTTS_ECUDA;
    const int L = lv ? w.H : w.T;
    m.nkv = (L + ATT2_KT_MAX - 1) / ATT2_KT_MAX;
    m.KT = (int)align_up((L + m.nkv - 1) / m.nkv, 16);
    if (make_map(&m.k2, b + w.k, rows, 128, 128, m.KT)) return MTTS_ECUDA;
    if (make_map(&m.v2, b + w.v, rows, 128, 128, m.KT)) return MTTS_ECUDA;
  }
  const uint64_t rT = w.rowsT, rH = w.rowsH;
  if (make_tmap(&P.x0, b + w.x0, rT, w.cinp, w.cinp, 128)) return MTTS_ECUDA;
  if (make_tmap(&P.skip0, b + w.skip0, rT, C, C, 128)) return MTTS_ECUDA;
  if (make_tmap(&P.skip0_pair, b + w.skip0, rH, 2 * C, 2 * C, 128)) return MTTS_ECUDA;  // rows (2m, 2m+1) side by side
  if (make_tmap(&P.xD0, b + w.xD0, rH, C, C, 128)) return MTTS_ECUDA;
  if (make_tmap(&P.skip1, b + w.skip1, rH, C, C, 128)) return MTTS_ECUDA;
  if (make_tmap(&P.xD1, b + w.xD1, rH, C, C, 128)) return MTTS_ECUDA;
  if (make_tmap(&P.xM0, b + w.xM0, rH, C, C, 128)) return MTTS_ECUDA;
  if (make_tmap(&P.xM1, b + w.xM1, rH, C, C, 128)) return MTTS_ECUDA;
  if (make_tmap(&P.xU0s, b + w.xU0s, rH, C, C, 128)) return MTTS_ECUDA;
  if (make_tmap(&P.xU0, b + w.xU0, rT, C, C, 128)) return MTTS_ECUDA;
  if (make_tmap(&P.xU1s, b + w.xU1s, rT, C, C, 128)) return MTTS_ECUDA;
  if (make_tmap(&P.xF, b + w.xF, rT, C, C, 128)) return MTTS_ECUDA;
  CUDA_TRY(cudaMemsetAsync(ws, 0, w.total, stream));
  auto res = h->plans.emplace(key, P);
  *out = &res.first->second;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------
static void prof_mark(MttsHandle* h) {
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, h->prof_stream);
  h->prof_events.push_back(e);
}
// Every kernel launch goes through can_launch() ... launched(): launch counting, the debug launch
// limit, and (when enabled) a CUDA-event pair around the launch.
static bool can_launch(MttsHandle* h, int kind = MTTS_KIND_OTHER, double flops = 0.0) {
  if (h->launch_limit >= 0 && h->launch_count >= h->launch_limit) return false;
  ++h->launch_count;
  if (h->profiling) {
    h->prof_kind.push_back(kind);
    h->prof_flops.push_back(flops);
    prof_mark(h);
  }
  return true;
}
static void launched(MttsHandle* h) {
  if (h->profiling) prof_mark(h);
}

// Launch with (optional) programmatic stream serialization: the kernel may start while its
// predecessor drains; every kernel of the solve executes griddepcontrol.wait before touching
// global memory (ptx.cuh).  Works both eagerly and under stream capture (programmatic graph edges).
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(const MttsHandle* h, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                            cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (h->use_pdl && !h->profiling) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// CTA-pair launch: cluster of 2 (+ programmatic stream serialization like launch_k)
template <typename... KArgs, typename... Args>
static cudaError_t launch_k_pair(const MttsHandle* h, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                 cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (h->use_pdl && !h->profiling) ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// Grid of a persistent one-CTA-per-SM launch over `tiles` work units.  One solve at a time: every SM.  With `lanes` solves in
// flight (mtts_set_lanes) a launch takes its SHARE of the SMs instead: a B = 64 kernel has 173 (level T) or 87 (level T/2)
// row tiles, i.e. 1.17 / 0.59 tiles per CTA on 148 SMs, and each CTA's fixed time -- prologue, first-operand latency, the
// last tile's epilogue, exit: ~5 us next to 3.5 us of MMAs per tile -- is SM time no other lane can use.  On 44 SMs the same
// kernel runs 4 / 2 full waves (the epilogue of tile i under the main loop of tile i + 1, one prologue per 4 tiles) while the
// other lanes' kernels run next to it: 4.98 -> 5.7-5.9 M mel-frames/s with four lanes (profiles/r05n_sm_share_sweep.txt).
// The share may be oversubscribed by up to a quarter (kernels of different lanes rarely all peak together: 44 x 4 = 176);
// inside that window the grid with the least wave-quantisation waste wins (43 CTAs for 173 tiles = 5 waves: -6 %).
static int lane_grid(const MttsHandle* h, int tiles, int ctas_per_sm = 1) {
  const int all = h->num_sms * ctas_per_sm;
  if (h->lanes <= 1 || tiles <= 0) return tiles < all ? tiles : all;
  const int lo = (all + h->lanes - 1) / h->lanes, hi = (all * 5 / 4 + h->lanes - 1) / h->lanes;
  if (tiles <= lo) return tiles;
  int best = lo;
  double best_w = 1e30;
  for (int g = lo; g <= hi && g <= all; ++g) {
    const int waves = (tiles + g - 1) / g;
    const double w = (double)waves * g / tiles;
    if (w < best_w * 0.99 || (w <= best_w * 1.0001 && g > best)) { if (w < best_w) best_w = w; best = g; }
  }
  return best < tiles ? best : tiles;
}

// GroupNorm-apply pass (elementwise.cuh): rows staged by one bulk copy per block, or the register-staged form
template <int MODE>
static cudaError_t launch_gn(const MttsHandle* h, const GnParams& g, int B, cudaStream_t stream) {
  // Bulk-staged 64-row blocks (gn_apply2_kernel) when throughput counts: the launch fills every SM about three times over
  // (B >= ~80 at T = 344), or the caller keeps several solves in flight on several handles (mtts_set_chains(h, 1)) -- three
  // B = 64 solves in flight: +2.4 %.  The register-staged kernel (every row requested by its own warp at kernel entry, no
  // barrier, no TMA round trip) has the shorter chain and wins for ONE small solve at a time (3.83 vs 3.65 M frames/s).
  const int blocks64 = ((g.Lp + 63) / 64) * B;
  const bool bulk = h->gn_mode ? h->gn_mode == 2 : (blocks64 >= 3 * h->num_sms || (h->nsub_override == 1 && blocks64 >= h->num_sms));
  if (bulk) return launch_k(h, gn_apply2_kernel<MODE, 64>, dim3((g.Lp + 63) / 64, B), dim3(GN2_THREADS), gn2_smem_bytes<MODE, 64>(), stream, g);
  return launch_k(h, gn_apply_kernel<MODE>, dim3((g.Lp + GN_ROWS - 1) / GN_ROWS, B), dim3(GN_THREADS), 0, stream, g);
}

template <int EPI>
static int set_gemm_pair_attr() {
  CUDA_TRY(cudaFuncSetAttribute(gemm_tc_kernel<256, EPI, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                GemmSmem<256, EPI, 1, 2>::TOTAL));
  return 0;
}

template <int BN, int EPI, int KSUB = 1>
static int set_gemm_attr() {
  CUDA_TRY(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI, KSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                GemmSmem<BN, EPI, KSUB>::TOTAL));
  return 0;
}

// output tensor maps of a 256-wide STATS / PLAIN launch (GemmParams::tma_out); any other launch passes copies of an operand map
struct OutMaps {
  CUtensorMap out, res;
  int on = 0;
};
template <int BN, int EPI>
static int make_out_maps(const MttsHandle* h, OutMaps* o, const CUtensorMap& dummy, const GemmParams& p) {
  o->out = dummy; o->res = dummy; o->on = 0;
  if constexpr (BN == 256 && (EPI == EPI_STATS || EPI == EPI_PLAIN)) {
    if (!h->tma_out || !p.out || (p.ldo % 8) != 0 || (reinterpret_cast<uintptr_t>(p.out) & 15)) return 0;
    if (make_out_map(&o->out, p.out, (uint64_t)p.M, (uint64_t)p.n_tiles * BN, (uint64_t)p.ldo)) return MTTS_ECUDA;
    if (EPI == EPI_STATS && p.res_chunk0 > 0) {
      if (reinterpret_cast<uintptr_t>(p.res_out) & 15) return 0;
      if (make_out_map(&o->res, p.res_out, (uint64_t)p.M, (uint64_t)p.n_tiles * BN, (uint64_t)p.ldo)) return MTTS_ECUDA;
    }
    o->on = 1;
  }
  return 0;
}

template <int BN, int EPI, int KSUB>
static int launch_gemm_maps(MttsHandle* h, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& wmap,
                            const GemmParams& p, cudaStream_t stream, double aflops) {
  if (!can_launch(h, MTTS_KIND_GEMM, aflops)) return 0;
  OutMaps om;
  if (int e = make_out_maps<BN, EPI>(h, &om, a0, p)) return e;
  const int m_tiles = (p.M + GEMM_BM - 1) / GEMM_BM;
  const int tiles = p.m_major ? m_tiles : m_tiles * p.n_tiles;   // m_major: one CTA per row tile, all its N tiles
  int grid = lane_grid(h, tiles);
  GemmParams pp = p;
  pp.tl = nullptr;
  pp.tl2 = h->tl2_buf;
  pp.w_hint = h->w_hint ? 1 : 0;
  pp.a_prefetch = h->a_prefetch ? 1 : 0;
  pp.pdl_late = h->pdl_late ? 1 : 0;
  if (h->tl_buf && h->tl_count < h->tl_max) pp.tl = h->tl_buf + (size_t)(h->tl_count++) * 148 * 16;
  pp.tma_out = om.on;
  if (h->dbg) pp.dbg = h->dbg;
  CUDA_TRY(launch_k(h, gemm_tc_kernel<BN, EPI, KSUB>, dim3(grid), dim3(GEMM_THREADS), GemmSmem<BN, EPI, KSUB>::TOTAL, stream, a0, a1,
                    wmap, om.out, om.res, pp));
  launched(h);
  return 0;
}

// 256-wide N tiles use the 2-D maps (one 64-column K chunk per stage), 128-wide ones the 3-D maps (two chunks)
template <int BN, int EPI>
static int launch_gemm(MttsHandle* h, const TMap& a0, const TMap& a1, const TMap& wmap, const GemmParams& p,
                       cudaStream_t stream, double aflops = 0.0) {
  if constexpr (BN == 128) {
    for (int i = 0; i < p.num_segs; ++i)
      if ((p.seg[i].nchunks & 1) || ((p.seg[i].col0 / 64) & 1)) return fail(MTTS_EINVAL, "128-wide N tile needs whole 128-column K pairs");
    return launch_gemm_maps<128, EPI, 2>(h, a0.d3, a1.d3, wmap.d3, p, stream, aflops);
  } else {
    if constexpr (BN == 256 && (EPI == EPI_STATS || EPI == EPI_PLAIN)) {
      int chunks = 0;
      for (int i = 0; i < p.num_segs; ++i) chunks += p.seg[i].nchunks;
      // k3 conv over one or two sources (+ res_conv): the three taps share one (128 + 2)-row activation tile per K chunk
      int nsrc = 0;
      if (h->tap3) {
        nsrc = (p.num_segs == 3 || p.num_segs == 4) ? 1 : ((p.num_segs == 6 || p.num_segs == 8) ? 2 : 0);
        const bool has_res = nsrc && p.num_segs == 4 * nsrc;
        bool ok = nsrc > 0;
        int CH = 0;
        for (int s = 0; ok && s < nsrc; ++s) CH += p.seg[s].nchunks;
        for (int t = 0; ok && t < (has_res ? 4 : 3); ++t)
          for (int s = 0; ok && s < nsrc; ++s) {
            const GemmSeg& g = p.seg[t * nsrc + s];
            ok = g.src == s && g.row_shift == (t < 3 ? t - 1 : 0) && g.col0 == p.seg[s].col0 && g.nchunks == p.seg[s].nchunks;
          }
        ok = ok && p.res_chunk0 == (has_res ? 3 * CH : 0);
        if (!ok) nsrc = 0;
      }
      // tcgen05 cta_group::2: a CTA pair per 256-row tile, each CTA staging half of the weight tile.  Pays off for long K
      // (tools/gemm_repeat.py at 692 row tiles: K=1536 1221 vs 1111 TFLOP/s, K=768 1090 vs 1035, K=256 518 vs 597); with
      // tap sharing on top the pair pulls 65 KB per K chunk and row tile from L2 instead of 96 KB
      if (h->cta_pairs && chunks >= h->pair_min_chunks) {
        if (!can_launch(h, MTTS_KIND_GEMM, aflops)) return 0;
        const int m_tiles = (p.M + GEMM_BM - 1) / GEMM_BM;
        const int units = ((m_tiles + 1) / 2) * p.n_tiles;
        const int pairs = (lane_grid(h, 2 * units) + 1) / 2;   // CTA pairs inside this lane's share of the SMs
        GemmParams pp = p;
        pp.tl = nullptr; pp.tl2 = h->tl2_buf; pp.m_major = 0;
        pp.w_hint = h->w_hint ? 1 : 0; pp.a_prefetch = h->a_prefetch ? 1 : 0; pp.pdl_late = h->pdl_late ? 1 : 0;
        const bool t3 = nsrc > 0 && h->pair_tap3;
        pp.tap3 = t3 ? nsrc : 0;
        OutMaps om;
        if (int e = make_out_maps<256, EPI>(h, &om, a0.d2, p)) return e;
        pp.tma_out = om.on;
        CUDA_TRY(launch_k_pair(h, gemm_tc_kernel<256, EPI, 1, 2>, dim3(2 * pairs), dim3(GEMM_THREADS), GemmSmem<256, EPI, 1, 2>::TOTAL,
                               stream, t3 ? a0.d2t : a0.d2, t3 ? (nsrc == 2 ? a1.d2t : a0.d2t) : a1.d2, wmap.d2h, om.out, om.res, pp));
        launched(h);
        return 0;
      }
      if (nsrc > 0) {
        GemmParams q = p;
        q.tap3 = nsrc;
        return launch_gemm_maps<256, EPI, 1>(h, a0.d2t, nsrc == 2 ? a1.d2t : a0.d2t, wmap.d2, q, stream, aflops);
      }
    }
    return launch_gemm_maps<BN, EPI, 1>(h, a0.d2, a1.d2, wmap.d2, p, stream, aflops);
  }
}

static void segs_taps(GemmParams& p, int ntaps, const int* shifts, int cols0, int cols1) {
  p.num_segs = 0;
  for (int t = 0; t < ntaps; ++t) {
    p.seg[p.num_segs++] = GemmSeg{0, shifts[t], 0, cols0 / 64};
    if (cols1) p.seg[p.num_segs++] = GemmSeg{1, shifts[t], 0, cols1 / 64};
  }
}
static const int kTaps3[3] = {-1, 0, 1};
static const int kTap1[1] = {0};

struct LevelCtx {
  int lv, L, Lp, rows;
  const float* mask;
  const int* rowb;
  const int* npad;
};

// One resnet + transformer stage (reference ResnetBlock1D :785-790 + BasicTransformerBlock :733-744).
static int run_stage(MttsHandle* h, Plan& P, int s, const LevelCtx& lc, const TMap& in0, const TMap& in1,
                     __half* out, int t_off, int t_stride, cudaStream_t stream) {
  const WsLayout& w = P.w;
  const StageW& sw = h->st[s];
  char* ws = P.ws;
  const char* ar = h->arena;
  const int C = h->cfg.channels;
  const LevelMaps& lm = P.lv[lc.lv];
  auto H = [&](size_t off) { return reinterpret_cast<__half*>(ws + off); };
  auto F = [&](size_t off) { return reinterpret_cast<const float*>(ar + off); };
  float* part = reinterpret_cast<float*>(ws + w.part);
  const float* te6 = reinterpret_cast<const float*>(ws + w.te6);

  const double fr = 2.0 * w.B * (double)lc.L;   // algorithmic FLOPs = fr * N * K (valid rows, unpadded K/N)
  const int ci_real = (s == 0) ? h->cfg.in_channels : sw.src_cols[0] + sw.src_cols[1];
  GemmParams base{};
  base.M = lc.rows; base.rowb = lc.rowb; base.Lp = lc.Lp; base.mask_mul = 1; base.mask_nstep = 0;
  base.stats_part = part; base.S = w.S; base.ldo = C; base.ldr = C;

  // conv1 (k3) -> y with GroupNorm partial sums, and res_conv (1x1) -> res from the same activation tiles:
  // K chunks = [3 taps x sources | 1 x sources], the last group accumulates into the second TMEM accumulator
  {
    GemmParams p = base;
    segs_taps(p, 3, kTaps3, sw.src_cols[0], sw.src_cols[1]);
    int conv_chunks = 0;
    for (int i = 0; i < p.num_segs; ++i) conv_chunks += p.seg[i].nchunks;
    p.seg[p.num_segs++] = GemmSeg{0, 0, 0, sw.src_cols[0] / 64};
    if (sw.src_cols[1]) p.seg[p.num_segs++] = GemmSeg{1, 0, 0, sw.src_cols[1] / 64};
    p.res_chunk0 = conv_chunks; p.res_bias = F(sw.res_b); p.res_out = H(w.res);
    p.bias = F(sw.c1_b); p.out = H(w.y);
    p.n_tiles = 1;
    if (int e = launch_gemm<256, EPI_STATS>(h, in0, in1, sw.m_c1, p, stream, fr * C * 4 * ci_real)) return e;
  }
  // h1 = (Mish(GN(y))*m + temb)*m
  {
    GnParams g{};
    g.y = H(w.y); g.stats_part = part; g.S = w.S; g.L = lc.L; g.Lp = lc.Lp;
    g.gamma = F(sw.gn1_g); g.beta = F(sw.gn1_b); g.rowmask = lc.mask;
    g.temb = te6 + (size_t)s * C; g.t_off = t_off; g.t_stride = t_stride; g.t_ld = 6 * C; g.out = H(w.h1);
    if (can_launch(h, MTTS_KIND_NORM)) { CUDA_TRY(launch_gn<0>(h, g, w.B, stream)); launched(h); }
  }
  // conv2 (k3) -> y, partial sums
  {
    GemmParams p = base;
    segs_taps(p, 3, kTaps3, C, 0);
    p.bias = F(sw.c2_b); p.out = H(w.y);
    p.n_tiles = 1;
    if (int e = launch_gemm<256, EPI_STATS>(h, lm.h1, lm.h1, sw.m_c2, p, stream, fr * C * 3 * C)) return e;
  }
  // x_r = Mish(GN(y))*m + res ; a = LN1(x_r)
  {
    GnParams g{};
    g.y = H(w.y); g.stats_part = part; g.S = w.S; g.L = lc.L; g.Lp = lc.Lp;
    g.gamma = F(sw.gn2_g); g.beta = F(sw.gn2_b); g.rowmask = lc.mask;
    g.out = H(w.xr); g.res = H(w.res); g.ln_g = F(sw.ln1_g); g.ln_b = F(sw.ln1_b); g.out2 = H(w.a);
    if (can_launch(h, MTTS_KIND_NORM)) { CUDA_TRY(launch_gn<1>(h, g, w.B, stream)); launched(h); }
  }
  // q | k | v
  {
    if (can_launch(h, MTTS_KIND_GEMM, fr * 384 * C)) {
      QkvParams qp{};
      qp.M = lc.rows; qp.q = H(w.q); qp.k = H(w.k); qp.v = H(w.v); qp.w_hint = h->w_hint ? 1 : 0; qp.pdl_late = h->pdl_late ? 1 : 0;
      qp.tl = h->tail_tl;   // debug stamps share the tail kernel's buffer (tools/qkv_timeline.py stops before the first tail launch)
      const int tiles = (lc.rows + 127) / 128;
      if (h->qkv_pairs) {   // a CTA pair per 256 rows, each CTA staging half of every weight piece
        const int pairs = (lane_grid(h, 2 * ((tiles + 1) / 2)) + 1) / 2;
        CUDA_TRY(launch_k_pair(h, qkv_kernel<2>, dim3(2 * pairs), dim3(QKV_THREADS), QKV_SMEM, stream, lm.a.d2, sw.m_qkv.d2h, qp));
      } else {
        const int grid = lane_grid(h, tiles);
        CUDA_TRY(launch_k(h, qkv_kernel<1>, dim3(grid), dim3(QKV_THREADS), QKV_SMEM, stream, lm.a.d2, sw.m_qkv.d2, qp));
      }
      launched(h);
    }
  }
  // attention -> o
  if (can_launch(h, MTTS_KIND_ATTN, 512.0 * w.B * (double)lc.L * lc.L)) {
    const int items = ((lc.L + 127) / 128) * 2 * w.B;
    dim3 grid(lane_grid(h, items, 2));   // persistent: two CTAs per SM walk the (query tile, head, utterance) items
    Attn2Params ap{};
    ap.B = w.B;
    ap.L = lc.L; ap.Lp = lc.Lp; ap.KT = lm.KT; ap.nkv = lm.nkv; ap.rowmask = lc.mask; ap.npad = lc.npad;
    ap.v = H(w.v); ap.out = H(w.o); ap.pdl_late = h->pdl_late ? 1 : 0;
    CUDA_TRY(launch_k(h, attention3_kernel, grid, dim3(ATT3_THREADS), ATT3_SMEM, stream, lm.q, lm.k2, lm.v2, ap));
    launched(h);
  }
  // x_a = x_r + o Wo^T + b_o ; c = LN3(x_a) ; out = (x_a + SnakeBeta(c W1^T + b1) W2^T + b2) * m   -- one kernel
  if (can_launch(h, MTTS_KIND_GEMM, fr * C * 128 + 2.0 * fr * 4 * C * C)) {
    TailParams tp{};
    tp.M = lc.rows; tp.xr = H(w.xr); tp.b_o = F(sw.o_b); tp.ln_g = F(sw.ln3_g); tp.ln_b = F(sw.ln3_b);
    tp.b1 = F(sw.ff1_b); tp.sn_a = F(sw.sn_a); tp.sn_ib = F(sw.sn_ib); tp.b2 = F(sw.ff2_b);
    tp.rowmask = lc.mask; tp.out = out; tp.w_hint = h->w_hint ? 1 : 0; tp.pdl_late = h->pdl_late ? 1 : 0;
    tp.tl = h->tail_tl;
    const int tiles = (lc.rows + 127) / 128;
    const int grid = lane_grid(h, tiles);
    if (h->tail_pairs) {
      const int units = (tiles + 1) / 2;
      const int pairs = (lane_grid(h, 2 * units) + 1) / 2;   // CTA pairs inside this lane's share of the SMs
      CUDA_TRY(launch_k_pair(h, ff_tail_kernel<2>, dim3(2 * pairs), dim3(TAIL_THREADS), TAIL_SMEM, stream, lm.o3, sw.m_wo_h, sw.t_ff1_h,
                             sw.m_ff2_h, tp));
    } else {
      CUDA_TRY(launch_k(h, ff_tail_kernel<1>, dim3(grid), dim3(TAIL_THREADS), TAIL_SMEM, stream, lm.o3, sw.m_wo, sw.t_ff1, sw.m_ff2, tp));
    }
    launched(h);
  }
  return 0;
}

// One estimator evaluation on the operand buffer X0 (z | mu | spks already staged, masked).
// Writes zout = (zbase ? zbase + zscale * v : v) in (B, 80, T) fp32 and, if upd_x0, refreshes the
// z channels of X0 with zout * mask for the next evaluation.
static int run_estimator(MttsHandle* h, Plan& P, int t_off, int t_stride, float* zout, const float* zbase, float zscale,
                         bool upd_x0, cudaStream_t stream) {
  const WsLayout& w = P.w;
  char* ws = P.ws;
  const char* ar = h->arena;
  const int C = h->cfg.channels;
  auto H = [&](size_t off) { return reinterpret_cast<__half*>(ws + off); };
  auto F = [&](size_t off) { return reinterpret_cast<const float*>(ar + off); };
  LevelCtx lT{0, w.T, w.LpT, w.rowsT, reinterpret_cast<const float*>(ws + w.maskT),
              reinterpret_cast<const int*>(ws + w.rowbT), reinterpret_cast<const int*>(ws + w.npadT)};
  LevelCtx lH{1, w.H, w.LpH, w.rowsH, reinterpret_cast<const float*>(ws + w.maskH),
              reinterpret_cast<const int*>(ws + w.rowbH), reinterpret_cast<const int*>(ws + w.npadH)};
  float* part = reinterpret_cast<float*>(ws + w.part);

  auto level_conv = [&](const TMap& in, const TMap& wmap, size_t bias, const LevelCtx& lc, __half* out,
                        int mode) -> int {
    GemmParams p{};
    p.rowb = lc.rowb; p.Lp = lc.Lp; p.ldr = C; p.bias = F(bias); p.out = out; p.rowmask = lc.mask;
    p.mask_mul = 1; p.mask_nstep = 0; p.ldo = C; p.n_tiles = 1; p.M = lc.rows;
    if (mode == 0) {  // k3 s1
      segs_taps(p, 3, kTaps3, C, 0);
    } else if (mode == 1) {  // k3 s2 on the row-pair view: x[2m-1], x[2m], x[2m+1]
      p.num_segs = 3;
      p.seg[0] = GemmSeg{0, -1, C, C / 64};
      p.seg[1] = GemmSeg{0, 0, 0, C / 64};
      p.seg[2] = GemmSeg{0, 0, C, C / 64};
    } else {  // ConvTranspose k4 s2: M = input rows, output viewed as [rowsH, 2C], mask per output row
      segs_taps(p, 3, kTaps3, C, 0);
      p.n_tiles = 2; p.ldo = 2 * C; p.mask_mul = 2; p.mask_nstep = 1; p.M = lH.rows;
    }
    // k3 convs: out rows * 256 * 768; ConvTranspose: B*H input rows * 512 outputs * 512 (two 2-tap phases)
    const double af = (mode == 2) ? 2.0 * w.B * (double)w.H * 512 * 512 : 2.0 * w.B * (double)lc.L * C * 3 * C;
    return launch_gemm<256, EPI_PLAIN>(h, in, in, wmap, p, stream, af);
  };

  // down 0 @T
  if (int e = run_stage(h, P, 0, lT, P.x0, P.x0, H(w.skip0), t_off, t_stride, stream)) return e;
  if (int e = level_conv(P.skip0_pair, h->m_down0, h->b_down0, lH, H(w.xD0), 1)) return e;
  // down 1 @T/2
  if (int e = run_stage(h, P, 1, lH, P.xD0, P.xD0, H(w.skip1), t_off, t_stride, stream)) return e;
  if (int e = level_conv(P.skip1, h->m_down1, h->b_down1, lH, H(w.xD1), 0)) return e;
  // mid
  if (int e = run_stage(h, P, 2, lH, P.xD1, P.xD1, H(w.xM0), t_off, t_stride, stream)) return e;
  if (int e = run_stage(h, P, 3, lH, P.xM0, P.xM0, H(w.xM1), t_off, t_stride, stream)) return e;
  // up 0 @T/2 : cat[x, skip1]
  if (int e = run_stage(h, P, 4, lH, P.xM1, P.skip1, H(w.xU0s), t_off, t_stride, stream)) return e;
  {
    LevelCtx lc = lT;  // mask of the OUTPUT rows (level T), indexed 2*r + phase
    if (int e = level_conv(P.xU0s, h->m_up0, h->b_up0, lc, H(w.xU0), 2)) return e;
  }
  // up 1 @T : cat[x, skip0]
  if (int e = run_stage(h, P, 5, lT, P.xU0, P.skip0, H(w.xU1s), t_off, t_stride, stream)) return e;
  if (int e = level_conv(P.xU1s, h->m_up1, h->b_up1, lT, H(w.xF), 0)) return e;
  // final block + projection + ODE update
  {
    GemmParams p{};
    p.M = lT.rows; p.rowb = lT.rowb; p.Lp = lT.Lp; p.stats_part = part; p.S = w.S; p.ldo = C; p.ldr = C;
    segs_taps(p, 3, kTaps3, C, 0);
    p.bias = F(h->b_fin); p.out = H(w.y);
    p.n_tiles = 1;
    if (int e = launch_gemm<256, EPI_STATS>(h, P.xF, P.xF, h->m_fin, p, stream, 2.0 * w.B * (double)w.T * C * 3 * C)) return e;
    GnParams g{};
    g.y = H(w.y); g.stats_part = part; g.S = w.S; g.L = lT.L; g.Lp = lT.Lp;
    g.gamma = F(h->gnf_g); g.beta = F(h->gnf_b); g.rowmask = lT.mask; g.temb = nullptr; g.out = H(w.h1);
    if (can_launch(h, MTTS_KIND_NORM)) { CUDA_TRY(launch_gn<0>(h, g, w.B, stream)); launched(h); }
    GemmParams f{};
    f.M = lT.rows; f.rowb = lT.rowb; f.Lp = lT.Lp; f.rowmask = lT.mask; f.mask_mul = 1;
    segs_taps(f, 1, kTap1, C, 0);
    f.n_tiles = 1; f.bias = F(h->b_proj);
    f.zout = zout; f.zbase = zbase; f.zscale = zscale; f.x0 = upd_x0 ? H(w.x0) : nullptr; f.ldx0 = w.cinp;
    f.T = w.T; f.n_valid = h->cfg.out_channels;
    if (int e = launch_gemm<128, EPI_FINAL>(h, P.lv[0].h1, P.lv[0].h1, h->m_proj, f, stream, 2.0 * w.B * (double)w.T * h->cfg.out_channels * C)) return e;
  }
  return 0;
}

// The data-independent time path (reference model.py:753-762, :828-832, :780): te6[i][6*256] for the n_t time
// values in ws.tvals.  Inside a solve it depends only on (weights, n_timesteps, solver), so it is computed once
// per plan and cached (outside the CUDA graph).
static int run_time_table(MttsHandle* h, Plan& P, int n_t, cudaStream_t stream) {
  const WsLayout& w = P.w;
  char* ws = P.ws;
  const char* ar = h->arena;
  const int C = h->cfg.channels, Cin = h->cfg.in_channels, TD = 4 * C;
  auto Fw = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
  auto Fa = [&](size_t off) { return reinterpret_cast<const float*>(ar + off); };
  if (can_launch(h)) {
    sinus_emb_kernel<<<(n_t * (Cin / 2) + 255) / 256, 256, 0, stream>>>(Fw(w.tvals), Fa(h->freqs), n_t, Cin / 2, Fw(w.te_e));
    CUDA_TRY(cudaGetLastError());
    launched(h);
  }
  if (can_launch(h)) {
    small_linear_kernel<<<(TD + 7) / 8, 256, 0, stream>>>(Fw(w.te_e), Fa(h->tw1), Fa(h->tb1), Fw(w.te_h1), n_t, Cin, TD, 1);
    CUDA_TRY(cudaGetLastError());
    launched(h);
  }
  if (can_launch(h)) {  // Mish applied here is the nn.Mish at the head of every ResnetBlock1D.mlp (:780)
    small_linear_kernel<<<(TD + 7) / 8, 256, 0, stream>>>(Fw(w.te_h1), Fa(h->tw2), Fa(h->tb2), Fw(w.te_h2), n_t, TD, TD, 2);
    CUDA_TRY(cudaGetLastError());
    launched(h);
  }
  if (can_launch(h)) {
    small_linear_kernel<<<(6 * C + 7) / 8, 256, 0, stream>>>(Fw(w.te_h2), Fa(h->mlpW), Fa(h->mlpB), Fw(w.te6), n_t, TD,
                                                              6 * C, 0);
    CUDA_TRY(cudaGetLastError());
    launched(h);
  }
  return 0;
}

// masks, row maps and the first conv's operand buffer
static int run_prologue(MttsHandle* h, Plan& P, const float* x, const float* mu, const float* mask, const float* spks,
                        cudaStream_t stream) {
  const WsLayout& w = P.w;
  char* ws = P.ws;
  auto Fw = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
  if (can_launch(h)) {
    mask_prep_kernel<<<w.B, 256, 0, stream>>>(mask, w.T, w.H, w.LpT, w.LpH, Fw(w.maskT), Fw(w.maskH), reinterpret_cast<int*>(ws + w.rowbT),
                                              reinterpret_cast<int*>(ws + w.rowbH), reinterpret_cast<int*>(ws + w.npadT),
                                              reinterpret_cast<int*>(ws + w.npadH));
    CUDA_TRY(cudaGetLastError());
    launched(h);
  }
  if (can_launch(h)) {
    dim3 grid((w.LpT + 31) / 32, w.B);
    prep_x0_kernel<<<grid, 256, w.cinp * 33 * sizeof(float), stream>>>(
        x, mu, spks, Fw(w.maskT), w.T, w.LpT, h->cfg.out_channels, h->nspk, w.cinp, reinterpret_cast<__half*>(ws + w.x0), 0);
    CUDA_TRY(cudaGetLastError());
    launched(h);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// utterance chains: a solve of B utterances is split into nsub independent sub-batches
// ------------------------------------------------------------------------------------------------
struct Chunk { int b0, nb; size_t ws_off; };

static int pick_nsub(const MttsHandle* h, int B, int T) {
  int n = h->nsub_override;
  if (n <= 0) {
    // two chains once each still fills most SMs with one row tile per CTA at level T: the second chain's
    // kernels run in the SMs the first leaves idle (wave tails, T/2-level kernels); more chains only add
    // launches (measured: profiles/r01_chain_sweep.txt)
    const long rows = (long)B * (T + 2);
    n = rows >= 16384 ? 2 : 1;
  }
  if (n > 8) n = 8;
  if (n > B) n = B;
  if (n < 1) n = 1;
  return n;
}

// Workspace = [single-call estimator layout for (B, T)] [chain 0 layout] [chain 1 layout] ...
// The regions are disjoint: each plan zero-fills its region once and relies on its guard rows
// staying zero afterwards.
static bool make_chunks(const MttsHandle* h, int B, int T, int nsub, std::vector<Chunk>* out, size_t* total) {
  out->clear();
  WsLayout full;
  if (!ws_layout(h, B, T, &full)) return false;
  size_t off = align_up(full.total, 1024);
  int b0 = 0;
  for (int i = 0; i < nsub; ++i) {
    const int nb = B / nsub + (i < B % nsub ? 1 : 0);
    WsLayout w;
    if (!ws_layout(h, nb, T, &w)) return false;
    out->push_back(Chunk{b0, nb, off});
    off += align_up(w.total, 1024);
    b0 += nb;
  }
  *total = off;
  return true;
}

static int ensure_side_streams(MttsHandle* h, int nsub) {
  if (!h->ev_fork) CUDA_TRY(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  while ((int)h->side.size() < nsub - 1) {
    cudaStream_t st;
    cudaEvent_t ev;
    CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    h->side.push_back(st);
    h->ev_join.push_back(ev);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* mtts_last_error(void) { return g_err.c_str(); }
const char* mtts_version(void) { return "matcha_tts_b200 0.1 (sm_100a, tcgen05/TMA)"; }

int mtts_create(const MttsConfig* cfg, int device, MttsHandle** out) {
  if (!cfg || !out) return fail(MTTS_EINVAL, "null argument");
  if (cfg->channels != 256 || cfg->heads != 2 || cfg->head_dim != 64 || cfg->out_channels != 80 ||
      cfg->n_mid_blocks != 2)
    return fail(MTTS_EINVAL, "unsupported architecture: need channels=256, heads=2, head_dim=64, out_channels=80, "
                             "n_mid_blocks=2");
  if (cfg->in_channels < 2 * cfg->out_channels || cfg->in_channels > 256 || (cfg->in_channels % 16) != 0)
    return fail(MTTS_EINVAL, "unsupported in_channels: need 160 <= in_channels <= 256, multiple of 16");
  MttsHandle* h = new MttsHandle();
  h->cfg = *cfg;
  h->device = device;
  h->cinp = (int)align_up(cfg->in_channels, 64);   // whole 64-column K chunks: 160 -> 192 (three chunks per tap, not four), 224 -> 256
  h->nspk = cfg->in_channels - 2 * cfg->out_channels;
  h->num_sms = 148;
  if (const char* e = getenv("MTTS_NO_PDL")) h->use_pdl = !(e[0] == '1');
  if (const char* e = getenv("MTTS_PAIRS")) h->cta_pairs = (e[0] == '1');
  if (const char* e = getenv("MTTS_PAIR_TAP3")) h->pair_tap3 = (e[0] == '1');
  if (const char* e = getenv("MTTS_PAIR_MIN_CHUNKS")) h->pair_min_chunks = atoi(e);
  if (const char* e = getenv("MTTS_NSUB")) h->nsub_override = atoi(e);
  if (const char* e = getenv("MTTS_NO_TAP3")) h->tap3 = !(e[0] == '1');
  if (const char* e = getenv("MTTS_GN_REGS")) if (e[0] == '1') h->gn_mode = 1;
  if (const char* e = getenv("MTTS_GN_BULK")) if (e[0] == '1') h->gn_mode = 2;
  if (const char* e = getenv("MTTS_SOLVE_DBG")) h->dbg = atoi(e);
  if (const char* e = getenv("MTTS_NO_TMA_OUT")) h->tma_out = !(e[0] == '1');
  if (const char* e = getenv("MTTS_TAIL_PAIRS")) h->tail_pairs = (e[0] == '1');
  if (const char* e = getenv("MTTS_QKV_PAIRS")) h->qkv_pairs = (e[0] == '1');
  build_tables(h);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0) {
    // on a GPU box: bind the device and opt the kernels into their dynamic shared memory sizes
    if (device < 0 || device >= ndev) { delete h; return fail(MTTS_EINVAL, "device index out of range"); }
    DeviceGuard dg(device);   // the caller's current device is restored on return
    cudaDeviceProp prop;
    if (!dg.ok || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
      delete h;
      return fail(MTTS_ECUDA, "cudaSetDevice / cudaGetDeviceProperties failed");
    }
    if (prop.major != 10) {
      delete h;
      return fail(MTTS_ECUDA, "this library contains sm_100a code only; device is not Blackwell (cc 10.x)");
    }
    h->num_sms = prop.multiProcessorCount;
    if (const char* e2 = getenv("MTTS_LANES")) { const int n = atoi(e2); if (n >= 1 && n <= 16) h->lanes = n; }
    int e = 0;
    e |= set_gemm_pair_attr<EPI_STATS>(); e |= set_gemm_pair_attr<EPI_PLAIN>();
    e |= set_gemm_attr<256, EPI_STATS>(); e |= set_gemm_attr<256, EPI_PLAIN>();
    e |= set_gemm_attr<128, EPI_FINAL, 2>();
    if (cudaFuncSetAttribute(qkv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, QKV_SMEM) != cudaSuccess) e = 1;
    if (cudaFuncSetAttribute(qkv_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, QKV_SMEM) != cudaSuccess) e = 1;
    if (cudaFuncSetAttribute(attention3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT3_SMEM) != cudaSuccess) e = 1;
    if (cudaFuncSetAttribute(gn_apply2_kernel<0, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, gn2_smem_bytes<0, 64>()) != cudaSuccess) e = 1;
    if (cudaFuncSetAttribute(gn_apply2_kernel<1, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, gn2_smem_bytes<1, 64>()) != cudaSuccess) e = 1;
    if (cudaFuncSetAttribute(ff_tail_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL_SMEM) != cudaSuccess) e = 1;
    if (cudaFuncSetAttribute(ff_tail_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL_SMEM) != cudaSuccess) e = 1;
    if (e) { delete h; return fail(MTTS_ECUDA, "cudaFuncSetAttribute(max dynamic smem) failed: " + g_err); }
  } else {
    cudaGetLastError();  // no GPU: tables/sizes still work (used by the CPU-side tests); compute calls will fail
  }
  *out = h;
  return 0;
}

void mtts_destroy(MttsHandle* h) {
  if (!h) return;
  DeviceGuard dg(h->device);
  for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second.first);
  for (cudaStream_t st : h->side) cudaStreamDestroy(st);
  for (cudaEvent_t ev : h->ev_join) cudaEventDestroy(ev);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  delete h;
}

int mtts_num_weights(const MttsHandle* h) { return h ? (int)h->entries.size() : 0; }
const char* mtts_weight_name(const MttsHandle* h, int idx) {
  if (!h || idx < 0 || idx >= (int)h->entries.size()) return nullptr;
  return h->entries[idx].name.c_str();
}
int64_t mtts_weight_numel(const MttsHandle* h, int idx) {
  if (!h || idx < 0 || idx >= (int)h->entries.size()) return -1;
  return h->entries[idx].numel;
}
size_t mtts_weight_arena_bytes(const MttsHandle* h) { return h ? h->arena_bytes : 0; }

int mtts_set_weight_arena(MttsHandle* h, void* dev_arena, size_t bytes, void* stream) {
  if (!h || !dev_arena) return fail(MTTS_EINVAL, "null argument");
  if (bytes < h->arena_bytes) return fail(MTTS_ENOMEM, "weight arena too small (see mtts_weight_arena_bytes)");
  if ((reinterpret_cast<uintptr_t>(dev_arena) & 255) != 0) return fail(MTTS_EINVAL, "arena must be 256-byte aligned");
  DEVICE_GUARD(h);
  h->arena = static_cast<char*>(dev_arena);
  CUDA_TRY(cudaMemsetAsync(dev_arena, 0, h->arena_bytes, static_cast<cudaStream_t>(stream)));
  for (auto& e : h->entries) e.loaded = false;
  h->plans.clear();
  h->ws_nsub.clear();
  for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second.first);   // they hold the old arena pointer and weight maps by value
  h->graphs.clear();
  return build_weight_maps(h);
}

int mtts_load_weight(MttsHandle* h, int idx, const float* src, int64_t numel, void* stream_) {
  if (!h || !src) return fail(MTTS_EINVAL, "null argument");
  if (!h->arena) return fail(MTTS_ESTATE, "mtts_set_weight_arena must be called first");
  if (idx < 0 || idx >= (int)h->entries.size()) return fail(MTTS_EINVAL, "weight index out of range");
  WEntry& e = h->entries[idx];
  if (numel != e.numel)
    return fail(MTTS_EINVAL, "size mismatch for " + e.name + ": expected " + std::to_string(e.numel) + " elements, got " +
                                 std::to_string(numel));
  DEVICE_GUARD(h);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  for (const PackOp& o : e.ops) {
    if (o.kind == 0) {
      const long total = (long)o.N * o.C;
      pack2d_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
          src, reinterpret_cast<__half*>(h->arena + o.dst), o.N, o.C, o.sn, o.sc, o.off, o.n_off, o.ldd, o.k_off, o.scale);
    } else {
      packf_kernel<<<(o.N + 255) / 256, 256, 0, stream>>>(src, reinterpret_cast<float*>(h->arena + o.dst), o.N, o.mode, o.scale);
    }
    CUDA_TRY(cudaGetLastError());
  }
  e.loaded = true;
  return 0;
}

int mtts_weights_loaded(const MttsHandle* h) {
  if (!h) return 0;
  for (const auto& e : h->entries)
    if (!e.loaded) return 0;
  return 1;
}

size_t mtts_workspace_bytes(const MttsHandle* h, int B, int T) {
  WsLayout w;
  if (!h || !ws_layout(h, B, T, &w)) return 0;
  // the single-call estimator uses one layout for the whole batch; the solver splits the batch into
  // chains, each with its own (smaller) layout placed after it
  std::vector<Chunk> ch;
  size_t total = 0;
  if (!make_chunks(h, B, T, pick_nsub(h, B, T), &ch, &total)) return 0;
  return total;
}

// drop the plans and graphs that live in [workspace, workspace + bytes)
static void forget_workspace(MttsHandle* h, const void* workspace, size_t workspace_bytes) {
  const char* lo = static_cast<const char*>(workspace);
  const char* hi = lo + workspace_bytes;
  for (auto it = h->plans.begin(); it != h->plans.end();) {
    const char* w = static_cast<const char*>(std::get<0>(it->first));
    if (w >= lo && w < hi) it = h->plans.erase(it);
    else ++it;
  }
  for (auto it = h->graphs.begin(); it != h->graphs.end();) {
    const char* w = static_cast<const char*>(it->first.ws);
    if (w >= lo && w < hi) { cudaGraphExecDestroy(it->second.first); it = h->graphs.erase(it); }
    else ++it;
  }
}

static int check_ready(MttsHandle* h) {
  if (!h) return fail(MTTS_EINVAL, "null handle");
  if (!h->arena || !h->maps_ready) return fail(MTTS_ESTATE, "weight arena not set");
  if (!mtts_weights_loaded(h)) {
    for (const auto& e : h->entries)
      if (!e.loaded) return fail(MTTS_ESTATE, "weight not loaded: " + e.name);
  }
  return 0;
}

int mtts_estimator_forward(MttsHandle* h, const float* x, const float* mu, const float* mask, const float* t,
                           const float* spks, float* out, void* workspace, size_t workspace_bytes, int B, int T,
                           void* stream_) {
  if (int e = check_ready(h)) return e;
  if (!x || !mu || !mask || !t || !out || !workspace) return fail(MTTS_EINVAL, "null tensor argument");
  if ((h->nspk > 0) != (spks != nullptr))
    return fail(MTTS_EINVAL, "spks must be given iff in_channels > 2*out_channels");
  DEVICE_GUARD(h);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Plan* P;
  if (int e = get_plan(h, workspace, workspace_bytes, B, T, stream, &P)) return e;
  h->launch_count = 0;
  CUDA_TRY(cudaMemcpyAsync(P->ws + P->w.tvals, t, sizeof(float) * B, cudaMemcpyDeviceToDevice, stream));
  P->te_n = -1;  // per-utterance times: not a solver table
  if (int e = run_time_table(h, *P, B, stream)) return e;
  if (int e = run_prologue(h, *P, x, mu, mask, spks, stream)) return e;
  return run_estimator(h, *P, /*t_off=*/0, /*t_stride=*/1, out, nullptr, 1.f, false, stream);
}

static int enqueue_solve(MttsHandle* h, Plan& P, float* z, const float* mu, const float* mask, const float* spks, int n,
                         int solver, cudaStream_t stream) {
  const WsLayout& w = P.w;
  if (P.te_n != n || P.te_solver != solver) return fail(MTTS_ESTATE, "internal: time-embedding table not prepared");
  if (int e = run_prologue(h, P, z, mu, mask, spks, stream)) return e;
  const float dt = (float)(1.0 / (double)n);
  float* zmid = reinterpret_cast<float*>(P.ws + w.zmid);
  for (int i = 0; i < n; ++i) {
    if (solver == MTTS_SOLVER_EULER) {
      if (int e = run_estimator(h, P, i, 0, z, z, dt, true, stream)) return e;
    } else {
      if (int e = run_estimator(h, P, 2 * i, 0, zmid, z, dt * 0.5f, true, stream)) return e;
      if (int e = run_estimator(h, P, 2 * i + 1, 0, z, z, dt, true, stream)) return e;
    }
  }
  return 0;
}

// all chains of one solve: chain 0 on `stream`, the others on forked side streams, joined at the end
static int enqueue_solve_chains(MttsHandle* h, const std::vector<Chunk>& chunks, std::vector<Plan*>& plans, float* z,
                                const float* mu, const float* mask, const float* spks, int n, int solver, int T,
                                cudaStream_t stream) {
  const int nsub = (int)chunks.size();
  const size_t NF = h->cfg.out_channels;
  if (nsub > 1) {
    CUDA_TRY(cudaEventRecord(h->ev_fork, stream));
    for (int i = 1; i < nsub; ++i) CUDA_TRY(cudaStreamWaitEvent(h->side[i - 1], h->ev_fork, 0));
  }
  for (int i = 0; i < nsub; ++i) {
    const Chunk& c = chunks[i];
    cudaStream_t st = i == 0 ? stream : h->side[i - 1];
    if (int e = enqueue_solve(h, *plans[i], z + (size_t)c.b0 * NF * T, mu + (size_t)c.b0 * NF * T, mask + (size_t)c.b0 * T,
                              spks ? spks + (size_t)c.b0 * h->nspk : nullptr, n, solver, st))
      return e;
  }
  for (int i = 1; i < nsub; ++i) {
    CUDA_TRY(cudaEventRecord(h->ev_join[i - 1], h->side[i - 1]));
    CUDA_TRY(cudaStreamWaitEvent(stream, h->ev_join[i - 1], 0));
  }
  return 0;
}

int mtts_euler_solve(MttsHandle* h, float* z, const float* mu, const float* mask, const float* spks, int n, int solver,
                     void* workspace, size_t workspace_bytes, int B, int T, int use_graph, void* stream_) {
  if (int e = check_ready(h)) return e;
  if (!z || !mu || !mask || !workspace) return fail(MTTS_EINVAL, "null tensor argument");
  if ((h->nspk > 0) != (spks != nullptr))
    return fail(MTTS_EINVAL, "spks must be given iff in_channels > 2*out_channels");
  if (n < 1 || 2 * n > kMaxTimes) return fail(MTTS_EINVAL, "n_timesteps out of range [1, 1024]");
  if (solver != MTTS_SOLVER_EULER && solver != MTTS_SOLVER_MIDPOINT) return fail(MTTS_EINVAL, "unknown solver");
  DEVICE_GUARD(h);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool debug_mode = h->launch_limit >= 0 || h->profiling;   // per-launch introspection: one chain
  const int nsub = debug_mode ? 1 : pick_nsub(h, B, T);
  {
    // The chain partition is laid over the workspace; plans (and graphs) built for another partition of the same
    // memory claim regions the new one overwrites (zero guard rows, time tables): forget them first.
    auto it = h->ws_nsub.find(workspace);
    if (it != h->ws_nsub.end() && it->second != nsub) forget_workspace(h, workspace, workspace_bytes);
    h->ws_nsub[workspace] = nsub;
  }
  std::vector<Chunk> chunks;
  size_t need = 0;
  if (!make_chunks(h, B, T, nsub, &chunks, &need)) return fail(MTTS_EINVAL, "unsupported shape: need 1 <= B <= 2048 and T >= 1");
  if (workspace_bytes < need) return fail(MTTS_ENOMEM, "workspace too small (see mtts_workspace_bytes)");
  if (int e = ensure_side_streams(h, nsub)) return e;
  std::vector<Plan*> plans(nsub);
  for (int i = 0; i < nsub; ++i) {
    const size_t avail = (i + 1 < nsub ? chunks[i + 1].ws_off : need) - chunks[i].ws_off;
    if (int e = get_plan(h, static_cast<char*>(workspace) + chunks[i].ws_off, avail, chunks[i].nb, T, stream, &plans[i])) return e;
  }
  h->launch_count = 0;
  for (int i = 0; i < nsub; ++i) {   // time-embedding tables: once per (plan, n_timesteps, solver), outside the graph
    Plan& P = *plans[i];
    if (P.te_n == n && P.te_solver == solver) continue;
    const int n_t = solver == MTTS_SOLVER_MIDPOINT ? 2 * n : n;
    if (can_launch(h)) {
      solver_times_kernel<<<(n + 127) / 128, 128, 0, stream>>>(reinterpret_cast<float*>(P.ws + P.w.tvals), n,
                                                               solver == MTTS_SOLVER_MIDPOINT);
      CUDA_TRY(cudaGetLastError());
      launched(h);
    }
    if (int e = run_time_table(h, P, n_t, stream)) return e;
    P.te_n = n; P.te_solver = solver;
  }
  if (!use_graph || debug_mode) return enqueue_solve_chains(h, chunks, plans, z, mu, mask, spks, n, solver, T, stream);

  GraphKey key{z, mu, mask, spks, workspace, B, T, n, solver};
  auto it = h->graphs.find(key);
  if (it == h->graphs.end()) {
    cudaGraph_t graph = nullptr;
    h->launch_count = 0;   // count the kernels inside the graph only (the cached table kernels ran above)
    CUDA_TRY(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
    int e = enqueue_solve_chains(h, chunks, plans, z, mu, mask, spks, n, solver, T, stream);
    cudaError_t ce = cudaStreamEndCapture(stream, &graph);
    if (e) { if (graph) cudaGraphDestroy(graph); return e; }
    if (ce != cudaSuccess) return fail(MTTS_ECUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
    cudaGraphExec_t exec = nullptr;
    ce = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) return fail(MTTS_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce));
    if (h->graphs.size() >= 64) {  // bound the cache
      for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second.first);
      h->graphs.clear();
    }
    it = h->graphs.emplace(key, std::make_pair(exec, h->launch_count)).first;
  }
  h->launch_count = it->second.second;  // kernels inside the (cached) graph
  CUDA_TRY(cudaGraphLaunch(it->second.first, stream));
  return 0;
}

int mtts_release_workspace(MttsHandle* h, const void* workspace, size_t workspace_bytes) {
  if (!h || !workspace) return fail(MTTS_EINVAL, "null argument");
  DEVICE_GUARD(h);
  forget_workspace(h, workspace, workspace_bytes);
  h->ws_nsub.erase(workspace);
  return 0;
}

int mtts_set_chains(MttsHandle* h, int n) {
  if (!h) return fail(MTTS_EINVAL, "null handle");
  if (n < 0 || n > 8) return fail(MTTS_EINVAL, "chains must be in [0, 8]");
  if (n != h->nsub_override) {   // the captured graphs embed the chain structure, the plans the chain partition
    if (!h->graphs.empty()) {
      DEVICE_GUARD(h);
      for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second.first);
      h->graphs.clear();
    }
    h->plans.clear();
    h->ws_nsub.clear();
  }
  h->nsub_override = n;
  return 0;
}

int mtts_set_lanes(MttsHandle* h, int lanes) {
  if (!h) return fail(MTTS_EINVAL, "null handle");
  if (lanes < 1 || lanes > 16) return fail(MTTS_EINVAL, "lanes must be in [1, 16]");
  if (lanes != h->lanes && !h->graphs.empty()) {   // the captured graphs embed the launch grids
    DEVICE_GUARD(h);
    for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second.first);
    h->graphs.clear();
  }
  h->lanes = lanes;
  return 0;
}

int mtts_debug_lane_grid(const MttsHandle* h, int work_units, int ctas_per_sm) {
  if (!h || work_units < 0 || ctas_per_sm < 1 || ctas_per_sm > 2) return MTTS_EINVAL;
  return lane_grid(h, work_units, ctas_per_sm);
}

int mtts_last_launch_count(const MttsHandle* h) { return h ? h->launch_count : 0; }

int mtts_debug_profile_begin(MttsHandle* h, void* stream) {
  if (!h) return fail(MTTS_EINVAL, "null handle");
  DEVICE_GUARD(h);
  for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
  h->prof_events.clear(); h->prof_kind.clear(); h->prof_flops.clear();
  h->prof_stream = static_cast<cudaStream_t>(stream);
  h->profiling = true;
  return 0;
}

int mtts_debug_profile_end(MttsHandle* h, int max_entries, float* ms, int* kind, double* flops) {
  if (!h) return fail(MTTS_EINVAL, "null handle");
  DEVICE_GUARD(h);
  h->profiling = false;
  if (!h->prof_events.empty()) CUDA_TRY(cudaEventSynchronize(h->prof_events.back()));
  const int n = (int)h->prof_kind.size();
  for (int i = 0; i < n && i < max_entries; ++i) {
    float t = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&t, h->prof_events[2 * i], h->prof_events[2 * i + 1]));
    if (ms) ms[i] = t;
    if (kind) kind[i] = h->prof_kind[i];
    if (flops) flops[i] = h->prof_flops[i];
  }
  for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
  h->prof_events.clear(); h->prof_kind.clear(); h->prof_flops.clear();
  return n;
}

int mtts_debug_set_tile_timeline(MttsHandle* h, void* dev_buf) {
  if (!h) return fail(MTTS_EINVAL, "null handle");
  h->tl2_buf = static_cast<long long*>(dev_buf);
  return 0;
}

int mtts_debug_set_timeline(MttsHandle* h, void* dev_buf, int max_launches) {
  if (!h) return fail(MTTS_EINVAL, "null handle");
  h->tl_buf = static_cast<long long*>(dev_buf);
  h->tl_max = dev_buf ? max_launches : 0;
  h->tl_count = 0;
  return 0;
}

int mtts_debug_set_tail_timeline(MttsHandle* h, void* dev_buf) {
  if (!h) return fail(MTTS_EINVAL, "null handle");
  h->tail_tl = static_cast<long long*>(dev_buf);
  return 0;
}

int mtts_debug_set_launch_limit(MttsHandle* h, int n) {
  if (!h) return fail(MTTS_EINVAL, "null handle");
  h->launch_limit = n;
  return 0;
}

int64_t mtts_debug_buffer_offset(const MttsHandle* h, int B, int T, int level, const char* name) {
  WsLayout w;
  if (!h || !name || !ws_layout(h, B, T, &w)) return -1;
  (void)level;
  const std::map<std::string, size_t> m = {
      {"maskT", w.maskT}, {"maskH", w.maskH}, {"rowbT", w.rowbT}, {"rowbH", w.rowbH}, {"npadT", w.npadT},
      {"npadH", w.npadH}, {"tvals", w.tvals}, {"te_e", w.te_e},   {"te_h1", w.te_h1}, {"te_h2", w.te_h2},
      {"te6", w.te6},     {"part", w.part},   {"x0", w.x0},       {"y", w.y},         {"res", w.res},
      {"h1", w.h1},       {"xr", w.xr},       {"a", w.a},         {"q", w.q},
      {"k", w.k},         {"o", w.o},         {"v", w.v},       {"skip0", w.skip0},
      {"xD0", w.xD0},     {"skip1", w.skip1}, {"xD1", w.xD1},     {"xM0", w.xM0},     {"xM1", w.xM1},
      {"xU0s", w.xU0s},   {"xU0", w.xU0},     {"xU1s", w.xU1s},   {"xF", w.xF},       {"zmid", w.zmid}};
  auto it = m.find(name);
  return it == m.end() ? -1 : (int64_t)it->second;
}

int mtts_debug_gemm(MttsHandle* h, const void* A, const void* W, const float* bias, void* out, int rows, int C, int N,
                    int ntaps, const int* shifts, void* stream_) {
  if (!h || !A || !W || !out) return fail(MTTS_EINVAL, "null argument");
  if (C % 64 || N % 256 || N > 1024 || ntaps < 1 || ntaps > GEMM_MAX_SEGS || rows < 1) return fail(MTTS_EINVAL, "bad gemm shape");
  if (int e = init_encode()) return e;
  DEVICE_GUARD(h);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GemmParams p{};
  p.M = rows; p.bias = bias; p.out = static_cast<__half*>(out); p.ldo = N; p.mask_mul = 1;
  if (const char* e = getenv("MTTS_DBG")) p.dbg = atoi(e);
  segs_taps(p, ntaps, shifts, C, 0);
  const int saved = h->launch_limit;
  h->launch_limit = -1;
  int e;
  {
    TMap ta, tw;   // production 256-wide path (CTA pairs with MTTS_PAIRS=1)
    if (make_tmap(&ta, A, rows, C, C, 128) || make_tmap(&tw, W, N, (uint64_t)ntaps * C, (uint64_t)ntaps * C, 256)) return MTTS_ECUDA;
    p.n_tiles = N / 256;
    e = launch_gemm<256, EPI_PLAIN>(h, ta, ta, tw, p, stream);
  }
  h->launch_limit = saved;
  return e;
}

}  // extern "C"

#include "mtts_text.inc"
#include "mtts_voc.inc"
