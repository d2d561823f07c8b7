/* mtts.h -- C ABI of the B200-native CFM decoder (Matcha-TTS inference hot path).
 *
 * The reference (Lounes78/matcha-tts) has NO FFI/plugin layer: the path sits behind plain
 * nn.Module methods.  This header is the boundary a maintainer binds instead of the PyTorch
 * bodies of
 *     Decoder.forward(x, mask, mu, t, spks, cond)                    reference model.py:964-1048
 *     BASECFM.forward Euler / midpoint loop                          reference model.py:1084-1109
 * (see INTEGRATION.md for the ctypes stub used by matcha_tts_b200/model.py).
 *
 * Conventions
 *   - plain C, no torch types: device pointers + sizes + an opaque CUDA stream handle (cudaStream_t
 *     passed as void*).  All tensors are contiguous fp32 in the reference's own layouts:
 *     x / mu / out / z : (B, out_channels, T)   mask : (B, 1, T) as 0/1 floats   t : (B,)
 *     spks : (B, in_channels - 2*out_channels) or NULL.
 *   - the caller owns every buffer (weights arena, workspace, tensors); the library never
 *     allocates device memory, never synchronises, and only enqueues work on the given stream.
 *   - every function returns 0 on success or a negative MTTS_E* code; mtts_last_error() gives
 *     the message of the last failure on the calling thread.  Shape violations are errors --
 *     there is no silent fallback and no CPU path.
 *   - one handle per (process, device); a handle is not thread-safe (neither is the reference).
 */
#ifndef MTTS_H_
#define MTTS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTTS_OK 0
#define MTTS_EINVAL (-1)   /* bad argument / unsupported shape */
#define MTTS_ECUDA (-2)    /* a CUDA runtime / driver call failed */
#define MTTS_ESTATE (-3)   /* call order violated (weights not loaded, arena not set, ...) */
#define MTTS_ENOMEM (-4)   /* caller-provided buffer too small */

#define MTTS_SOLVER_EULER 0
#define MTTS_SOLVER_MIDPOINT 1

typedef struct MttsHandle MttsHandle;

/* Hyper-parameters of the estimator; mirrors Decoder.__init__ (reference model.py:835-962) for the
 * architecture the reference instantiates (main.py:63-79): two U-Net levels of `channels` width,
 * one transformer block per stage, SnakeBeta feed-forward, GroupNorm(8).
 * Supported: channels == 256, heads == 2, head_dim == 64, out_channels == 80,
 * in_channels in {160 .. 256} and a multiple of 16 (160 = LJSpeech, 224 = multi-speaker). */
typedef struct MttsConfig {
  int in_channels;
  int out_channels;
  int channels;
  int heads;
  int head_dim;
  int n_mid_blocks;
} MttsConfig;

/* ---- lifetime ------------------------------------------------------------------------------ */
int mtts_create(const MttsConfig* cfg, int device, MttsHandle** out);
void mtts_destroy(MttsHandle* h);
const char* mtts_last_error(void);
const char* mtts_version(void);

/* ---- weights: replaces nn.Module.load_state_dict for `decoder.estimator.*` (model.py App. B keys,
 * checkpoint load at reference main.py:94-121).  The table lists every state-dict key the estimator
 * owns, in a fixed order; entries whose name starts with '@' are host-derived constants
 * ("@time_freqs": exp(arange(C/2) * -ln(1e4)/(C/2-1)), reference model.py:757-758).
 * mtts_load_weight() converts one fp32 tensor in the reference's layout into the packed fp16/fp32
 * kernel layouts inside the caller-provided arena. */
int mtts_num_weights(const MttsHandle* h);
const char* mtts_weight_name(const MttsHandle* h, int idx);
int64_t mtts_weight_numel(const MttsHandle* h, int idx);
size_t mtts_weight_arena_bytes(const MttsHandle* h);
int mtts_set_weight_arena(MttsHandle* h, void* dev_arena, size_t bytes, void* stream);
int mtts_load_weight(MttsHandle* h, int idx, const float* dev_src, int64_t numel, void* stream);
int mtts_weights_loaded(const MttsHandle* h); /* 1 when every table entry has been loaded */

/* ---- workspace ----------------------------------------------------------------------------- */
/* Bytes of scratch needed for a batch of B utterances padded to T frames (1 <= B <= 2048, T >= 1; odd T follows the
 * reference's nearest-resize crop after the ConvTranspose, model.py:1027-1028).
 * Returns 0 on invalid shapes. */
size_t mtts_workspace_bytes(const MttsHandle* h, int B, int T);

/* Forget everything cached for a workspace the caller is about to free or reuse for something else (the per-shape plan
 * with its tensor maps, and the CUDA graphs captured over it).  The library zero-fills a workspace once, when it first
 * sees the (pointer, B, T) triple, and relies on its guard rows staying zero afterwards: a caller that recycles
 * workspace memory must call this before handing the same address back.  Work already enqueued must have completed. */
int mtts_release_workspace(MttsHandle* h, const void* workspace, size_t workspace_bytes);

/* ---- compute ------------------------------------------------------------------------------- */
/* One estimator call: out = Decoder.forward(x, mask, mu, t, spks)     (reference model.py:964) */
int mtts_estimator_forward(MttsHandle* h, const float* x, const float* mu, const float* mask, const float* t,
                           const float* spks, float* out, void* workspace, size_t workspace_bytes, int B, int T,
                           void* stream);

/* The ODE solve of BASECFM.forward (reference model.py:1086-1104) on a caller-provided initial
 * state: z <- z + dt * v(z, i/n) for i < n (Euler), or the midpoint rule.  z_inout holds z_0 =
 * randn * temperature on entry (drawn by the host exactly like model.py:1085) and z_n on return.
 * use_graph != 0 captures the whole solve into a CUDA graph (cached per shape/pointers). */
int mtts_euler_solve(MttsHandle* h, float* z_inout, const float* mu, const float* mask, const float* spks,
                     int n_timesteps, int solver, void* workspace, size_t workspace_bytes, int B, int T,
                     int use_graph, void* stream);

/* Utterance chains per solve: a solve of B utterances is split into n independent sub-batches on forked streams (graph
 * branches) whose latency-bound kernels overlap.  0 = heuristic (two chains for large batches -- best when ONE solve is
 * in flight); a host that keeps several solves in flight on several handles/streams sets 1 (the solves overlap each
 * other instead).  Takes effect for the following mtts_euler_solve / mtts_workspace_bytes calls. */
int mtts_set_chains(MttsHandle* h, int n);

/* Solves the caller keeps in flight at a time on as many handles / streams (default 1).  With lanes > 1 every persistent
 * launch of this handle sizes its grid for its share of the SMs (148 x 5/4 / lanes at most, the grid inside that window
 * with the least wave-quantisation waste) instead of all of them: a batch-64 kernel then runs full waves on ~44 SMs next
 * to the other lanes' kernels rather than 1.17 tiles per CTA on every SM.  Call it together with mtts_set_chains(h, 1);
 * results are bit-identical for every value.  Takes effect for the following calls (cached graphs are dropped). */
int mtts_set_lanes(MttsHandle* h, int lanes);

/* The grid (CTAs) a persistent launch over `work_units` tiles gets under the current mtts_set_lanes value, for kernels with
 * `ctas_per_sm` (1 or 2) resident CTAs per SM -- the rule above, exposed for tests and capacity planning. */
int mtts_debug_lane_grid(const MttsHandle* h, int work_units, int ctas_per_sm);

/* Number of kernels enqueued by the last estimator_forward / euler_solve call on this handle. */
int mtts_last_launch_count(const MttsHandle* h);

/* Per-launch device timing for bench.py's roofline: between begin and end every kernel the library
 * launches on `stream` is bracketed by a CUDA-event pair (graph capture is bypassed meanwhile).
 * end() synchronises on the last event and returns the number of launches recorded, filling up to
 * max_entries of: milliseconds, kind (MTTS_KIND_*), algorithmic FLOPs of the launch. */
#define MTTS_KIND_GEMM 0   /* tcgen05 implicit-GEMM (convs + linears) */
#define MTTS_KIND_ATTN 1   /* attention */
#define MTTS_KIND_NORM 2   /* GroupNorm-apply / Mish / LayerNorm pass */
#define MTTS_KIND_OTHER 3  /* masks, operand staging, time embedding */
int mtts_debug_profile_begin(MttsHandle* h, void* stream);
int mtts_debug_profile_end(MttsHandle* h, int max_entries, float* ms, int* kind, double* flops);

/* In-kernel timeline of the GEMM launches (tools/gemm_timeline.py): launch i of the following calls
 * writes [148 CTAs][16] int64 stamps (clock64 at entry / setup done / dependency wait passed / first
 * operands landed / last MMA issued / first accumulator ready / epilogue done / exit, and globaltimer
 * ns at [8] entry, [9] wait passed, [10] exit) at dev_buf + i*148*16.  NULL switches it off. */
int mtts_debug_set_timeline(MttsHandle* h, void* dev_buf, int max_launches);
/* Same for the fused transformer-tail kernel: [148 CTAs][128] clock64 stamps of each CTA's first tile
 * ([0,64) MMA issuer, [64,128) one epilogue warp; see tools/tail_timeline.py); every tail launch overwrites it. */
int mtts_debug_set_tail_timeline(MttsHandle* h, void* dev_buf);

/* ---- introspection used by the parity tests -------------------------------------------------- */
/* Stop the estimator after `n` kernel launches (n < 0: run everything). */
/* Per-tile stamps of the following GEMM launches (tools/gemm_tiles.py): dev_buf = [148][64] int64, or NULL to stop. */
int mtts_debug_set_tile_timeline(MttsHandle* h, void* dev_buf);
int mtts_debug_set_launch_limit(MttsHandle* h, int n);
/* Byte offset / row count / column count of a named intermediate inside the workspace for (B, T);
 * level 0 = T frames, 1 = T/2 frames.  Returns -1 for an unknown name. */
int64_t mtts_debug_buffer_offset(const MttsHandle* h, int B, int T, int level, const char* name);
/* Generic implicit-GEMM entry for unit tests: out[r, n] = bias[n] + sum_taps A[r + shift_i, :] . W[n, i*C:(i+1)*C]
 * A: (rows, C) fp16, W: (N, ntaps*C) fp16, out: (rows, N) fp16; C % 64 == 0, N % 128 == 0. */
int mtts_debug_gemm(MttsHandle* h, const void* A, const void* W, const float* bias, void* out, int rows, int C, int N,
                    int ntaps, const int* shifts, void* stream);

/* ---- the step before the path: TextEncoder + duration predictor ----------------------
 * Replaces the body of TextEncoder.forward(x, x_lengths, spks) -> (mu, logw, x_mask)          reference model.py:500-535
 * (embedding, ConvReluNorm prenet :168-207, RoPE transformer encoder :244-438, proj_m, DurationPredictor :209-235).
 * Same conventions as above: caller-owned buffers, no allocation, no synchronisation, errors as codes.
 *   tokens  (B, T_x) int64 symbol ids          lengths (B,) int64          spks (B, spk_emb_dim) fp32, already embedded, or NULL
 *   mu (B, n_feats, T_x) fp32      logw (B, 1, T_x) fp32      x_mask (B, 1, T_x) fp32 0/1
 * Supported: n_channels and n_channels + spk_emb_dim (when n_spks > 1) multiples of 64 and <= 256; (width / n_heads) in
 * {32, 64, 96, 128}; filter_channels in {256, 512, 768, 1024}; filter_channels_dp = 256; kernel_size = kernel_size_dp = 3
 * (the prenet's kernel 5 and 3 layers are fixed by the reference, model.py:463-471); n_feats <= 256.
 * The weight table lists the reference's state-dict keys of `TextEncoder` (MatchaTTS prefix `encoder.`) in a fixed order;
 * "@rope_theta" is host-derived: 1 / 10000^(arange(0, d, 2) / d), d = (width / n_heads) / 2 (model.py:262, :319-320). */
typedef struct MttsTextHandle MttsTextHandle;
typedef struct MttsTextConfig {
  int n_vocab;
  int n_feats;
  int n_channels;
  int filter_channels;
  int n_heads;
  int n_layers;
  int kernel_size;
  int prenet;              /* 0 / 1 */
  int filter_channels_dp;
  int kernel_size_dp;
  int n_spks;
  int spk_emb_dim;
} MttsTextConfig;

int mtts_text_create(const MttsTextConfig* cfg, int device, MttsTextHandle** out);
void mtts_text_destroy(MttsTextHandle* h);
int mtts_text_num_weights(const MttsTextHandle* h);
const char* mtts_text_weight_name(const MttsTextHandle* h, int idx);
int64_t mtts_text_weight_numel(const MttsTextHandle* h, int idx);
size_t mtts_text_weight_arena_bytes(const MttsTextHandle* h);
int mtts_text_set_weight_arena(MttsTextHandle* h, void* dev_arena, size_t bytes, void* stream);
int mtts_text_load_weight(MttsTextHandle* h, int idx, const float* dev_src, int64_t numel, void* stream);
int mtts_text_weights_loaded(const MttsTextHandle* h);
size_t mtts_text_workspace_bytes(const MttsTextHandle* h, int B, int T_x);
int mtts_text_release_workspace(MttsTextHandle* h, const void* workspace, size_t workspace_bytes);
/* use_graph != 0 captures the call's ~57 launches into a CUDA graph (cached per pointers / shape; the stream must be
 * capturable, i.e. not the legacy default stream) */
int mtts_text_encoder_forward(MttsTextHandle* h, const int64_t* tokens, const int64_t* lengths, const float* spks, float* mu,
                              float* logw, float* x_mask, void* workspace, size_t workspace_bytes, int B, int T_x, int use_graph,
                              void* stream);
int mtts_text_last_launch_count(const MttsTextHandle* h);
/* introspection used by the parity tests: stop after n kernel launches (n < 0: run everything); byte offset of a named
 * intermediate ("X", "X1", "Y", "H", "O", "QKV", "F": fp16 rows of 256 / 768 / filter_channels columns, T_x + 2 rows per
 * utterance) inside the workspace for (B, T_x), -1 for an unknown name */
int mtts_text_debug_set_launch_limit(MttsTextHandle* h, int n);
int64_t mtts_text_debug_buffer_offset(const MttsTextHandle* h, int B, int T_x, const char* name);

/* ---- the step after the path: HiFi-GAN generator -----------------------------------------
 * Replaces the body of Generator.forward(mel) -> wav                               reference hifigan/models.py:181-195
 * (conv_pre, four ConvTranspose1d upsampling stages each followed by three ResBlock1 :14-98 whose outputs are averaged,
 * conv_post, tanh; configuration hifigan/config.py v1 -- what main.py:134-150 loads).
 * Same conventions as above: caller-owned buffers, no allocation, no synchronisation, errors as codes.
 *   mel (B, num_mels, T) fp32          wav (B, 1, hop * T) fp32 in (-1, 1), hop = product of the upsampling rates (256)
 * Every utterance is vocoded over all T frames (the reference passes the padded mel, main.py:197).
 * Supported: 1..4 upsampling stages with even rate u and kernel size 2u (padding u / 2), upsample_initial_channel a
 * multiple of 256 with upsample_initial_channel / 2^n_ups == 32; 1..3 resblocks per stage with odd kernel sizes and 1..3
 * dilations each, (kernel - 1) * dilation <= 50; num_mels a multiple of 8, <= 128.
 * The weight table lists the reference's state-dict keys AFTER remove_weight_norm() (models.py:197-205): "conv_pre.weight",
 * "ups.0.weight" ... "resblocks.11.convs2.2.bias", "conv_post.bias".  A weight-normed checkpoint (`*.weight_g`, `*.weight_v`,
 * what main.py:146-147 loads) is folded by the host first: weight = g * v / |v| with the norm over all but the first axis. */
typedef struct MttsVocHandle MttsVocHandle;
typedef struct MttsVocConfig {
  int num_mels;                         /* 80 */
  int upsample_initial_channel;         /* 512 */
  int n_ups;                            /* 4 */
  int upsample_rates[4];                /* 8 8 2 2 */
  int upsample_kernel_sizes[4];         /* 16 16 4 4 */
  int n_resblocks;                      /* 3 */
  int resblock_kernel_sizes[3];         /* 3 7 11 */
  int n_dilations;                      /* 3 */
  int resblock_dilation_sizes[3][3];    /* 1 3 5 for each resblock */
} MttsVocConfig;

int mtts_voc_create(const MttsVocConfig* cfg, int device, MttsVocHandle** out);
void mtts_voc_destroy(MttsVocHandle* h);
int mtts_voc_num_weights(const MttsVocHandle* h);
const char* mtts_voc_weight_name(const MttsVocHandle* h, int idx);
int64_t mtts_voc_weight_numel(const MttsVocHandle* h, int idx);
size_t mtts_voc_weight_arena_bytes(const MttsVocHandle* h);
int mtts_voc_set_weight_arena(MttsVocHandle* h, void* dev_arena, size_t bytes, void* stream);   /* 1024-byte aligned */
int mtts_voc_load_weight(MttsVocHandle* h, int idx, const float* dev_src, int64_t numel, void* stream);
int mtts_voc_weights_loaded(const MttsVocHandle* h);
int mtts_voc_hop_length(const MttsVocHandle* h);
size_t mtts_voc_workspace_bytes(const MttsVocHandle* h, int B, int T);
int mtts_voc_release_workspace(MttsVocHandle* h, const void* workspace, size_t workspace_bytes);
/* use_graph != 0 captures the call's 79 launches into a CUDA graph (cached per pointers / shape; the stream must be
 * capturable, i.e. not the legacy default stream) */
int mtts_voc_generator_forward(MttsVocHandle* h, const float* mel, float* wav, void* workspace, size_t workspace_bytes, int B, int T,
                               int use_graph, void* stream);
int mtts_voc_last_launch_count(const MttsVocHandle* h);
/* introspection used by the parity tests and the bench: stop after n kernel launches (n < 0: run everything); byte offset
 * of a named intermediate ("mel16", "a_in", "xa", "t_act", "r_act", "xs": channels-last fp16 [B][frames][C]
 * of the current level) inside the workspace for (B, T), -1 for an unknown name; per-launch timing like
 * mtts_debug_profile_begin / _end */
int mtts_voc_debug_set_launch_limit(MttsVocHandle* h, int n);
int64_t mtts_voc_debug_buffer_offset(const MttsVocHandle* h, int B, int T, const char* name);
int mtts_voc_debug_profile_begin(MttsVocHandle* h, void* stream);
int mtts_voc_debug_profile_end(MttsVocHandle* h, int max_entries, float* ms, int* kind, double* flops);

/* ---- Denoiser (reference hifigan/denoiser.py:12-68): torch.stft / torch.istft with n_fft = win_length = 1024, hop 256,
 * periodic Hann window, centred reflect-padded frames -- the configuration the reference constructs (filter_length 1024,
 * n_overlap 4) -- as hand-written fp32 FFT kernels.  Handle-free: they run on `device` and restore the caller's device.
 *   audio (B, n) fp32, n > 512          F = mtts_stft_frames(n) = 1 + n / 256 frames
 *   mtts_stft_magnitude:  mag (B, 513, F) = |stft(audio)|                                                    denoiser.py:29-39
 *   mtts_denoiser_forward: out (B, 256 * (F - 1)) = istft((|X| - bias_spec * strength)+ * exp(i arg X))      denoiser.py:62-68
 *                          bias_spec: 513 floats (denoiser.py:60); workspace: mtts_stft_workspace_bytes(B, n), 16-byte aligned */
int mtts_stft_frames(int n);
size_t mtts_stft_workspace_bytes(int B, int n);
int mtts_stft_magnitude(int device, const float* audio, float* mag, int B, int n, void* stream);
int mtts_denoiser_forward(int device, const float* audio, const float* bias_spec, float strength, float* out, void* workspace,
                          size_t workspace_bytes, int B, int n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MTTS_H_ */
