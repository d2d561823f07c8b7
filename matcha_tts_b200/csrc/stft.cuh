// Short-time Fourier kernels of the waveglow-style Denoiser (reference hifigan/denoiser.py:12-68): torch.stft / torch.istft with
// n_fft = win_length = 1024, hop 256, periodic Hann window, centred frames with reflect padding -- the only configuration the
// reference constructs (filter_length = 1024, n_overlap = 4) -- as hand-written fp32 FFTs in shared memory.
//
//   stft_denoise_kernel   one block per (frame, utterance): window the reflect-padded frame, 1024-point FFT, then either
//                         |X| -> mag (B, 513, F) (the bias spectrum, denoiser.py:56-60) or the denoising step
//                         X' = X * max(|X| - bias * strength, 0) / |X|   ( == (|X| - b s)+ * (cos phi, sin phi), :64-67 )
//                         followed by the inverse FFT of the Hermitian extension, the synthesis window, -> frames (B, F, 1024)
//   istft_ola_kernel      overlap-add of the four frames that cover a sample, in a fixed order, divided by the window
//                         envelope sum w^2, centre padding trimmed (torch.istft: length = hop * (F - 1))
//
// The FFT is an in-place radix-2 decimation-in-time over 1024 complex points (bit-reversed load, ten stages of 512
// butterflies on 256 threads, twiddles from a shared table built with sincospif): ~2e-7 relative error, nothing here is
// performance-critical (22 k frames per 64 x 344-frame batch).
#pragma once
#include <cuda_runtime.h>

#include "ptx.cuh"

namespace mtts {

constexpr int STFT_N = 1024;
constexpr int STFT_HOP = 256;
constexpr int STFT_BINS = STFT_N / 2 + 1;
constexpr int STFT_THREADS = 256;

__device__ __forceinline__ int stft_bitrev10(int i) { return (int)(__brev((unsigned)i) >> 22); }

// ten radix-2 stages over buf[1024] (already in bit-reversed order); tw[k] = exp(-2 pi i k / 1024), k < 512
__device__ __forceinline__ void stft_fft_stages(float2* buf, const float2* tw) {
#pragma unroll 1
  for (int s = 1; s <= 10; ++s) {
    const int half = 1 << (s - 1);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int j = threadIdx.x + r * STFT_THREADS;
      const int pos = j & (half - 1);
      const int i0 = ((j >> (s - 1)) << s) + pos, i1 = i0 + half;
      const float2 w = tw[pos << (10 - s)];
      const float2 a = buf[i0], b = buf[i1];
      const float2 t = make_float2(w.x * b.x - w.y * b.y, w.x * b.y + w.y * b.x);
      buf[i0] = make_float2(a.x + t.x, a.y + t.y);
      buf[i1] = make_float2(a.x - t.x, a.y - t.y);
    }
  }
  __syncthreads();
}

// audio (B, n); frame f covers padded samples [f * hop, f * hop + 1024), padded[i] = audio[reflect(i - 512)]
__global__ void __launch_bounds__(STFT_THREADS)
stft_denoise_kernel(const float* __restrict__ audio, int n, int F, const float* __restrict__ bias, float strength,
                    float* __restrict__ mag_out, float* __restrict__ frames_out) {
  __shared__ float2 buf[STFT_N];
  __shared__ float2 buf2[STFT_N];
  __shared__ float2 tw[STFT_N / 2];
  __shared__ float win[STFT_N];
  const int f = blockIdx.x, b = blockIdx.y;
  for (int k = threadIdx.x; k < STFT_N / 2; k += STFT_THREADS) {
    float sn, cs;
    sincospif((float)k / 512.f, &sn, &cs);
    tw[k] = make_float2(cs, -sn);
  }
  for (int i = threadIdx.x; i < STFT_N; i += STFT_THREADS) win[i] = 0.5f - 0.5f * cospif((float)i / 512.f);   // periodic Hann
  __syncthreads();
  const float* x = audio + (size_t)b * n;
  for (int i = threadIdx.x; i < STFT_N; i += STFT_THREADS) {
    int j = f * STFT_HOP + i - STFT_N / 2;
    if (j < 0) j = -j;
    if (j >= n) j = 2 * (n - 1) - j;
    buf[stft_bitrev10(i)] = make_float2(x[j] * win[i], 0.f);
  }
  stft_fft_stages(buf, tw);
  if (mag_out) {
    for (int k = threadIdx.x; k < STFT_BINS; k += STFT_THREADS) {
      const float2 v = buf[k];
      mag_out[((size_t)b * STFT_BINS + k) * F + f] = sqrtf(v.x * v.x + v.y * v.y);
    }
  }
  if (!frames_out) return;
  // X' on bins 0..512, Hermitian extension, conjugated for the inverse transform: ifft(X) = conj(fft(conj(X))) / N
  for (int k = threadIdx.x; k < STFT_BINS; k += STFT_THREADS) {
    float2 v = buf[k];
    const float mag = sqrtf(v.x * v.x + v.y * v.y);
    const float sc = mag > 0.f ? fmaxf(mag - bias[k] * strength, 0.f) / mag : 0.f;
    v.x *= sc; v.y *= sc;
    if (k == 0 || k == STFT_N / 2) v.y = 0.f;              // irfft ignores the imaginary part of the DC and Nyquist bins
    buf2[stft_bitrev10(k)] = make_float2(v.x, -v.y);        // conj(X'[k])
    if (k > 0 && k < STFT_N / 2) buf2[stft_bitrev10(STFT_N - k)] = make_float2(v.x, v.y);   // conj(X'[N - k]) = X'[k]
  }
  stft_fft_stages(buf2, tw);
  float* out = frames_out + ((size_t)b * F + f) * STFT_N;
  for (int i = threadIdx.x; i < STFT_N; i += STFT_THREADS) out[i] = buf2[i].x * (1.f / STFT_N) * win[i];
}

// out (B, hop * (F - 1)): sample i sits at padded position p = i + 512, covered by frames f with 0 <= p - f * hop < 1024
__global__ void istft_ola_kernel(const float* __restrict__ frames, int F, float* __restrict__ out, int n_out) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const int p = i + STFT_N / 2;
  const int f_hi = min(F - 1, p / STFT_HOP);
  const int f_lo = p >= STFT_N ? (p - STFT_N) / STFT_HOP + 1 : 0;
  float acc = 0.f, env = 0.f;
  for (int f = f_lo; f <= f_hi; ++f) {
    const int k = p - f * STFT_HOP;
    const float w = 0.5f - 0.5f * cospif((float)k / 512.f);
    acc += frames[((size_t)b * F + f) * STFT_N + k];
    env = fmaf(w, w, env);
  }
  out[(size_t)b * n_out + i] = acc / env;
}

}  // namespace mtts
