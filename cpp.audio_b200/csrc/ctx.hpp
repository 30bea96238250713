// ctx.hpp — the per-GPU context behind the C ABI (include/cpp_audio_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/cpp_audio_b200.h"

namespace ca {
// one planned analysis: parameters, FFT length and the window table on the device
struct StftPlan {
  bool planned = false;
  ca_stft_params prm{};
  int n_fft = 0;
  float* d_window = nullptr;  // [n_fft] window, zero padded, pre-scaled (see stft.cu)
  float thr_sq = 0.f;
  bool freq_as_offset = false;  // internal (round trip, FP32 A/B form): peak_freq receives the sub-bin offset
  // FP64 form of the analysis (stft64.cu; the round trip's default): window [n_fft] in double, twiddles W_{n_fft}^k, k <= n_fft/2
  double* d_window64 = nullptr;
  double2* d_tw64 = nullptr;
  int n64 = 0;
  // internal (vocoder, FFT lengths >= 8192): per-frame band amplitudes sqrt(sum of |X|^2 over the bins [band_begin[b], band_begin[b+1])),
  // summed from the kernel's shared-memory copy of the spectrum, so the |X|^2 rows need not go through HBM; set before each
  // stft_analyze call (band_amps: [frames of the call][n_bands])
  const int32_t* band_begin = nullptr;
  int n_bands = 0;
  float* band_amps = nullptr;
};
}  // namespace ca

struct ca_ctx {
  int device = 0;
  int sm_count = 0;
  int64_t launches = 0;
  std::string last_error;
  cudaStream_t own_stream = nullptr;   // used by the *_host entry points
  cudaStream_t own_stream2 = nullptr;  // second lane of the copy/compute pipeline of ca_stft_analyze_host
  cudaStream_t pipe_streams[8] = {};  // stream-group lanes of the round trip / vocoder (created on first use)

  // analysis plans: the caller's (ca_stft_plan) and the private ones of the round trip and the vocoder, so that a
  // ca_resynth_roundtrip_host / ca_vocoder_host call never disturbs what ca_stft_analyze was planned for
  ca::StftPlan plan_user, plan_rt, plan_voc;
  uint64_t plan_generation = 0;  // ca_stft_plan calls so far

  int64_t rt_truncated_frames = 0;  // last round trip: frames with more peaks than stft.max_peaks (ca_resynth_truncated_frames)

  // generic FFT scratch (twiddle tables, four-step buffers)
  float2* d_fft_tw = nullptr;
  int fft_tw_n = 0;
  float2* d_fft_work = nullptr;
  size_t fft_work_bytes = 0;

  // synthesis plan
  int64_t osc_max_voices = 0;
  int32_t osc_max_partials = 0;
  void* d_osc_partials = nullptr;  // PartialParams[max_voices][padded partials]
  void* d_osc_voices = nullptr;    // VoiceParams[max_voices]
  void* d_osc_umma = nullptr;      // operand tables of the tcgen05 kernel, 128 KB per voice (allocated on first use)
  size_t osc_umma_bytes = 0;

  // scratch for the *_host entry points
  void* d_scratch[16] = {};
  size_t scratch_bytes[16] = {};
  // pinned host staging of the round trip (peak lists coming down, segments going up)
  void* h_pinned[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t pinned_bytes[6] = {0, 0, 0, 0, 0, 0};
};

namespace ca {

// arguments of the analysis kernels (stft.cu: the fused 4096-point kernel; rfftw.cu: the fused kernels of the other lengths)
struct StftArgs {
  const float* x;
  int64_t n_samples;
  int64_t stream_stride;
  int64_t frames_per_stream;
  int64_t total_frames;
  int32_t n_streams;
  int32_t hop;
  int32_t window;       // samples actually read per frame (<= 4096); the rest is zero padding
  const float* win;     // [4096] window * 1/(2W), zero beyond `window`
  float* sqmag;         // nullable
  int64_t sqmag_row_stride;
  int32_t* peak_bin;    // nullable
  float* peak_freq;
  float* peak_db;
  int32_t* peak_count;
  int32_t max_peaks;
  float thr_below;  // largest float strictly below the sqmag threshold (-1 when there is no threshold)
  float bin_to_hz;
  int32_t freq_as_offset;  // internal: peak_freq receives the sub-bin offset p (round trip computes k+p in double)
  const int32_t* band_begin;  // internal (vocoder): see StftPlan
  int32_t n_bands;
  float* band_amps;
  int32_t step_streams, step_frames;  // stft4096p_kernel: its grid stride (frames) split as streams * frames_per_stream + frames
  long long* dbg;          // CA_STFT_TIMELINE builds: per-phase clock64 stamps of a few warps
};

inline int fail(ca_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->last_error = msg;
  return code;
}

#define CA_CUDA_OK(ctx, expr)                                                                          \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess)                                                                             \
      return ca::fail((ctx), CA_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));         \
  } while (0)

inline int ensure_scratch(ca_ctx* ctx, int slot, size_t bytes) {
  if (ctx->scratch_bytes[slot] >= bytes) return CA_OK;
  if (ctx->d_scratch[slot]) cudaFree(ctx->d_scratch[slot]);
  ctx->d_scratch[slot] = nullptr;
  ctx->scratch_bytes[slot] = 0;
  CA_CUDA_OK(ctx, cudaMalloc(&ctx->d_scratch[slot], bytes));
  ctx->scratch_bytes[slot] = bytes;
  return CA_OK;
}

inline int ensure_pinned(ca_ctx* ctx, int slot, size_t bytes) {
  if (ctx->pinned_bytes[slot] >= bytes) return CA_OK;
  if (ctx->h_pinned[slot]) cudaFreeHost(ctx->h_pinned[slot]);
  ctx->h_pinned[slot] = nullptr;
  ctx->pinned_bytes[slot] = 0;
  bytes += bytes / 4;  // headroom: the segment count varies from call to call
  CA_CUDA_OK(ctx, cudaHostAlloc(&ctx->h_pinned[slot], bytes, cudaHostAllocDefault));
  ctx->pinned_bytes[slot] = bytes;
  return CA_OK;
}

// entry points implemented per kernel family
int stft_plan(ca_ctx* ctx, StftPlan& plan, const ca_stft_params* p);
int stft_analyze(ca_ctx* ctx, const StftPlan& plan, const float* d_x, int64_t n_samples, int64_t n_streams, int64_t stream_stride,
                 float* d_sqmag, int64_t sqmag_row_stride, int32_t* d_peak_bin, float* d_peak_freq, float* d_peak_db,
                 int32_t* d_peak_count, cudaStream_t stream);
int rfft_batched(ca_ctx* ctx, const float* d_in, int n_fft, int64_t batch, float* d_out, cudaStream_t stream);
int irfft_batched(ca_ctx* ctx, const float* d_in, int n_fft, int64_t batch, float* d_out, float scale, cudaStream_t stream);
int rfft_warp(ca_ctx* ctx, const float* d_in, int n_fft, int64_t batch, float* d_out, cudaStream_t stream);  // N <= 4096
int oscbank_plan(ca_ctx* ctx, int64_t max_voices, int32_t max_partials);
int oscbank_render(ca_ctx* ctx, const ca_oscbank_desc* desc, float* d_out, cudaStream_t stream, int eased = -1);
int measure_fp32_peak(ca_ctx* ctx, double* tflops);
// h_x: float32 samples, or int16 PCM when pcm16 is set (converted on the device: x / 32768, exact)
int roundtrip_host(ca_ctx* ctx, const ca_roundtrip_params* params, const void* h_x, bool pcm16, int64_t n_samples, int64_t n_streams,
                   float* h_out, int64_t* h_stats);
int tracks_open(ca_ctx* ctx, const ca_tracks_desc* desc, ca_tracks** out);
int tracks_post(ca_tracks* tp, const ca_track_event* events, int64_t n);
int tracks_pull(ca_tracks* tp, int64_t n_samples, float* d_out, cudaStream_t st);
int tracks_pull_host(ca_tracks* tp, int64_t n_samples, float* h_out);
ca_ctx* tracks_ctx(const ca_tracks* tp);
int tracks_state(const ca_tracks* tp, int64_t state[5]);
void tracks_close(ca_tracks* tp);
int tracks_render(ca_ctx* ctx, const ca_tracks_desc* desc, float* d_out, int64_t* h_stats, cudaStream_t st);
int tracks_render_host(ca_ctx* ctx, const ca_tracks_desc* desc, float* h_out, int64_t* h_stats);
int sound_engine_render_host(ca_ctx* ctx, const ca_sound_engine_desc* desc, float* h_out);
int sound_engine_specs(const ca_sound_engine_gen* gen, ca_ramp_spec* specs, int32_t capacity);
int waveform_harmonics(int32_t waveform, int32_t n, float* volume, float* phase);
int ramp_render_host(ca_ctx* ctx, const ca_ramp_desc* desc, float* h_out);
int vocoder_host(ca_ctx* ctx, const ca_vocoder_params* prm, const float* h_modulator, const float* h_carrier, int64_t n_samples,
                 int64_t n_streams, float* h_out);
// fused window -> FFT -> |X|^2 -> peaks for n_fft = 1024, 2048, 8192, 16384, 32768 (rfftw.cu); CA_ERR_UNSUPPORTED otherwise
int stft_fused_generic(ca_ctx* ctx, const StftArgs& a, int n_fft, bool vec2, cudaStream_t stream);
void stft_release(ca_ctx* ctx);
// FP64 form of the analysis, peaks only (stft64.cu): n_fft 1024 .. 16384. d_peak_off = sub-bin offset p, d_peak_db = dB, both double.
int stft64_plan(ca_ctx* ctx, StftPlan& plan);
bool stft64_supported(const StftPlan& plan);
int stft64_peaks(ca_ctx* ctx, const StftPlan& plan, const float* d_x, int64_t n_samples, int64_t n_streams, int64_t stream_stride,
                 int32_t* d_peak_bin, double* d_peak_off, double* d_peak_db, int32_t* d_peak_count, cudaStream_t stream);
void rfft_release(ca_ctx* ctx);
void oscbank_release(ca_ctx* ctx);

}  // namespace ca
