// capi.cu — the extern "C" layer declared in include/cpp_audio_b200.h: context, argument checks, host-buffer wrappers.
#include <stdlib.h>
#include <math.h>

#include <algorithm>
#include <new>

#include "ctx.hpp"

using namespace ca;

namespace {
// every entry point that takes a ctx runs on the ctx's device whatever the calling thread's current device is, and
// leaves the thread's current device as it found it
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(const ca_ctx* ctx) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != ctx->device) cudaSetDevice(ctx->device);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};
}  // namespace

extern "C" {

const char* ca_version(void) { return "cpp.audio_b200 0.1 (sm_100a)"; }

int ca_init(int device, ca_ctx** out) {
  if (!out) return CA_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return CA_ERR_CUDA;
  ca_ctx* ctx = new (std::nothrow) ca_ctx();
  if (!ctx) return CA_ERR_NOMEM;
  ctx->device = device;
  cudaDeviceProp prop{};
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    delete ctx;
    return CA_ERR_CUDA;
  }
  if (prop.major != 10) {  // the kernels are built for sm_100a only; there is no fallback
    delete ctx;
    return CA_ERR_UNSUPPORTED;
  }
  ctx->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->own_stream2, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return CA_ERR_CUDA;
  }
  *out = ctx;
  return CA_OK;
}

void ca_destroy(ca_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  stft_release(ctx);
  rfft_release(ctx);
  oscbank_release(ctx);
  for (void* d : ctx->d_scratch)
    if (d) cudaFree(d);
  for (int i = 0; i < 6; ++i)
    if (ctx->h_pinned[i]) cudaFreeHost(ctx->h_pinned[i]);
  for (cudaStream_t ps : ctx->pipe_streams)
    if (ps) cudaStreamDestroy(ps);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->own_stream2) cudaStreamDestroy(ctx->own_stream2);
  delete ctx;
}

const char* ca_last_error(const ca_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }
int ca_sm_count(const ca_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
int64_t ca_kernel_launches(const ca_ctx* ctx) { return ctx ? ctx->launches : 0; }

void* ca_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}
void ca_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int ca_fft_length_for(int n_samples, int zero_padding_factor) {
  long long want = static_cast<long long>(n_samples) * zero_padding_factor;
  int n = 1;
  while (n < want) n <<= 1;
  return n;
}

int64_t ca_stft_num_frames(int64_t n_samples, int window, int hop) {
  if (window <= 0 || hop <= 0 || n_samples < window) return 0;
  return 1 + (n_samples - window) / hop;
}

int ca_rfft_batched(ca_ctx* ctx, const float* d_in, int n_fft, int64_t batch, float* d_out, ca_stream stream) {
  if (!ctx) return CA_ERR_INVALID;
  DeviceGuard guard(ctx);
  return rfft_batched(ctx, d_in, n_fft, batch, d_out, static_cast<cudaStream_t>(stream));
}

int ca_irfft_batched(ca_ctx* ctx, const float* d_in, int n_fft, int64_t batch, float* d_out, float scale, ca_stream stream) {
  if (!ctx) return CA_ERR_INVALID;
  DeviceGuard guard(ctx);
  return irfft_batched(ctx, d_in, n_fft, batch, d_out, scale, static_cast<cudaStream_t>(stream));
}

int ca_rfft_batched_host(ca_ctx* ctx, const float* h_in, int n_fft, int64_t batch, float* h_out) {
  if (!ctx) return CA_ERR_INVALID;
  DeviceGuard guard(ctx);
  if (!h_in || !h_out || batch < 0 || n_fft < 2) return fail(ctx, CA_ERR_INVALID, "ca_rfft_batched_host: bad argument");
  const size_t in_bytes = sizeof(float) * static_cast<size_t>(batch) * n_fft;
  const size_t out_bytes = sizeof(float) * 2 * static_cast<size_t>(batch) * (n_fft / 2 + 1);
  int rc = ensure_scratch(ctx, 0, in_bytes);
  if (rc) return rc;
  rc = ensure_scratch(ctx, 1, out_bytes);
  if (rc) return rc;
  cudaStream_t s = ctx->own_stream;
  CA_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_scratch[0], h_in, in_bytes, cudaMemcpyHostToDevice, s));
  rc = rfft_batched(ctx, static_cast<const float*>(ctx->d_scratch[0]), n_fft, batch, static_cast<float*>(ctx->d_scratch[1]), s);
  if (rc) return rc;
  CA_CUDA_OK(ctx, cudaMemcpyAsync(h_out, ctx->d_scratch[1], out_bytes, cudaMemcpyDeviceToHost, s));
  CA_CUDA_OK(ctx, cudaStreamSynchronize(s));
  return CA_OK;
}

int ca_stft_plan(ca_ctx* ctx, const ca_stft_params* params) {
  if (!ctx) return CA_ERR_INVALID;
  DeviceGuard guard(ctx);
  ++ctx->plan_generation;
  return stft_plan(ctx, ctx->plan_user, params);
}

uint64_t ca_stft_plan_generation(const ca_ctx* ctx) { return ctx ? ctx->plan_generation : 0;
}

int ca_stft_analyze(ca_ctx* ctx, const float* d_x, int64_t n_samples, int64_t n_streams, int64_t stream_stride,
                    float* d_sqmag, int64_t sqmag_row_stride, int32_t* d_peak_bin, float* d_peak_freq, float* d_peak_db,
                    int32_t* d_peak_count, ca_stream stream) {
  if (!ctx) return CA_ERR_INVALID;
  DeviceGuard guard(ctx);
  return stft_analyze(ctx, ctx->plan_user, d_x, n_samples, n_streams, stream_stride, d_sqmag, sqmag_row_stride, d_peak_bin, d_peak_freq,
                      d_peak_db, d_peak_count, static_cast<cudaStream_t>(stream));
}

int ca_stft_analyze_host(ca_ctx* ctx, const float* h_x, int64_t n_samples, int64_t n_streams, int64_t stream_stride,
                         float* h_sqmag, int64_t sqmag_row_stride, int32_t* h_peak_bin, float* h_peak_freq,
                         float* h_peak_db, int32_t* h_peak_count) {
  if (!ctx) return CA_ERR_INVALID;
  DeviceGuard guard(ctx);
  if (!ctx->plan_user.planned) return fail(ctx, CA_ERR_NOT_PLANNED, "ca_stft_analyze_host: call ca_stft_plan first");
  if (!h_x || n_streams < 0 || n_samples < 0 || stream_stride < n_samples)
    return fail(ctx, CA_ERR_INVALID, "ca_stft_analyze_host: bad input shape");
  const bool any_peak = h_peak_bin || h_peak_freq || h_peak_db || h_peak_count;
  if (any_peak && !(h_peak_bin && h_peak_freq && h_peak_db && h_peak_count))
    return fail(ctx, CA_ERR_INVALID, "ca_stft_analyze_host: peak outputs must be all null or all non-null");
  const int64_t frames = ca_stft_num_frames(n_samples, ctx->plan_user.prm.window, ctx->plan_user.prm.hop);
  const size_t rows = static_cast<size_t>(frames) * n_streams;
  if (rows == 0) return CA_OK;
  const size_t x_bytes = sizeof(float) * static_cast<size_t>(n_streams) * stream_stride;
  const size_t sq_bytes = h_sqmag ? sizeof(float) * rows * sqmag_row_stride : 0;
  const size_t pk_bytes = any_peak ? 4 * rows * ctx->plan_user.prm.max_peaks : 0;
  const size_t cnt_bytes = any_peak ? 4 * rows : 0;
  int rc = ensure_scratch(ctx, 0, x_bytes);
  if (!rc && sq_bytes) rc = ensure_scratch(ctx, 1, sq_bytes);
  if (!rc && any_peak) rc = ensure_scratch(ctx, 2, pk_bytes ? pk_bytes : 4);
  if (!rc && any_peak) rc = ensure_scratch(ctx, 3, pk_bytes ? pk_bytes : 4);
  if (!rc && any_peak) rc = ensure_scratch(ctx, 4, pk_bytes ? pk_bytes : 4);
  if (!rc && any_peak) rc = ensure_scratch(ctx, 5, cnt_bytes);
  if (rc) return rc;
  // Copy/compute pipeline over two streams: the work is cut into chunks (groups of whole streams, or frame ranges of a
  // stream with their window-hop halo); chunk c+1's H2D overlaps chunk c's kernel and D2H (separate copy engines).
  const int W = ctx->plan_user.prm.window, hop = ctx->plan_user.prm.hop, P = ctx->plan_user.prm.max_peaks;
  float* d_x = static_cast<float*>(ctx->d_scratch[0]);
  float* d_sq = h_sqmag ? static_cast<float*>(ctx->d_scratch[1]) : nullptr;
  int32_t* d_pb = any_peak ? static_cast<int32_t*>(ctx->d_scratch[2]) : nullptr;
  float* d_pf = any_peak ? static_cast<float*>(ctx->d_scratch[3]) : nullptr;
  float* d_pd = any_peak ? static_cast<float*>(ctx->d_scratch[4]) : nullptr;
  int32_t* d_pc = any_peak ? static_cast<int32_t*>(ctx->d_scratch[5]) : nullptr;
  cudaStream_t lanes[2] = {ctx->own_stream, ctx->own_stream2};
  // 16 chunks: what the pipeline cannot overlap is the first chunk's copy up and the last chunk's kernel and copy down (with 8
  // chunks 1.3 of the 7.4 ms of a 32-stream config-2 step; CA_E2E_CHUNKS overrides, A/B)
  static const int64_t target_chunks = [] {
    const char* e = getenv("CA_E2E_CHUNKS");
    const int v = e ? atoi(e) : 16;
    return static_cast<int64_t>(v > 0 ? v : 16);
  }();
  int64_t streams_per_chunk = n_streams >= target_chunks ? (n_streams + target_chunks - 1) / target_chunks : 1;
  int64_t ranges_per_stream = n_streams >= target_chunks ? 1 : (target_chunks + n_streams - 1) / n_streams;
  if (ranges_per_stream > frames) ranges_per_stream = frames;
  // 16-byte aligned spans keep the vector / bulk-copy kernels eligible
  const bool can_split_frames = (hop % 4 == 0) && (stream_stride % 4 == 0);
  if (!can_split_frames) ranges_per_stream = 1;
  int lane_i = 0;
  for (int64_t s0 = 0; s0 < n_streams; s0 += streams_per_chunk) {
    const int64_t ns = std::min(streams_per_chunk, n_streams - s0);
    for (int64_t r = 0; r < ranges_per_stream; ++r) {
      const int64_t f0 = frames * r / ranges_per_stream, f1 = frames * (r + 1) / ranges_per_stream;
      if (f1 <= f0) continue;
      cudaStream_t s = lanes[lane_i++ & 1];
      const int64_t a0 = f0 * hop, a1 = (f1 - 1) * hop + W;  // sample span of the frame range
      const int64_t row0 = s0 * frames + f0;
      int rcc;
      if (ranges_per_stream == 1) {  // whole streams: one contiguous copy
        const size_t off = static_cast<size_t>(s0) * stream_stride;
        const size_t n = static_cast<size_t>(ns - 1) * stream_stride + n_samples;
        CA_CUDA_OK(ctx, cudaMemcpyAsync(d_x + off, h_x + off, sizeof(float) * n, cudaMemcpyHostToDevice, s));
        rcc = stft_analyze(ctx, ctx->plan_user, d_x + off, n_samples, ns, stream_stride, d_sq ? d_sq + row0 * sqmag_row_stride : nullptr, sqmag_row_stride,
                           d_pb ? d_pb + row0 * P : nullptr, d_pf ? d_pf + row0 * P : nullptr, d_pd ? d_pd + row0 * P : nullptr,
                           d_pc ? d_pc + row0 : nullptr, s);
      } else {  // ns == 1: a frame range of stream s0
        const size_t off = static_cast<size_t>(s0) * stream_stride + a0;
        CA_CUDA_OK(ctx, cudaMemcpyAsync(d_x + off, h_x + off, sizeof(float) * (a1 - a0), cudaMemcpyHostToDevice, s));
        rcc = stft_analyze(ctx, ctx->plan_user, d_x + off, a1 - a0, 1, a1 - a0, d_sq ? d_sq + row0 * sqmag_row_stride : nullptr, sqmag_row_stride,
                           d_pb ? d_pb + row0 * P : nullptr, d_pf ? d_pf + row0 * P : nullptr, d_pd ? d_pd + row0 * P : nullptr,
                           d_pc ? d_pc + row0 : nullptr, s);
      }
      if (rcc) return rcc;
      const int64_t nrows = ns * (ranges_per_stream == 1 ? frames : f1 - f0);
      if (h_sqmag)
        CA_CUDA_OK(ctx, cudaMemcpyAsync(h_sqmag + row0 * sqmag_row_stride, d_sq + row0 * sqmag_row_stride,
                                        sizeof(float) * nrows * sqmag_row_stride, cudaMemcpyDeviceToHost, s));
      if (any_peak) {
        if (P > 0) {
          CA_CUDA_OK(ctx, cudaMemcpyAsync(h_peak_bin + row0 * P, d_pb + row0 * P, 4 * nrows * P, cudaMemcpyDeviceToHost, s));
          CA_CUDA_OK(ctx, cudaMemcpyAsync(h_peak_freq + row0 * P, d_pf + row0 * P, 4 * nrows * P, cudaMemcpyDeviceToHost, s));
          CA_CUDA_OK(ctx, cudaMemcpyAsync(h_peak_db + row0 * P, d_pd + row0 * P, 4 * nrows * P, cudaMemcpyDeviceToHost, s));
        }
        CA_CUDA_OK(ctx, cudaMemcpyAsync(h_peak_count + row0, d_pc + row0, 4 * nrows, cudaMemcpyDeviceToHost, s));
      }
    }
  }
  CA_CUDA_OK(ctx, cudaStreamSynchronize(lanes[0]));
  CA_CUDA_OK(ctx, cudaStreamSynchronize(lanes[1]));
  return CA_OK;
}

int ca_oscbank_plan(ca_ctx* ctx, int64_t max_voices, int32_t max_partials) {
  if (!ctx) return CA_ERR_INVALID;
  DeviceGuard guard(ctx);
  return oscbank_plan(ctx, max_voices, max_partials);
}

int ca_oscbank_render(ca_ctx* ctx, const ca_oscbank_desc* desc, float* d_out, ca_stream stream) {
  if (!ctx) return CA_ERR_INVALID;
  DeviceGuard guard(ctx);
  int eased = -1;
  if (desc) {
    if (desc->env_kind < CA_ENV_UNKNOWN || desc->env_kind > CA_ENV_MIXED) return fail(ctx, CA_ERR_INVALID, "ca_oscbank_render: bad env_kind");
    eased = desc->env_kind == CA_ENV_ALL_LINEAR ? 0 : desc->env_kind == CA_ENV_ALL_EASED ? 2 : desc->env_kind == CA_ENV_MIXED ? 1 : -1;
  }
  return oscbank_render(ctx, desc, d_out, static_cast<cudaStream_t>(stream), eased);
}

int ca_oscbank_render_host(ca_ctx* ctx, const ca_oscbank_desc* desc, float* h_out) {
  if (!ctx) return CA_ERR_INVALID;
  DeviceGuard guard(ctx);
  if (!desc || !h_out) return fail(ctx, CA_ERR_INVALID, "ca_oscbank_render_host: null argument");
  const ca_oscbank_desc& h = *desc;
  if (h.n_voices < 0 || h.n_partials < 1 || h.n_samples < 0 || !h.f0_hz || !h.volume || !h.partial_volume ||
      !h.partial_phase || !h.env || (h.n_env != 1 && h.n_env != h.n_voices))
    return fail(ctx, CA_ERR_INVALID, "ca_oscbank_render_host: bad descriptor");
  int n_eased = 0;
  for (int64_t i = 0; i < h.n_env; ++i) {
    const ca_ahdsr& e = h.env[i];
    n_eased += (e.attack_itp | e.decay_itp | e.release_itp) != 0;
    if (e.attack_itp < 0 || e.attack_itp >= CA_ITP_COUNT || e.decay_itp < 0 || e.decay_itp >= CA_ITP_COUNT ||
        e.release_itp < 0 || e.release_itp >= CA_ITP_COUNT)
      return fail(ctx, CA_ERR_INVALID, "ca_oscbank_render_host: envelope interpolation is not a ca_itp value");
  }
  if (h.n_voices == 0 || h.n_samples == 0) return CA_OK;
  const size_t nv = static_cast<size_t>(h.n_voices), np = static_cast<size_t>(h.n_partials);
  // one parameter block: f0 | volume | start_angle | pan | key_delay | release_at | pvol | pphase | env
  size_t off = 0;
  auto take = [&off](size_t bytes) {
    size_t o = off;
    off += (bytes + 15) / 16 * 16;
    return o;
  };
  const size_t o_f0 = take(4 * nv), o_vol = take(4 * nv), o_sa = take(4 * nv), o_pan = take(4 * nv), o_kd = take(4 * nv),
               o_ra = take(8 * nv), o_pv = take(4 * nv * np), o_pp = take(4 * nv * np),
               o_env = take(sizeof(ca_ahdsr) * static_cast<size_t>(h.n_env));
  const size_t n_out = h.mix == CA_MIX_NONE ? nv * h.n_samples : static_cast<size_t>(h.n_samples) * (h.mix == CA_MIX_STEREO ? 2 : 1);
  int rc = ensure_scratch(ctx, 6, off);
  if (!rc) rc = ensure_scratch(ctx, 7, sizeof(float) * n_out);
  if (rc) return rc;
  cudaStream_t s = ctx->own_stream;
  char* base = static_cast<char*>(ctx->d_scratch[6]);
  auto up = [&](size_t o, const void* src, size_t bytes) -> cudaError_t {
    return src ? cudaMemcpyAsync(base + o, src, bytes, cudaMemcpyHostToDevice, s) : cudaSuccess;
  };
  CA_CUDA_OK(ctx, up(o_f0, h.f0_hz, 4 * nv));
  CA_CUDA_OK(ctx, up(o_vol, h.volume, 4 * nv));
  CA_CUDA_OK(ctx, up(o_sa, h.start_angle, 4 * nv));
  CA_CUDA_OK(ctx, up(o_pan, h.pan, 4 * nv));
  CA_CUDA_OK(ctx, up(o_kd, h.key_delay, 4 * nv));
  CA_CUDA_OK(ctx, up(o_ra, h.release_at, 8 * nv));
  CA_CUDA_OK(ctx, up(o_pv, h.partial_volume, 4 * nv * np));
  CA_CUDA_OK(ctx, up(o_pp, h.partial_phase, 4 * nv * np));
  CA_CUDA_OK(ctx, up(o_env, h.env, sizeof(ca_ahdsr) * static_cast<size_t>(h.n_env)));
  ca_oscbank_desc d = h;
  d.f0_hz = reinterpret_cast<const float*>(base + o_f0);
  d.volume = reinterpret_cast<const float*>(base + o_vol);
  d.start_angle = h.start_angle ? reinterpret_cast<const float*>(base + o_sa) : nullptr;
  d.pan = h.pan ? reinterpret_cast<const float*>(base + o_pan) : nullptr;
  d.key_delay = h.key_delay ? reinterpret_cast<const int32_t*>(base + o_kd) : nullptr;
  d.release_at = h.release_at ? reinterpret_cast<const int64_t*>(base + o_ra) : nullptr;
  d.partial_volume = reinterpret_cast<const float*>(base + o_pv);
  d.partial_phase = reinterpret_cast<const float*>(base + o_pp);
  d.env = reinterpret_cast<const ca_ahdsr*>(base + o_env);
  if (h.n_voices > ctx->osc_max_voices || h.n_partials > ctx->osc_max_partials) {
    rc = oscbank_plan(ctx, h.n_voices, h.n_partials);
    if (rc) return rc;
  }
  rc = oscbank_render(ctx, &d, static_cast<float*>(ctx->d_scratch[7]), s, n_eased == 0 ? 0 : n_eased == h.n_env ? 2 : 1);
  if (rc) return rc;
  CA_CUDA_OK(ctx, cudaMemcpyAsync(h_out, ctx->d_scratch[7], sizeof(float) * n_out, cudaMemcpyDeviceToHost, s));
  CA_CUDA_OK(ctx, cudaStreamSynchronize(s));
  return CA_OK;
}

int ca_resynth_roundtrip_host(ca_ctx* ctx, const ca_roundtrip_params* params, const float* h_x, int64_t n_samples,
                              int64_t n_streams, float* h_out, int64_t* h_stats) {
  if (!ctx) return CA_ERR_INVALID;
  DeviceGuard guard(ctx);
  return roundtrip_host(ctx, params, h_x, false, n_samples, n_streams, h_out, h_stats);
}

int ca_resynth_roundtrip_pcm16_host(ca_ctx* ctx, const ca_roundtrip_params* params, const int16_t* h_x, int64_t n_samples,
                                    int64_t n_streams, float* h_out, int64_t* h_stats) {
  if (!ctx) return CA_ERR_INVALID;
  DeviceGuard guard(ctx);
  return roundtrip_host(ctx, params, h_x, true, n_samples, n_streams, h_out, h_stats);
}

int64_t ca_resynth_truncated_frames(const ca_ctx* ctx) { return ctx ? ctx->rt_truncated_frames : 0; }

int ca_oscbank_render_tracks(ca_ctx* ctx, const ca_tracks_desc* desc, float* d_out, int64_t* h_stats, ca_stream stream) {
  if (!ctx) return CA_ERR_INVALID;
  DeviceGuard guard(ctx);
  return tracks_render(ctx, desc, d_out, h_stats, static_cast<cudaStream_t>(stream));
}

int ca_oscbank_render_tracks_host(ca_ctx* ctx, const ca_tracks_desc* desc, float* h_out, int64_t* h_stats) {
  if (!ctx) return CA_ERR_INVALID;
  DeviceGuard guard(ctx);
  return tracks_render_host(ctx, desc, h_out, h_stats);
}

int ca_tracks_open(ca_ctx* ctx, const ca_tracks_desc* desc, ca_tracks** tracks) {
  if (!ctx) return CA_ERR_INVALID;
  return tracks_open(ctx, desc, tracks);
}
int ca_tracks_post(ca_tracks* tracks, const ca_track_event* events, int64_t n_events) { return tracks_post(tracks, events, n_events); }
int ca_tracks_pull(ca_tracks* tracks, int64_t n_samples, float* d_out, ca_stream stream) {
  if (!tracks) return CA_ERR_INVALID;
  DeviceGuard guard(tracks_ctx(tracks));
  return tracks_pull(tracks, n_samples, d_out, static_cast<cudaStream_t>(stream));
}
int ca_tracks_pull_host(ca_tracks* tracks, int64_t n_samples, float* h_out) {
  if (!tracks) return CA_ERR_INVALID;
  DeviceGuard guard(tracks_ctx(tracks));
  return tracks_pull_host(tracks, n_samples, h_out);
}
int ca_tracks_state(const ca_tracks* tracks, int64_t state[5]) { return tracks_state(tracks, state); }
void ca_tracks_close(ca_tracks* tracks) { tracks_close(tracks); }

int ca_sound_engine_render_host(ca_ctx* ctx, const ca_sound_engine_desc* desc, float* h_out) {
  if (!ctx) return CA_ERR_INVALID;
  DeviceGuard guard(ctx);
  return sound_engine_render_host(ctx, desc, h_out);
}
int ca_sound_engine_specs(const ca_sound_engine_gen* gen, ca_ramp_spec* specs, int32_t capacity) { return sound_engine_specs(gen, specs, capacity); }
int ca_waveform_harmonics(int32_t waveform, int32_t n_harmonics, float* volume, float* phase) {
  return waveform_harmonics(waveform, n_harmonics, volume, phase);
}

int ca_vocoder_host(ca_ctx* ctx, const ca_vocoder_params* params, const float* h_modulator, const float* h_carrier,
                    int64_t n_samples, int64_t n_streams, float* h_out) {
  if (!ctx) return CA_ERR_INVALID;
  DeviceGuard guard(ctx);
  return vocoder_host(ctx, params, h_modulator, h_carrier, n_samples, n_streams, h_out);
}

int ca_ramp_render_host(ca_ctx* ctx, const ca_ramp_desc* desc, float* h_out) {
  if (!ctx) return CA_ERR_INVALID;
  DeviceGuard guard(ctx);
  return ramp_render_host(ctx, desc, h_out);
}

int ca_measure_fp32_peak(ca_ctx* ctx, double* tflops) {
  if (!ctx) return CA_ERR_INVALID;
  DeviceGuard guard(ctx);
  return measure_fp32_peak(ctx, tflops);
}

}  // extern "C"
