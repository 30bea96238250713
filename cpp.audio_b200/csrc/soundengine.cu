// soundengine.cu — gen.voice voices: SoundEngine (include/soundengine.h:240-468, :738-803) rendered on the GPU.
//
// Reference: a voice of voice::Impl_ (include/gen.voice.h:1093-1189) is VolumeAdjusted<SoundEngine>; the engine owns three ramp
// elements VolumeAdjusted<Enveloped<FreqCtrl_<Mix<LoudnessVolumeAdjusted<Sine>>, LogRamp>, AHDSR>> (soundengine.h:245-257,
// MixOf :209-216) and a list of ramp specs that a Markov chain fills at note-on (:496-736). Every sample, orchestrate_algos
// (:744-772) decides whether the next spec starts: it is copied to a free ramp (playNextSpec, :774-803), which fades in over
// xfade_len samples while the ramp before it fades out; step_algos / imag (:412-464) step and sum the sounding ramps.
//
// Here: the orchestration is control rate, so the host walks it once per voice and turns every spec into a PLAY: the sample at
// which the ramp starts, the number of steps before its release starts, its LogRamp and its volume. One warp renders one voice,
// play after play, 32 samples per step:
//   frequency   LogRamp in closed form per leg (tracks.cuh Leg): position over duration u_i = u0 + Ku expm1(L i), emitted
//               increment from + delta * e(u_i) with e one of the 22 interpolations (ca_itp) -- the position dynamics do not
//               depend on the interpolation (audioelement.h:2537-2540)
//   phase       running sum of the increments: a warp prefix sum in FP64 (no closed form for eased ramps)
//   loudness    LoudnessVolumeAdjusted (audioelement.h:1275-1316): target = equal-loudness amplitude of the current frequency
//               (include/loudness.h), one-pole low-pass whose rate is the current increment (BaseVolumeAdjusted::step,
//               :1195-1216; oracle/SPEC.md S8): a first-order recurrence with varying coefficients = a warp scan of affine maps
//   envelope    AHDSR{xfade_len, LINEAR, 0, 0, LINEAR, xfade_len, LINEAR, 1} (soundengine.h:790), closed form in the step index
// The plays of a voice overlap (cross-fades): the warp adds them into the voice's output one after the other, so the sum has a
// fixed order.
#include <math.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "common.cuh"
#include "ctx.hpp"
#include "pitch.cuh"
#include "tracks.cuh"

namespace ca {

namespace {

struct Play {
  double from, to;  // angle increments (freq_to_angle_increment in float, include/sound.functions.h:58-62)
  double dur, c_start;
  int64_t s0;        // output sample of the ramp's first step (relative to the voice's row)
  int32_t n_press;   // ramp steps before the release starts
  int32_t attack, release;
  int32_t itp;
  float volume;      // spec volume * velocity * sine gain
};

struct VoiceRec {
  int32_t play_begin, play_end;
  float gl, gr;
  int64_t out_base;  // first element of the voice's output (floats; stereo: pairs)
};

struct EngineArgs {
  const Play* plays;
  const VoiceRec* voices;
  int64_t n_voices;
  int64_t n_samples;
  float* out;
  int32_t stereo;
  int32_t sample_rate;
  int32_t loudness_on;
  float log_ratio, max_db;
  float freqs[kLoudN], elv[kLoudN];
};

__device__ double ease64(int type, double u) {  // e(u) of itp::interpolate = b + c e(t/d); oracle/SPEC.md S7
  if (type == CA_ITP_LINEAR) return u;
  const int fam = (type - 1) / 3, mode = (type - 1) % 3;
  auto ipow = [](double x, int p) {
    double r = 1;
    for (int i = 0; i < p; ++i) r *= x;
    return r;
  };
  const double kHalfPi = 1.5707963267948966192313216916398;
  if (fam <= 3) {
    const int p = fam + 2;
    if (mode == 0) return ipow(u, p);
    if (mode == 1) return 1.0 - ipow(1.0 - u, p);
    return u < 0.5 ? 0.5 * ipow(2.0 * u, p) : 1.0 - 0.5 * ipow(2.0 - 2.0 * u, p);
  }
  if (fam == 4) {
    if (mode == 0) return 1.0 - cos(u * kHalfPi);
    if (mode == 1) return sin(u * kHalfPi);
    return 0.5 * (1.0 - cos(2.0 * kHalfPi * u));
  }
  if (fam == 5) {
    if (mode == 0) return u == 0.0 ? 0.0 : exp2(10.0 * (u - 1.0));
    if (mode == 1) return u == 1.0 ? 1.0 : 1.0 - exp2(-10.0 * u);
    if (u == 0.0) return 0.0;
    if (u == 1.0) return 1.0;
    return u < 0.5 ? 0.5 * exp2(10.0 * (2.0 * u - 1.0)) : 0.5 * (2.0 - exp2(-10.0 * (2.0 * u - 1.0)));
  }
  auto root = [](double x) { return sqrt(x > 0.0 ? x : 0.0); };
  if (mode == 0) return 1.0 - root(1.0 - u * u);
  if (mode == 1) return root(1.0 - (u - 1.0) * (u - 1.0));
  const double w = 2.0 * u;
  return w < 1.0 ? 0.5 * (1.0 - root(1.0 - w * w)) : 0.5 * (root(1.0 - (w - 2.0) * (w - 2.0)) + 1.0);
}

// equal_loudness_volume_from_freq (include/loudness.h:258-275), float as the reference
__device__ float loudness_amp(const EngineArgs& a, float freq) {
  int lo = 0, hi = kLoudN - 1;
  do {
    const int c = (lo + hi) / 2;
    if (freq < a.freqs[c]) hi = c;
    else lo = c;
  } while (hi - lo >= 2);
  float db;
  if (freq <= a.freqs[lo]) db = a.elv[lo];
  else if (freq >= a.freqs[hi]) db = a.elv[hi];
  else {
    const float ratio = (freq - a.freqs[lo]) / (a.freqs[hi] - a.freqs[lo]);
    db = ratio * a.elv[hi] + (1.f - ratio) * a.elv[lo];
  }
  if (db > a.max_db) return 1.f;
  return powf(10.f, a.log_ratio * (db - a.max_db) / 20.f);
}

__global__ void __launch_bounds__(128) sound_engine_kernel(const __grid_constant__ EngineArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t v = static_cast<int64_t>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (v >= a.n_voices) return;
  const VoiceRec vr = a.voices[v];
  constexpr unsigned kFull = 0xffffffffu;
  for (int pi = vr.play_begin; pi < vr.play_end; ++pi) {
    const Play p = a.plays[pi];
    const int64_t n_total = min(static_cast<int64_t>(p.n_press) + p.release, a.n_samples - p.s0);
    const float top = static_cast<float>(min(p.n_press, p.attack)) / static_cast<float>(p.attack);  // envelope value when the release starts
    double lf = p.from, lt = p.to, cs = p.c_start;
    Leg leg;
    leg.start(lf, lt, p.dur, cs);
    int64_t li = 0;      // step index inside the leg
    double phase = 0.0;  // setStartAngle(start_angle), start_angle = 0 (soundengine.h:320-323, gen.voice.h:1100)
    float y_carry = 0.f;
    for (int64_t j = 0; j < n_total;) {
      if (li >= leg.m) {  // the ramp turns around (audioelement.h:2523-2526)
        const double t = lf;
        lf = lt;
        lt = t;
        cs = 0.0;
        leg.start(lf, lt, p.dur, cs);
        li = 0;
      }
      const int chunk = static_cast<int>(min(min(static_cast<int64_t>(32), leg.m - li), n_total - j));
      const bool on = lane < chunk;
      const int64_t i = li + lane;
      double f = lf;
      if (!leg.steady) {
        const double x = expm1(leg.L * static_cast<double>(i));
        f = p.itp == CA_ITP_LINEAR ? leg.from1 + leg.K1 * x : lf + (lt - lf) * ease64(p.itp, leg.u0 + leg.Ku * x);
      }
      // ---- phase: inclusive prefix sum of the increments ----
      double s = on ? f : 0.0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double o = __shfl_up_sync(kFull, s, d);
        if (lane >= d) s += o;
      }
      double ph = phase + s;
      ph -= 2.0 * floor(0.5 * ph);
      phase += __shfl_sync(kFull, s, 31);
      phase -= 2.0 * floor(0.5 * phase);
      const float ff = static_cast<float>(f);
      float tone = alias_gain(ff) * sinpif(static_cast<float>(ph));
      // ---- loudness volume: y_j = (1 - a_j) y_{j-1} + a_j t_j, y_0 = t_0 ----
      if (a.loudness_on) {
        const float target = loudness_amp(a, static_cast<float>(f * a.sample_rate / 2));
        const float k = 1.f - expf(-3.14159265358979f * fabsf(ff));
        float A = (j + lane == 0) ? 0.f : 1.f - k, B = (j + lane == 0) ? target : k * target;
        if (!on) {
          A = 1.f;
          B = 0.f;
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const float Ap = __shfl_up_sync(kFull, A, d), Bp = __shfl_up_sync(kFull, B, d);
          if (lane >= d) {
            B = fmaf(A, Bp, B);
            A *= Ap;
          }
        }
        const float y = fmaf(A, y_carry, B);
        y_carry = __shfl_sync(kFull, y, 31);
        tone *= y;
      }
      // ---- envelope ----
      const int c = static_cast<int>(j) + lane + 1;  // ramp step, 1-based
      float env;
      if (c <= p.n_press) env = c <= p.attack ? static_cast<float>(c) / static_cast<float>(p.attack) : 1.f;
      else {
        const int r = c - p.n_press;
        env = r >= p.release ? 0.f : top - top * (static_cast<float>(r) / static_cast<float>(p.release));
      }
      if (on) {
        const float y = p.volume * (tone * env);
        const int64_t n = p.s0 + j + lane;
        if (a.stereo) {
          float2* o = reinterpret_cast<float2*>(a.out + vr.out_base) + n;
          float2 cur = *o;
          cur.x = fmaf(vr.gl, y, cur.x);
          cur.y = fmaf(vr.gr, y, cur.y);
          *o = cur;
        } else {
          a.out[vr.out_base + n] += y;
        }
      }
      j += chunk;
      li += chunk;
    }
    __syncwarp();  // the next play adds into samples this one wrote
  }
}

// the voices of a mixed bus are rendered into rows of their own and summed here, in voice order (a fixed order of additions)
__global__ void __launch_bounds__(256) mix_rows_kernel(const float* __restrict__ rows, int64_t n_rows, int64_t row_len, float* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= row_len) return;
  float acc = 0.f;
  for (int64_t r = 0; r < n_rows; ++r) acc += rows[r * row_len + i];
  out[i] = acc;
}

// orchestrate_algos / playNextSpec over the whole life of one voice -> plays
void voice_plays(const ca_sound_engine_desc& d, int64_t v, std::vector<Play>& out) {
  const int sr = d.sample_rate;
  const int32_t min_dt = sr / 1000;
  const int32_t A = std::max(d.xfade_len, min_dt), R = A;  // setAHDSR (audioelement.h:923-952); no min-change floor: nothing sets it
  const int64_t key_on = d.key_on ? std::max<int64_t>(d.key_on[v], 0) : 0;
  const int64_t key_off = d.key_off ? d.key_off[v] : -1;
  const float velocity = d.velocity ? d.velocity[v] : 1.f;
  int64_t free_at[3] = {0, 0, 0};  // engine step from which the ramp is seen in EnvelopeDone2
  int64_t q = key_on;              // engine step (= output sample) at which the next spec may start
  for (int k = d.spec_begin[v]; k < d.spec_begin[v + 1]; ++k) {
    const ca_ramp_spec& sp = d.specs[k];
    // the free ramp with the highest index (get_ramps, soundengine.h:277-295); none: wait for the first one to finish
    int ramp = -1;
    for (int r = 0; r < 3; ++r)
      if (free_at[r] <= q) ramp = r;
    if (ramp < 0) {
      q = std::min({free_at[0], free_at[1], free_at[2]});
      for (int r = 0; r < 3; ++r)
        if (free_at[r] <= q) ramp = r;
    }
    if (key_off >= 0 && q >= key_off) break;  // the note was released: nothing starts any more (goOn, :357-365)
    if (q >= d.n_samples) break;
    int32_t D = std::max(static_cast<int32_t>(.5f + sp.duration) - d.xfade_len, 0);
    int64_t next_q;
    if (D == 0) {  // pressed and released at once: the note is skipped (audioelement.h:774-777), the ramp idles through EnvelopeDone1
      free_at[ramp] = q + 17;
      next_q = q + 1;
      q = next_q;
      continue;
    }
    next_q = q + D + (sp.silence_follows ? d.articulative_pause : 0);
    if (key_off >= 0 && key_off > q && key_off - q < D) D = static_cast<int32_t>(key_off - q);  // onKeyReleased(0) while pressed
    Play p{};
    p.from = static_cast<double>(2 * sp.from_hz / sr);
    p.to = static_cast<double>(2 * sp.to_hz / sr);
    p.dur = static_cast<double>(sp.duration);
    p.c_start = static_cast<double>(sp.start_sample);
    p.s0 = q;
    p.n_press = D;
    p.attack = A;
    p.release = R;
    p.itp = sp.itp;
    p.volume = sp.volume * velocity * d.sine_gain;
    out.push_back(p);
    free_at[ramp] = q + D + R + 17;
    q = next_q;
  }
}

}  // namespace

int sound_engine_render_host(ca_ctx* ctx, const ca_sound_engine_desc* desc, float* h_out) {
  if (!desc || !h_out || desc->n_voices < 0 || desc->n_samples < 0 || desc->sample_rate <= 0 || !desc->spec_begin || (!desc->specs && desc->n_voices > 0))
    return fail(ctx, CA_ERR_INVALID, "ca_sound_engine_render_host: null argument or negative size");
  const ca_sound_engine_desc d = *desc;
  if (d.mix != CA_MIX_NONE && d.mix != CA_MIX_MONO && d.mix != CA_MIX_STEREO)
    return fail(ctx, CA_ERR_INVALID, "ca_sound_engine_render_host: mix is not a ca_mix value");
  if (d.xfade_len < 0 || d.articulative_pause < 0 || d.loudness_ref_index < 0 || d.loudness_ref_index >= kLoudN)
    return fail(ctx, CA_ERR_INVALID, "ca_sound_engine_render_host: bad engine parameters");
  for (int64_t v = 0; v < d.n_voices; ++v) {
    if (d.spec_begin[v + 1] < d.spec_begin[v] || d.spec_begin[v + 1] - d.spec_begin[v] > 30)
      return fail(ctx, CA_ERR_INVALID, "ca_sound_engine_render_host: a voice holds 0 .. 30 ramp specs (RampSpecs::n_specs)");
    for (int k = d.spec_begin[v]; k < d.spec_begin[v + 1]; ++k) {
      const ca_ramp_spec& sp = d.specs[k];
      if (!(sp.from_hz > 0.f) || !(sp.to_hz > 0.f) || !(sp.duration > 0.f) || sp.start_sample < 0.f || sp.itp < 0 || sp.itp >= CA_ITP_COUNT)
        return fail(ctx, CA_ERR_INVALID, "ca_sound_engine_render_host: a ramp spec needs positive frequencies and duration and a ca_itp value");
    }
  }
  const bool stereo = d.mix == CA_MIX_STEREO, per_voice = d.mix == CA_MIX_NONE;
  const size_t row = static_cast<size_t>(d.n_samples) * (stereo ? 2 : 1);
  const size_t out_elems = row * (per_voice ? static_cast<size_t>(d.n_voices) : 1);
  if (out_elems == 0) return CA_OK;

  std::vector<Play> plays;
  std::vector<VoiceRec> voices(d.n_voices);
  for (int64_t v = 0; v < d.n_voices; ++v) {
    voices[v].play_begin = static_cast<int32_t>(plays.size());
    voice_plays(d, v, plays);
    voices[v].play_end = static_cast<int32_t>(plays.size());
    double gl = 1, gr = 1;
    if (stereo) {  // stereo(pan), include/stereo.h:18-27
      double pan = d.pan ? static_cast<double>(d.pan[v]) : 0.0;
      pan = std::min(1.0, std::max(-1.0, pan));
      gl = cos(0.25 * 3.14159265358979323846 * (pan + 1.0));
      gr = sin(0.25 * 3.14159265358979323846 * (pan + 1.0));
    }
    voices[v].gl = static_cast<float>(gl);
    voices[v].gr = static_cast<float>(gr);
    voices[v].out_base = static_cast<int64_t>(v * row);
  }
  const size_t rows_elems = row * static_cast<size_t>(d.n_voices);
  int rc = ensure_scratch(ctx, 6, std::max<size_t>(16, sizeof(Play) * plays.size()));
  if (!rc) rc = ensure_scratch(ctx, 1, std::max<size_t>(16, sizeof(VoiceRec) * voices.size()));
  if (!rc) rc = ensure_scratch(ctx, 7, sizeof(float) * out_elems);
  if (!rc && !per_voice) rc = ensure_scratch(ctx, 8, std::max<size_t>(16, sizeof(float) * rows_elems));
  if (rc) return rc;
  cudaStream_t st = ctx->own_stream;
  float* d_rows = static_cast<float*>(per_voice ? ctx->d_scratch[7] : ctx->d_scratch[8]);
  CA_CUDA_OK(ctx, cudaMemsetAsync(d_rows, 0, sizeof(float) * rows_elems, st));
  if (!plays.empty()) {
    CA_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_scratch[6], plays.data(), sizeof(Play) * plays.size(), cudaMemcpyHostToDevice, st));
    CA_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_scratch[1], voices.data(), sizeof(VoiceRec) * voices.size(), cudaMemcpyHostToDevice, st));
    EngineArgs a{};
    a.plays = static_cast<const Play*>(ctx->d_scratch[6]);
    a.voices = static_cast<const VoiceRec*>(ctx->d_scratch[1]);
    a.n_voices = d.n_voices;
    a.n_samples = d.n_samples;
    a.out = d_rows;
    a.stereo = stereo ? 1 : 0;
    a.sample_rate = d.sample_rate;
    a.loudness_on = d.loudness_on ? 1 : 0;
    a.log_ratio = d.loudness_compensation;
    {
      double pitch[kLoudN];
      // phons_to_index (include/loudness.h:220-229): contours exist for 20, 30, .. 100 phons
      int level = static_cast<int>(d.loudness_level * .1f) - 2;
      level = std::max(0, std::min(8, level));
      loudness_tables((level + 2) * 10.f, pitch, a.elv);
      static const float freqs[kLoudN] = {20.f,  25.f,  31.5f, 40.f,   50.f,   63.f,   80.f,   100.f,  125.f,  160.f,
                                          200.f, 250.f, 315.f, 400.f,  500.f,  630.f,  800.f,  1000.f, 1250.f, 1600.f,
                                          2000.f, 2500.f, 3150.f, 4000.f, 5000.f, 6300.f, 8000.f, 10000.f, 12500.f};
      for (int i = 0; i < kLoudN; ++i) a.freqs[i] = freqs[i];
      a.max_db = a.elv[d.loudness_ref_index];
    }
    sound_engine_kernel<<<static_cast<unsigned>((d.n_voices + 3) / 4), 128, 0, st>>>(a);
    CA_CUDA_OK(ctx, cudaGetLastError());
    ctx->launches += 1;
  }
  if (!per_voice) {
    mix_rows_kernel<<<static_cast<unsigned>((row + 255) / 256), 256, 0, st>>>(d_rows, d.n_voices, static_cast<int64_t>(row), static_cast<float*>(ctx->d_scratch[7]));
    CA_CUDA_OK(ctx, cudaGetLastError());
    ctx->launches += 1;
  }
  CA_CUDA_OK(ctx, cudaMemcpyAsync(h_out, ctx->d_scratch[7], sizeof(float) * out_elems, cudaMemcpyDeviceToHost, st));
  CA_CUDA_OK(ctx, cudaStreamSynchronize(st));
  return CA_OK;
}

}  // namespace ca
