// oracle/oracle_capi.cpp — C entry points over oracle.hpp so tests/bench can drive the CPU oracle through ctypes.
// TEST INFRASTRUCTURE ONLY (see header of oracle.hpp). PARITY: analysis arithmetic unpinned, synthesis pinned.
#include "oracle.hpp"

#include <cstring>
#include <thread>

using namespace orc;

namespace {

template <class F>
void parallel_for(int64_t n, int threads, F&& f) {
  threads = std::max(1, threads);
  if (threads == 1 || n <= 1) {
    for (int64_t i = 0; i < n; ++i) f(i);
    return;
  }
  std::vector<std::thread> pool;
  int64_t per = (n + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    int64_t lo = t * per, hi = std::min(n, lo + per);
    if (lo >= hi) break;
    pool.emplace_back([lo, hi, &f] { for (int64_t i = lo; i < hi; ++i) f(i); });
  }
  for (auto& th : pool) th.join();
}

template <class T>
void rfft_batched(const T* in, int n, int64_t batch, T* out, int threads) {
  int64_t bins = n / 2 + 1;
  int nt = std::max(1, threads);
  int64_t per = (batch + nt - 1) / nt;
  parallel_for(nt, nt, [&](int64_t t) {
    RealFft<T> fft(n);
    std::vector<std::complex<T>> X(bins);
    for (int64_t b = t * per; b < std::min(batch, (t + 1) * per); ++b) {
      fft.forward(in + b * n, X.data());
      for (int64_t k = 0; k < bins; ++k) {
        out[(b * bins + k) * 2 + 0] = X[k].real();
        out[(b * bins + k) * 2 + 1] = X[k].imag();
      }
    }
  });
}

}  // namespace

extern "C" {

int orc_hardware_threads() { return static_cast<int>(std::thread::hardware_concurrency()); }

// forward real FFT, unscaled, out = interleaved (re,im) [batch][n/2+1]
void orc_rfft_f64(const double* in, int n, int64_t batch, double* out, int threads) { rfft_batched<double>(in, n, batch, out, threads); }
void orc_rfft_f32(const float* in, int n, int64_t batch, float* out, int threads) { rfft_batched<float>(in, n, batch, out, threads); }

int64_t orc_num_frames(int64_t n_samples, int window, int hop) { return num_frames(n_samples, window, hop); }
int orc_fft_length_for(int n, int zpf) { return fft_length_for(n, zpf); }

void orc_window_f64(int window_type, double sigma_factor, int window, double* out) {
  auto hw = make_half_window<double>(static_cast<Window>(window_type), sigma_factor, window / 2);
  for (int n = 0; n < window; ++n) out[n] = window_at(hw, n);
}

// Streaming analysis of n_streams independent streams, exactly as the reference drives it: one sample at a time through
// PeriodicFFT::feed (source/rt.resynth.lib.cpp:1624-1668), SqMagFftOperation<Gaussian,double> per full window,
// extractLocalMaxFreqsMags per frame (source/rt.resynth.lib.cpp:1592-1595). Everything in double.
// Outputs (all nullable): sqmag [streams][frames][n_fft/2+1]; peaks [streams][frames][max_peaks]; counts [streams][frames].
// peak_count holds the number of peaks FOUND (may exceed max_peaks; only the first max_peaks are stored).
// Returns frames per stream.
int64_t orc_stft_analyze(const float* x, int64_t n_samples, int64_t n_streams, int64_t stream_stride, int window, int hop,
                         int window_type, double sigma_factor, int sample_rate, double min_db, int max_peaks,
                         double* sqmag, int32_t* peak_bin, double* peak_freq, double* peak_db, int32_t* peak_count,
                         int threads) {
  int64_t frames = num_frames(n_samples, window, hop);
  int n_fft = fft_length_for(window, 1);
  int64_t bins = n_fft / 2 + 1;
  parallel_for(n_streams, threads, [&](int64_t s) {
    PeriodicFft<SqMagOp<double>> pfft(static_cast<Window>(window_type), sigma_factor);
    std::vector<FreqMag<double>> freqmags;
    freqmags.reserve(200);
    int64_t frame = 0;
    pfft.set_lambdas([&] { return window; }, [&] { return hop; },
                     [&](int, const std::vector<double>& sq) {
                       int64_t row = s * frames + frame;
                       if (sqmag) std::memcpy(sqmag + row * bins, sq.data(), sizeof(double) * bins);
                       if (peak_count || peak_bin || peak_freq || peak_db) {
                         extract_local_max_freqs_mags<double>(sample_rate, sq.data(), n_fft, min_db, freqmags);
                         if (peak_count) peak_count[row] = static_cast<int32_t>(freqmags.size());
                         int n = std::min<int>(max_peaks, static_cast<int>(freqmags.size()));
                         for (int i = 0; i < n; ++i) {
                           if (peak_bin) peak_bin[row * max_peaks + i] = freqmags[i].bin;
                           if (peak_freq) peak_freq[row * max_peaks + i] = freqmags[i].freq;
                           if (peak_db) peak_db[row * max_peaks + i] = freqmags[i].mag_db;
                         }
                       }
                       ++frame;
                     });
    const float* xs = x + s * stream_stride;
    for (int64_t i = 0; i < n_samples; ++i) pfft.feed(xs[i]);
  });
  return frames;
}

// Harmonic-bank voices (VolumeAdjusted<MultiEnveloped<SineOscillatorAlgo,AHDSR>>), each rendered by the reference's block
// structure: blocks of <=16 frames (include/simple_audio_context.h:80-89), step() then imag() (include/gen.crtp.h:358-370).
//   f0_hz[v]          note frequency (float, as carried by Event::noteOn, include/events.h)
//   volume[v]         volume target = baseVolume*velocity (include/gen.crtp.h:425)
//   start_angle[v]    voice start angle a0 (units of pi rad) handed to MultiEnveloped::setStartAngle
//   key_delay[v]      onKeyPressed(delay)
//   release_at[v]     sample index before whose step onKeyReleased(0) is called; <0 = never released
//   partial_volume/phase [v][p]
//   env               n_env == 1 (shared) or n_voices
//   pan               nullable; used only when mix != 0
// out: mix==0 -> [voices][samples]; mix==1 -> interleaved stereo [samples][2]; mix==2 -> mono [samples]
void orc_oscbank_render(int64_t n_voices, int n_partials, int64_t n_samples, int sample_rate, const float* f0_hz,
                        const float* volume, const float* start_angle, const int32_t* key_delay, const int64_t* release_at,
                        const float* partial_volume, const float* partial_phase, const Ahdsr* env, int64_t n_env,
                        const float* pan, int mix, double* out, int threads) {
  auto make_voice = [&](int64_t v, BankVoice<double>& voice) {
    std::vector<HarmonicProps> props(n_partials);
    for (int p = 0; p < n_partials; ++p) props[p] = {partial_phase[v * n_partials + p], partial_volume[v * n_partials + p]};
    voice.set_harmonics(props.data(), n_partials);
    voice.try_acquire();
    voice.forget_past_signals();
    voice.set_volume_target(volume[v]);
    voice.set_ahdsr(env[n_env == 1 ? 0 : v], sample_rate);
    voice.set_angle_increments(static_cast<double>(freq_to_angle_increment<float>(f0_hz[v], sample_rate)));
    voice.set_start_angle(start_angle ? start_angle[v] : 0.f);
    voice.on_key_pressed(key_delay ? key_delay[v] : 0);
  };
  if (mix == 0) {
    parallel_for(n_voices, threads, [&](int64_t v) {
      BankVoice<double> voice;
      make_voice(v, voice);
      double* o = out + v * n_samples;
      int64_t rel = release_at ? release_at[v] : -1;
      for (int64_t i0 = 0; i0 < n_samples; i0 += kFramesPerBuffer) {
        int n = static_cast<int>(std::min<int64_t>(kFramesPerBuffer, n_samples - i0));
        for (int i = 0; i < n; ++i) {
          if (i0 + i == rel) voice.on_key_released(0);
          o[i0 + i] = 0;
          if (!voice.rt_active()) continue;
          voice.step();
          if (!voice.rt_active()) continue;
          o[i0 + i] = voice.imag();
        }
      }
    });
    return;
  }
  // mixed bus: voices are partitioned over threads, each thread renders its partition voice-major in 16-frame blocks
  // into a private bus, then buses are summed in thread order.
  int n_out = mix == 1 ? 2 : 1;
  int nt = std::max(1, threads);
  std::vector<std::vector<double>> buses(nt, std::vector<double>(static_cast<size_t>(n_samples) * n_out, 0.0));
  int64_t per = (n_voices + nt - 1) / nt;
  parallel_for(nt, nt, [&](int64_t t) {
    int64_t lo = t * per, hi = std::min(n_voices, lo + per);
    if (lo >= hi) return;
    std::vector<BankVoice<double>> voices(hi - lo);
    std::vector<double> gl(hi - lo, 1.0), gr(hi - lo, 1.0);
    for (int64_t v = lo; v < hi; ++v) {
      make_voice(v, voices[v - lo]);
      if (mix == 1 && pan) stereo_gains(pan[v], gl[v - lo], gr[v - lo]);
      else if (mix == 1) stereo_gains(0.0, gl[v - lo], gr[v - lo]);
    }
    double* bus = buses[t].data();
    for (int64_t i0 = 0; i0 < n_samples; i0 += kFramesPerBuffer) {
      int n = static_cast<int>(std::min<int64_t>(kFramesPerBuffer, n_samples - i0));
      for (int64_t v = lo; v < hi; ++v) {
        auto& voice = voices[v - lo];
        int64_t rel = release_at ? release_at[v] : -1;
        for (int i = 0; i < n; ++i) {
          if (i0 + i == rel) voice.on_key_released(0);
          if (!voice.rt_active()) break;
          voice.step();
          if (!voice.rt_active()) break;
          double y = voice.imag();
          if (mix == 1) {
            bus[2 * (i0 + i) + 0] += gl[v - lo] * y;
            bus[2 * (i0 + i) + 1] += gr[v - lo] * y;
          } else {
            bus[i0 + i] += y;
          }
        }
      }
    }
  });
  size_t total = static_cast<size_t>(n_samples) * n_out;
  for (size_t i = 0; i < total; ++i) {
    double s = 0;
    for (int t = 0; t < nt; ++t) s += buses[t][i];
    out[i] = s;
  }
}

// gen.voice ramp voices (LogRamp-controlled sine with AHDSR envelope and constant volume): out [voices][samples] (mix 0)
// or their sum [samples] (mix 2). env_ai[v]: the angle increment that sets the envelope's minimum change durations.
void orc_ramp_render(int64_t n_voices, int64_t n_samples, int sample_rate, const float* from_hz, const float* to_hz,
                     const float* duration, const float* start_sample, const float* volume, const float* start_angle,
                     const int32_t* key_delay, const int64_t* release_at, const Ahdsr* env, int mix, double* out, int threads) {
  std::vector<double> tmp;
  if (mix != 0) {
    tmp.assign(static_cast<size_t>(n_voices) * n_samples, 0.0);
    for (int64_t i = 0; i < n_samples; ++i) out[i] = 0;
  }
  double* dst = mix == 0 ? out : tmp.data();
  parallel_for(n_voices, threads, [&](int64_t v) {
    RampVoice<double> voice;
    const double from = static_cast<double>(freq_to_angle_increment<float>(from_hz[v], sample_rate));
    const double to = static_cast<double>(freq_to_angle_increment<float>(to_hz[v], sample_rate));
    voice.try_acquire();
    voice.initialize(*env, sample_rate, volume[v]);
    voice.set_angle_increments(from);
    voice.forget_past_signals();
    voice.setup_ramp(from, to, duration[v], start_sample ? start_sample[v] : 0.f);
    voice.set_start_angle(start_angle ? start_angle[v] : 0.f);
    voice.on_key_pressed(key_delay ? key_delay[v] : 0);
    double* o = dst + v * n_samples;
    const int64_t rel = release_at ? release_at[v] : -1;
    for (int64_t i = 0; i < n_samples; ++i) {
      if (i == rel) voice.on_key_released(0);
      o[i] = 0;
      if (!voice.rt_active()) continue;
      voice.step();
      if (!voice.rt_active()) continue;
      o[i] = voice.imag();
    }
  });
  if (mix != 0)
    for (int64_t v = 0; v < n_voices; ++v)
      for (int64_t i = 0; i < n_samples; ++i) out[i] += tmp[v * n_samples + i];
}

// Envelope alone (known-answer tests): value after each of n_steps steps, note pressed with `delay`,
// released (delay 0) just before step index release_at (<0: never), for an oscillator of angle increment ai.
void orc_envelope_trace(const Ahdsr* env, int sample_rate, double angle_increment, int32_t key_delay, int64_t release_at,
                        int64_t n_steps, double* values, int32_t* states) {
  EnvelopedOsc<double> e;
  e.env.set_ahdsr(*env, sample_rate);
  e.env.try_acquire();
  e.forget_past_signals();
  e.set_angle_increments(angle_increment);
  e.forget_past_signals();  // ImplCRTP::onEvent calls forgetPastSignals before setupAudioElement (gen.crtp.h:423,436); the
                            // min change duration is latched at the next safe point (startPressed), so order is irrelevant
  e.env.on_key_pressed(key_delay);
  for (int64_t i = 0; i < n_steps; ++i) {
    if (i == release_at) e.env.on_key_released(0);
    e.step();
    if (values) values[i] = e.env.value();
    if (states) states[i] = static_cast<int32_t>(e.env.state());
  }
}

double orc_freq_aliasing_multiplicator(double ai) { return freq_aliasing_multiplicator<double>(ai); }

// Single oscillator trace: y[n] = imag after step n+1 (step precedes read, include/gen.crtp.h:359-368)
void orc_sine_trace(double start_angle, double angle_increment, int64_t n, double* out) {
  SineOsc<double> o;
  o.set_angle_increments(angle_increment);
  o.set_start_angle(start_angle);
  for (int64_t i = 0; i < n; ++i) {
    o.step();
    out[i] = o.imag();
  }
}

// Analysis -> pitch pipeline -> resynthesis of n_streams independent streams (BASELINE config 4), each through the offline
// loop of the reference. out: [streams][samples][2] double. stats (nullable): [streams][2] = notes started, notes dropped.
// the optional stages alone (shift, harmonize, autotune, shift, harmonize; rt.resynth.lib.cpp:1687-1707) on one list;
// pitch / volume hold n entries and have room for `cap`; returns the resulting count
int orc_pitch_stages(const PitchOptions* o, double* pitch, double* volume, int n, int cap) {
  std::vector<PitchVolume> reduced(n), tuned, tmp;
  for (int i = 0; i < n; ++i) reduced[i] = {pitch[i], volume[i]};
  shift_pitches(o->shift_pre_autotune, reduced);
  harmonize_pitches(o->harmonize_pre_autotune, tmp, reduced);
  autotune_pitches(*o, reduced, tuned);
  shift_pitches(o->shift_post_autotune, tuned);
  harmonize_pitches(o->harmonize_post_autotune, tmp, tuned);
  const int m = static_cast<int>(std::min<size_t>(tuned.size(), cap));
  for (int i = 0; i < m; ++i) {
    pitch[i] = tuned[i].midipitch;
    volume[i] = tuned[i].volume;
  }
  return static_cast<int>(tuned.size());
}

// tracking[3] (nullable): nearby_distance_tones, min_volume, max_track_pitches; max_channels <= 0: the reference's 254
void orc_roundtrip_full(const float* x, int64_t n_samples, int64_t n_streams, int sample_rate, int window, int hop, double min_db,
                        float gain, float stereo_spread, const Ahdsr* env, const PitchOptions* opts, const double* tracking,
                        int max_channels, double* out, int64_t* stats, int threads) {
  parallel_for(n_streams, threads, [&](int64_t s) {
    RoundTripParams p;
    p.sample_rate = sample_rate;
    p.window = window;
    p.hop = hop;
    p.min_db = min_db;
    p.gain = gain;
    p.stereo_spread = stereo_spread;
    if (env) p.env = *env;
    if (opts) p.opts = *opts;
    if (tracking) {
      p.nearby_distance_tones = tracking[0];
      p.min_volume = tracking[1];
      p.max_track_pitches = tracking[2];
    }
    if (max_channels > 0) p.max_channels = max_channels;
    RoundTrip<double> rt(p);
    rt.run(x + s * n_samples, n_samples, out + s * n_samples * 2);
    if (stats) {
      stats[2 * s] = rt.notes_started();
      stats[2 * s + 1] = rt.dropped_notes();
    }
  });
}

// A note schedule rendered by the rt.resynth voice pool (VoicePool, oracle.hpp). out: [n_samples][2], or [n_samples] when mono.
void orc_render_tracks(const TrackEvent* events, int64_t n_events, int64_t n_samples, int sample_rate, int glide_samples,
                       const Ahdsr* env, int max_channels, float stereo_spread, int mono, double* out, int64_t* stats) {
  render_tracks<double>(events, n_events, n_samples, sample_rate, glide_samples, *env, max_channels, stereo_spread, mono != 0, out, stats);
}

// One gen.voice voice (VolumeAdjusted<SoundEngine>, include/gen.voice.h:1093-1110) playing `specs` from sample key_on on, released at
// key_off (< 0: never), ACCUMULATED into out ([n_samples][2] when stereo, else [n_samples]). The voice-level volume is the velocity
// (base volume 1: include/audioelement.h:1278, a constant target, so its filter is the identity).
void orc_sound_engine_render(const RampSpec* specs, int n_specs, const SoundEngineParams* prm, int64_t key_on, int64_t key_off,
                             float velocity, int stereo, float pan, int64_t n_samples, double* out) {
  SoundEngineVoice<double> v(*prm, specs, n_specs, stereo != 0, pan);
  bool on = false;
  for (int64_t i = std::max<int64_t>(key_on, 0); i < n_samples; ++i) {
    if (i == std::max<int64_t>(key_on, 0)) {
      v.on_key_pressed();
      on = true;
    }
    if (on && key_off >= 0 && i == key_off) v.on_key_released(0);
    v.step();
    if (stereo) {
      out[2 * i] += velocity * v.imag(0);
      out[2 * i + 1] += velocity * v.imag(1);
    } else {
      out[i] += velocity * v.imag(0);
    }
  }
}

float orc_equal_loudness_volume_from_freq(float freq, int ref_index, float log_ratio, float phons) {
  return EqualLoudness(phons).volume_from_freq(freq, ref_index, log_ratio);
}

// equal_loudness_volume_db(pitches, pitch, phons_to_index(phons)) of include/loudness.h
float orc_equal_loudness_db(double pitch, float phons) { return EqualLoudness(phons).volume_db(pitch); }

// test hooks of the start-phase rule (oracle.hpp sync_phase / sync_count)
void orc_set_sync_phase(int on) { sync_phase() = on != 0; }
int64_t orc_sync_count(int reset) {
  const int64_t v = sync_count();
  if (reset) sync_count() = 0;
  return v;
}

void orc_roundtrip_opts(const float* x, int64_t n_samples, int64_t n_streams, int sample_rate, int window, int hop, double min_db,
                        float gain, float stereo_spread, const Ahdsr* env, const PitchOptions* opts, double* out, int64_t* stats,
                        int threads) {
  orc_roundtrip_full(x, n_samples, n_streams, sample_rate, window, hop, min_db, gain, stereo_spread, env, opts, nullptr, 0, out, stats, threads);
}

void orc_roundtrip(const float* x, int64_t n_samples, int64_t n_streams, int sample_rate, int window, int hop, double min_db,
                   float gain, float stereo_spread, const Ahdsr* env, double* out, int64_t* stats, int threads) {
  orc_roundtrip_opts(x, n_samples, n_streams, sample_rate, window, hop, min_db, gain, stereo_spread, env, nullptr, out, stats, threads);
}

// Vocoder (source/rt.resynth.lib.vocoder.cpp) over n_streams independent (modulator, carrier) pairs, one sample at a time.
// out [streams][samples] double. Returns the stride in samples.
int orc_vocoder(const float* modulator, const float* carrier, int64_t n_samples, int64_t n_streams, int sample_rate,
                float modulator_window_seconds, float stride_seconds, int count_bands, float min_freq, float max_freq,
                float vol_modulator, float vol_carrier, float vol_vocoder, double* out, int threads) {
  VocoderSetup su;
  su.modulator_window_size_seconds = modulator_window_seconds;
  su.stride_seconds = stride_seconds;
  su.count_bands = count_bands;
  su.min_freq = min_freq;
  su.max_freq = max_freq;
  parallel_for(n_streams, threads, [&](int64_t s) {
    Vocoder<double> v(sample_rate, su, vol_modulator, vol_carrier, vol_vocoder);
    for (int64_t n = 0; n < n_samples; ++n) out[s * n_samples + n] = v.feed(modulator[s * n_samples + n], carrier[s * n_samples + n]);
  });
  return vocoder_good_stride(stride_seconds, sample_rate);
}

}  // extern "C"
