// pitch.cuh — the pitch pipeline of RtResynth::step (source/rt.resynth.lib.cpp:1670-1759) and the note bookkeeping of
// synthesize_sounds (:264-382) as functions shared by the device kernels (pitch.cu) and the host A/B path (roundtrip.cu).
//
//   peaks of one frame --pitch_prefix()--> "tuned" list (stateless: source/rt.resynth.lib.algo.cpp:10-230)
//   tuned list + played notes --tracking--> NoteOn / NoteChange / NoteOff  (algo.cpp:232-440, frame-sequential)
//   sounding notes --emit_segment()--> one closed-form Segment per note and render tile (tracks.cuh)
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#include <climits>

#include "tracks.cuh"

#ifdef __CUDACC__
#define CA_HD __host__ __device__
#else
#define CA_HD
#endif

namespace ca {

constexpr int kMaxNotes = 256;   // voice pool bound of the device pipeline (reference: 254 channels, rt.resynth.lib.cpp:201-210)
constexpr int kMaxTuned = 256;   // peak capacity per frame of the device pipeline (reference reserve: 200, :1102)

// Equal-loudness contours: the ISO 226:2003 tables as include/loudness.h:9-149 holds them, and the contour arithmetic of
// compute_equal_loudness_volume (:188-197, float). order_pitches_by_perceived_loudness (source/rt.resynth.lib.algo.cpp:377-411)
// ranks the new notes by volume / equal_loudness_volume_db(pitches, pitch, phons_to_index(60)) (rt.resynth.lib.cpp:1716-1724).
constexpr int kLoudN = 29;
inline void loudness_tables(float phons, double (&pitch)[kLoudN], float (&elv)[kLoudN]) {
  static const float freqs[kLoudN] = {20.f,  25.f,  31.5f, 40.f,   50.f,   63.f,   80.f,   100.f,  125.f,  160.f,
                                      200.f, 250.f, 315.f, 400.f,  500.f,  630.f,  800.f,  1000.f, 1250.f, 1600.f,
                                      2000.f, 2500.f, 3150.f, 4000.f, 5000.f, 6300.f, 8000.f, 10000.f, 12500.f};
  static const float alpha_f[kLoudN] = {.532f, .506f, .480f, .455f, .432f, .409f, .387f, .367f, .349f, .330f, .315f, .301f, .288f, .276f, .267f,
                                        .259f, .253f, .250f, .246f, .244f, .243f, .243f, .243f, .242f, .242f, .245f, .254f, .271f, .301f};
  static const float Lu[kLoudN] = {-31.6f, -27.2f, -23.0f, -19.1f, -15.9f, -13.0f, -10.3f, -8.1f, -6.2f, -4.5f, -3.1f, -2.0f, -1.1f, -0.4f, 0.f,
                                   .3f,    .5f,    0.f,    -2.7f,  -4.1f,  -1.0f,  1.7f,   2.5f,  1.2f,  -2.1f, -7.1f, -11.2f, -10.7f, -3.1f};
  static const float Tf[kLoudN] = {78.5f, 68.7f, 59.5f, 51.1f, 44.0f, 37.5f, 31.5f, 26.5f, 22.1f, 17.9f, 14.4f, 11.4f, 8.6f, 6.2f, 4.4f,
                                   3.0f,  2.2f,  2.4f,  3.5f,  1.7f,  -1.3f, -4.2f, -6.0f, -5.4f, -1.5f, 6.0f,  12.6f, 13.9f, 12.3f};
  for (int i = 0; i < kLoudN; ++i) {
    pitch[i] = 69. + 12. * log2(static_cast<double>(freqs[i]) / 440.f);  // include/loudness.h:43-54, include/midi.h:224-232
    const float Af = 4.47e-3f * (powf(10.f, .025f * phons) - 1.14f) + powf(.4f * powf(10.f, ((Tf[i] + Lu[i]) * .1f) - 9.f), alpha_f[i]);
    elv[i] = 94.f - Lu[i] + (10.f / alpha_f[i]) * logf(Af) / logf(10.f);
  }
}

// Everything the pipeline needs, resolved once on the host.
struct PitchParams {
  int32_t W, hop, sr, P;       // P: entries per frame of the peak arrays (capacity of a frame's raw peak list)
  int32_t PT;                  // entries per frame of the tuned arrays: min(stft.max_peaks, kMaxTuned)
  int32_t cap_raw;             // peaks the analysis keeps per frame (= P, except in the host form, which compacts the lists)
  double bin_hz;
  float nearby, min_volume, max_track, gain, stereo_spread;
  int32_t A, H, D, R;          // AHDSREnvelopeBase::setAHDSR (include/audioelement.h:923-952)
  float sustain;
  double cap;                  // setMaxFilterIncrement(2/stride), source/rt.resynth.lib.cpp:106
  double semitone;             // pow(2., 1./12.) as the host's libm evaluates it
  double decay_full;           // exp(-pi * cap * hop): volume low-pass decay over a whole frame interval at the capped rate
  int32_t max_channels;
  int64_t n_samples, frames;
  // optional stages between reduce_pitches and track_pitches (source/rt.resynth.lib.cpp:1687-1707); opt_active == 0: none
  int32_t opt_active;
  int32_t tcap;                // capacity of a frame's list at every stage = stft.max_peaks (independent of the stride P)
  float shift_pre, harm_pre, shift_post, harm_post;
  int32_t autotune;            // ca_autotune
  float at_max_pitch, at_tolerance;
  int32_t n_table;
  const double* table;         // allowed pitches / scale pitches; a HOST pointer in the host form, a DEVICE pointer in the kernels
  double scale_root;
  double loud_pitch[kLoudN];   // the 60-phon equal-loudness contour in dB over midi pitch (loudness_tables)
  float loud_elv[kLoudN];
};

inline PitchParams make_pitch_params(const ca_roundtrip_params& prm, int64_t n_samples, int64_t frames, int peak_stride) {
  PitchParams p{};
  p.W = prm.stft.window;
  p.hop = prm.stft.hop;
  p.sr = prm.stft.sample_rate;
  p.P = peak_stride;
  p.cap_raw = peak_stride;
  p.PT = prm.stft.max_peaks < kMaxTuned ? prm.stft.max_peaks : kMaxTuned;
  if (p.PT < 1) p.PT = 1;
  p.bin_hz = static_cast<double>(p.sr) / ca_fft_length_for(p.W, 1);
  p.nearby = prm.nearby_distance_tones;
  p.min_volume = prm.min_volume;
  p.max_track = prm.max_track_pitches;
  p.gain = prm.gain;
  p.stereo_spread = prm.stereo_spread;
  const int32_t min_dt = p.sr / 1000;
  const bool has_decay = prm.env.sustain < 0.999999;
  p.sustain = has_decay ? fminf(fmaxf(prm.env.sustain, 0.f), 1.f) : 1.f;
  p.A = prm.env.attack > min_dt ? prm.env.attack : min_dt;
  p.H = prm.env.hold > 0 ? prm.env.hold : 0;
  p.D = has_decay ? (prm.env.decay > min_dt ? prm.env.decay : min_dt) : 0;
  p.R = prm.env.release > min_dt ? prm.env.release : min_dt;
  p.cap = 2.0 / p.hop;
  p.semitone = pow(2., 1. / 12.);
  p.decay_full = exp(-3.14159265358979323846 * (p.cap * p.hop));
  p.max_channels = prm.max_channels;
  p.n_samples = n_samples;
  p.frames = frames;
  const ca_pitch_options& o = prm.opts;
  p.opt_active = (o.shift_pre_autotune != 0.f || o.harmonize_pre_autotune != 0.f || o.shift_post_autotune != 0.f ||
                  o.harmonize_post_autotune != 0.f || o.autotune != CA_AUTOTUNE_NONE)
                     ? 1
                     : 0;
  p.tcap = p.PT;
  p.shift_pre = o.shift_pre_autotune;
  p.harm_pre = o.harmonize_pre_autotune;
  p.shift_post = o.shift_post_autotune;
  p.harm_post = o.harmonize_post_autotune;
  p.autotune = o.autotune;
  p.at_max_pitch = o.autotune_max_pitch;
  p.at_tolerance = o.autotune_tolerance_pitches;
  p.n_table = o.n_table;
  p.table = o.table;
  p.scale_root = o.scale_root_pitch;
  loudness_tables(60.f, p.loud_pitch, p.loud_elv);  // phons_to_index(60.f), rt.resynth.lib.cpp:1717
  return p;
}

// render tiles of one stream: the interval up to the next frame is one tile, or several when the hop exceeds kMaxTile
inline int64_t tiles_per_stream(const PitchParams& p) {
  int64_t t = 0;
  for (int64_t k = 0; k < p.frames; ++k) {
    const int64_t n_k = static_cast<int64_t>(p.W) - 1 + k * p.hop;
    const int64_t len = p.hop < p.n_samples - n_k ? p.hop : p.n_samples - n_k;
    t += len <= 0 ? 1 : (len + kMaxTile - 1) / kMaxTile;
  }
  return t;
}

CA_HD inline float hash_unit(uint64_t id, uint64_t salt) {  // SPEC S10
  uint64_t z = id * 0x9E3779B97F4A7C15ull + salt * 0xBF58476D1CE4E5B9ull + 0x94D049BB133111EBull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return static_cast<float>(static_cast<double>(z >> 11) * (2.0 / 9007199254740992.0) - 1.0);
}

// equal_loudness_volume_db (include/loudness.h:232-244) over closest() (:155-186)
CA_HD inline float equal_loudness_db(const PitchParams& p, double pitch) {
  int lo = 0, hi = kLoudN - 1;
  do {
    const int c = (lo + hi) / 2;
    if (pitch < p.loud_pitch[c]) hi = c;
    else lo = c;
  } while (hi - lo >= 2);
  if (pitch <= p.loud_pitch[lo]) return p.loud_elv[lo];
  if (pitch >= p.loud_pitch[hi]) return p.loud_elv[hi];
  const float ratio = static_cast<float>((pitch - p.loud_pitch[lo]) / (p.loud_pitch[hi] - p.loud_pitch[lo]));
  return ratio * p.loud_elv[hi] + (1.f - ratio) * p.loud_elv[lo];
}

CA_HD inline int32_t min_change_for(double ai) {  // Enveloped::setAngleIncrements, include/audioelement.h:218-225
  const double period = ai == 0.0 ? 1.0 : 2.0 / fabs(ai);
  return static_cast<int32_t>(0.5f + 2.5 * period);
}

// One entry of the list the stateless prefix of the pipeline produces for a frame, with the event parameters that follow
// from it (synthesize_sounds, source/rt.resynth.lib.cpp:264-382) already evaluated.
struct __align__(16) Tuned {
  double pitch;        // midi pitch (tracking key)
  double vol;          // summed volume of the aggregated peaks; on leaving pitch_prefix: their perceived loudness (the ranking key)
  double ai;           // angle increment of the note frequency (freq_to_angle_increment evaluated in float)
  float volume;        // event volume: gain * vol / 0.1
  int32_t min_change;  // min_change_for(ai)
};

// ---- optional stages of the stateless prefix: shift / harmonize / autotune (source/rt.resynth.lib.algo.cpp:186-229, 307-371) ----
// They work on (pitch, volume) pairs kept in the raw bytes of the frame's Tuned row: tcap entries of 32 bytes hold the list
// (pairs [0, tcap)) and the work vector of the merge (pairs [tcap, 2 tcap)). Lists are capped at tcap entries at every stage
// (the reference's vectors grow; its peak list is capped the same way at max_peaks).
struct PV {
  double pitch, vol;
};
CA_HD inline PV pv_ld(const unsigned char* raw, int i) {
  PV q;
  memcpy(&q, raw + 16 * static_cast<size_t>(i), 16);
  return q;
}
CA_HD inline void pv_st(unsigned char* raw, int i, PV q) { memcpy(raw + 16 * static_cast<size_t>(i), &q, 16); }

// std::lower_bound's bisection, written out: harmonize_pitches runs it on a vector whose tail is not ordered, so the
// probe sequence is part of the behaviour (find_closest_pitch, source/rt.resynth.lib.autotune.cpp:189-218)
template <class Get>
CA_HD inline int closest_index(double pitch, int n, Get&& get) {
  if (n == 0) return -1;
  int first = 0, len = n;
  while (len > 0) {
    const int half = len >> 1;
    if (get(first + half) < pitch) {
      first += half + 1;
      len -= half + 1;
    } else {
      len = half;
    }
  }
  if (first == n) return n - 1;
  if (first == 0) return 0;
  const float d_lb = static_cast<float>(fabs(pitch - get(first)));
  const float d_prev = static_cast<float>(fabs(pitch - get(first - 1)));
  return d_lb < d_prev ? first : first - 1;
}

CA_HD inline int harmonize_pairs(float amount, unsigned char* raw, int n, int cap) {  // algo.cpp:318-371
  if (!amount) return n;
  const int sz = n;
  for (int i = 0; i < sz; ++i) {
    const PV p = pv_ld(raw, i);
    const double harmonized = p.pitch + amount;
    const int c = closest_index(harmonized, n, [&](int j) { return pv_ld(raw, j).pitch; });
    PV q = pv_ld(raw, c);
    if (static_cast<float>(fabs(q.pitch - harmonized)) <= 0.0001f) {
      q.vol += p.vol;
      pv_st(raw, c, q);
    } else if (n < cap) {
      pv_st(raw, n++, PV{harmonized, p.vol});
    }
  }
  // merge of the runs [0, sz) and [sz, n) by pitch into the work half, ties keep the first run first; then copy back
  int a = 0, b = sz, w = cap;
  while (a < sz && b < n) {
    const PV x = pv_ld(raw, a), y = pv_ld(raw, b);
    if (y.pitch < x.pitch) {
      pv_st(raw, w++, y);
      ++b;
    } else {
      pv_st(raw, w++, x);
      ++a;
    }
  }
  while (a < sz) pv_st(raw, w++, pv_ld(raw, a++));
  while (b < n) pv_st(raw, w++, pv_ld(raw, b++));
  for (int i = 0; i < n; ++i) pv_st(raw, i, pv_ld(raw, cap + i));
  return n;
}

// MusicalScalePitches::closest_pitch, source/rt.resynth.lib.autotune.cpp:96-146
CA_HD inline double scale_closest(const double* scale, int n_scale, double root, double pitch) {
  const double octave_dist = (pitch - root) / 12;
  const int octaves = octave_dist >= 0. ? static_cast<int>(octave_dist) : static_cast<int>(octave_dist) - 1;
  const double rel = (pitch - octaves * 12) - root;
  int i = 0;
  for (; i < n_scale - 1; ++i)
    if (rel < 0.5 * (scale[i] + scale[i + 1])) break;
  return pitch - (rel - scale[i]);
}

CA_HD inline int autotune_pairs(const PitchParams& p, unsigned char* raw, int n) {  // algo.cpp:186-229; in place, output index <= input index
  int o = 0;
  for (int i = 0; i < n; ++i) {
    const PV pv = pv_ld(raw, i);
    double t = pv.pitch;
    if (p.autotune != CA_AUTOTUNE_NONE) {
      bool has = true;
      if (pv.pitch <= p.at_max_pitch) {
        if (p.autotune == CA_AUTOTUNE_TABLE) {
          const int c = closest_index(pv.pitch, p.n_table, [&](int j) { return p.table[j]; });
          has = c >= 0;
          if (has) t = p.table[c];
        } else {
          t = scale_closest(p.table, p.n_table, p.scale_root, pv.pitch);
        }
        if (has && t > p.at_max_pitch) t = pv.pitch;  // keeps the output sorted
      }
      if (!has || !(fabs(t - pv.pitch) <= p.at_tolerance)) continue;
    }
    if (o > 0) {
      PV last = pv_ld(raw, o - 1);
      if (fabs(last.pitch - t) < 0.0001f) {
        last.vol += pv.vol;
        pv_st(raw, o - 1, last);
        continue;
      }
    }
    pv_st(raw, o++, PV{t, pv.vol});
  }
  return o;
}

static_assert(sizeof(Tuned) == 32 && sizeof(PV) == 16, "a Tuned row holds two (pitch, volume) pairs per entry");

// Peaks of one frame (ascending frequency) -> tuned list. Returns the number of entries written to `out` (<= cnt, or
// <= tcap with the optional stages; `out` then has room for tcap entries).
//   freq -> midi pitch 69 + 12 log2(f/440) (include/midi.h:226-231), dB -> volume 10^(dB/20);
//   aggregation of peaks closer than nearby_distance_tones (interval diameter), volume-weighted pitch, summed volume,
//   min_volume cut (source/rt.resynth.lib.algo.cpp:10-230); entries that land on the same pitch are merged
//   (autotune_pitches with the identity transform); OPTS: shift, harmonize, autotune, shift, harmonize in between.
// R: float (peak lists of the float32 analysis kernels) or double (stft64.cu).
template <bool OPTS, class R>
CA_HD inline int pitch_prefix_t(const PitchParams& p, const int32_t* peak_bin, const R* peak_off, const R* peak_db, int cnt,
                                Tuned* out, int* cut) {
  int T = 0;
  const int room = p.PT;  // entries of `out`; what does not fit is dropped and counted in *cut
  bool open = false;
  double mn = 0, mx = 0, sv = 0, spv = 0;
  unsigned char* raw = reinterpret_cast<unsigned char*>(out);
  auto close = [&] {
    if (open && !(sv < p.min_volume)) {
      const double pitch = spv / sv;
      if (OPTS) {
        if (T < room) pv_st(raw, T++, PV{pitch, sv});
        else ++*cut;
      } else if (T > 0 && fabs(out[T - 1].pitch - pitch) < 0.0001f) {
        out[T - 1].vol += sv;
      } else if (T < room) {
        out[T].pitch = pitch;
        out[T].vol = sv;
        ++T;
      } else {
        ++*cut;
      }
    }
    open = false;
  };
  for (int i = 0; i < cnt; ++i) {
    const double freq = (static_cast<double>(peak_bin[i]) + static_cast<double>(peak_off[i])) * p.bin_hz;
    if (!(freq > 0)) continue;
    const double pitch = 69. + 12. * log2(freq / 440.f);
    const double vol = pow(10.0, static_cast<double>(peak_db[i]) / 20.0);
    if (open && fmax(fmax(mn, mx), pitch) - fmin(fmin(mn, mx), pitch) > p.nearby) close();
    if (!open) {
      open = true;
      mn = mx = pitch;
      sv = spv = 0;
    } else {
      mn = fmin(mn, pitch);
      mx = fmax(mx, pitch);
    }
    sv += vol;
    spv += pitch * vol;
  }
  close();
  if (OPTS) {
    const int cap = p.tcap;
    if (p.shift_pre != 0.f)
      for (int i = 0; i < T; ++i) {  // shift_pitches, algo.cpp:307-316
        PV q = pv_ld(raw, i);
        q.pitch += p.shift_pre;
        pv_st(raw, i, q);
      }
    T = harmonize_pairs(p.harm_pre, raw, T, cap);
    T = autotune_pairs(p, raw, T);
    if (p.shift_post != 0.f)
      for (int i = 0; i < T; ++i) {
        PV q = pv_ld(raw, i);
        q.pitch += p.shift_post;
        pv_st(raw, i, q);
      }
    T = harmonize_pairs(p.harm_post, raw, T, cap);
    // pairs -> Tuned entries, last first: entry i covers the bytes of pairs 2i and 2i+1, which are >= i
    for (int i = T - 1; i >= 0; --i) {
      const PV q = pv_ld(raw, i);
      Tuned t;
      t.pitch = q.pitch;
      t.vol = q.vol;
      t.ai = 0;
      t.volume = 0;
      t.min_change = 0;
      memcpy(raw + 32 * static_cast<size_t>(i), &t, 32);
    }
  }
  for (int i = 0; i < T; ++i) {
    const float new_freq = static_cast<float>(440.f * pow(p.semitone, out[i].pitch - 69.));
    out[i].volume = static_cast<float>(p.gain * out[i].vol / 0.1f);
    out[i].ai = static_cast<double>(2 * new_freq / p.sr);  // freq_to_angle_increment in float
    out[i].min_change = min_change_for(out[i].ai);
    // from here on `vol` is the key the new notes are ranked by: the perceived loudness, as the float the reference sorts
    out[i].vol = static_cast<double>(static_cast<float>(out[i].vol / equal_loudness_db(p, out[i].pitch)));
  }
  return T;
}

template <class R>
CA_HD inline int pitch_prefix(const PitchParams& p, const int32_t* peak_bin, const R* peak_off, const R* peak_db, int cnt,
                              Tuned* out, int* cut) {
  return p.opt_active ? pitch_prefix_t<true, R>(p, peak_bin, peak_off, peak_db, cnt, out, cut)
                      : pitch_prefix_t<false, R>(p, peak_bin, peak_off, peak_db, cnt, out, cut);
}

struct Note {
  int64_t id = 0;
  double pitch = 0;      // current midi pitch (tracking key)
  Glide glide;
  double phase = 0;      // units of pi rad, [0,2)
  double f_last = 0;     // last emitted increment
  double vol_x = 0, vol_y = 0;
  int64_t n_on = 0;
  int32_t attack = 0, hold = 0, decay = -1, release = 0;
  int32_t next_min_change = 0;
  int64_t n_off = -1;    // sample of the first release step
  double top = 0;
  float gl = 1, gr = 1;
  bool tracked = true;   // still in played_pitches
  int32_t chan = 0;      // channel of the voice pool (ImplCRTP channels, include/gen.crtp.h:221-233): lowest free index at note-on
};

// setPhase<SynchronizePhase::Yes> (include/gen.crtp.h:127-156, instantiated at source/rt.resynth.lib.cpp:201-210): a new note
// takes over the oscillator phase of the first other RT-active channel, in channel order, whose angleIncrements() equals its
// own. angleIncrements() (FreqCtrl_, include/audioelement.h:3117-3124) is the increment the channel's OSCILLATOR last
// stepped with when that is non-zero -- the oscillator forgets nothing when a channel is reused (:2387-2388), so for a
// channel keyed in this frame it is still the last increment of the channel's previous note -- and the glide's start
// value (= the note's own increment) on a channel that never sounded.
CA_HD inline double chan_angle_increments(double osc_ai, double ctrl_from) { return osc_ai != 0.0 ? osc_ai : ctrl_from; }

CA_HD inline double env_pressed_value(const Note& n, int64_t c, float sustain) {
  if (c < 1) return 0.0;
  if (c <= n.attack) return static_cast<double>(c) / n.attack;
  c -= n.attack;
  if (c <= n.hold) return 1.0;
  c -= n.hold;
  const int32_t d = n.decay < 0 ? 0 : n.decay;
  if (c <= d) return 1.0 + (static_cast<double>(sustain) - 1.0) * static_cast<double>(c) / d;
  return sustain;
}

// NoteOff -> onKeyReleased(0), include/audioelement.h:769-781, :836-841
CA_HD inline void note_off(const PitchParams& p, Note& n, int64_t n_k) {
  n.tracked = false;
  n.n_off = n_k;
  n.top = env_pressed_value(n, n_k - n.n_on, p.sustain);
  n.release = n.next_min_change > p.R ? n.next_min_change : p.R;
}

// NoteChange, include/gen.crtp.h:595-618
CA_HD inline void note_change(const PitchParams& p, Note& n, const Tuned& t) {
  n.pitch = t.pitch;
  n.vol_x = static_cast<double>(0.1f * t.volume);
  n.next_min_change = t.min_change;
  n.glide.start(n.f_last, t.ai, p.hop);  // InterpolatedFreq::setAngleIncrements, include/audioelement.h:2752-2762
}

// NoteOn (include/gen.crtp.h:382-594; start phase and pan from a hash of the note id, SPEC S10)
CA_HD inline Note note_on(const PitchParams& p, const Tuned& t, int64_t id, int64_t n_k) {
  Note n;
  n.id = id;
  n.pitch = t.pitch;
  n.vol_x = n.vol_y = static_cast<double>(0.1f * t.volume);
  n.n_on = n_k;
  n.next_min_change = t.min_change;
  n.attack = n.next_min_change > p.A ? n.next_min_change : p.A;
  n.hold = p.H;
  n.decay = -1;
  n.glide.start(t.ai, t.ai, p.hop);
  n.f_last = t.ai;
  double a0 = static_cast<double>(hash_unit(static_cast<uint64_t>(id), 2));
  a0 -= 2.0 * floor(0.5 * a0);
  n.phase = a0;
  double pan = static_cast<double>(p.stereo_spread * hash_unit(static_cast<uint64_t>(id), 1));
  pan = fmin(1.0, fmax(-1.0, pan));
  n.gl = static_cast<float>(cos(0.25 * 3.14159265358979323846 * (pan + 1.0)));
  n.gr = static_cast<float>(sin(0.25 * 3.14159265358979323846 * (pan + 1.0)));
  return n;
}

// a released note leaves the pool once its envelope reached Done2: release steps + 16 more (include/audioelement.h:738-749)
CA_HD inline bool note_done(const Note& n, int64_t n_next) { return n.n_off >= 0 && n_next >= n.n_off + n.release + 17; }

// The closed-form description of note `n` over the tile [n_k_tile, n_k_tile + len) and the note state after it.
CA_HD inline Segment emit_segment(const PitchParams& p, Note& n, int64_t n_k_tile, int len) {
  Segment g{};
  const double cap = p.cap;
  const int64_t i0 = n.glide.i;
  g.phase_start = n.phase;
  g.from = n.glide.from;
  g.to = n.glide.to;
  g.K = n.glide.K;
  g.L = n.glide.L;
  g.c1 = n.glide.c1;
  g.P_m1 = n.glide.P_m1;
  g.m = n.glide.m;
  g.i0 = static_cast<int32_t>(i0 < INT_MAX / 2 ? i0 : INT_MAX / 2);
  const double P_before = n.glide.P(i0 - 1);
  g.P_before = P_before;
  if (i0 >= n.glide.m) {  // steady for the whole tile: re-base so that the step index stays small
    g.i0 = 0;
    g.m = 0;
    g.P_m1 = 0;  // P(-1)
    g.P_before = 0;
  }
  // volume regime over the tile
  const double fa = n.glide.f_at(i0), fb = n.glide.f_at(i0 + len - 1);
  g.cap = static_cast<float>(cap);
  g.vol_x = static_cast<float>(n.vol_x);
  g.vol_dy = static_cast<float>(n.vol_y - n.vol_x);
  double mv_end;
  bool full_cap = false;  // capped rate over a whole frame interval: the decay factor is a constant of the run
  const double adv_end = n.glide.P(i0 + len - 1) - P_before;
  if (fmin(fa, fb) >= cap) {
    g.regime = 0;
    mv_end = cap * len;
    full_cap = len == p.hop;
  } else if (fmax(fa, fb) < cap) {
    g.regime = 1;
    mv_end = adv_end;
  } else {
    // the glide crosses the cap inside the tile: first step on the other side of it (f is monotone along a glide, so the
    // steps on the starting side form a prefix and a bisection finds its length)
    const bool f_first = fa < cap;
    int lo = 0, hi = len - 1;  // invariant: step lo is on the starting side, step hi is not
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if ((n.glide.f_at(i0 + mid) < cap) == f_first) lo = mid;
      else hi = mid;
    }
    const int istar = hi;
    g.istar = istar;
    const double adv_star = n.glide.P(i0 + istar - 1) - P_before;
    if (f_first) {
      g.regime = 2;
      g.m_star = static_cast<float>(adv_star);
      mv_end = adv_star + cap * (len - istar);
    } else {
      g.regime = 3;
      g.m_star = static_cast<float>(cap * istar - adv_star);
      mv_end = cap * istar + (adv_end - adv_star);
    }
  }
  // envelope
  g.c0 = static_cast<int32_t>(n_k_tile - n.n_on + 1);
  if (n.decay < 0 && (n_k_tile + len - n.n_on) > static_cast<int64_t>(n.attack) + n.hold)  // decay latched when entered
    n.decay = p.D == 0 ? 0 : (n.next_min_change > p.D ? n.next_min_change : p.D);
  g.attack = n.attack;
  g.hold = n.hold;
  g.decay = n.decay < 0 ? 0 : n.decay;
  g.sustain = p.sustain;
  g.release = n.release > 0 ? n.release : 1;
  const int64_t rb = n.n_off - n_k_tile;
  g.rel_base = n.n_off < 0 ? INT_MAX : static_cast<int32_t>(rb > -(1ll << 30) ? rb : -(1ll << 30));
  g.top = static_cast<float>(n.top);
  g.gl = n.gl;
  g.gr = n.gr;
  // advance
  n.phase += adv_end;
  n.phase -= 2.0 * floor(0.5 * n.phase);
  n.f_last = fb;
  n.glide.i = i0 + len;
  // decay of the volume low-pass over the tile. Single precision is enough here: the factor scales the (small, shrinking)
  // distance of the filter state from its target, which reaches the output as a float32 anyway, and a double exp() costs ~330
  // clocks on the frame-sequential path of the segments warp.
  n.vol_y = n.vol_x + (n.vol_y - n.vol_x) *
                          (full_cap ? p.decay_full : static_cast<double>(expf(static_cast<float>(-3.14159265358979323846 * mv_end))));
  return g;
}

}  // namespace ca

#include <algorithm>
#include <vector>
namespace ca {
// Host form of the voice pool's channels (ImplCRTP, include/gen.crtp.h:221-233): which note of `notes` sits on each channel
// (-1: free) and the increment the channel's oscillator last stepped with (it survives the note).
struct HostChannels {
  std::vector<int> note;
  std::vector<double> osc_ai;
  explicit HostChannels(int n) : note(n > 0 ? n : 1, -1), osc_ai(n > 0 ? n : 1, 0.0) {}
  void rebuild(const std::vector<Note>& notes) {
    std::fill(note.begin(), note.end(), -1);
    for (size_t i = 0; i < notes.size(); ++i) note[notes[i].chan] = static_cast<int>(i);
  }
};
// NoteOn on the first free channel (include/gen.crtp.h:398-410) with the start-phase rule of setPhase (:127-156).
// Returns the index of the new note in `notes`, or -1 when every channel is taken (DROPPED_NOTE, :412-414).
inline int pool_note_on(const PitchParams& p, std::vector<Note>& notes, HostChannels& ch, const Tuned& t, int64_t id, int64_t n_k) {
  if (static_cast<int>(notes.size()) >= p.max_channels) return -1;  // every entry of `notes` still holds a channel
  Note nn = note_on(p, t, id, n_k);
  int c = 0;
  while (ch.note[c] >= 0) ++c;
  nn.chan = c;
  const double ai = chan_angle_increments(ch.osc_ai[c], nn.glide.from);
  for (int o = 0; o < p.max_channels; ++o) {
    if (o == c || ch.note[o] < 0) continue;
    const Note& other = notes[ch.note[o]];
    if (chan_angle_increments(ch.osc_ai[o], other.glide.from) == ai) {
      nn.phase = other.phase;  // synchronizeAngles: the other oscillator's phasor as of its last step
      break;
    }
  }
  notes.push_back(nn);
  ch.note[c] = static_cast<int>(notes.size()) - 1;
  return ch.note[c];
}
}  // namespace ca

struct ca_ctx;
namespace ca {
// pitch.cu: tuned lists (pitch_prefix_kernel), then tracking and segments (pitch_track3_kernel, three warps per stream)
int pitch_prefix_device(ca_ctx* ctx, const PitchParams& p, int64_t rows, const int32_t* d_peak_bin, const void* d_peak_off, const void* d_peak_db,
                        bool peaks_f64, const int32_t* d_peak_count, unsigned long long* d_truncated, Tuned* d_tuned, int32_t* d_tcount,
                        cudaStream_t st);
size_t pitch_track_state_bytes();
int pitch_track_device(ca_ctx* ctx, const PitchParams& p, int64_t n_streams, const Tuned* d_tuned, const int32_t* d_tcount, int64_t k_begin,
                       int64_t k_end, unsigned char* d_state, Segment* d_segs, int64_t seg_capacity, int64_t n_tiles_per_stream,
                       int64_t* d_tile_begin, int64_t* d_tile_end, int64_t* d_tile_s0, int32_t* d_tile_len, int64_t* d_stats,
                       unsigned long long* d_counter, int32_t* d_overflow, cudaStream_t st);
// d_peak_off / d_peak_db: float arrays, or double arrays when peaks_f64 is set (the FP64 analysis of stft64.cu)
int pitch_pipeline_device(ca_ctx* ctx, const PitchParams& p, int64_t n_streams, const int32_t* d_peak_bin, const void* d_peak_off,
                          const void* d_peak_db, bool peaks_f64, const int32_t* d_peak_count, unsigned long long* d_truncated, Tuned* d_tuned, int32_t* d_tcount, Segment* d_segs,
                          int64_t seg_capacity, int64_t n_tiles_per_stream, int64_t* d_tile_begin, int64_t* d_tile_end,
                          int64_t* d_tile_s0, int32_t* d_tile_len, int64_t* d_stats, unsigned long long* d_counter,
                          int32_t* d_overflow, bool run_prefix, cudaStream_t st, bool reset_counters = true);
}  // namespace ca
