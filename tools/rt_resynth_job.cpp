// rt_resynth_job — the offline rt.resynth job (source/rt.resynth.lib.cpp:1185-1235, source/rt.resynth.lib.params.cpp:
// 183-386) on the B200 path: a job file names a preset, a voice WAV file and an output file; the voice is analysed and
// resynthesised (ca_resynth_roundtrip_host); when the preset gives the voice, the carrier or the vocoder a volume, the vocoder
// stage (voice = modulator, carrier file = carrier: ca_vocoder_host) is mixed into both channels as the reference's compute
// callbacks do (rt.resynth.lib.cpp:1378-1420, rt.resynth.lib.vocoder.cpp:777-812); the result is written as IEEE-float stereo.
//
//   rt_resynth_job job.json [more job files...]          every job of the command line goes to the GPU in ONE batch
//   rt_resynth_job --voice in.wav --out out.wav [--preset preset.json]
//   rt_resynth_job --dry-run job.json                    parse and print the derived parameters, no GPU needed
//
// Job file:    {"preset_file": "...", "input_voice_file": "...", "input_carrier_file": "", "output_file": "...",
//               "post": "none" | "limit"}                                     (RtResynthOfflineJobConfig, params.cpp:183-212)
// Preset file: {"bool_params": {..}, "enum_params": {..}, "int32_params": {..}, "uint64_params": {..},
//               "float_params": {"window_size_seconds": 0.1814, ...}}         (Preset, params.cpp:57-141)
// The float parameters of the analysis -> resynthesis path are honoured (names of rt.resynth.lib.cpp:1944-2040): window and
// stride, min_volume, nearby_distance_tones, max_track_pitches, stereo_spread, env_*, analysis_volume, and of the vocoder:
// voice_volume, carrier_volume, vocoder_volume, vocoder_modulator_window_size_seconds, vocoder_stride_seconds,
// vocoder_count_bands (int32), vocoder_min_freq, vocoder_max_freq; and the pitch stages of RtResynth::step (:1687-1707):
// pitch_shift_*, pitch_harmonize_*, use_autotune with autotune_type / scale mode / root note / transpose / chord bits /
// chord frequency layout / interval factor / tolerance / max pitch, turned into ca_pitch_options the way mkAutotuneFunction
// (:1762-1869) builds its transform. The real-time carrier synth is outside this path. "limit" post-processing (an external
// compressor in the reference) clips to [-1, 1].
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../include/cpp_audio_b200.h"
#include "mini_json.hpp"
#include "wav_io.hpp"

namespace {

struct Job {
  std::string preset_file, voice_file, carrier_file, output_file, post = "none";
};

struct Derived {  // what the preset means at a sample rate (RtResynth getters, rt.resynth.lib.cpp:484-487, :1576-1578)
  int window = 0, hop = 0;
  ca_roundtrip_params prm{};  // prm.opts.table stays null here: the table lives in `table` and is attached at the call
  std::vector<double> table;  // autotune: allowed pitches (Chord / FixedSizeIntervals) or the scale's pitches (MusicalScale)
  ca_vocoder_params voc{};  // Vocoder::Params, rt.resynth.lib.cpp:1378-1395
};

std::string slurp(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("cannot open '" + path + "'");
  std::ostringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

Job read_job(const std::string& path) {
  const mini_json::Value j = mini_json::parse(slurp(path));
  Job job;
  job.preset_file = j.at("preset_file").as_string();
  job.voice_file = j.at("input_voice_file").as_string();
  job.carrier_file = j.at("input_carrier_file").as_string();
  job.output_file = j.at("output_file").as_string();
  job.post = j.at("post").as_string();
  if (job.post != "none" && job.post != "limit") throw std::runtime_error("bad encoding of 'post'");  // params.cpp:170-178
  return job;
}

float preset_float(const mini_json::Value* preset, const char* name, float dflt) {
  if (!preset) return dflt;
  const mini_json::Value* fp = preset->find("float_params");
  const mini_json::Value* v = fp ? fp->find(name) : nullptr;
  if (!v) {
    std::fprintf(stderr, "Preset param '%s' not found\n", name);  // Params::read, params.cpp:19-27
    return dflt;
  }
  return static_cast<float>(v->as_number());
}

int preset_int(const mini_json::Value* preset, const char* name, int dflt) {
  if (!preset) return dflt;
  const mini_json::Value* ip = preset->find("int32_params");
  const mini_json::Value* v = ip ? ip->find(name) : nullptr;
  if (!v) {
    std::fprintf(stderr, "Preset param '%s' not found\n", name);
    return dflt;
  }
  return static_cast<int>(v->as_number());
}

Derived derive(const mini_json::Value* preset, int sample_rate) {
  Derived d;
  // defaults: rt.resynth.lib.cpp:940-961
  const float ws = preset_float(preset, "window_size_seconds", 0.1814f);
  const float stride = preset_float(preset, "window_center_stride_seconds", 0.09f);
  d.window = 2 * std::max(1, static_cast<int>(0.5 * ws * sample_rate));          // getEvenWindowSizeFrames, :484-487
  d.hop = std::max(1, static_cast<int>(0.5f + stride * sample_rate));            // :1577-1578
  ca_roundtrip_params& p = d.prm;
  p.stft.window = d.window;
  p.stft.hop = d.hop;
  p.stft.window_type = CA_WINDOW_GAUSSIAN;
  p.stft.sigma_factor = 4.f;  // periodicfft.cpp:288
  p.stft.sample_rate = sample_rate;
  p.min_volume = preset_float(preset, "min_volume", 0.0001f);
  // The reference thresholds on volume (reduce_pitches, rt.resynth.lib.algo.cpp:166-184), not on the spectrum, and its peak list
  // is unbounded (reserve(200) at rt.resynth.lib.cpp:1102 grows): a spectrum of n_fft points has at most n_fft/4 strict local
  // maxima, so that capacity never cuts a list.
  p.stft.min_db = -INFINITY;
  p.stft.max_peaks = std::max(200, ca_fft_length_for(d.window, 1) / 4);
  p.nearby_distance_tones = preset_float(preset, "nearby_distance_tones", 0.4f);
  p.max_track_pitches = preset_float(preset, "max_track_pitches", 1.f);
  p.stereo_spread = preset_float(preset, "stereo_spread", 1.f);
  p.gain = preset_float(preset, "analysis_volume", 1.f);
  auto seconds = [&](const char* name) { return static_cast<int32_t>(0.5f + preset_float(preset, name, 0.f) * sample_rate); };
  p.env.attack = seconds("env_attack_seconds");
  p.env.hold = seconds("env_hold_seconds");
  p.env.decay = seconds("env_decay_seconds");
  p.env.release = seconds("env_release_seconds");
  p.env.sustain = preset_float(preset, "env_sustain_level", 1.f);
  p.max_channels = 254;
  p.host_threads = 0;
  // vocoder stage (defaults: rt.resynth.lib.cpp:986-994)
  d.voc.sample_rate = sample_rate;
  d.voc.modulator_window_size_seconds = preset_float(preset, "vocoder_modulator_window_size_seconds", 0.10f);
  d.voc.stride_seconds = preset_float(preset, "vocoder_stride_seconds", 0.005f);
  d.voc.count_bands = preset_int(preset, "vocoder_count_bands", 5);
  d.voc.min_freq = preset_float(preset, "vocoder_min_freq", 100.f);
  d.voc.max_freq = preset_float(preset, "vocoder_max_freq", 20000.f);
  d.voc.vol_modulator = preset_float(preset, "voice_volume", 0.f);
  d.voc.vol_carrier = preset_float(preset, "carrier_volume", 0.1f);
  d.voc.vol_vocoder = preset_float(preset, "vocoder_volume", 0.f);
  // shift / harmonize / autotune stages of RtResynth::step (rt.resynth.lib.cpp:1687-1707; defaults :950-976)
  ca_pitch_options& o = p.opts;
  auto quiet_float = [&](const char* name, float dflt) {
    const mini_json::Value* fp = preset ? preset->find("float_params") : nullptr;
    const mini_json::Value* v = fp ? fp->find(name) : nullptr;
    return v ? static_cast<float>(v->as_number()) : dflt;
  };
  auto quiet_num = [&](const char* section, const char* name, double dflt) {
    const mini_json::Value* sp = preset ? preset->find(section) : nullptr;
    const mini_json::Value* v = sp ? sp->find(name) : nullptr;
    return v ? v->as_number() : dflt;
  };
  o.shift_pre_autotune = quiet_float("pitch_shift_pre_autotune", 0.f);
  o.shift_post_autotune = quiet_float("pitch_shift_post_autotune", 0.f);
  o.harmonize_pre_autotune = quiet_float("pitch_harmonize_pre_autotune", 0.f);
  o.harmonize_post_autotune = quiet_float("pitch_harmonize_post_autotune", 0.f);
  if (quiet_num("bool_params", "use_autotune", 0) != 0) {  // mkAutotuneFunction, rt.resynth.lib.cpp:1762-1869
    o.autotune_tolerance_pitches = quiet_float("autotune_tolerance_pitches", 100.f);
    o.autotune_max_pitch = static_cast<float>(static_cast<int>(quiet_num("int32_params", "autotune_max_pitch", 150)));
    const int type = static_cast<int>(quiet_num("enum_params", "autotune_type", 0));  // MusicalScale, Chord, FixedSizeIntervals
    const int root_note = static_cast<int>(quiet_num("enum_params", "autotune_musical_scale_root_note", 0));  // Note::Do = 0 ... Si = 11
    const int transpose = static_cast<int>(quiet_num("int32_params", "autotune_root_note_halftones_transpose", 0));
    constexpr int kLa = 9;  // include/midi.h:10-23
    if (type == 0) {
      static const double scales[3][8] = {{0, 2, 4, 5, 7, 9, 11, 12}, {0, 2, 3, 5, 7, 8, 10, 12}, {0, 2, 3, 5, 7, 8, 11, 12}};  // autotune.cpp:142-175
      const int mode = static_cast<int>(quiet_num("enum_params", "autotune_musical_scale_mode", 0));
      if (mode < 0 || mode > 2) throw std::runtime_error("bad autotune_musical_scale_mode");
      o.autotune = CA_AUTOTUNE_SCALE;
      d.table.assign(scales[mode], scales[mode] + 8);
      o.scale_root_pitch = 69. + transpose + (root_note - kLa);  // A_pitch + transpose + half_tones_distance(La, root), :1862-1865
    } else if (type == 1) {
      int offset = root_note;  // half_tones_distance(Do, root) is never negative
      offset += transpose;
      const double root_pitch = 69. + (0 - kLa) + 12 + offset;  // C_pitch + offset, :1775-1777
      const uint64_t chord = static_cast<uint64_t>(quiet_num("uint64_params", "autotune_bit_chord", 0b10010001));
      const int freqs = static_cast<int>(quiet_num("enum_params", "autotune_chord_frequencies", 1));  // SingleFreq, Harmonics, OctavePeriodic
      if (freqs == 0 || freqs == 2) {
        const int lo = freqs == 0 ? 0 : -5, hi = freqs == 0 ? 0 : 5;
        for (int octave = lo; octave <= hi; ++octave)
          for (int i = 0; i < 64; ++i)
            if ((chord >> i) & 1) d.table.push_back(root_pitch + i + 12 * octave);
      } else if (freqs == 1) {
        for (int h = 0; h < 36; ++h) {  // compute_harmonic_pitch_adds, include/midi.h:278-287
          const double add = static_cast<int>(0.5 + 12. * std::log2(static_cast<float>(h + 1)));
          for (int i = 0; i < 64; ++i)
            if ((chord >> i) & 1) {
              d.table.push_back(root_pitch + add + i);
              d.table.push_back(root_pitch - add + i);
            }
        }
      } else {
        throw std::runtime_error("bad autotune_chord_frequencies");
      }
      std::sort(d.table.begin(), d.table.end());
      o.autotune = CA_AUTOTUNE_TABLE;
    } else if (type == 2) {
      const double offset = root_note + transpose;
      const int factor = static_cast<int>(quiet_num("int32_params", "autotune_factor", 2));
      d.table.push_back(offset);
      if (factor > 0) {
        for (double val = offset - factor; val > 0.; val -= factor) d.table.push_back(val);
        for (double val = offset + factor; val < 151.; val += factor) d.table.push_back(val);  // max_audible_midi_pitch, midi.h:134
      }
      std::sort(d.table.begin(), d.table.end());
      o.autotune = CA_AUTOTUNE_TABLE;
    } else {
      throw std::runtime_error("bad autotune_type");
    }
    o.n_table = static_cast<int32_t>(d.table.size());
  }
  return d;
}

void print_derived(const Job& job, const Derived& d, int sample_rate, size_t frames) {
  std::printf("{\"voice\": \"%s\", \"output\": \"%s\", \"sample_rate\": %d, \"frames\": %zu, \"window\": %d, \"hop\": %d, "
              "\"fft_length\": %d, \"min_volume\": %g, \"nearby_distance_tones\": %g, \"max_track_pitches\": %g, "
              "\"stereo_spread\": %g, \"gain\": %g, \"env\": [%d, %d, %d, %d, %g], \"post\": \"%s\", "
              "\"pitch_options\": {\"shift\": [%g, %g], \"harmonize\": [%g, %g], \"autotune\": %d, \"max_pitch\": %g, "
              "\"tolerance\": %g, \"n_table\": %d, \"table_first\": %g, \"table_last\": %g, \"scale_root_pitch\": %g}}\n",
              job.voice_file.c_str(), job.output_file.c_str(), sample_rate, frames, d.window, d.hop, ca_fft_length_for(d.window, 1),
              d.prm.min_volume, d.prm.nearby_distance_tones, d.prm.max_track_pitches, d.prm.stereo_spread, d.prm.gain,
              d.prm.env.attack, d.prm.env.hold, d.prm.env.decay, d.prm.env.release, d.prm.env.sustain, job.post.c_str(),
              d.prm.opts.shift_pre_autotune, d.prm.opts.shift_post_autotune, d.prm.opts.harmonize_pre_autotune,
              d.prm.opts.harmonize_post_autotune, d.prm.opts.autotune, d.prm.opts.autotune_max_pitch,
              d.prm.opts.autotune_tolerance_pitches, d.prm.opts.n_table, d.table.empty() ? 0. : d.table.front(),
              d.table.empty() ? 0. : d.table.back(), d.prm.opts.scale_root_pitch);
}

bool same_params(const Derived& a, const Derived& b) { return std::memcmp(&a.prm, &b.prm, sizeof(a.prm)) == 0 && a.table == b.table; }

// the vocoder stage contributes when the voice is mixed in, or when there is a carrier and it (or its vocoded form) is audible
bool vocoder_active(const Derived& d, bool has_carrier) {
  return d.voc.vol_modulator != 0.f || (has_carrier && (d.voc.vol_carrier != 0.f || d.voc.vol_vocoder != 0.f));
}

}  // namespace

int main(int argc, char** argv) {
  try {
    std::vector<Job> jobs;
    bool dry = false, allow_truncation = false;
    Job cli;
    for (int i = 1; i < argc; ++i) {
      const std::string a = argv[i];
      auto next = [&]() -> std::string {
        if (i + 1 >= argc) throw std::runtime_error("missing value after " + a);
        return argv[++i];
      };
      if (a == "--dry-run") dry = true;
      else if (a == "--allow-truncation") allow_truncation = true;
      else if (a == "--voice") cli.voice_file = next();
      else if (a == "--out") cli.output_file = next();
      else if (a == "--preset") cli.preset_file = next();
      else if (a == "--post") cli.post = next();
      else jobs.push_back(read_job(a));
    }
    if (!cli.voice_file.empty()) jobs.push_back(cli);
    if (jobs.empty()) {
      std::fprintf(stderr, "usage: rt_resynth_job [--dry-run] [--allow-truncation] job.json... | --voice in.wav --out out.wav [--preset p.json] [--post none|limit]\n");
      return 2;
    }
    struct Loaded {
      Job job;
      Derived d;
      int sample_rate;
      std::vector<float> voice;
      std::vector<float> carrier;  // empty: no carrier file (the reference then feeds zeros, params.cpp:298-303)
    };
    std::vector<Loaded> loaded;
    for (const Job& job : jobs) {
      if (job.voice_file.empty()) throw std::runtime_error("must have a voice file (carrier-only jobs have nothing to analyse)");
      if (job.output_file.empty()) throw std::runtime_error("no output file");  // params.cpp:344-346
      mini_json::Value preset;
      const bool has_preset = !job.preset_file.empty();
      if (has_preset) {
        try {
          preset = mini_json::parse(slurp(job.preset_file));
        } catch (const std::exception& e) {
          std::fprintf(stderr, "Reading preset file '%s': %s\n", job.preset_file.c_str(), e.what());  // params.cpp:262-268
          throw;
        }
      }
      const wav_io::Audio audio = wav_io::read(job.voice_file);
      Loaded l{job, derive(has_preset ? &preset : nullptr, audio.sample_rate), audio.sample_rate, wav_io::as_one_float(audio), {}};
      if (!job.carrier_file.empty()) {
        const wav_io::Audio car = wav_io::read(job.carrier_file);
        if (car.sample_rate != audio.sample_rate) throw std::runtime_error("sample rate mismatch between carrier and voice");  // params.cpp:270-274
        l.carrier = wav_io::as_one_float(car);
      }
      print_derived(job, l.d, l.sample_rate, l.voice.size());
      loaded.push_back(std::move(l));
    }
    if (dry) return 0;

    ca_ctx* ctx = nullptr;
    if (int rc = ca_init(0, &ctx)) {
      std::fprintf(stderr, "ca_init failed (%d): this tool needs a B200; there is no CPU fallback\n", rc);
      return 1;
    }
    // jobs with identical parameters and sample rate are batched as streams of one call (zero padded to the longest)
    std::vector<bool> done(loaded.size(), false);
    for (size_t i = 0; i < loaded.size(); ++i) {
      if (done[i]) continue;
      std::vector<size_t> group;
      size_t longest = 0;
      for (size_t k = i; k < loaded.size(); ++k)
        if (!done[k] && loaded[k].sample_rate == loaded[i].sample_rate && same_params(loaded[k].d, loaded[i].d)) {
          group.push_back(k);
          done[k] = true;
          longest = std::max(longest, loaded[k].voice.size());
        }
      // the reference keeps running until the output has been silent for one second (rt.resynth.lib.cpp:1228-1234)
      const size_t n = longest + loaded[i].sample_rate + loaded[i].d.window;
      float* x = static_cast<float*>(ca_host_alloc(sizeof(float) * n * group.size()));
      float* y = static_cast<float*>(ca_host_alloc(sizeof(float) * 2 * n * group.size()));
      std::vector<int64_t> stats(2 * group.size());
      if (!x || !y) throw std::runtime_error("pinned allocation failed");
      std::memset(x, 0, sizeof(float) * n * group.size());
      for (size_t g = 0; g < group.size(); ++g)
        std::memcpy(x + g * n, loaded[group[g]].voice.data(), sizeof(float) * loaded[group[g]].voice.size());
      ca_roundtrip_params prm = loaded[i].d.prm;
      prm.opts.table = loaded[i].d.table.empty() ? nullptr : loaded[i].d.table.data();
      if (int rc = ca_resynth_roundtrip_host(ctx, &prm, x, static_cast<int64_t>(n), static_cast<int64_t>(group.size()), y, stats.data())) {
        std::fprintf(stderr, "ca_resynth_roundtrip_host failed (%d): %s\n", rc, ca_last_error(ctx));
        return 1;
      }
      if (const long long cut = static_cast<long long>(ca_resynth_truncated_frames(ctx))) {
        std::fprintf(stderr,
                     "rt_resynth_job: %lld analysis frames held more than 256 pitches above min_volume; the excess (highest pitches) was "
                     "dropped, which the reference would not do. %s\n",
                     cut, allow_truncation ? "Continuing (--allow-truncation)." : "Raise min_volume in the preset, or pass --allow-truncation.");
        if (!allow_truncation) return 3;
      }
      // vocoder stage, one call per job that uses it: mixed[n] goes to both channels (rt.resynth.lib.vocoder.cpp:803-810)
      for (size_t g = 0; g < group.size(); ++g) {
        const Loaded& l = loaded[group[g]];
        if (!vocoder_active(l.d, !l.carrier.empty())) continue;
        float* car = static_cast<float*>(ca_host_alloc(sizeof(float) * n));
        float* mix = static_cast<float*>(ca_host_alloc(sizeof(float) * n));
        if (!car || !mix) throw std::runtime_error("pinned allocation failed");
        std::memset(car, 0, sizeof(float) * n);
        std::memcpy(car, l.carrier.data(), sizeof(float) * std::min(n, l.carrier.size()));
        if (int rc = ca_vocoder_host(ctx, &l.d.voc, x + g * n, car, static_cast<int64_t>(n), 1, mix)) {
          std::fprintf(stderr, "ca_vocoder_host failed (%d): %s\n", rc, ca_last_error(ctx));
          return 1;
        }
        float* out = y + 2 * g * n;
        for (size_t k = 0; k < n; ++k) {
          out[2 * k] += mix[k];
          out[2 * k + 1] += mix[k];
        }
        ca_host_free(car);
        ca_host_free(mix);
      }
      for (size_t g = 0; g < group.size(); ++g) {
        const Loaded& l = loaded[group[g]];
        float* out = y + 2 * g * n;
        size_t frames = l.voice.size() + l.sample_rate;  // input length + the second of silence that ends the reference loop
        frames = std::min(frames, n);
        if (l.job.post == "limit")
          for (size_t k = 0; k < 2 * frames; ++k) out[k] = std::min(1.f, std::max(-1.f, out[k]));
        wav_io::write_float32(l.job.output_file, l.sample_rate, 2, out, frames);
        std::fprintf(stderr, "%s: %zu frames written, %lld notes started, %lld dropped\n", l.job.output_file.c_str(), frames,
                     static_cast<long long>(stats[2 * g]), static_cast<long long>(stats[2 * g + 1]));
      }
      ca_host_free(x);
      ca_host_free(y);
    }
    ca_destroy(ctx);
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "rt_resynth_job: %s\n", e.what());
    return 1;
  }
}
