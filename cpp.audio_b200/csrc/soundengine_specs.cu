// soundengine_specs.cu — the ramp specs of a gen.voice note, as SoundEngine's Markov chains produce them at note-on
// (include/soundengine.h:496-736: play(), create_birds / create_robot / create_sweep; :883-1003: initialize_*, do_initialize),
// and the harmonic tables of the non-sine oscillators. Control-rate host code: a note is a few dozen draws.
//
// The lambdas attached to the chains' nodes are in the reference tree and are restated here. The MarkovChain class itself and
// the random generator (mersenne<SEEDED::Yes>()) are in cpp.algorithms, which is absent: oracle/SPEC.md S15 fixes
//   step_normalized(u)  the outgoing transitions of the current node, probabilities rescaled to sum 1, picked by u in [0,1)
//   step(u)             the raw probabilities; u beyond their sum: stay (no lambda runs)
//   a move from a to b  runs a's lambda with LEAVE_NODE, then b's with ENTER_NODE
//   draws               a splitmix64 stream seeded by the caller; uniform floats from its top 24 bits
#include <math.h>

#include <algorithm>
#include <functional>
#include <vector>

#include "ctx.hpp"

namespace ca {

namespace {

struct Rng {
  uint64_t s;
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  float unit() { return static_cast<float>(next() >> 40) * (1.f / 16777216.f); }  // [0,1)
  float uniform(float a, float b) { return a + (b - a) * unit(); }
};

enum Move { ENTER_NODE, LEAVE_NODE };

struct Chain {
  struct Edge {
    int to;
    float p;
  };
  std::vector<std::function<void(Move)>> on;
  std::vector<std::vector<Edge>> out;
  int cur = 0;
  int emplace(std::function<void(Move)> f) {
    on.push_back(std::move(f));
    out.emplace_back();
    return static_cast<int>(on.size()) - 1;
  }
  void transition(int a, int b, float p) { out[a].push_back({b, p}); }
  void move(int to, bool exec) {
    if (exec) on[cur](LEAVE_NODE);
    cur = to;
    if (exec) on[cur](ENTER_NODE);
  }
  void step(float u, bool normalized, bool exec) {
    const auto& e = out[cur];
    float sum = 0;
    for (auto& x : e) sum += x.p;
    if (e.empty() || sum <= 0) return;
    float acc = 0;
    const float target = normalized ? u * sum : u;
    for (auto& x : e) {
      acc += x.p;
      if (target < acc) {
        move(x.to, exec);
        return;
      }
    }
    if (normalized) move(e.back().to, exec);  // u * sum rounded up to the total
  }
};

int ms_to_frames(float ms, int sr) { return static_cast<int>(0.5f + (static_cast<float>(sr) / 1000.f) * ms); }  // sound.functions.h:27-33
float f2ai(float f, int sr) { return 2 * f / static_cast<float>(sr); }                                           // :58-62
float ai2f(float ai, int sr) { return ai * static_cast<float>(sr) / 2; }
double transpose(double freq, int n) { return freq * pow(pow(2., 1. / 12.), static_cast<double>(n)); }          // midi.h:269-272

struct Builder {
  const ca_sound_engine_gen& g;
  Rng rng;
  std::vector<ca_ramp_spec> specs;  // at most 30 (RampSpecs::n_specs)
  float state_freq = 0.f, state_factor = 0.f;
  float freq1_robot = 0, freq2_robot = 0, vol1 = 1.f, vol2 = 1.f;

  ca_ramp_spec* next_for_build() {
    if (specs.size() >= 30) return nullptr;
    specs.emplace_back();
    return &specs.back();
  }
  // LogRamp::setup keeps angle increments; the spec carries Hz, so go through the increment to round as the reference does
  static void setup(ca_ramp_spec* s, float from_ai, float to_ai, float n_frames, float start, int itp, int sr) {
    s->from_hz = ai2f(from_ai, sr);
    s->to_hz = ai2f(to_ai, sr);
    s->duration = n_frames;
    s->start_sample = start;
    s->itp = itp;
    s->volume = 1.f;
    s->silence_follows = 1;
  }
  float scaled_frames(float length) {
    length *= powf(2.f, rng.uniform(g.length_exp_min, g.length_exp_max));
    return static_cast<float>(ms_to_frames(length, g.sample_rate));
  }
  // SoundEngine::play, soundengine.h:496-563
  void play(float freq1, float freq2) {
    const float n_frames = scaled_frames(g.length_ms);
    if (n_frames <= 0) return;
    const bool had_current = !specs.empty();
    const ca_ramp_spec current = had_current ? specs.back() : ca_ramp_spec{};
    ca_ramp_spec* spec = next_for_build();
    if (!spec) return;
    if (state_freq != freq1) {
      const float scatter = 1.f + g.freq_scatter;
      state_factor = rng.uniform(1.f / scatter, scatter);
    }
    state_freq = freq2;
    freq1 *= state_factor;
    freq2 *= state_factor;
    setup(spec, f2ai(freq1, g.sample_rate), f2ai(freq2, g.sample_rate), n_frames, 0.f, g.itp, g.sample_rate);
    if (g.xfade_freq == 0 || !had_current) return;
    float from_inc = f2ai(current.to_hz, g.sample_rate);
    const float to_inc = f2ai(spec->from_hz, g.sample_rate);
    if (g.xfade_freq == 2 || from_inc - to_inc != 0.f) {  // the new spec starts at another frequency: glide to it first
      const ca_ramp_spec moved = *spec;
      if (ca_ramp_spec* spec2 = next_for_build()) {
        *spec2 = moved;
        spec = &specs[specs.size() - 2];
        if (from_inc == to_inc) from_inc *= 1.00001f;
        setup(spec, from_inc, to_inc, static_cast<float>(g.freq_xfade_samples), 0.f, g.freq_itp, g.sample_rate);
      } else {
        specs.pop_back();  // cancel_last_ramp
      }
    }
  }
  void robot_spec(float f_from, float f_to, float n_frames, float vol, bool silence) {
    if (ca_ramp_spec* s = next_for_build()) {
      setup(s, f2ai(f_from, g.sample_rate), f2ai(f_to, g.sample_rate), n_frames, g.phase_ratio1 * n_frames, g.itp, g.sample_rate);
      s->volume = vol;
      s->silence_follows = silence ? 1 : 0;
    }
  }
};

}  // namespace

int sound_engine_specs(const ca_sound_engine_gen* gen, ca_ramp_spec* out, int32_t capacity) {
  if (!gen || !out || capacity < 0 || gen->sample_rate <= 0 || gen->mode < 0 || gen->mode > 2) return CA_ERR_INVALID;
  if (!(gen->base_freq > 0.f) && gen->mode != 2) return CA_ERR_INVALID;  // do_initialize :931-934
  if (ms_to_frames(gen->length_ms, gen->sample_rate) <= 0) return CA_ERR_INVALID;  // :925-929
  Builder b{*gen, Rng{0x5eedull ^ (static_cast<uint64_t>(gen->seed) << 17)}, {}};
  const ca_sound_engine_gen& g = *gen;
  Chain mc;
  if (g.mode == 0) {  // create_birds, :565-596
    const int n1 = mc.emplace([](Move) {});
    const int n2 = mc.emplace([&](Move m) {
      if (m == ENTER_NODE) b.play(g.base_freq * 4, g.base_freq * 3);
      else b.play(g.base_freq * 2, g.base_freq * 4);
    });
    const int n3 = mc.emplace([&](Move m) {
      if (m == ENTER_NODE) b.play(g.base_freq * 4, g.base_freq * 3);
    });
    mc.transition(n1, n2, 0.5f);
    mc.transition(n2, n1, 0.015f);
    mc.transition(n1, n3, 0.5f);
    mc.transition(n3, n1, 0.015f);
    mc.transition(n3, n2, 0.885f);
  } else if (g.mode == 1) {  // initialize_robot :909-944 + create_robot :598-704
    const float scatter = 1.f + g.freq_scatter;
    constexpr float detune = 0.985f;
    b.freq1_robot = b.rng.uniform(g.base_freq / scatter, g.base_freq * scatter);
    b.freq2_robot = b.rng.uniform(b.freq1_robot * detune, b.freq1_robot / detune);
    b.vol1 = b.vol2 = 1.f;
    if (!(b.rng.next() & 1)) {
      b.freq1_robot = static_cast<float>(transpose(b.freq1_robot, static_cast<int>(g.d1)));
      b.vol1 = powf(g.har_att, g.d1);
    } else {
      b.freq2_robot = static_cast<float>(transpose(b.freq2_robot, static_cast<int>(g.d2)));
      b.vol2 = powf(g.har_att, g.d2);
    }
    const int n0 = mc.emplace([&](Move m) {
      if (m == LEAVE_NODE) {
        const float nf = b.scaled_frames(g.length_ms);
        b.robot_spec(b.freq1_robot, b.freq1_robot, nf, b.vol1, false);
        b.robot_spec(b.freq2_robot, b.freq2_robot, nf, b.vol2, true);
      }
    });
    const int n1 = mc.emplace([](Move) {});
    const int n2 = mc.emplace([&](Move m) {
      if (m == ENTER_NODE) {
        const float nf = b.scaled_frames(g.length_ms);
        b.robot_spec(b.freq2_robot, b.freq2_robot, nf, b.vol2, true);
        const float f2 = static_cast<float>(transpose(b.freq2_robot, 2)), f4 = static_cast<float>(transpose(b.freq2_robot, 4));
        b.robot_spec(f2, f2, nf, b.vol2, true);
        b.robot_spec(f4, f4, nf, b.vol2, true);
      }
    });
    const int n3 = mc.emplace([&](Move m) {
      if (m == ENTER_NODE) {
        const float nf = b.scaled_frames(2.f * g.length_ms);
        b.robot_spec(b.freq2_robot, b.freq1_robot, nf, std::min(b.vol2, b.vol1), true);
      }
    });
    mc.transition(n0, n1, 1.f);
    mc.transition(n1, n2, 0.2f);
    mc.transition(n2, n1, 0.1f);
    mc.transition(n1, n3, 0.2f);
    mc.transition(n3, n1, 1.f);
  } else {  // initialize_sweep :883-895 + create_sweep :706-724
    b.freq1_robot = g.low_freq;
    b.freq2_robot = g.high_freq;
    if (!(g.low_freq > 0.f) || !(g.high_freq > 0.f)) return CA_ERR_INVALID;
    const int n0 = mc.emplace([&](Move m) {
      if (m == LEAVE_NODE) {
        const float nf = b.scaled_frames(g.length_ms);
        if (ca_ramp_spec* s = b.next_for_build()) {
          Builder::setup(s, f2ai(b.freq1_robot, g.sample_rate), f2ai(b.freq2_robot, g.sample_rate), nf, g.phase_ratio1 * nf, g.itp, g.sample_rate);
        }
      }
    });
    const int n1 = mc.emplace([](Move) {});
    mc.transition(n0, n1, 1.f);
  }
  // do_initialize, :946-980 (sweep: start node 0, no pre-tries, one normalised step, no additional tries)
  const int start = g.mode == 2 ? 0 : g.start_node, pre = g.mode == 2 ? 0 : g.pre_tries, path = g.mode == 2 ? 1 : g.min_path_length,
            extra = g.mode == 2 ? 0 : g.additional_tries;
  if (start < 0 || start >= static_cast<int>(mc.on.size()) || pre < 0 || path < 0 || extra < 0) return CA_ERR_INVALID;
  mc.cur = start;
  for (int i = 0; i < pre; ++i) mc.step(b.rng.unit(), true, false);
  for (int i = 0; i < path; ++i) mc.step(b.rng.unit(), true, true);
  for (int i = 0; i < extra; ++i) mc.step(b.rng.unit(), false, true);
  const int n = static_cast<int>(std::min<size_t>(b.specs.size(), static_cast<size_t>(capacity)));
  for (int i = 0; i < n; ++i) out[i] = b.specs[i];
  return n;
}

// Fourier series of saw / square / triangle over the angle a in [0,2) (units of pi), include/sound.functions.h:86-139:
//   saw(a)      = a on [0,1], a - 2 on (1,2]                    = (2/pi) sum_h (-1)^(h+1)/h sin(h pi a)
//   square(a)   = -1 on (0.5,1.5), +1 elsewhere = sgn cos(pi a)  = (4/pi) sum_{h odd} (-1)^((h-1)/2)/h cos(h pi a)
//   triangle(a) = 0 -> 1 -> -1 -> 0 through 0.5, 1.5            = (8/pi^2) sum_{h odd} (-1)^((h-1)/2)/h^2 sin(h pi a)
// as sum_h volume sin(h pi a + pi phase): a negative coefficient is a phase of 1, a cosine a phase of 1/2.
int waveform_harmonics(int32_t waveform, int32_t n, float* volume, float* phase) {
  if (!volume || !phase || n < 0 || waveform < CA_WAVE_SAW || waveform > CA_WAVE_TRIANGLE) return CA_ERR_INVALID;
  const double pi = 3.14159265358979323846;
  for (int i = 0; i < n; ++i) {
    const int h = i + 1;
    double c = 0, ph = 0;
    if (waveform == CA_WAVE_SAW) {
      c = 2.0 / (pi * h);
      ph = (h % 2) ? 0.0 : 1.0;
    } else if (h % 2) {
      const bool neg = ((h - 1) / 2) % 2;
      if (waveform == CA_WAVE_SQUARE) {
        c = 4.0 / (pi * h);
        ph = neg ? 1.5 : 0.5;
      } else {
        c = 8.0 / (pi * pi * h * h);
        ph = neg ? 1.0 : 0.0;
      }
    }
    volume[i] = static_cast<float>(c);
    phase[i] = static_cast<float>(ph);
  }
  return CA_OK;
}

}  // namespace ca
