// adapters_test.cpp — drives the C++ host adapters (cpp.audio_b200/host/*.hpp) the way reference code drives its own
// PeriodicFFT / ComputeFunc seams, and checks them against the CPU oracle (TEST INFRASTRUCTURE, oracle/oracle.hpp).
// Built and run by tests/test_gpu_cpp_adapters.py on the GPU box. Prints "OK ..." lines; exit code 0 on success.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../cpp.audio_b200/host/cuda_periodic_fft.hpp"
#include "../../cpp.audio_b200/host/cuda_resynth_synth.hpp"
#include "../../cpp.audio_b200/host/cuda_sine_bank.hpp"
#include "../../oracle/oracle.hpp"

extern "C" void orc_oscbank_render(int64_t n_voices, int n_partials, int64_t n_samples, int sample_rate, const float* f0_hz,
                                   const float* volume, const float* start_angle, const int32_t* key_delay,
                                   const int64_t* release_at, const float* partial_volume, const float* partial_phase,
                                   const orc::Ahdsr* env, int64_t n_env, const float* pan, int mix, double* out, int threads);

static int fails = 0;
#define EXPECT(cond, ...)            \
  do {                               \
    if (!(cond)) {                   \
      std::printf("FAIL: " __VA_ARGS__); \
      std::printf("\n");             \
      ++fails;                       \
    }                                \
  } while (0)

static std::vector<float> harmonic_tone(int n, double f0, int sr) {
  std::vector<float> x(n);
  double peak = 0;
  std::vector<double> t(n);
  for (int i = 0; i < n; ++i) {
    double s = 0;
    for (int h = 1; h <= 12; ++h) s += (1.0 / h) * std::sin(2 * orc::kPi * f0 * h * i / sr + 0.37 * h);
    t[i] = s;
    peak = std::max(peak, std::fabs(s));
  }
  for (int i = 0; i < n; ++i) x[i] = static_cast<float>(0.9 * t[i] / peak);
  return x;
}

static void test_periodic_fft(ca_ctx* ctx) {
  using namespace ca_host;
  int const sr = 48000, W = 4096, hop = 256, n = 48000;
  auto x = harmonic_tone(n, 220.0, sr);

  // oracle: the reference structure, one sample at a time
  std::vector<std::vector<double>> ref_sq;
  std::vector<std::vector<orc::FreqMag<double>>> ref_pk;
  orc::PeriodicFft<orc::SqMagOp<double>> opf(orc::Window::Gaussian, 4.0);
  std::vector<orc::FreqMag<double>> tmp;
  opf.set_lambdas([&] { return W; }, [&] { return hop; }, [&](int, const std::vector<double>& sq) {
    ref_sq.push_back(sq);
    orc::extract_local_max_freqs_mags<double>(sr, sq.data(), 4096, -80.0, tmp);
    ref_pk.push_back(tmp);
  });
  for (float s : x) opf.feed(s);

  // adapter: same calls as reference code (setLambdas + feed), batched on the GPU
  AnalysisConfig cfg;
  cfg.sample_rate = sr;
  PeriodicFFT<SqMagFftOperation<Window::Gaussian, double>> pf(1 << 14, ctx, cfg);
  pf.set_batch_frames(50);
  size_t frame = 0;
  double max_err = 0, max_ferr = 0;
  bool bins_ok = true;
  pf.setLambdas([&] { return W; }, [&] { return hop; }, [&](int const stride, FrequenciesSqMag<double> const& r) {
    EXPECT(stride == hop, "stride %d", stride);
    if (frame >= ref_sq.size()) { ++frame; return; }
    for (size_t k = 0; k < r.frequencies_sqmag.size(); ++k)
      max_err = std::max(max_err, std::fabs(std::sqrt(r.frequencies_sqmag[k]) - std::sqrt(ref_sq[frame][k])));
    auto const& pk = pf.last_peaks();
    if (pk.size() != ref_pk[frame].size()) bins_ok = false;
    else
      for (size_t i = 0; i < pk.size(); ++i) {
        max_ferr = std::max(max_ferr, std::fabs(pk[i].freq - ref_pk[frame][i].freq) / ref_pk[frame][i].freq);
        if (std::lround(pk[i].freq * 4096.0 / sr) != ref_pk[frame][i].bin && std::fabs(pk[i].freq * 4096.0 / sr - ref_pk[frame][i].bin) > 0.51) bins_ok = false;
      }
    ++frame;
  });
  for (float s : x) pf.feed(s);
  pf.flush();
  EXPECT(frame == ref_sq.size(), "frames %zu vs %zu", frame, ref_sq.size());
  EXPECT(max_err < 1e-5, "|X| error %g", max_err);
  EXPECT(bins_ok, "peak lists differ");
  EXPECT(max_ferr < 1e-5, "peak frequency error %g", max_ferr);
  std::printf("OK periodic_fft frames=%zu max|X|err=%.3g max rel freq err=%.3g\n", frame, max_err, max_ferr);

  // the per-frame operator concept (init_data / operator() / get_results)
  SqMagFftOperation<Window::Gaussian, double> op(1 << 14, ctx, cfg);
  EXPECT(op.init_data(W) == true, "first init_data must report a change");
  EXPECT(op.init_data(W) == false, "second init_data must not");
  std::vector<double> frame0(x.begin(), x.begin() + W);
  op(frame0.begin(), frame0.end());
  double e2 = 0;
  for (size_t k = 0; k < ref_sq[0].size(); ++k)
    e2 = std::max(e2, std::fabs(std::sqrt(op.get_results().frequencies_sqmag[k]) - std::sqrt(ref_sq[0][k])));
  EXPECT(op.get_results().get_fft_length() == 4096, "fft length");
  EXPECT(e2 < 1e-5, "operator() error %g", e2);
  std::printf("OK sqmag_operator max|X|err=%.3g\n", e2);
}

// PeriodicFFT<FftOperation<double>>, the vocoder carrier's analysis (source/rt.resynth.lib.vocoder.cpp:544): frames of 2 * stride
// samples every stride, rectangular window, zero padded to the FFT length, complex bins.
static void test_fft_operation(ca_ctx* ctx) {
  using namespace ca_host;
  int const W = 482, hop = 241, n = 20000;  // the vocoder's default stride at 48 kHz -> 512-point frames
  auto x = harmonic_tone(n, 330.0, 48000);
  std::vector<std::vector<std::complex<double>>> ref;
  orc::PeriodicFft<orc::FftOp<double>> opf;
  opf.set_lambdas([&] { return W; }, [&] { return hop; }, [&](int, const std::vector<std::complex<double>>& b) { ref.push_back(b); });
  for (float s : x) opf.feed(s);
  PeriodicFFT<FftOperation<double>> pf(1 << 10, ctx);
  pf.set_batch_frames(16);
  size_t frame = 0;
  double err = 0, peak = 0;
  pf.setLambdas([&] { return W; }, [&] { return hop; }, [&](int const stride, RealFBins<double> const& r) {
    EXPECT(stride == hop && r.get_fft_length() == 512 && r.bins.size() == 257, "shape");
    if (frame < ref.size())
      for (size_t k = 0; k < r.bins.size(); ++k) {
        err = std::max(err, std::abs(r.bins[k] - ref[frame][k]));
        peak = std::max(peak, std::abs(ref[frame][k]));
      }
    ++frame;
  });
  for (float s : x) pf.feed(s);
  pf.flush();
  EXPECT(frame == ref.size() && frame > 70, "frames %zu vs %zu", frame, ref.size());
  EXPECT(err < 1e-5 * peak, "bins error %g (peak %g)", err, peak);   // unscaled transform: relative to the largest bin
  FftOperation<double> op(1 << 10, ctx);
  op.init_data(W);
  std::vector<double> f0(x.begin(), x.begin() + W);
  op(f0.begin(), f0.end());
  double e2 = 0;
  for (size_t k = 0; k < ref[0].size(); ++k) e2 = std::max(e2, std::abs(op.get_results().bins[k] - ref[0][k]));
  EXPECT(e2 < 1e-5 * peak, "operator() error %g", e2);
  std::printf("OK fft_operation frames=%zu max bin err=%.3g (largest bin %.3g)\n", frame, err, peak);
}

static void test_sine_bank(ca_ctx* ctx) {
  using namespace ca_host;
  int const sr = 48000, n = 40000, np = 8;
  std::vector<harmonicProperties_t> har(np);
  for (int p = 0; p < np; ++p) har[p] = {0.1f * p, 1.0f / (p + 1)};
  ca_ahdsr env{200, 0, 50, 400, 0, 1500, 0, 0.6f};
  CudaSineBank<2> bank(ctx, sr, 8, har, env, 4096);
  struct N { float f, vel, pan, sa; int64_t on, off; };
  std::vector<N> notes = {{220.f, 1.0f, -0.5f, 0.0f, 0, 9000}, {330.5f, 0.7f, 0.5f, 0.25f, 1234, 30000}, {1760.f, 0.5f, 0.0f, -0.5f, 7000, -1}};
  for (size_t i = 0; i < notes.size(); ++i) {
    auto r = bank.onEvent(mkNoteOn(NoteId{static_cast<int64_t>(i)}, notes[i].f, notes[i].vel, notes[i].on), notes[i].pan, notes[i].sa);
    EXPECT(r == onEventResult::OK, "note on dropped");
    if (notes[i].off >= 0) bank.onEvent(mkNoteOff(NoteId{static_cast<int64_t>(i)}, notes[i].off));
  }
  std::vector<double> out(static_cast<size_t>(n) * 2, 0.0);
  for (int i = 0; i < n; i += 16) bank.compute(out.data() + static_cast<size_t>(i) * 2, std::min(16, n - i));  // 16-frame blocks like SimpleAudioOutContext

  // oracle
  size_t nv = notes.size();
  std::vector<float> f0(nv), vol(nv), sa(nv), pan(nv), pv(nv * np), pp(nv * np);
  std::vector<int32_t> kd(nv);
  std::vector<int64_t> ra(nv);
  for (size_t v = 0; v < nv; ++v) {
    f0[v] = notes[v].f; vol[v] = 0.1f * notes[v].vel; sa[v] = notes[v].sa; pan[v] = notes[v].pan;
    kd[v] = static_cast<int32_t>(notes[v].on); ra[v] = notes[v].off;
    for (int p = 0; p < np; ++p) { pv[v * np + p] = har[p].volume; pp[v * np + p] = har[p].phase; }
  }
  orc::Ahdsr oenv{env.attack, 0, env.hold, env.decay, 0, env.release, 0, env.sustain};
  std::vector<double> ref(static_cast<size_t>(n) * 2, 0.0);
  orc_oscbank_render(static_cast<int64_t>(nv), np, n, sr, f0.data(), vol.data(), sa.data(), kd.data(), ra.data(), pv.data(), pp.data(), &oenv, 1,
                     pan.data(), 1, ref.data(), 1);
  double err = 0, peak = 0;
  for (size_t i = 0; i < out.size(); ++i) { err = std::max(err, std::fabs(out[i] - ref[i])); peak = std::max(peak, std::fabs(ref[i])); }
  EXPECT(err < 1e-6, "bank error %g", err);
  EXPECT(peak > 0.05, "silent reference?");
  std::printf("OK sine_bank frames=%d max err=%.3g (peak %.3g)\n", n, err, peak);

  // voice pool exhaustion reports DROPPED_NOTE like ImplCRTP::onEvent (gen.crtp.h:412-414)
  CudaSineBank<1> small(ctx, sr, 2, har, env, 1024);
  EXPECT(small.onEvent(mkNoteOn(NoteId{1}, 100.f, 1.f, 0)) == onEventResult::OK, "1");
  EXPECT(small.onEvent(mkNoteOn(NoteId{2}, 200.f, 1.f, 0)) == onEventResult::OK, "2");
  EXPECT(small.onEvent(mkNoteOn(NoteId{3}, 300.f, 1.f, 0)) == onEventResult::DROPPED_NOTE, "3 must be dropped");
  std::printf("OK dropped_note\n");
}

// The rt.resynth synth through the reference's own calls: onEvent(sample_rate, Event, out, chans, maybeTime) + compute(buffer, n)
// in 16-frame blocks, events posted just before the block they belong to (the offline loop, rt.resynth.lib.cpp:1204-1234),
// against the oracle's VoicePool stepping the reference elements one sample at a time.
static void test_resynth_synth(ca_ctx* ctx) {
  using namespace ca_host::resynth;
  int const sr = 48000, n = 48000, stride = 256;
  struct E { int64_t at; int64_t id; int type; float f, v; };  // type: 0 on, 1 off, 2 change (orc::TrackEvent)
  std::vector<E> ev = {{0, 1, 0, 440.f, 0.9f},      {2560, 1, 2, 466.16f, 0.7f}, {2816, 1, 2, 493.88f, 0.8f}, {4096, 2, 0, 1320.f, 0.5f},
                       {8192, 1, 1, 0, 0},          {8704, 3, 0, 493.88f, 0.6f}, {12800, 2, 2, 1100.f, 0.9f}, {20480, 2, 1, 0, 0},
                       {20992, 4, 0, 3000.f, 0.4f}, {30000 / 16 * 16, 3, 1, 0, 0}};
  ResynthElementInitializer init;
  init.sample_rate = sr;
  init.stride = stride;
  init.stereo_spread = 0.7f;
  init.ahdsr = ca_ahdsr{300, 0, 100, 500, 0, 4000, 0, 0.7f};
  CudaResynthSynth<2> synth(ctx, 127, 16);
  synth.setSynchronousElementInitializer(init);
  NoOut out_stub;
  NoChans chans_stub;
  std::vector<double> out(static_cast<size_t>(n) * 2, 0.0);
  size_t next = 0;
  for (int i = 0; i < n; i += 16) {
    for (; next < ev.size() && ev[next].at <= i; ++next) {
      E const& e = ev[next];
      Event const x = e.type == 0 ? mkNoteOn(NoteId{e.id}, e.f, e.v) : e.type == 1 ? mkNoteOff(NoteId{e.id}) : mkNoteChange(NoteId{e.id}, e.v, e.f);
      EXPECT(synth.onEvent(sr, x, out_stub, chans_stub, std::nullopt) == onEventResult::OK, "event %zu dropped", next);
    }
    synth.compute(out.data() + static_cast<size_t>(i) * 2, std::min(16, n - i));
  }
  std::vector<orc::TrackEvent> oev;
  for (auto const& e : ev) oev.push_back({e.at, e.id, e.type, e.f, e.v, 0});
  orc::Ahdsr oenv{300, 0, 100, 500, 0, 4000, 0, 0.7f};
  std::vector<double> ref(static_cast<size_t>(n) * 2, 0.0);
  int64_t stats[2];
  orc::render_tracks<double>(oev.data(), static_cast<int64_t>(oev.size()), n, sr, stride, oenv, 254, 0.7f, false, ref.data(), stats);
  double err = 0, peak = 0;
  for (size_t i = 0; i < out.size(); ++i) { err = std::max(err, std::fabs(out[i] - ref[i])); peak = std::max(peak, std::fabs(ref[i])); }
  EXPECT(err < 1e-5, "resynth synth error %g", err);
  EXPECT(peak > 0.05 && stats[0] == 4 && stats[1] == 0, "reference: peak %g notes %lld", peak, static_cast<long long>(stats[0]));
  std::printf("OK resynth_synth frames=%d max err=%.3g (peak %.3g)\n", n, err, peak);

  // a timestamped event lands on its sample; allNotesOff releases what is on; a full pool reports DROPPED_NOTE
  CudaResynthSynth<1> mono(ctx, 1, 64);  // 2 channels
  mono.setSynchronousElementInitializer(init);
  EXPECT(mono.onEvent(sr, mkNoteOn(NoteId{1}, 500.f, 1.f), TimestampAndSource{1000000000ull * 1000 / sr, 0}) == onEventResult::OK, "timed note on");
  EXPECT(mono.onEvent(sr, mkNoteOn(NoteId{2}, 700.f, 1.f)) == onEventResult::OK, "second channel");
  EXPECT(mono.onEvent(sr, mkNoteOn(NoteId{3}, 900.f, 1.f)) == onEventResult::DROPPED_NOTE, "third note must be dropped");
  std::vector<double> m(4096, 0.0);
  for (int i = 0; i < 2048; i += 16) mono.compute(m.data() + i, 16);
  mono.allNotesOff();
  bool sounding = true;
  for (int i = 2048; i < 4096; i += 16) sounding = mono.compute(m.data() + i, 16);
  std::vector<orc::TrackEvent> mev = {{1000, 1, 0, 500.f, 1.f, 0}, {1000, 2, 0, 700.f, 1.f, 0}, {2048, 1, 1, 0, 0, 0}, {2048, 2, 1, 0, 0, 0}};
  std::vector<double> mref(4096, 0.0);
  orc::render_tracks<double>(mev.data(), 4, 4096, sr, stride, oenv, 2, 0.7f, true, mref.data(), stats);
  double merr = 0;
  for (size_t i = 0; i < m.size(); ++i) merr = std::max(merr, std::fabs(m[i] - mref[i]));
  EXPECT(merr < 1e-5, "timed events / allNotesOff error %g", merr);
  EXPECT(mono.dropped_notes() == 1, "dropped count");
  (void)sounding;
  std::printf("OK resynth_synth timed events, allNotesOff, dropped note: max err=%.3g\n", merr);
}

int main() {
  ca_ctx* ctx = nullptr;
  int rc = ca_init(0, &ctx);
  if (rc != CA_OK) {
    std::printf("FAIL: ca_init %d (a B200 is required; there is no CPU fallback)\n", rc);
    return 2;
  }
  test_periodic_fft(ctx);
  test_fft_operation(ctx);
  test_sine_bank(ctx);
  test_resynth_synth(ctx);
  ca_destroy(ctx);
  std::printf(fails ? "FAILED %d\n" : "ALL OK %d\n", fails);
  return fails ? 1 : 0;
}
