// Host-side run of the product's stateless pitch prefix (cpp.audio_b200/csrc/pitch.cuh, the source the GPU kernels compile
// too): reads one frame's peak list and the pitch options from stdin, prints the tuned list. No GPU is touched, so the
// CPU test suite can compare the product's shift / harmonize / autotune stages with the oracle (tests/test_pitch_prefix_host.py).
//   stdin:  P tcap  shift_pre harm_pre shift_post harm_post  autotune max_pitch tolerance root  n_table table...  cnt (bin off db)...
//   stdout: T, then "pitch vol volume ai min_change" per entry (17 significant digits)
#include <cstdio>
#include <vector>

#include "cpp_audio_b200.h"
#include "pitch.cuh"

int main() {
  ca_roundtrip_params prm{};
  int P = 0, tcap = 0, n_table = 0, cnt = 0;
  ca_pitch_options& o = prm.opts;
  double root = 0;
  if (scanf("%d %d %f %f %f %f %d %f %f %lf %d", &P, &tcap, &o.shift_pre_autotune, &o.harmonize_pre_autotune, &o.shift_post_autotune,
            &o.harmonize_post_autotune, &o.autotune, &o.autotune_max_pitch, &o.autotune_tolerance_pitches, &root, &n_table) != 11)
    return 2;
  std::vector<double> table(n_table > 0 ? n_table : 1);
  for (int i = 0; i < n_table; ++i)
    if (scanf("%lf", &table[i]) != 1) return 2;
  o.n_table = n_table;
  o.table = table.data();
  o.scale_root_pitch = root;
  if (scanf("%d", &cnt) != 1) return 2;
  std::vector<int32_t> bin(cnt > 0 ? cnt : 1);
  std::vector<float> off(bin.size()), db(bin.size());
  for (int i = 0; i < cnt; ++i)
    if (scanf("%d %f %f", &bin[i], &off[i], &db[i]) != 3) return 2;
  prm.stft.window = 4096;
  prm.stft.hop = 256;
  prm.stft.window_type = 1;
  prm.stft.sigma_factor = 4.f;
  prm.stft.sample_rate = 48000;
  prm.stft.max_peaks = tcap;
  prm.nearby_distance_tones = 0.4f;
  prm.min_volume = 1e-4f;
  prm.max_track_pitches = 1.f;
  prm.gain = 1.f;
  prm.env.sustain = 1.f;
  prm.max_channels = 254;
  const ca::PitchParams pp = ca::make_pitch_params(prm, 48000, 172, P);
  std::vector<ca::Tuned> row(tcap > P ? tcap : P);
  int cut = 0;
  const int T = ca::pitch_prefix(pp, bin.data(), off.data(), db.data(), cnt, row.data(), &cut);
  printf("%d\n", T);
  for (int i = 0; i < T; ++i) printf("%.17g %.17g %.9g %.17g %d\n", row[i].pitch, row[i].vol, row[i].volume, row[i].ai, row[i].min_change);
  return 0;
}
