// roundtrip.cu — analysis -> pitch pipeline -> resynthesis of whole streams (BASELINE config 4).
//
// Reference: the offline RtResynth loop (source/rt.resynth.lib.cpp:1185-1235). Per analysis frame, RtResynth::step
// (:1670-1759) runs the pitch pipeline of source/rt.resynth.lib.algo.cpp on the peak list and synthesize_sounds (:264-382)
// posts NoteOn / NoteChange / NoteOff to a pool of ResynthElement voices (:5-18); the voices are stepped one sample at a
// time (oscillator with a per-sample frequency glide, include/audioelement.h:2706-2817; volume low-pass :1195-1216;
// AHDSR envelope :666-1099; stereo pan include/stereo.h:18-27).
//
// Here:
//   1. K2 (stft.cu) extracts the peaks of every frame of every stream on the GPU.
//   2. Host threads (one stream each) run the frame-sequential pitch pipeline and note bookkeeping — control-rate code,
//      O(peaks) per frame — and emit, per frame interval ("tile" = hop samples) and per sounding note, a SEGMENT: the
//      closed-form description of that note over the tile. Every per-sample recurrence of the reference voice has a
//      closed form in the step index (the glide's time warp cur' = alpha*cur + beta is affine, the low-pass and the envelope
//      are first order / piecewise linear), so a segment is restartable at tile boundaries and tiles are independent.
//   3. tracks_kernel renders the tiles: one warp per tile, lanes = consecutive samples, phases and glide sums in FP64
//      (the reference runs in double and phase errors accumulate over the whole stream), the rest in FP32.
//
// Unpinned pieces (oracle/SPEC.md): S8 one-pole volume filter, S9 loudness ordering = raw volume, S10 start phase / pan
// from a hash of the note id instead of the external mersenne<> generator.
#include <math.h>

#include <stdio.h>
#include <stdlib.h>

#include <string.h>

#include <algorithm>
#include <chrono>
#include <climits>
#include <thread>
#include <vector>

#include "common.cuh"
#include "ctx.hpp"
#include "tracks.cuh"

namespace ca {

namespace {


// ---------------------------------------------------------------------------------------------------------------
// Host side: pitch pipeline + note bookkeeping of one stream -> segments
// ---------------------------------------------------------------------------------------------------------------

inline float hash_unit(uint64_t id, uint64_t salt) {  // SPEC S10
  uint64_t z = id * 0x9E3779B97F4A7C15ull + salt * 0xBF58476D1CE4E5B9ull + 0x94D049BB133111EBull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return static_cast<float>(static_cast<double>(z >> 11) * (2.0 / 9007199254740992.0) - 1.0);
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
};

struct StreamResult {
  std::vector<Segment> segs;
  std::vector<int64_t> tile_seg_begin;  // per tile (a frame interval longer than kMaxTile samples is cut into several tiles)
  std::vector<int64_t> tile_n0;         // first sample of the tile (stream-relative)
  std::vector<int32_t> tile_len;
  int64_t notes_started = 0, notes_dropped = 0;
};

inline int32_t min_change_for(double ai) {  // Enveloped::setAngleIncrements, include/audioelement.h:218-225
  const double period = ai == 0.0 ? 1.0 : 2.0 / fabs(ai);
  return static_cast<int32_t>(0.5f + 2.5 * period);
}

double env_pressed_value(const Note& n, int64_t c, float sustain) {
  if (c < 1) return 0.0;
  if (c <= n.attack) return static_cast<double>(c) / n.attack;
  c -= n.attack;
  if (c <= n.hold) return 1.0;
  c -= n.hold;
  const int32_t d = n.decay < 0 ? 0 : n.decay;
  if (c <= d) return 1.0 + (static_cast<double>(sustain) - 1.0) * static_cast<double>(c) / d;
  return sustain;
}

// peak_stride: entries per frame in the (compacted) peak arrays
void run_stream(const ca_roundtrip_params& prm, int64_t n_samples, int64_t frames, const int32_t* peak_bin,
                const float* peak_off, const float* peak_db, const int32_t* peak_count, int peak_stride, StreamResult& res) {
  const int W = prm.stft.window, hop = prm.stft.hop, sr = prm.stft.sample_rate, P = peak_stride;
  const int n_fft = ca_fft_length_for(W, 1);
  const double bin_hz = static_cast<double>(sr) / n_fft;
  // AHDSREnvelopeBase::setAHDSR (include/audioelement.h:923-952)
  const int32_t min_dt = sr / 1000;
  const bool has_decay = prm.env.sustain < 0.999999;
  const float sustain = has_decay ? std::min(std::max(prm.env.sustain, 0.f), 1.f) : 1.f;
  const int32_t A = std::max(prm.env.attack, min_dt), H = std::max(prm.env.hold, 0);
  const int32_t D = has_decay ? std::max(prm.env.decay, min_dt) : 0, R = std::max(prm.env.release, min_dt);
  const double cap = 2.0 / hop;  // setMaxFilterIncrement(2/stride), source/rt.resynth.lib.cpp:106

  std::vector<Note> notes;       // sounding notes (tracked or releasing)
  std::vector<int> played;       // indices into notes of tracked notes, sorted by pitch
  struct PV { double pitch, vol; };
  std::vector<PV> pv, reduced, tuned;
  std::vector<int> change, order, next_played;
  std::vector<char> keep;

  res.tile_seg_begin.clear();
  res.tile_n0.clear();
  res.tile_len.clear();
  res.tile_seg_begin.reserve(frames + 1);
  for (int64_t k = 0; k < frames; ++k) {
    const int64_t n_k = static_cast<int64_t>(W) - 1 + k * hop;  // events of frame k act from this sample on
    const int len = static_cast<int>(std::min<int64_t>(hop, n_samples - n_k));
    // ---------------- pitch pipeline (source/rt.resynth.lib.algo.cpp) ----------------
    pv.clear();
    const int cnt = std::min(peak_count[k], P);
    for (int i = 0; i < cnt; ++i) {
      const double freq = (static_cast<double>(peak_bin[k * P + i]) + static_cast<double>(peak_off[k * P + i])) * bin_hz;
      if (freq > 0) pv.push_back({69. + 12. * log2(freq / 440.f), pow(10.0, static_cast<double>(peak_db[k * P + i]) / 20.0)});
    }
    reduced.clear();
    {
      bool open = false;
      double mn = 0, mx = 0, sv = 0, spv = 0;
      auto close = [&] {
        if (open && !(sv < prm.min_volume)) reduced.push_back({spv / sv, sv});
        open = false;
      };
      for (auto& q : pv) {
        if (open && std::max({mn, mx, q.pitch}) - std::min({mn, mx, q.pitch}) > prm.nearby_distance_tones) close();
        if (!open) {
          open = true;
          mn = mx = q.pitch;
          sv = spv = 0;
        } else {
          mn = std::min(mn, q.pitch);
          mx = std::max(mx, q.pitch);
        }
        sv += q.vol;
        spv += q.pitch * q.vol;
      }
      close();
    }
    tuned.clear();
    for (auto& q : reduced) {
      if (!tuned.empty() && fabs(tuned.back().pitch - q.pitch) < 0.0001f) tuned.back().vol += q.vol;
      else tuned.push_back(q);
    }
    change.assign(tuned.size(), -1);
    keep.assign(played.size(), 0);
    {
      size_t it = 0;
      for (size_t i = 0; i < tuned.size(); ++i) {
        for (; it < played.size(); ++it) {
          const double pp = notes[played[it]].pitch;
          if (pp < tuned[i].pitch - prm.max_track_pitches) continue;
          if (pp <= tuned[i].pitch + prm.max_track_pitches) {
            change[i] = static_cast<int>(it);
            keep[it] = 1;
            ++it;
          }
          break;
        }
      }
    }
    order.resize(tuned.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = static_cast<int>(i);
    std::stable_sort(order.begin(), order.end(),
                     [&](int x, int y) { return static_cast<float>(tuned[x].vol) > static_cast<float>(tuned[y].vol); });
    // ---------------- events (synthesize_sounds, source/rt.resynth.lib.cpp:264-382) ----------------
    for (size_t i = 0; i < played.size(); ++i)
      if (!keep[i]) {  // NoteOff -> onKeyReleased(0), include/audioelement.h:769-781, :836-841
        Note& n = notes[played[i]];
        n.tracked = false;
        n.n_off = n_k;
        n.top = env_pressed_value(n, n_k - n.n_on, sustain);
        n.release = std::max(n.next_min_change, R);
      }
    next_played.clear();
    for (size_t i = 0; i < played.size(); ++i)
      if (keep[i]) next_played.push_back(played[i]);
    for (int idx : order) {
      const float new_freq = static_cast<float>(440.f * pow(pow(2., 1. / 12.), tuned[idx].pitch - 69.));
      const float volume = static_cast<float>(prm.gain * tuned[idx].vol / 0.1f);
      const double ai = static_cast<double>(2 * new_freq / sr);  // freq_to_angle_increment in float
      if (change[idx] >= 0) {                                      // NoteChange, include/gen.crtp.h:595-618
        Note& n = notes[played[change[idx]]];
        n.pitch = tuned[idx].pitch;
        n.vol_x = static_cast<double>(0.1f * volume);
        n.next_min_change = min_change_for(ai);
        n.glide.start(n.f_last, ai, hop);                          // InterpolatedFreq::setAngleIncrements, :2752-2762
      } else {
        if (volume <= 0) continue;
        int sounding = 0;
        for (auto& n : notes) sounding += 1;  // every entry of `notes` still holds a channel
        if (sounding >= prm.max_channels) {
          ++res.notes_dropped;
          ++res.notes_started;
          continue;
        }
        ++res.notes_started;
        Note n;
        n.id = res.notes_started;
        n.pitch = tuned[idx].pitch;
        n.vol_x = n.vol_y = static_cast<double>(0.1f * volume);
        n.n_on = n_k;
        n.next_min_change = min_change_for(ai);
        n.attack = std::max(n.next_min_change, A);
        n.hold = H;
        n.decay = -1;
        n.glide.start(ai, ai, hop);
        n.f_last = ai;
        double a0 = static_cast<double>(hash_unit(static_cast<uint64_t>(n.id), 2));
        a0 -= 2.0 * floor(0.5 * a0);
        n.phase = a0;
        double pan = static_cast<double>(prm.stereo_spread * hash_unit(static_cast<uint64_t>(n.id), 1));
        pan = std::min(1.0, std::max(-1.0, pan));
        n.gl = static_cast<float>(cos(0.25 * 3.14159265358979323846 * (pan + 1.0)));
        n.gr = static_cast<float>(sin(0.25 * 3.14159265358979323846 * (pan + 1.0)));
        notes.push_back(n);
        next_played.push_back(static_cast<int>(notes.size()) - 1);
      }
    }
    std::stable_sort(next_played.begin(), next_played.end(), [&](int x, int y) { return notes[x].pitch < notes[y].pitch; });
    played.swap(next_played);

    // ---------------- one segment per sounding note and tile, then advance the note states ----------------
    // (the interval up to the next frame is one tile, or several when the hop exceeds kMaxTile samples)
    const int len_frame = len;
    for (int off = 0; off < std::max(len_frame, 0) || off == 0; off += kMaxTile) {
      const int len = std::min(kMaxTile, len_frame - off);
      const int64_t n_k_tile = n_k + off;
      res.tile_seg_begin.push_back(static_cast<int64_t>(res.segs.size()));
      res.tile_n0.push_back(n_k_tile);
      res.tile_len.push_back(std::max(len, 0));
      if (len <= 0) break;
      for (auto& n : notes) {
        Segment g{};
        const int64_t i0 = n.glide.i;
        g.phase_start = n.phase;
        g.from = n.glide.from;
        g.to = n.glide.to;
        g.K = n.glide.K;
        g.L = n.glide.L;
        g.c1 = n.glide.c1;
        g.P_m1 = n.glide.P_m1;
        g.m = n.glide.m;
        g.i0 = static_cast<int32_t>(std::min<int64_t>(i0, INT_MAX / 2));
        g.P_before = n.glide.P(i0 - 1);
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
        const double adv_end = n.glide.P(i0 + len - 1) - n.glide.P(i0 - 1);
        if (std::min(fa, fb) >= cap) {
          g.regime = 0;
          mv_end = cap * len;
        } else if (std::max(fa, fb) < cap) {
          g.regime = 1;
          mv_end = adv_end;
        } else {  // the glide crosses the cap inside the tile: find the switch by walking the (at most hop) steps
          int istar = 0;
          const bool f_first = fa < cap;
          while (istar < len && ((n.glide.f_at(i0 + istar) < cap) == f_first)) ++istar;
          g.istar = istar;
          const double adv_star = n.glide.P(i0 + istar - 1) - n.glide.P(i0 - 1);
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
          n.decay = D == 0 ? 0 : std::max(n.next_min_change, D);
        g.attack = n.attack;
        g.hold = n.hold;
        g.decay = n.decay < 0 ? 0 : n.decay;
        g.sustain = sustain;
        g.release = n.release > 0 ? n.release : 1;
        g.rel_base = n.n_off < 0 ? INT_MAX : static_cast<int32_t>(std::max<int64_t>(n.n_off - n_k_tile, -(1ll << 30)));
        g.top = static_cast<float>(n.top);
        g.gl = n.gl;
        g.gr = n.gr;
        res.segs.push_back(g);
        // advance
        n.phase += adv_end;
        n.phase -= 2.0 * floor(0.5 * n.phase);
        n.f_last = fb;
        n.glide.i = i0 + len;
        n.vol_y = n.vol_x + (n.vol_y - n.vol_x) * exp(-3.14159265358979323846 * mv_end);
      }
    }
    if (len_frame > 0) {
      // a released note leaves the pool once its envelope reached Done2: release steps + 16 more (:738-749)
      const int64_t n_next = n_k + len_frame;
      size_t w = 0;
      std::vector<int> remap(notes.size(), -1);
      for (size_t i = 0; i < notes.size(); ++i) {
        const Note& n = notes[i];
        const bool done = n.n_off >= 0 && n_next >= n.n_off + n.release + 17;
        if (!done) {
          remap[i] = static_cast<int>(w);
          if (w != i) notes[w] = notes[i];
          ++w;
        }
      }
      notes.resize(w);
      for (auto& pi : played) pi = remap[pi];
    }
  }
  res.tile_seg_begin.push_back(static_cast<int64_t>(res.segs.size()));
}

}  // namespace

int roundtrip_host(ca_ctx* ctx, const ca_roundtrip_params* params, const float* h_x, int64_t n_samples, int64_t n_streams,
                   float* h_out, int64_t* h_stats) {
  if (!params || !h_x || !h_out || n_samples < 0 || n_streams < 0)
    return fail(ctx, CA_ERR_INVALID, "ca_resynth_roundtrip_host: null argument or negative size");
  ca_roundtrip_params prm = *params;
  if (prm.max_channels <= 0) prm.max_channels = 254;
  if (prm.stft.max_peaks <= 0) prm.stft.max_peaks = 200;
  int rc = stft_plan(ctx, &prm.stft);
  if (rc) return rc;
  const int W = prm.stft.window, hop = prm.stft.hop, P = prm.stft.max_peaks;
  const int64_t frames = ca_stft_num_frames(n_samples, W, hop);
  const size_t total = static_cast<size_t>(n_streams) * n_samples;
  if (total == 0) return CA_OK;
  const size_t rows = static_cast<size_t>(frames) * n_streams;
  static const bool timing = getenv("CA_RT_TIMING") != nullptr;  // stage times on stderr (profiles/time_roundtrip.py)
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "roundtrip: %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };

  // ---- 1. peaks on the GPU (sub-bin offsets instead of Hz so that the host forms k+p in double) ----
  // The input goes up in one copy, K2 runs on the whole batch, and only the USED part of the peak lists comes back: the
  // capacity per frame is max_peaks (200 in the reference), a frame of music holds a few dozen. Host-side buffers are
  // pinned and owned by the context (no page faults, DMA at PCIe speed, no allocation after the first call).
  cudaStream_t st = ctx->own_stream;
  const size_t out_bytes = sizeof(float) * 2 * total;
  rc = ensure_scratch(ctx, 0, sizeof(float) * total);
  if (!rc) rc = ensure_scratch(ctx, 7, out_bytes);
  for (int i = 2; i <= 4 && !rc; ++i) rc = ensure_scratch(ctx, i, std::max<size_t>(16, sizeof(float) * rows * P));
  if (!rc) rc = ensure_scratch(ctx, 5, std::max<size_t>(16, sizeof(int32_t) * rows));
  if (!rc) rc = ensure_pinned(ctx, 0, std::max<size_t>(16, sizeof(int32_t) * rows));
  if (rc) return rc;
  int32_t* h_cnt = static_cast<int32_t*>(ctx->h_pinned[0]);
  int cmax = 0;
  if (rows) {
    float* d_x = static_cast<float*>(ctx->d_scratch[0]);
    CA_CUDA_OK(ctx, cudaMemcpyAsync(d_x, h_x, sizeof(float) * total, cudaMemcpyHostToDevice, st));
    ctx->stft_freq_as_offset = true;
    rc = stft_analyze(ctx, d_x, n_samples, n_streams, n_samples, nullptr, 0, static_cast<int32_t*>(ctx->d_scratch[2]),
                      static_cast<float*>(ctx->d_scratch[3]), static_cast<float*>(ctx->d_scratch[4]),
                      static_cast<int32_t*>(ctx->d_scratch[5]), st);
    ctx->stft_freq_as_offset = false;
    if (rc) return rc;
    CA_CUDA_OK(ctx, cudaMemcpyAsync(h_cnt, ctx->d_scratch[5], sizeof(int32_t) * rows, cudaMemcpyDeviceToHost, st));
    CA_CUDA_OK(ctx, cudaStreamSynchronize(st));
    for (size_t r = 0; r < rows; ++r) cmax = std::max(cmax, h_cnt[r]);
    cmax = std::min(cmax, P);
  }
  const int PS = std::max(cmax, 1);  // compacted entries per frame
  for (int i = 1; i <= 3 && !rc; ++i) rc = ensure_pinned(ctx, i, std::max<size_t>(16, sizeof(float) * rows * PS));
  if (rc) return rc;
  if (rows && cmax > 0)
    for (int i = 0; i < 3; ++i)
      CA_CUDA_OK(ctx, cudaMemcpy2DAsync(ctx->h_pinned[1 + i], sizeof(float) * PS, ctx->d_scratch[2 + i], sizeof(float) * P,
                                        sizeof(float) * PS, rows, cudaMemcpyDeviceToHost, st));
  CA_CUDA_OK(ctx, cudaMemsetAsync(ctx->d_scratch[7], 0, out_bytes, st));  // samples before the first frame are silent
  CA_CUDA_OK(ctx, cudaStreamSynchronize(st));
  const int32_t* pbin = static_cast<const int32_t*>(ctx->h_pinned[1]);
  const float* poff = static_cast<const float*>(ctx->h_pinned[2]);
  const float* pdb = static_cast<const float*>(ctx->h_pinned[3]);
  lap("analysis (H2D, K2, D2H peaks)");

  // ---- 2. pitch pipeline on host threads, one stream at a time per thread ----
  std::vector<StreamResult> results(n_streams);
  int nt = prm.host_threads > 0 ? prm.host_threads : static_cast<int>(std::thread::hardware_concurrency());
  nt = std::max(1, std::min<int>(nt, static_cast<int>(std::max<int64_t>(n_streams, 1))));
  auto parallel = [&](auto&& body) {
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t)
      pool.emplace_back([&, t] {
        for (int64_t s = t; s < n_streams; s += nt) body(s);
      });
    for (auto& th : pool) th.join();
  };
  parallel([&](int64_t s) {
    run_stream(prm, n_samples, frames, pbin + s * frames * PS, poff + s * frames * PS, pdb + s * frames * PS, h_cnt + s * frames, PS,
               results[s]);
  });
  lap("pitch pipeline (host)");

  // ---- 3. flatten (in parallel, into pinned memory) and render ----
  std::vector<int64_t> tile_base(n_streams + 1, 0);
  for (int64_t s = 0; s < n_streams; ++s) tile_base[s + 1] = tile_base[s] + static_cast<int64_t>(results[s].tile_len.size());
  const int64_t n_tiles = tile_base[n_streams];
  std::vector<int64_t> seg_base(n_streams + 1, 0);
  for (int64_t s = 0; s < n_streams; ++s) seg_base[s + 1] = seg_base[s] + static_cast<int64_t>(results[s].segs.size());
  const int64_t n_segs = seg_base[n_streams];
  const size_t b_begin = sizeof(int64_t) * (n_tiles + 1), b_s0 = sizeof(int64_t) * std::max<int64_t>(n_tiles, 1),
               b_len = sizeof(int32_t) * std::max<int64_t>(n_tiles, 1);
  rc = ensure_pinned(ctx, 4, std::max<size_t>(16, sizeof(Segment) * n_segs));
  if (!rc) rc = ensure_pinned(ctx, 5, b_begin + b_s0 + b_len);
  if (rc) return rc;
  Segment* h_segs = static_cast<Segment*>(ctx->h_pinned[4]);
  int64_t* tile_begin = static_cast<int64_t*>(ctx->h_pinned[5]);
  int64_t* tile_s0 = tile_begin + (n_tiles + 1);
  int32_t* tile_len = reinterpret_cast<int32_t*>(tile_s0 + std::max<int64_t>(n_tiles, 1));
  parallel([&](int64_t s) {
    const int64_t base = seg_base[s], tb = tile_base[s];
    const StreamResult& r = results[s];
    for (size_t k = 0; k < r.tile_len.size(); ++k) {
      tile_begin[tb + k] = base + r.tile_seg_begin[k];
      tile_s0[tb + k] = s * n_samples + r.tile_n0[k];
      tile_len[tb + k] = r.tile_len[k];
    }
    if (!r.segs.empty()) memcpy(h_segs + base, r.segs.data(), sizeof(Segment) * r.segs.size());
    if (h_stats) {
      h_stats[2 * s] = r.notes_started;
      h_stats[2 * s + 1] = r.notes_dropped;
    }
  });
  tile_begin[n_tiles] = n_segs;
  lap("flatten segments");

  rc = ensure_scratch(ctx, 6, std::max<size_t>(16, sizeof(Segment) * n_segs));
  if (!rc) rc = ensure_scratch(ctx, 1, b_begin + b_s0 + b_len);
  if (rc) return rc;
  if (n_tiles > 0 && n_segs > 0) {
    CA_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_scratch[6], h_segs, sizeof(Segment) * n_segs, cudaMemcpyHostToDevice, st));
    CA_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_scratch[1], tile_begin, b_begin + b_s0 + b_len, cudaMemcpyHostToDevice, st));
    const int64_t* d_begin = static_cast<const int64_t*>(ctx->d_scratch[1]);
    const int64_t* d_s0 = d_begin + (n_tiles + 1);
    const int32_t* d_len = reinterpret_cast<const int32_t*>(d_s0 + std::max<int64_t>(n_tiles, 1));
    TileArgs ta{static_cast<const Segment*>(ctx->d_scratch[6]), d_begin, d_s0, d_len, n_tiles, static_cast<float*>(ctx->d_scratch[7]), 0};
    tracks_kernel<false><<<static_cast<unsigned>((n_tiles + 3) / 4), 128, 0, st>>>(ta);
    CA_CUDA_OK(ctx, cudaGetLastError());
    ctx->launches += 1;
  }
  if (timing) {
    CA_CUDA_OK(ctx, cudaStreamSynchronize(st));
    lap("H2D segments + tracks kernel");
  }
  CA_CUDA_OK(ctx, cudaMemcpyAsync(h_out, ctx->d_scratch[7], out_bytes, cudaMemcpyDeviceToHost, st));
  CA_CUDA_OK(ctx, cudaStreamSynchronize(st));
  lap("D2H output");
  return CA_OK;
}

}  // namespace ca
