// ramp.cu — gen.voice ramp voices: a sine oscillator whose frequency is driven by LogRamp
// (include/audioelement.h:2464-2572: the ramp position advances by C*f per sample, f linear in the position, and the ramp
// bounces between its bounds) inside Enveloped<FreqCtrl_<Sine, LogRamp>, AHDSR> with a constant volume — the inner
// element of voice::Impl_ / SoundEngine (include/soundengine.h:245-257, 412-464; include/gen.voice.h:1130-1189), whose
// up-to-three cross-faded ramps per voice are a sum of such elements with staggered key-press / key-release times.
//
// Not a per-sample loop: inside one leg of the ramp the position obeys cur' = (1+lambda)*cur + beta, so the emitted
// increment and its running sum (the oscillator phase) are exponentials of the step index (tracks.cuh). The host cuts
// every voice into legs and the legs into tiles of <= 1024 samples (a few transcendental calls per tile), the GPU
// renders the tiles with one warp each, phases in FP64.
#include <math.h>

#include <algorithm>
#include <climits>
#include <thread>
#include <vector>

#include "common.cuh"
#include "ctx.hpp"
#include "tracks.cuh"

namespace ca {

namespace {

inline int32_t min_change_for_ramp(double ai) {  // Enveloped::setAngleIncrements, include/audioelement.h:218-225
  const double period = ai == 0.0 ? 1.0 : 2.0 / fabs(ai);
  return static_cast<int32_t>(0.5f + 2.5 * period);
}

struct VoiceTiles {
  std::vector<Segment> segs;
  std::vector<int64_t> s0;
  std::vector<int32_t> len;
};

void build_voice(const ca_ramp_desc& d, int64_t v, int64_t row0, VoiceTiles& out) {
  const int sr = d.sample_rate;
  const double from = static_cast<double>(2 * d.from_hz[v] / sr), to = static_cast<double>(2 * d.to_hz[v] / sr);  // float, :58-62
  const double dur = static_cast<double>(d.duration[v]);
  const double c_start = d.start_sample ? std::max(0.0, static_cast<double>(d.start_sample[v])) : 0.0;
  const int64_t kd = d.key_delay ? d.key_delay[v] : 0;
  const int64_t rel = d.release_at ? d.release_at[v] : -1;
  // AHDSREnvelopeBase::setAHDSR (include/audioelement.h:923-952) with the floors of Enveloped (:216-223, :1002-1004, :1066-1084)
  const int32_t min_dt = sr / 1000;
  const bool has_decay = d.env.sustain < 0.999999;
  const float sustain = has_decay ? std::min(std::max(d.env.sustain, 0.f), 1.f) : 1.f;
  const int32_t mc = min_change_for_ramp(from);
  const int32_t A = std::max(mc, std::max(d.env.attack, min_dt)), H = std::max(d.env.hold, 0);
  const int32_t D = has_decay ? std::max(mc, std::max(d.env.decay, min_dt)) : 0;
  const int32_t R = std::max(mc, std::max(d.env.release, min_dt));
  auto pressed = [&](int64_t c) -> double {
    if (c < 1) return 0.0;
    if (c <= A) return static_cast<double>(c) / A;
    c -= A;
    if (c <= H) return 1.0;
    c -= H;
    if (c <= D) return 1.0 + (static_cast<double>(sustain) - 1.0) * static_cast<double>(c) / D;
    return sustain;
  };
  if (rel >= 0 && rel - kd <= 0) return;  // released before the key delay elapsed: the note never sounds (:774-777)
  const double top = rel >= 0 ? pressed(rel - kd) : 0.0;
  const int64_t n_end = rel >= 0 ? std::min<int64_t>(d.n_samples, rel + R) : d.n_samples;  // silent once the release is over

  double phase = d.start_angle ? static_cast<double>(d.start_angle[v]) : 0.0;
  phase -= 2.0 * floor(0.5 * phase);
  double lf = from, lt = to, cs = c_start;
  int64_t n = 0;
  Leg leg;
  while (n < n_end) {
    leg.start(lf, lt, dur, cs);
    int64_t i = 0;
    while (i < leg.m && n < n_end) {
      const int len = static_cast<int>(std::min<int64_t>({static_cast<int64_t>(kMaxTile), leg.m - i, n_end - n}));
      Segment g{};
      g.phase_start = phase;
      const double P_before = leg.steady ? 0.0 : leg.P(i - 1);
      if (leg.steady) {  // re-based constant increment: P_j = (j+1)*to
        g.i0 = 0;
        g.m = 0;
        g.to = leg.from1;
      } else {
        g.from = leg.from1;
        g.K = leg.K1;
        g.L = leg.L;
        g.c1 = leg.c1;
        g.i0 = static_cast<int32_t>(std::min<int64_t>(i, INT_MAX / 4));
        g.m = INT_MAX;
        g.P_before = P_before;
      }
      g.vol_x = d.volume ? d.volume[v] : 1.f;
      g.vol_dy = 0.f;
      g.regime = 0;
      g.cap = 0.f;
      g.c0 = static_cast<int32_t>(std::max<int64_t>(n - kd + 1, -(1ll << 30)));
      g.attack = A;
      g.hold = H;
      g.decay = D;
      g.sustain = sustain;
      g.release = R;
      g.rel_base = rel < 0 ? INT_MAX : static_cast<int32_t>(std::max<int64_t>(rel - n, -(1ll << 30)));
      g.top = static_cast<float>(top);
      g.gl = 1.f;
      g.gr = 0.f;
      out.segs.push_back(g);
      out.s0.push_back(row0 + n);
      out.len.push_back(len);
      const double adv = leg.steady ? static_cast<double>(len) * leg.from1 : leg.P(i + len - 1) - P_before;
      phase += adv;
      phase -= 2.0 * floor(0.5 * phase);
      i += len;
      n += len;
    }
    std::swap(lf, lt);  // the ramp turns around (:2523-2526)
    cs = 0.0;
    if (leg.m == 0 && leg.steady) break;
  }
}

}  // namespace

int ramp_render_host(ca_ctx* ctx, const ca_ramp_desc* desc, float* h_out) {
  if (!desc || !h_out || desc->n_voices < 0 || desc->n_samples < 0 || desc->sample_rate <= 0)
    return fail(ctx, CA_ERR_INVALID, "ca_ramp_render_host: null argument or negative size");
  if (!desc->from_hz || !desc->to_hz || !desc->duration)
    return fail(ctx, CA_ERR_INVALID, "ca_ramp_render_host: from_hz / to_hz / duration are required");
  if (desc->mix != CA_MIX_NONE && desc->mix != CA_MIX_MONO)
    return fail(ctx, CA_ERR_UNSUPPORTED, "ca_ramp_render_host: mix must be CA_MIX_NONE or CA_MIX_MONO");
  const ca_ramp_desc d = *desc;
  if (d.env.attack_itp != CA_ITP_LINEAR || d.env.decay_itp != CA_ITP_LINEAR || d.env.release_itp != CA_ITP_LINEAR)
    return fail(ctx, CA_ERR_UNSUPPORTED, "ca_ramp_render_host: the ramp voices render LINEAR envelopes only");
  for (int64_t v = 0; v < d.n_voices; ++v)
    if (!(d.from_hz[v] > 0.f) || !(d.to_hz[v] > 0.f) || !(d.duration[v] > 0.f))
      return fail(ctx, CA_ERR_INVALID, "ca_ramp_render_host: frequencies and durations must be positive (LogRamp asserts it)");
  const bool mixed = d.mix == CA_MIX_MONO;
  const size_t out_elems = static_cast<size_t>(mixed ? 1 : d.n_voices) * d.n_samples;
  if (out_elems == 0) return CA_OK;

  // ---- tiles, built per voice on host threads ----
  std::vector<VoiceTiles> per(d.n_voices);
  {
    int nt = static_cast<int>(std::thread::hardware_concurrency());
    nt = std::max(1, std::min<int>(nt, static_cast<int>(std::max<int64_t>(d.n_voices, 1))));
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t)
      pool.emplace_back([&, t] {
        for (int64_t v = t; v < d.n_voices; v += nt) build_voice(d, v, mixed ? 0 : v * d.n_samples, per[v]);
      });
    for (auto& th : pool) th.join();
  }
  int64_t n_tiles = 0;
  for (auto& p : per) n_tiles += static_cast<int64_t>(p.segs.size());
  const size_t b_begin = sizeof(int64_t) * (n_tiles + 1), b_s0 = sizeof(int64_t) * std::max<int64_t>(n_tiles, 1),
               b_len = sizeof(int32_t) * std::max<int64_t>(n_tiles, 1);
  int rc = ensure_pinned(ctx, 4, std::max<size_t>(16, sizeof(Segment) * n_tiles));
  if (!rc) rc = ensure_pinned(ctx, 5, b_begin + b_s0 + b_len);
  if (!rc) rc = ensure_scratch(ctx, 6, std::max<size_t>(16, sizeof(Segment) * n_tiles));
  if (!rc) rc = ensure_scratch(ctx, 1, b_begin + b_s0 + b_len);
  if (!rc) rc = ensure_scratch(ctx, 7, sizeof(float) * out_elems);
  if (rc) return rc;
  Segment* h_segs = static_cast<Segment*>(ctx->h_pinned[4]);
  int64_t* tile_begin = static_cast<int64_t*>(ctx->h_pinned[5]);
  int64_t* tile_s0 = tile_begin + (n_tiles + 1);
  int32_t* tile_len = reinterpret_cast<int32_t*>(tile_s0 + std::max<int64_t>(n_tiles, 1));
  int64_t at = 0;
  for (auto& p : per) {  // one segment per tile
    for (size_t k = 0; k < p.segs.size(); ++k) {
      h_segs[at] = p.segs[k];
      tile_begin[at] = at;
      tile_s0[at] = p.s0[k];
      tile_len[at] = p.len[k];
      ++at;
    }
  }
  tile_begin[n_tiles] = n_tiles;

  cudaStream_t st = ctx->own_stream;
  CA_CUDA_OK(ctx, cudaMemsetAsync(ctx->d_scratch[7], 0, sizeof(float) * out_elems, st));
  if (n_tiles > 0) {
    CA_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_scratch[6], h_segs, sizeof(Segment) * n_tiles, cudaMemcpyHostToDevice, st));
    CA_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_scratch[1], tile_begin, b_begin + b_s0 + b_len, cudaMemcpyHostToDevice, st));
    const int64_t* d_begin = static_cast<const int64_t*>(ctx->d_scratch[1]);
    const int64_t* d_s0 = d_begin + (n_tiles + 1);
    const int32_t* d_len = reinterpret_cast<const int32_t*>(d_s0 + std::max<int64_t>(n_tiles, 1));
    TileArgs ta{static_cast<const Segment*>(ctx->d_scratch[6]), d_begin, d_s0, d_len, n_tiles, static_cast<float*>(ctx->d_scratch[7]),
                mixed ? 1 : 0};
    tracks_kernel<true><<<static_cast<unsigned>((n_tiles + 3) / 4), 128, 0, st>>>(ta);
    CA_CUDA_OK(ctx, cudaGetLastError());
    ctx->launches += 1;
  }
  CA_CUDA_OK(ctx, cudaMemcpyAsync(h_out, ctx->d_scratch[7], sizeof(float) * out_elems, cudaMemcpyDeviceToHost, st));
  CA_CUDA_OK(ctx, cudaStreamSynchronize(st));
  return CA_OK;
}

}  // namespace ca
