// tracks_render.cu — ca_oscbank_render_tracks: a note schedule (NoteOn / NoteChange / NoteOff at sample times) rendered on the
// rt.resynth voice pool. The stand-alone form of the synthesis half of the round trip, for a host that owns its note schedule.
//
// Reference: ImplCRTP::onEvent (include/gen.crtp.h:382-622; Event of include/events.h:12-107) posts the events to a pool of
// ResynthElement channels = StereoPanned<VolumeAdjusted<Enveloped<FreqCtrl_<SineOscillatorAlgo, InterpolatedFreq>, AHDSR>>>
// (source/rt.resynth.lib.cpp:5-18) initialised by ResynthElementInitializer (:96-118: AHDSR, volume filter limit 2/stride,
// glide over `stride` samples, random pan); SimpleAudioOutContext::step applies the events of a sample before the pool steps
// to produce it (include/simple_audio_context.h:69-112) and ImplCRTP::compute sums the channels (include/gen.crtp.h:350-378).
//
// Here: the host walks the schedule once (control rate: one visit per event) and cuts every sounding note into SEGMENTS, the
// closed-form description of the note over a tile of at most 1024 samples (tracks.cuh, pitch.cuh: the same note arithmetic the
// round trip's host form uses); tracks_kernel renders the tiles, one warp each.
#include <algorithm>
#include <string>
#include <vector>

#include "common.cuh"
#include "ctx.hpp"
#include "pitch.cuh"
#include "tracks.cuh"

namespace ca {

namespace {

struct Tiles {
  std::vector<Segment> segs;
  std::vector<int64_t> begin;  // [tiles + 1]
  std::vector<int64_t> s0;     // first sample of the tile
  std::vector<int32_t> len;
};

// schedule -> tiles and segments. stats[0] = notes started, stats[1] = notes dropped (no free channel).
int schedule_segments(ca_ctx* ctx, const ca_tracks_desc& d, Tiles& out, int64_t* stats) {
  ca_roundtrip_params rp{};
  rp.stft.window = 2;  // unused by the note arithmetic
  rp.stft.hop = d.glide_samples;
  rp.stft.sample_rate = d.sample_rate;
  rp.stereo_spread = d.stereo_spread;
  rp.env = d.env;
  rp.max_channels = d.max_channels;
  const PitchParams p = make_pitch_params(rp, d.n_samples, 0, 1);
  const bool mono = d.mix == CA_MIX_MONO;

  std::vector<Note> notes;
  HostChannels ch(p.max_channels);
  int64_t started = 0, dropped = 0;
  int64_t t = 0;  // samples [0, t) are covered by tiles
  out.begin.push_back(0);
  auto advance_to = [&](int64_t t_end) {
    while (t < t_end) {
      const int len = static_cast<int>(std::min<int64_t>(kMaxTile, t_end - t));
      for (auto& n : notes) {
        Segment g = emit_segment(p, n, t, len);
        if (mono) g.gl = 1.f;  // the element without StereoPanned (nOuts == 1)
        out.segs.push_back(g);
        ch.osc_ai[n.chan] = n.f_last;  // the oscillator has stepped: its increment is the last one emitted
      }
      out.begin.push_back(static_cast<int64_t>(out.segs.size()));
      out.s0.push_back(t);
      out.len.push_back(len);
      t += len;
    }
    // channels whose envelope reached EnvelopeDone2 are free for the events of sample t
    const size_t before = notes.size();
    notes.erase(std::remove_if(notes.begin(), notes.end(), [&](const Note& n) { return note_done(n, t); }), notes.end());
    if (notes.size() != before) ch.rebuild(notes);
  };
  int64_t last_at = 0;
  for (int64_t e = 0; e < d.n_events; ++e) {
    const ca_track_event& ev = d.events[e];
    if (ev.at_sample < last_at) return fail(ctx, CA_ERR_INVALID, "ca_oscbank_render_tracks: events must be ordered by at_sample");
    last_at = ev.at_sample;
    if (ev.at_sample >= d.n_samples) break;
    advance_to(std::max<int64_t>(ev.at_sample, 0));
    Tuned tu{};
    if (ev.type != CA_NOTE_OFF) {
      const float f = 2 * ev.frequency / static_cast<float>(d.sample_rate);  // freq_to_angle_increment in float (sound.functions.h:58-62)
      tu.ai = static_cast<double>(f);
      tu.volume = ev.velocity;
      tu.min_change = min_change_for(tu.ai);
    }
    switch (ev.type) {
      case CA_NOTE_ON:
        ++started;
        if (pool_note_on(p, notes, ch, tu, ev.noteid, t) < 0) ++dropped;
        break;
      case CA_NOTE_CHANGE:  // every channel that holds the id (include/gen.crtp.h:605-614)
        for (auto& n : notes)
          if (n.id == ev.noteid) {
            note_change(p, n, tu);
            if (n.n_on == t) n.glide.start(tu.ai, tu.ai, p.hop);  // not stepped yet: the control has no past value to glide from (:2752-2762)
          }
        break;
      case CA_NOTE_OFF: {  // the first channel, in channel order, that holds the id; released if its key is pressed (:540-590)
        Note* first = nullptr;
        for (auto& n : notes)
          if (n.id == ev.noteid && (!first || n.chan < first->chan)) first = &n;
        if (first && first->tracked) note_off(p, *first, t);
        break;
      }
      default:
        return fail(ctx, CA_ERR_INVALID, "ca_oscbank_render_tracks: event type is not a ca_event_type");
    }
  }
  advance_to(d.n_samples);
  if (stats) {
    stats[0] = started;
    stats[1] = dropped;
  }
  return CA_OK;
}

int check_desc(ca_ctx* ctx, const ca_tracks_desc* d) {
  if (!d || d->n_events < 0 || (d->n_events > 0 && !d->events) || d->n_samples < 0 || d->sample_rate <= 0 || d->glide_samples < 1)
    return fail(ctx, CA_ERR_INVALID, "ca_oscbank_render_tracks: bad descriptor");
  if (d->mix != CA_MIX_STEREO && d->mix != CA_MIX_MONO) return fail(ctx, CA_ERR_INVALID, "ca_oscbank_render_tracks: mix must be CA_MIX_STEREO or CA_MIX_MONO");
  if (d->env.attack_itp != CA_ITP_LINEAR || d->env.decay_itp != CA_ITP_LINEAR || d->env.release_itp != CA_ITP_LINEAR)
    return fail(ctx, CA_ERR_UNSUPPORTED, "ca_oscbank_render_tracks: LINEAR envelopes only (as rt.resynth.lib.cpp:75-81)");
  if (d->max_channels < 1 || d->max_channels > 65535) return fail(ctx, CA_ERR_INVALID, "ca_oscbank_render_tracks: max_channels must be within 1 .. 65535");
  return CA_OK;
}

}  // namespace

// d_out: device memory, [n_samples][2] (CA_MIX_STEREO) or [n_samples] (CA_MIX_MONO); every sample is written.
int tracks_render(ca_ctx* ctx, const ca_tracks_desc* desc, float* d_out, int64_t* h_stats, cudaStream_t st) {
  if (int rc = check_desc(ctx, desc)) return rc;
  if (!d_out) return fail(ctx, CA_ERR_INVALID, "ca_oscbank_render_tracks: null output");
  if (desc->n_samples == 0) {
    if (h_stats) h_stats[0] = h_stats[1] = 0;
    return CA_OK;
  }
  Tiles tl;
  if (int rc = schedule_segments(ctx, *desc, tl, h_stats)) return rc;
  const int64_t n_tiles = static_cast<int64_t>(tl.len.size());
  const size_t b_segs = sizeof(Segment) * std::max<size_t>(tl.segs.size(), 1);
  const size_t b_begin = sizeof(int64_t) * (n_tiles + 1), b_s0 = sizeof(int64_t) * n_tiles, b_len = sizeof(int32_t) * n_tiles;
  int rc = ensure_scratch(ctx, 6, b_segs);
  if (!rc) rc = ensure_scratch(ctx, 1, b_begin + b_s0 + b_len);
  if (rc) return rc;
  // (pageable sources: the runtime stages them before the calls return, so the vectors may go out of scope)
  if (!tl.segs.empty()) CA_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_scratch[6], tl.segs.data(), sizeof(Segment) * tl.segs.size(), cudaMemcpyHostToDevice, st));
  char* base = static_cast<char*>(ctx->d_scratch[1]);
  CA_CUDA_OK(ctx, cudaMemcpyAsync(base, tl.begin.data(), b_begin, cudaMemcpyHostToDevice, st));
  CA_CUDA_OK(ctx, cudaMemcpyAsync(base + b_begin, tl.s0.data(), b_s0, cudaMemcpyHostToDevice, st));
  CA_CUDA_OK(ctx, cudaMemcpyAsync(base + b_begin + b_s0, tl.len.data(), b_len, cudaMemcpyHostToDevice, st));
  TileArgs ta{static_cast<const Segment*>(ctx->d_scratch[6]), reinterpret_cast<const int64_t*>(base), reinterpret_cast<const int64_t*>(base + b_begin),
              reinterpret_cast<const int32_t*>(base + b_begin + b_s0), n_tiles, d_out, 0};
  const unsigned grid = static_cast<unsigned>((n_tiles + 3) / 4);
  if (desc->mix == CA_MIX_MONO) tracks_kernel<true><<<grid, 128, 0, st>>>(ta);
  else tracks_kernel<false><<<grid, 128, 0, st>>>(ta);
  CA_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches += 1;
  return CA_OK;
}

int tracks_render_host(ca_ctx* ctx, const ca_tracks_desc* desc, float* h_out, int64_t* h_stats) {
  if (int rc = check_desc(ctx, desc)) return rc;
  if (!h_out) return fail(ctx, CA_ERR_INVALID, "ca_oscbank_render_tracks_host: null output");
  const size_t bytes = sizeof(float) * static_cast<size_t>(desc->n_samples) * (desc->mix == CA_MIX_MONO ? 1 : 2);
  if (bytes == 0) {
    if (h_stats) h_stats[0] = h_stats[1] = 0;
    return CA_OK;
  }
  int rc = ensure_scratch(ctx, 7, bytes);
  if (rc) return rc;
  cudaStream_t st = ctx->own_stream;
  rc = tracks_render(ctx, desc, static_cast<float*>(ctx->d_scratch[7]), h_stats, st);
  if (rc) return rc;
  CA_CUDA_OK(ctx, cudaMemcpyAsync(h_out, ctx->d_scratch[7], bytes, cudaMemcpyDeviceToHost, st));
  CA_CUDA_OK(ctx, cudaStreamSynchronize(st));
  return CA_OK;
}

}  // namespace ca
