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
#include <climits>
#include <deque>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "ctx.hpp"
#include "pitch.cuh"
#include "tracks.cuh"

// The pool as an object (ca_tracks_*): events are posted as they become known, audio is pulled block by block; the one-shot
// entry points open a pool, post the whole schedule and pull once.
struct ca_tracks {
  ca_ctx* ctx = nullptr;
  ca::PitchParams p{};
  bool mono = false;
  std::vector<ca::Note> notes;
  ca::HostChannels ch{1};
  int64_t t = 0;  // position: samples [0, t) have been rendered
  std::deque<ca_track_event> pending;
  int64_t started = 0, dropped = 0;
};

namespace ca {

namespace {

struct Tiles {
  std::vector<Segment> segs;
  std::vector<int64_t> begin;  // [tiles + 1]
  std::vector<int64_t> s0;     // first sample of the tile, relative to the block
  std::vector<int32_t> len;
};

int check_desc(ca_ctx* ctx, const ca_tracks_desc* d) {
  if (!d || d->n_events < 0 || (d->n_events > 0 && !d->events) || d->n_samples < 0 || d->sample_rate <= 0 || d->glide_samples < 1)
    return fail(ctx, CA_ERR_INVALID, "ca_oscbank_render_tracks: bad descriptor");
  if (d->mix != CA_MIX_STEREO && d->mix != CA_MIX_MONO) return fail(ctx, CA_ERR_INVALID, "ca_oscbank_render_tracks: mix must be CA_MIX_STEREO or CA_MIX_MONO");
  if (d->env.attack_itp != CA_ITP_LINEAR || d->env.decay_itp != CA_ITP_LINEAR || d->env.release_itp != CA_ITP_LINEAR)
    return fail(ctx, CA_ERR_UNSUPPORTED, "ca_oscbank_render_tracks: LINEAR envelopes only (as rt.resynth.lib.cpp:75-81)");
  if (d->max_channels < 1 || d->max_channels > 65535) return fail(ctx, CA_ERR_INVALID, "ca_oscbank_render_tracks: max_channels must be within 1 .. 65535");
  return CA_OK;
}

void pool_setup(ca_tracks& tp, ca_ctx* ctx, const ca_tracks_desc& d) {
  ca_roundtrip_params rp{};
  rp.stft.window = 2;  // unused by the note arithmetic
  rp.stft.hop = d.glide_samples;
  rp.stft.sample_rate = d.sample_rate;
  rp.stereo_spread = d.stereo_spread;
  rp.env = d.env;
  rp.max_channels = d.max_channels;
  tp.ctx = ctx;
  tp.p = make_pitch_params(rp, 0, 0, 1);
  tp.mono = d.mix == CA_MIX_MONO;
  tp.ch = HostChannels(tp.p.max_channels);
}

int pool_post(ca_tracks& tp, const ca_track_event* ev, int64_t n) {
  int64_t last = tp.pending.empty() ? INT64_MIN : tp.pending.back().at_sample;
  for (int64_t e = 0; e < n; ++e) {
    if (ev[e].at_sample < last) return fail(tp.ctx, CA_ERR_INVALID, "ca_tracks_post: events must be ordered by at_sample");
    if (ev[e].type != CA_NOTE_ON && ev[e].type != CA_NOTE_OFF && ev[e].type != CA_NOTE_CHANGE)
      return fail(tp.ctx, CA_ERR_INVALID, "ca_tracks_post: event type is not a ca_event_type");
    last = ev[e].at_sample;
  }
  tp.pending.insert(tp.pending.end(), ev, ev + n);
  return CA_OK;
}

// the next n samples -> tiles and segments (tile positions relative to the block)
void pool_advance(ca_tracks& tp, int64_t n, Tiles& out) {
  const PitchParams& p = tp.p;
  const int64_t t_begin = tp.t, t_stop = tp.t + n;
  out.begin.push_back(0);
  auto advance_to = [&](int64_t t_end) {
    while (tp.t < t_end) {
      const int len = static_cast<int>(std::min<int64_t>(kMaxTile, t_end - tp.t));
      for (auto& nt : tp.notes) {
        Segment g = emit_segment(p, nt, tp.t, len);
        if (tp.mono) g.gl = 1.f;  // the element without StereoPanned (nOuts == 1)
        out.segs.push_back(g);
        tp.ch.osc_ai[nt.chan] = nt.f_last;  // the oscillator has stepped: its increment is the last one emitted
      }
      out.begin.push_back(static_cast<int64_t>(out.segs.size()));
      out.s0.push_back(tp.t - t_begin);
      out.len.push_back(len);
      tp.t += len;
    }
    // channels whose envelope reached EnvelopeDone2 are free for the events of sample t
    const size_t before = tp.notes.size();
    tp.notes.erase(std::remove_if(tp.notes.begin(), tp.notes.end(), [&](const Note& nt) { return note_done(nt, tp.t); }), tp.notes.end());
    if (tp.notes.size() != before) tp.ch.rebuild(tp.notes);
  };
  while (!tp.pending.empty() && tp.pending.front().at_sample < t_stop) {
    const ca_track_event ev = tp.pending.front();
    tp.pending.pop_front();
    advance_to(std::max(ev.at_sample, tp.t));  // an event posted late acts now (gen.crtp.h:503-510)
    const int64_t t = tp.t;
    Tuned tu{};
    if (ev.type != CA_NOTE_OFF) {
      const float f = 2 * ev.frequency / static_cast<float>(p.sr);  // freq_to_angle_increment in float (sound.functions.h:58-62)
      tu.ai = static_cast<double>(f);
      tu.volume = ev.velocity;
      tu.min_change = min_change_for(tu.ai);
    }
    if (ev.type == CA_NOTE_ON) {
      ++tp.started;
      if (pool_note_on(p, tp.notes, tp.ch, tu, ev.noteid, t) < 0) ++tp.dropped;
    } else if (ev.type == CA_NOTE_CHANGE) {  // every channel that holds the id (include/gen.crtp.h:605-614)
      for (auto& nt : tp.notes)
        if (nt.id == ev.noteid) {
          note_change(p, nt, tu);
          if (nt.n_on == t) nt.glide.start(tu.ai, tu.ai, p.hop);  // not stepped yet: the control has no past value to glide from (:2752-2762)
        }
    } else {  // NoteOff: the first channel, in channel order, that holds the id; released if its key is pressed (:540-590)
      Note* first = nullptr;
      for (auto& nt : tp.notes)
        if (nt.id == ev.noteid && (!first || nt.chan < first->chan)) first = &nt;
      if (first && first->tracked) note_off(p, *first, t);
    }
  }
  advance_to(t_stop);
}

// tiles -> device -> tracks_kernel into d_out ([n][2] or [n])
int render_tiles(ca_ctx* ctx, const Tiles& tl, bool mono, float* d_out, cudaStream_t st) {
  const int64_t n_tiles = static_cast<int64_t>(tl.len.size());
  if (n_tiles == 0) return CA_OK;
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
  if (mono) tracks_kernel<true><<<grid, 128, 0, st>>>(ta);
  else tracks_kernel<false><<<grid, 128, 0, st>>>(ta);
  CA_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches += 1;
  return CA_OK;
}

}  // namespace

int tracks_open(ca_ctx* ctx, const ca_tracks_desc* desc, ca_tracks** out) {
  if (!out) return fail(ctx, CA_ERR_INVALID, "ca_tracks_open: null argument");
  *out = nullptr;
  if (int rc = check_desc(ctx, desc)) return rc;
  ca_tracks* tp = new (std::nothrow) ca_tracks();
  if (!tp) return fail(ctx, CA_ERR_NOMEM, "ca_tracks_open: out of memory");
  pool_setup(*tp, ctx, *desc);
  if (int rc = pool_post(*tp, desc->events, desc->n_events)) {
    delete tp;
    return rc;
  }
  *out = tp;
  return CA_OK;
}

int tracks_post(ca_tracks* tp, const ca_track_event* events, int64_t n) {
  if (!tp || n < 0 || (n > 0 && !events)) return CA_ERR_INVALID;
  return pool_post(*tp, events, n);
}

// the next n_samples of the pool into device memory (asynchronous on st)
int tracks_pull(ca_tracks* tp, int64_t n_samples, float* d_out, cudaStream_t st) {
  if (!tp || n_samples < 0 || (n_samples > 0 && !d_out)) return CA_ERR_INVALID;
  if (n_samples == 0) return CA_OK;
  Tiles tl;
  pool_advance(*tp, n_samples, tl);
  return render_tiles(tp->ctx, tl, tp->mono, d_out, st);
}

int tracks_pull_host(ca_tracks* tp, int64_t n_samples, float* h_out) {
  if (!tp || n_samples < 0 || (n_samples > 0 && !h_out)) return CA_ERR_INVALID;
  if (n_samples == 0) return CA_OK;
  ca_ctx* ctx = tp->ctx;
  const size_t bytes = sizeof(float) * static_cast<size_t>(n_samples) * (tp->mono ? 1 : 2);
  int rc = ensure_scratch(ctx, 7, bytes);
  if (rc) return rc;
  cudaStream_t st = ctx->own_stream;
  rc = tracks_pull(tp, n_samples, static_cast<float*>(ctx->d_scratch[7]), st);
  if (rc) return rc;
  CA_CUDA_OK(ctx, cudaMemcpyAsync(h_out, ctx->d_scratch[7], bytes, cudaMemcpyDeviceToHost, st));
  CA_CUDA_OK(ctx, cudaStreamSynchronize(st));
  return CA_OK;
}

ca_ctx* tracks_ctx(const ca_tracks* tp) { return tp->ctx; }
int tracks_state(const ca_tracks* tp, int64_t state[5]) {
  if (!tp || !state) return CA_ERR_INVALID;
  state[0] = tp->t;
  state[1] = tp->started;
  state[2] = tp->dropped;
  state[3] = static_cast<int64_t>(tp->notes.size());
  state[4] = static_cast<int64_t>(tp->pending.size());
  return CA_OK;
}
void tracks_close(ca_tracks* tp) { delete tp; }

// d_out: device memory, [n_samples][2] (CA_MIX_STEREO) or [n_samples] (CA_MIX_MONO); every sample is written.
int tracks_render(ca_ctx* ctx, const ca_tracks_desc* desc, float* d_out, int64_t* h_stats, cudaStream_t st) {
  if (desc && desc->n_samples > 0 && !d_out) return fail(ctx, CA_ERR_INVALID, "ca_oscbank_render_tracks: null output");
  ca_tracks* tp = nullptr;
  int rc = tracks_open(ctx, desc, &tp);
  if (rc) return rc;
  rc = tracks_pull(tp, desc->n_samples, d_out, st);
  if (h_stats) {
    h_stats[0] = tp->started;
    h_stats[1] = tp->dropped;
  }
  delete tp;
  return rc;
}

int tracks_render_host(ca_ctx* ctx, const ca_tracks_desc* desc, float* h_out, int64_t* h_stats) {
  if (desc && desc->n_samples > 0 && !h_out) return fail(ctx, CA_ERR_INVALID, "ca_oscbank_render_tracks_host: null output");
  ca_tracks* tp = nullptr;
  int rc = tracks_open(ctx, desc, &tp);
  if (rc) return rc;
  rc = tracks_pull_host(tp, desc->n_samples, h_out);
  if (h_stats) {
    h_stats[0] = tp->started;
    h_stats[1] = tp->dropped;
  }
  delete tp;
  return rc;
}

}  // namespace ca
