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
#include "pitch.cuh"
#include "tracks.cuh"

namespace ca {

namespace {


// ---------------------------------------------------------------------------------------------------------------
// Host side: pitch pipeline + note bookkeeping of one stream -> segments
// ---------------------------------------------------------------------------------------------------------------

struct StreamResult {
  std::vector<Segment> segs;
  std::vector<int64_t> tile_seg_begin;  // per tile (a frame interval longer than kMaxTile samples is cut into several tiles)
  std::vector<int64_t> tile_n0;         // first sample of the tile (stream-relative)
  std::vector<int32_t> tile_len;
  int64_t notes_started = 0, notes_dropped = 0;
  int64_t truncated = 0;  // frames whose raw or tuned list did not fit its capacity
};

// The host form of the pipeline (A/B against the device kernels of pitch.cu; host_threads != 0 selects it).
// peak_stride: entries per frame in the (compacted) peak arrays
template <class R>
void run_stream(const PitchParams& p, const int32_t* peak_bin, const R* peak_off, const R* peak_db,
                const int32_t* peak_count, StreamResult& res) {
  const int P = p.P;
  std::vector<Note> notes;       // sounding notes (tracked or releasing)
  std::vector<int> played;       // indices into notes of tracked notes, sorted by pitch
  HostChannels ch(p.max_channels);
  std::vector<Tuned> tuned(std::max(std::max(P, p.opt_active ? p.tcap : 0), 1));
  std::vector<int> change, order, next_played;
  std::vector<char> keep;

  res.tile_seg_begin.clear();
  res.tile_n0.clear();
  res.tile_len.clear();
  res.tile_seg_begin.reserve(p.frames + 1);
  for (int64_t k = 0; k < p.frames; ++k) {
    const int64_t n_k = static_cast<int64_t>(p.W) - 1 + k * p.hop;  // events of frame k act from this sample on
    const int len_frame = static_cast<int>(std::min<int64_t>(p.hop, p.n_samples - n_k));
    // ---------------- stateless prefix (source/rt.resynth.lib.algo.cpp:10-230) ----------------
    int cut = peak_count[k] > p.cap_raw ? 1 : 0;
    const int T = pitch_prefix(p, peak_bin + k * P, peak_off + k * P, peak_db + k * P, std::min(peak_count[k], P), tuned.data(), &cut);
    res.truncated += cut != 0;
    // ---------------- tracking (greedy, monotone in pitch) ----------------
    change.assign(T, -1);
    keep.assign(played.size(), 0);
    {
      size_t it = 0;
      for (int i = 0; i < T; ++i) {
        for (; it < played.size(); ++it) {
          const double pp = notes[played[it]].pitch;
          if (pp < tuned[i].pitch - p.max_track) continue;
          if (pp <= tuned[i].pitch + p.max_track) {
            change[i] = static_cast<int>(it);
            keep[it] = 1;
            ++it;
          }
          break;
        }
      }
    }
    order.resize(T);
    for (int i = 0; i < T; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(),
                     [&](int x, int y) { return static_cast<float>(tuned[x].vol) > static_cast<float>(tuned[y].vol); });
    // ---------------- events (synthesize_sounds, source/rt.resynth.lib.cpp:264-382) ----------------
    for (size_t i = 0; i < played.size(); ++i)
      if (!keep[i]) note_off(p, notes[played[i]], n_k);
    next_played.clear();
    for (size_t i = 0; i < played.size(); ++i)
      if (keep[i]) next_played.push_back(played[i]);
    for (int idx : order) {
      if (change[idx] >= 0) {
        note_change(p, notes[played[change[idx]]], tuned[idx]);
      } else {
        if (tuned[idx].volume <= 0) continue;
        ++res.notes_started;
        if (pool_note_on(p, notes, ch, tuned[idx], res.notes_started, n_k) < 0) {
          ++res.notes_dropped;
          continue;
        }
        next_played.push_back(static_cast<int>(notes.size()) - 1);
      }
    }
    std::stable_sort(next_played.begin(), next_played.end(), [&](int x, int y) { return notes[x].pitch < notes[y].pitch; });
    played.swap(next_played);

    // ---------------- one segment per sounding note and tile, then advance the note states ----------------
    // (the interval up to the next frame is one tile, or several when the hop exceeds kMaxTile samples)
    for (int off = 0; off < std::max(len_frame, 0) || off == 0; off += kMaxTile) {
      const int len = std::min(kMaxTile, len_frame - off);
      const int64_t n_k_tile = n_k + off;
      res.tile_seg_begin.push_back(static_cast<int64_t>(res.segs.size()));
      res.tile_n0.push_back(n_k_tile);
      res.tile_len.push_back(std::max(len, 0));
      if (len <= 0) break;
      for (auto& n : notes) {
        res.segs.push_back(emit_segment(p, n, n_k_tile, len));
        ch.osc_ai[n.chan] = n.f_last;  // the oscillator has stepped: its increment is the last one emitted
      }
    }
    if (len_frame > 0) {
      const int64_t n_next = n_k + len_frame;
      size_t w = 0;
      std::vector<int> remap(notes.size(), -1);
      for (size_t i = 0; i < notes.size(); ++i) {
        if (!note_done(notes[i], n_next)) {
          remap[i] = static_cast<int>(w);
          if (w != i) notes[w] = notes[i];
          ++w;
        }
      }
      notes.resize(w);
      for (auto& pi : played) pi = remap[pi];
      ch.rebuild(notes);
    }
  }
  res.tile_seg_begin.push_back(static_cast<int64_t>(res.segs.size()));
}


// 16-bit PCM -> float as a WAV reader does it (x / 32768, exact in float32): the job's input crosses PCIe at 2 bytes per sample
__global__ void __launch_bounds__(256) pcm16_to_float_kernel(const int16_t* __restrict__ in, float* __restrict__ out, int64_t n) {
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 2;
  if (i + 1 < n) {
    const short2 v = *reinterpret_cast<const short2*>(in + i);
    *reinterpret_cast<float2*>(out + i) = make_float2(static_cast<float>(v.x) * (1.f / 32768.f), static_cast<float>(v.y) * (1.f / 32768.f));
  } else if (i < n) {
    out[i] = static_cast<float>(in[i]) * (1.f / 32768.f);
  }
}

// the same over a range of samples of every stream ([streams][stride] arrays)
__global__ void __launch_bounds__(256) pcm16_to_float_2d_kernel(const int16_t* __restrict__ in, float* __restrict__ out, int64_t w, int64_t stride) {
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 2;
  const int16_t* row_in = in + static_cast<int64_t>(blockIdx.y) * stride;
  float* row_out = out + static_cast<int64_t>(blockIdx.y) * stride;
  if (i < w) row_out[i] = static_cast<float>(row_in[i]) * (1.f / 32768.f);
  if (i + 1 < w) row_out[i + 1] = static_cast<float>(row_in[i + 1]) * (1.f / 32768.f);
}

// host samples -> d_x on stream st: float32 goes straight up, 16-bit PCM goes up at half the bytes and is converted on the device
int upload_samples(ca_ctx* ctx, const void* h_x, bool pcm16, size_t first, size_t count, float* d_x, cudaStream_t st) {
  if (!pcm16) {
    CA_CUDA_OK(ctx, cudaMemcpyAsync(d_x + first, static_cast<const float*>(h_x) + first, sizeof(float) * count, cudaMemcpyHostToDevice, st));
    return CA_OK;
  }
  int16_t* d_pcm = static_cast<int16_t*>(ctx->d_scratch[12]);
  CA_CUDA_OK(ctx, cudaMemcpyAsync(d_pcm + first, static_cast<const int16_t*>(h_x) + first, sizeof(int16_t) * count, cudaMemcpyHostToDevice, st));
  // (first is even whenever count is: streams of even length; an odd first only loses the 4-byte alignment of the pair loads)
  if (first % 2 == 0) {
    pcm16_to_float_kernel<<<static_cast<unsigned>((count / 2 + 256) / 256), 256, 0, st>>>(d_pcm + first, d_x + first, static_cast<int64_t>(count));
  } else {
    pcm16_to_float_kernel<<<1, 32, 0, st>>>(d_pcm + first, d_x + first, 1);
    pcm16_to_float_kernel<<<static_cast<unsigned>((count / 2 + 256) / 256), 256, 0, st>>>(d_pcm + first + 1, d_x + first + 1, static_cast<int64_t>(count) - 1);
  }
  CA_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches += 1;
  return CA_OK;
}

// make_pitch_params for the kernels: the autotune table (a host pointer in the parameters) goes to the device first
int device_pitch_params(ca_ctx* ctx, const ca_roundtrip_params& prm, int64_t n_samples, int64_t frames, int P, PitchParams* out) {
  *out = make_pitch_params(prm, n_samples, frames, P);
  if (out->autotune != CA_AUTOTUNE_NONE && out->n_table > 0) {
    const size_t bytes = sizeof(double) * static_cast<size_t>(out->n_table);
    const int rc = ensure_scratch(ctx, 11, bytes);
    if (rc) return rc;
    CA_CUDA_OK(ctx, cudaMemcpy(ctx->d_scratch[11], prm.opts.table, bytes, cudaMemcpyHostToDevice));
    out->table = static_cast<const double*>(ctx->d_scratch[11]);
  } else {
    out->table = nullptr;
  }
  return CA_OK;
}

// Device form of steps 2-3: tuned lists, tracking and segments by the kernels of pitch.cu, rendering, audio back. Nothing
// but the audio and 16 bytes of counters crosses PCIe, and the host does no per-frame work.
template <class Lap>
int roundtrip_device_tail(ca_ctx* ctx, const ca_roundtrip_params& prm, int64_t n_samples, int64_t n_streams, int64_t frames,
                          float* h_out, int64_t* h_stats, bool timing, bool use64, Lap&& lap) {
  cudaStream_t st = ctx->own_stream;
  const int P = prm.stft.max_peaks;
  PitchParams pp;
  if (int prc = device_pitch_params(ctx, prm, n_samples, frames, P, &pp)) return prc;
  const size_t rows = static_cast<size_t>(frames) * n_streams, total = static_cast<size_t>(n_streams) * n_samples;
  const size_t out_bytes = sizeof(float) * 2 * total;
  const int64_t tps = tiles_per_stream(pp), n_tiles = tps * n_streams;
  const size_t b_tiles = (3 * sizeof(int64_t) + sizeof(int32_t)) * static_cast<size_t>(n_tiles);
  const size_t b_misc = sizeof(int64_t) * 2 * n_streams + 24;  // stats, allocator counter, overflow flag, truncated frames
  int rc = ensure_scratch(ctx, 8, sizeof(Tuned) * rows * pp.PT);
  if (!rc) rc = ensure_scratch(ctx, 9, sizeof(int32_t) * rows);
  if (!rc) rc = ensure_scratch(ctx, 10, b_misc);
  if (!rc) rc = ensure_scratch(ctx, 1, b_tiles);
  if (!rc) rc = ensure_pinned(ctx, 0, b_misc);
  // first guess of the segment capacity: 48 sounding notes per tile (the run reports the exact need when that is short)
  const int64_t guess = std::max<int64_t>(n_tiles * std::min(prm.max_channels, 48), n_streams * 8192);
  if (!rc && ctx->scratch_bytes[6] < sizeof(Segment) * static_cast<size_t>(guess)) rc = ensure_scratch(ctx, 6, sizeof(Segment) * guess);
  if (rc) return rc;
  int64_t* d_begin = static_cast<int64_t*>(ctx->d_scratch[1]);
  int64_t* d_end = d_begin + n_tiles;
  int64_t* d_s0 = d_end + n_tiles;
  int32_t* d_len = reinterpret_cast<int32_t*>(d_s0 + n_tiles);
  int64_t* d_stats = static_cast<int64_t*>(ctx->d_scratch[10]);
  unsigned long long* d_counter = reinterpret_cast<unsigned long long*>(d_stats + 2 * n_streams);
  int32_t* d_overflow = reinterpret_cast<int32_t*>(d_counter + 1);
  unsigned long long* d_trunc = d_counter + 2;
  int64_t* h_misc = static_cast<int64_t*>(ctx->h_pinned[0]);
  CA_CUDA_OK(ctx, cudaMemsetAsync(d_trunc, 0, sizeof(unsigned long long), st));
  for (int attempt = 0; attempt < 2; ++attempt) {
    const int64_t capacity = static_cast<int64_t>(ctx->scratch_bytes[6] / sizeof(Segment));
    rc = pitch_pipeline_device(ctx, pp, n_streams, static_cast<const int32_t*>(ctx->d_scratch[2]), ctx->d_scratch[3], ctx->d_scratch[4], use64,
                               static_cast<const int32_t*>(ctx->d_scratch[5]), d_trunc, static_cast<Tuned*>(ctx->d_scratch[8]),
                               static_cast<int32_t*>(ctx->d_scratch[9]), static_cast<Segment*>(ctx->d_scratch[6]), capacity, tps,
                               d_begin, d_end, d_s0, d_len, d_stats, d_counter, d_overflow, attempt == 0, st);
    if (rc) return rc;
    CA_CUDA_OK(ctx, cudaMemcpyAsync(h_misc, d_stats, b_misc, cudaMemcpyDeviceToHost, st));
    if (timing) {
      CA_CUDA_OK(ctx, cudaStreamSynchronize(st));
      lap("pitch pipeline (device)");
    }
    CA_CUDA_OK(ctx, cudaMemsetAsync(ctx->d_scratch[7], 0, out_bytes, st));  // samples before the first frame are silent
    // Rendering and the copy of the audio overlap: the streams are rendered in groups, and each group's audio leaves on the
    // second stream while the next group renders.
    const int groups = timing ? 1 : static_cast<int>(std::min<int64_t>(4, n_streams));
    cudaStream_t st2 = ctx->own_stream2;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int g = 0; g < groups; ++g) {
      const int64_t s0 = n_streams * g / groups, s1 = n_streams * (g + 1) / groups;
      const int64_t t0 = s0 * tps, nt = (s1 - s0) * tps;
      TileArgs ta{static_cast<const Segment*>(ctx->d_scratch[6]), d_begin + t0, d_s0 + t0, d_len + t0, nt, static_cast<float*>(ctx->d_scratch[7]), 0, d_end + t0};
      tracks_kernel<false><<<static_cast<unsigned>((nt + 3) / 4), 128, 0, st>>>(ta);
      CA_CUDA_OK(ctx, cudaGetLastError());
      ctx->launches += 1;
      if (groups > 1) {
        CA_CUDA_OK(ctx, cudaEventCreateWithFlags(&ev[g], cudaEventDisableTiming));
        CA_CUDA_OK(ctx, cudaEventRecord(ev[g], st));
        CA_CUDA_OK(ctx, cudaStreamWaitEvent(st2, ev[g], 0));
        const size_t o0 = sizeof(float) * 2 * static_cast<size_t>(s0) * n_samples, o1 = sizeof(float) * 2 * static_cast<size_t>(s1) * n_samples;
        CA_CUDA_OK(ctx, cudaMemcpyAsync(reinterpret_cast<char*>(h_out) + o0, static_cast<char*>(ctx->d_scratch[7]) + o0, o1 - o0,
                                        cudaMemcpyDeviceToHost, st2));
      }
    }
    if (timing) {
      CA_CUDA_OK(ctx, cudaStreamSynchronize(st));
      lap("tracks kernel");
    }
    if (groups == 1) CA_CUDA_OK(ctx, cudaMemcpyAsync(h_out, ctx->d_scratch[7], out_bytes, cudaMemcpyDeviceToHost, st));
    CA_CUDA_OK(ctx, cudaStreamSynchronize(st));
    if (groups > 1) CA_CUDA_OK(ctx, cudaStreamSynchronize(st2));
    for (int g = 0; g < groups; ++g)
      if (ev[g]) cudaEventDestroy(ev[g]);
    lap("D2H output");
    const int64_t needed = h_misc[2 * n_streams];
    const int32_t overflow = *reinterpret_cast<const int32_t*>(h_misc + 2 * n_streams + 1);
    if (!overflow) break;
    if (attempt == 1) return fail(ctx, CA_ERR_CUDA, "round trip: segment buffer overflow after resizing");
    rc = ensure_scratch(ctx, 6, sizeof(Segment) * static_cast<size_t>(needed));
    if (rc) return rc;
  }
  if (h_stats) memcpy(h_stats, h_misc, sizeof(int64_t) * 2 * n_streams);
  ctx->rt_truncated_frames = h_misc[2 * n_streams + 2];
  return CA_OK;
}


// The device path as a pipeline over groups of streams: every group runs H2D -> K2 -> pitch pipeline -> rendering -> D2H on its
// own CUDA stream, so the copies of one group overlap the (latency-bound) tracking kernel of the others. Returns
// CA_OK with *overflowed = true when the segment buffer was too small (the caller grows it and takes the sequential path).
int roundtrip_device_pipelined(ca_ctx* ctx, const ca_roundtrip_params& prm, const void* h_x, bool pcm16, int64_t n_samples, int64_t n_streams,
                               int64_t frames, float* h_out, int64_t* h_stats, bool use64, bool* overflowed, int64_t* needed) {
  const int P = prm.stft.max_peaks;
  PitchParams pp;
  if (int prc = device_pitch_params(ctx, prm, n_samples, frames, P, &pp)) return prc;
  const size_t rows = static_cast<size_t>(frames) * n_streams, total = static_cast<size_t>(n_streams) * n_samples;
  const int64_t tps = tiles_per_stream(pp), n_tiles = tps * n_streams;
  const size_t b_tiles = (3 * sizeof(int64_t) + sizeof(int32_t)) * static_cast<size_t>(n_tiles);
  const size_t b_misc = sizeof(int64_t) * 2 * n_streams + 24;  // stats, allocator counter, overflow flag, truncated frames
  int rc = ensure_scratch(ctx, 8, sizeof(Tuned) * rows * pp.PT);
  if (!rc) rc = ensure_scratch(ctx, 9, sizeof(int32_t) * rows);
  if (!rc) rc = ensure_scratch(ctx, 10, b_misc);
  if (!rc) rc = ensure_scratch(ctx, 1, b_tiles);
  if (!rc) rc = ensure_pinned(ctx, 0, b_misc);
  const int64_t guess = std::max<int64_t>(n_tiles * std::min(prm.max_channels, 48), n_streams * 8192);
  if (!rc && ctx->scratch_bytes[6] < sizeof(Segment) * static_cast<size_t>(guess)) rc = ensure_scratch(ctx, 6, sizeof(Segment) * guess);
  if (rc) return rc;
  static const int max_groups = [] { const char* e = getenv("CA_RT_GROUPS"); const int v = e ? atoi(e) : 4; return v < 1 ? 1 : v > 8 ? 8 : v; }();
  const int groups = static_cast<int>(std::min<int64_t>(max_groups, n_streams));
  for (int g = 0; g < groups; ++g)
    if (!ctx->pipe_streams[g]) CA_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->pipe_streams[g], cudaStreamNonBlocking));
  float* d_x = static_cast<float*>(ctx->d_scratch[0]);
  int32_t* d_pbin = static_cast<int32_t*>(ctx->d_scratch[2]);
  char* d_poff = static_cast<char*>(ctx->d_scratch[3]);  // float or double entries (use64)
  char* d_pdb = static_cast<char*>(ctx->d_scratch[4]);
  const size_t esz = use64 ? sizeof(double) : sizeof(float);
  int32_t* d_pcnt = static_cast<int32_t*>(ctx->d_scratch[5]);
  Tuned* d_tuned = static_cast<Tuned*>(ctx->d_scratch[8]);
  int32_t* d_tcount = static_cast<int32_t*>(ctx->d_scratch[9]);
  float* d_out = static_cast<float*>(ctx->d_scratch[7]);
  int64_t* d_begin = static_cast<int64_t*>(ctx->d_scratch[1]);
  int64_t* d_end = d_begin + n_tiles;
  int64_t* d_s0 = d_end + n_tiles;
  int32_t* d_len = reinterpret_cast<int32_t*>(d_s0 + n_tiles);
  int64_t* d_stats = static_cast<int64_t*>(ctx->d_scratch[10]);
  unsigned long long* d_counter = reinterpret_cast<unsigned long long*>(d_stats + 2 * n_streams);
  int32_t* d_overflow = reinterpret_cast<int32_t*>(d_counter + 1);
  int64_t* h_misc = static_cast<int64_t*>(ctx->h_pinned[0]);
  const int64_t capacity = static_cast<int64_t>(ctx->scratch_bytes[6] / sizeof(Segment));
  // the allocator counter and the overflow flag are shared by the groups: reset once, before any group starts
  cudaStream_t s0 = ctx->pipe_streams[0];
  cudaEvent_t ready = nullptr;
  unsigned long long* d_trunc = d_counter + 2;
  CA_CUDA_OK(ctx, cudaMemsetAsync(d_counter, 0, 24, s0));
  CA_CUDA_OK(ctx, cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  CA_CUDA_OK(ctx, cudaEventRecord(ready, s0));
  for (int g = 0; g < groups && !rc; ++g) {
    cudaStream_t sg = ctx->pipe_streams[g];
    const int64_t a0 = n_streams * g / groups, a1 = n_streams * (g + 1) / groups, ns = a1 - a0;
    const size_t x0 = static_cast<size_t>(a0) * n_samples, r0 = static_cast<size_t>(a0) * frames;
    const int64_t t0 = a0 * tps, nt = ns * tps;
    if (g > 0) cudaStreamWaitEvent(sg, ready, 0);
    rc = upload_samples(ctx, h_x, pcm16, x0, static_cast<size_t>(ns) * n_samples, d_x, sg);
    if (rc) break;
    if (use64)
      rc = stft64_peaks(ctx, ctx->plan_rt, d_x + x0, n_samples, ns, n_samples, d_pbin + r0 * P, reinterpret_cast<double*>(d_poff + r0 * P * esz),
                        reinterpret_cast<double*>(d_pdb + r0 * P * esz), d_pcnt + r0, sg);
    else
      rc = stft_analyze(ctx, ctx->plan_rt, d_x + x0, n_samples, ns, n_samples, nullptr, 0, d_pbin + r0 * P, reinterpret_cast<float*>(d_poff + r0 * P * esz),
                        reinterpret_cast<float*>(d_pdb + r0 * P * esz), d_pcnt + r0, sg);
    if (rc) break;
    rc = pitch_pipeline_device(ctx, pp, ns, d_pbin + r0 * P, d_poff + r0 * P * esz, d_pdb + r0 * P * esz, use64, d_pcnt + r0, d_trunc, d_tuned + r0 * pp.PT, d_tcount + r0,
                               static_cast<Segment*>(ctx->d_scratch[6]), capacity, tps, d_begin + t0, d_end + t0, d_s0 + t0, d_len + t0,
                               d_stats + 2 * a0, d_counter, d_overflow, true, sg, false);
    if (rc) break;
    float* out_g = d_out + 2 * x0;  // tile_s0 holds group-relative sample indices
    cudaMemsetAsync(out_g, 0, sizeof(float) * 2 * ns * n_samples, sg);  // samples before the first frame are silent
    TileArgs ta{static_cast<const Segment*>(ctx->d_scratch[6]), d_begin + t0, d_s0 + t0, d_len + t0, nt, out_g, 0, d_end + t0};
    tracks_kernel<false><<<static_cast<unsigned>((nt + 3) / 4), 128, 0, sg>>>(ta);
    ctx->launches += 1;
    cudaMemcpyAsync(h_out + 2 * x0, out_g, sizeof(float) * 2 * ns * n_samples, cudaMemcpyDeviceToHost, sg);
  }
  cudaError_t err = cudaGetLastError();
  for (int g = 0; g < groups; ++g) {
    const cudaError_t e = cudaStreamSynchronize(ctx->pipe_streams[g]);
    if (err == cudaSuccess) err = e;
  }
  cudaEventDestroy(ready);
  if (rc) return rc;
  if (err != cudaSuccess) return fail(ctx, CA_ERR_CUDA, std::string("round trip (pipelined): ") + cudaGetErrorString(err));
  CA_CUDA_OK(ctx, cudaMemcpy(h_misc, d_stats, b_misc, cudaMemcpyDeviceToHost));
  *needed = h_misc[2 * n_streams];
  *overflowed = *reinterpret_cast<const int32_t*>(h_misc + 2 * n_streams + 1) != 0;
  if (!*overflowed && h_stats) memcpy(h_stats, h_misc, sizeof(int64_t) * 2 * n_streams);
  ctx->rt_truncated_frames = h_misc[2 * n_streams + 2];
  return CA_OK;
}

// The device path as a pipeline over TIME: the streams are cut into chunks of consecutive frames, and chunk c goes through
//   copy engine (up)   samples the chunk's frames read, all streams (one 2-D copy)
//   stream A           analysis + tuned lists of the chunk's frames
//   stream B           tracking of the chunk's frames (the tracker is sequential over frames; its state rests in global memory
//                      between launches, pitch.cu), one launch per chunk, in order
//   stream C + copy engine (down)   rendering of the chunk's tiles, then the chunk's audio goes back (one 2-D copy)
// so the audio of the first chunk is on its way to the host while the tracker is still walking the later ones. The tracker is
// latency-bound (a fixed ~1.5 us per frame whatever the number of streams) and the copy back is the largest transfer of the call:
// run one after the other they add up, overlapped the call takes the longer of the two. Returns CA_OK with *overflowed = true when
// the segment buffer was too small (the caller grows it and takes the sequential path).
int roundtrip_device_chunked(ca_ctx* ctx, const ca_roundtrip_params& prm, const void* h_x, bool pcm16, int64_t n_samples, int64_t n_streams,
                             int64_t frames, float* h_out, int64_t* h_stats, bool use64, bool* overflowed, int64_t* needed) {
  const int P = prm.stft.max_peaks, W = prm.stft.window, hop = prm.stft.hop;
  PitchParams pp;
  if (int prc = device_pitch_params(ctx, prm, n_samples, frames, P, &pp)) return prc;
  const size_t rows = static_cast<size_t>(frames) * n_streams;
  const int64_t tps = tiles_per_stream(pp), n_tiles = tps * n_streams;
  const size_t b_tiles = (3 * sizeof(int64_t) + sizeof(int32_t)) * static_cast<size_t>(n_tiles);
  const size_t b_misc = sizeof(int64_t) * 2 * n_streams + 24;  // stats, allocator counter, overflow flag, truncated frames
  const size_t state_stride = (pitch_track_state_bytes() + 15) / 16 * 16;
  int rc = ensure_scratch(ctx, 8, sizeof(Tuned) * rows * pp.PT);
  if (!rc) rc = ensure_scratch(ctx, 9, sizeof(int32_t) * rows);
  if (!rc) rc = ensure_scratch(ctx, 10, b_misc);
  if (!rc) rc = ensure_scratch(ctx, 1, b_tiles);
  if (!rc) rc = ensure_scratch(ctx, 13, state_stride * static_cast<size_t>(n_streams));
  if (!rc) rc = ensure_pinned(ctx, 0, b_misc);
  const int64_t guess = std::max<int64_t>(n_tiles * std::min(prm.max_channels, 48), n_streams * 8192);
  if (!rc && ctx->scratch_bytes[6] < sizeof(Segment) * static_cast<size_t>(guess)) rc = ensure_scratch(ctx, 6, sizeof(Segment) * guess);
  if (rc) return rc;
  static const int max_chunks = [] { const char* e = getenv("CA_RT_CHUNKS"); const int v = e ? atoi(e) : 16; return v < 1 ? 1 : v > 64 ? 64 : v; }();
  const int chunks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(max_chunks, frames / 64)));
  for (int g = 0; g < 4; ++g)
    if (!ctx->pipe_streams[g]) CA_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->pipe_streams[g], cudaStreamNonBlocking));
  cudaStream_t s_up = ctx->pipe_streams[0], s_ana = ctx->pipe_streams[1], s_trk = ctx->pipe_streams[2], s_out = ctx->pipe_streams[3];
  float* d_x = static_cast<float*>(ctx->d_scratch[0]);
  int32_t* d_pbin = static_cast<int32_t*>(ctx->d_scratch[2]);
  char* d_poff = static_cast<char*>(ctx->d_scratch[3]);  // float or double entries (use64)
  char* d_pdb = static_cast<char*>(ctx->d_scratch[4]);
  const size_t esz = use64 ? sizeof(double) : sizeof(float);
  int32_t* d_pcnt = static_cast<int32_t*>(ctx->d_scratch[5]);
  Tuned* d_tuned = static_cast<Tuned*>(ctx->d_scratch[8]);
  int32_t* d_tcount = static_cast<int32_t*>(ctx->d_scratch[9]);
  float* d_out = static_cast<float*>(ctx->d_scratch[7]);
  int64_t* d_begin = static_cast<int64_t*>(ctx->d_scratch[1]);
  int64_t* d_end = d_begin + n_tiles;
  int64_t* d_s0 = d_end + n_tiles;
  int32_t* d_len = reinterpret_cast<int32_t*>(d_s0 + n_tiles);
  int64_t* d_stats = static_cast<int64_t*>(ctx->d_scratch[10]);
  unsigned long long* d_counter = reinterpret_cast<unsigned long long*>(d_stats + 2 * n_streams);
  int32_t* d_overflow = reinterpret_cast<int32_t*>(d_counter + 1);
  unsigned long long* d_trunc = d_counter + 2;
  unsigned char* d_state = static_cast<unsigned char*>(ctx->d_scratch[13]);
  int64_t* h_misc = static_cast<int64_t*>(ctx->h_pinned[0]);
  const int64_t capacity = static_cast<int64_t>(ctx->scratch_bytes[6] / sizeof(Segment));
  // tiles before every frame (a frame interval is one tile, or several when the hop exceeds kMaxTile; the last ones are shorter)
  std::vector<int64_t> tile_of(frames + 1, 0);
  for (int64_t k = 0; k < frames; ++k) {
    const int64_t n_k = static_cast<int64_t>(W) - 1 + k * hop;
    const int64_t len = std::min<int64_t>(hop, n_samples - n_k);
    tile_of[k + 1] = tile_of[k] + (len <= 0 ? 1 : (len + kMaxTile - 1) / kMaxTile);
  }
  std::vector<cudaEvent_t> ev(3 * chunks + 1, nullptr);
  for (auto& e : ev) CA_CUDA_OK(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  auto cleanup = [&] {
    for (auto& e : ev)
      if (e) cudaEventDestroy(e);
  };
  cudaError_t err = cudaSuccess;
  auto ok = [&](cudaError_t e) {
    if (err == cudaSuccess && e != cudaSuccess) err = e;
  };
  // counters shared by the chunks are reset before anything else runs; silence before the first frame
  ok(cudaMemsetAsync(d_counter, 0, 24, s_up));
  ok(cudaEventRecord(ev[3 * chunks], s_up));
  ok(cudaStreamWaitEvent(s_ana, ev[3 * chunks], 0));
  ok(cudaStreamWaitEvent(s_trk, ev[3 * chunks], 0));
  ok(cudaMemsetAsync(d_out, 0, sizeof(float) * 2 * static_cast<size_t>(n_streams) * n_samples, s_out));
  const size_t in_sz = pcm16 ? sizeof(int16_t) : sizeof(float);
  int16_t* d_pcm = pcm16 ? static_cast<int16_t*>(ctx->d_scratch[12]) : nullptr;
  int64_t up_to = 0;  // samples [0, up_to) of every stream are on their way to the device
  for (int c = 0; c < chunks && !rc; ++c) {
    const int64_t k0 = frames * c / chunks, k1 = frames * (c + 1) / chunks, fc = k1 - k0;
    const size_t r0 = static_cast<size_t>(n_streams) * k0;  // rows before this chunk: its arrays are [streams][fc] from there
    // ---- up: the samples this chunk's frames read ----
    const int64_t need = c + 1 == chunks ? n_samples : std::min<int64_t>(n_samples, ((k1 - 1) * hop + W + 1) & ~int64_t(1));
    if (need > up_to) {
      const size_t w = static_cast<size_t>(need - up_to);
      if (pcm16) {
        ok(cudaMemcpy2DAsync(d_pcm + up_to, sizeof(int16_t) * n_samples, static_cast<const int16_t*>(h_x) + up_to, sizeof(int16_t) * n_samples,
                             sizeof(int16_t) * w, n_streams, cudaMemcpyHostToDevice, s_up));
        pcm16_to_float_2d_kernel<<<dim3(static_cast<unsigned>((w / 2 + 256) / 256), static_cast<unsigned>(n_streams)), 256, 0, s_up>>>(
            d_pcm + up_to, d_x + up_to, static_cast<int64_t>(w), n_samples);
        ctx->launches += 1;
      } else {
        ok(cudaMemcpy2DAsync(d_x + up_to, in_sz * n_samples, static_cast<const float*>(h_x) + up_to, in_sz * n_samples, in_sz * w, n_streams,
                             cudaMemcpyHostToDevice, s_up));
      }
      up_to = need;
    }
    ok(cudaEventRecord(ev[3 * c], s_up));
    // ---- analysis and tuned lists of frames [k0, k1) ----
    ok(cudaStreamWaitEvent(s_ana, ev[3 * c], 0));
    const int64_t span = (fc - 1) * hop + W;
    if (use64)
      rc = stft64_peaks(ctx, ctx->plan_rt, d_x + k0 * hop, span, n_streams, n_samples, d_pbin + r0 * P, reinterpret_cast<double*>(d_poff + r0 * P * esz),
                        reinterpret_cast<double*>(d_pdb + r0 * P * esz), d_pcnt + r0, s_ana);
    else
      rc = stft_analyze(ctx, ctx->plan_rt, d_x + k0 * hop, span, n_streams, n_samples, nullptr, 0, d_pbin + r0 * P,
                        reinterpret_cast<float*>(d_poff + r0 * P * esz), reinterpret_cast<float*>(d_pdb + r0 * P * esz), d_pcnt + r0, s_ana);
    if (rc) break;
    rc = pitch_prefix_device(ctx, pp, n_streams * fc, d_pbin + r0 * P, d_poff + r0 * P * esz, d_pdb + r0 * P * esz, use64, d_pcnt + r0, d_trunc,
                             d_tuned + r0 * pp.PT, d_tcount + r0, s_ana);
    if (rc) break;
    ok(cudaEventRecord(ev[3 * c + 1], s_ana));
    // ---- tracking of frames [k0, k1), every stream taking its state up where the chunk before left it ----
    ok(cudaStreamWaitEvent(s_trk, ev[3 * c + 1], 0));
    rc = pitch_track_device(ctx, pp, n_streams, d_tuned + r0 * pp.PT, d_tcount + r0, k0, k1, chunks > 1 ? d_state : nullptr,
                            static_cast<Segment*>(ctx->d_scratch[6]), capacity, tps, d_begin, d_end, d_s0, d_len, d_stats, d_counter, d_overflow, s_trk);
    if (rc) break;
    ok(cudaEventRecord(ev[3 * c + 2], s_trk));
    // ---- rendering of the chunk's tiles and its audio back to the host ----
    ok(cudaStreamWaitEvent(s_out, ev[3 * c + 2], 0));
    const int64_t t0 = tile_of[k0], nt = tile_of[k1] - t0;
    TileArgs ta{static_cast<const Segment*>(ctx->d_scratch[6]), d_begin, d_s0, d_len, nt * n_streams, d_out, 0, d_end, tps, t0, nt};
    tracks_kernel<false><<<static_cast<unsigned>((nt * n_streams + 3) / 4), 128, 0, s_out>>>(ta);
    ctx->launches += 1;
    const int64_t o0 = c == 0 ? 0 : std::min<int64_t>(n_samples, static_cast<int64_t>(W) - 1 + k0 * hop);
    const int64_t o1 = c + 1 == chunks ? n_samples : std::min<int64_t>(n_samples, static_cast<int64_t>(W) - 1 + k1 * hop);
    if (o1 > o0)
      ok(cudaMemcpy2DAsync(h_out + 2 * o0, sizeof(float) * 2 * n_samples, d_out + 2 * o0, sizeof(float) * 2 * n_samples, sizeof(float) * 2 * (o1 - o0),
                           n_streams, cudaMemcpyDeviceToHost, s_out));
  }
  ok(cudaGetLastError());
  for (int g = 0; g < 4; ++g) ok(cudaStreamSynchronize(ctx->pipe_streams[g]));
  cleanup();
  if (rc) return rc;
  if (err != cudaSuccess) return fail(ctx, CA_ERR_CUDA, std::string("round trip (time chunks): ") + cudaGetErrorString(err));
  CA_CUDA_OK(ctx, cudaMemcpy(h_misc, d_stats, b_misc, cudaMemcpyDeviceToHost));
  *needed = h_misc[2 * n_streams];
  *overflowed = *reinterpret_cast<const int32_t*>(h_misc + 2 * n_streams + 1) != 0;
  if (!*overflowed && h_stats) memcpy(h_stats, h_misc, sizeof(int64_t) * 2 * n_streams);
  ctx->rt_truncated_frames = h_misc[2 * n_streams + 2];
  return CA_OK;
}

}  // namespace

int roundtrip_host(ca_ctx* ctx, const ca_roundtrip_params* params, const void* h_x, bool pcm16, int64_t n_samples, int64_t n_streams,
                   float* h_out, int64_t* h_stats) {
  if (!params || !h_x || !h_out || n_samples < 0 || n_streams < 0)
    return fail(ctx, CA_ERR_INVALID, "ca_resynth_roundtrip_host: null argument or negative size");
  ca_roundtrip_params prm = *params;
  if (prm.env.attack_itp != CA_ITP_LINEAR || prm.env.decay_itp != CA_ITP_LINEAR || prm.env.release_itp != CA_ITP_LINEAR)
    return fail(ctx, CA_ERR_UNSUPPORTED, "ca_resynth_roundtrip_host: LINEAR envelopes only (as rt.resynth.lib.cpp:75-81)");
  if (prm.max_channels <= 0) prm.max_channels = 254;
  if (prm.stft.max_peaks <= 0) prm.stft.max_peaks = 200;
  if (prm.stft.max_peaks > 16384) return fail(ctx, CA_ERR_UNSUPPORTED, "ca_resynth_roundtrip_host: stft.max_peaks above 16384");
  {
    const ca_pitch_options& o = prm.opts;
    if (o.autotune != CA_AUTOTUNE_NONE && o.autotune != CA_AUTOTUNE_TABLE && o.autotune != CA_AUTOTUNE_SCALE)
      return fail(ctx, CA_ERR_INVALID, "ca_resynth_roundtrip_host: opts.autotune is not a ca_autotune value");
    if (o.autotune == CA_AUTOTUNE_TABLE && (o.n_table < 0 || o.n_table > 4096 || (o.n_table > 0 && !o.table)))
      return fail(ctx, CA_ERR_INVALID, "ca_resynth_roundtrip_host: opts.table must hold 0 ... 4096 allowed pitches");
    if (o.autotune == CA_AUTOTUNE_SCALE && (o.n_table < 2 || o.n_table > 16 || !o.table))
      return fail(ctx, CA_ERR_INVALID, "ca_resynth_roundtrip_host: opts.table must hold the 2 ... 16 pitches of the scale");
    if (o.autotune != CA_AUTOTUNE_NONE)
      for (int i = 1; i < o.n_table; ++i)
        if (!(o.table[i - 1] <= o.table[i]))
          return fail(ctx, CA_ERR_INVALID, "ca_resynth_roundtrip_host: opts.table must be ascending");
  }
  int rc = stft_plan(ctx, ctx->plan_rt, &prm.stft);
  if (rc) return rc;
  ctx->plan_rt.freq_as_offset = true;  // float32 form: sub-bin offsets instead of Hz, so that k + p is formed in double
  // The peaks come from the FP64 form of the analysis (stft64.cu: the reference analyses in double and the resynthesised
  // phase integrates the peak frequencies); CA_RT_F32_ANALYSIS=1 selects the float32 kernels for A/B, and they also serve the
  // shapes the FP64 kernel does not take (32768-point frames, odd hops).
  static const bool env_f32 = getenv("CA_RT_F32_ANALYSIS") != nullptr;
  const bool use64 = !env_f32 && stft64_supported(ctx->plan_rt) && n_samples % 2 == 0;
  if (use64) {
    rc = stft64_plan(ctx, ctx->plan_rt);
    if (rc) return rc;
  }
  const size_t esz = use64 ? sizeof(double) : sizeof(float);
  const int W = prm.stft.window, hop = prm.stft.hop, P = prm.stft.max_peaks;
  const int64_t frames = ca_stft_num_frames(n_samples, W, hop);
  const size_t total = static_cast<size_t>(n_streams) * n_samples;
  ctx->rt_truncated_frames = 0;
  if (total == 0) return CA_OK;
  const size_t rows = static_cast<size_t>(frames) * n_streams;
  static const bool timing = getenv("CA_RT_TIMING") != nullptr;  // stage times on stderr (profiles/time_roundtrip.py)
  // the pitch pipeline runs on the device (pitch.cu); host_threads != 0 or CA_RT_HOST_PITCH=1 selects the host form (A/B)
  static const bool env_host_pitch = getenv("CA_RT_HOST_PITCH") != nullptr;
  const bool host_pitch = env_host_pitch || prm.host_threads != 0;
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
  if (!rc && pcm16) rc = ensure_scratch(ctx, 12, sizeof(int16_t) * total);
  if (!rc) rc = ensure_scratch(ctx, 7, out_bytes);
  rc = rc ? rc : ensure_scratch(ctx, 2, std::max<size_t>(16, sizeof(int32_t) * rows * P));
  for (int i = 3; i <= 4 && !rc; ++i) rc = ensure_scratch(ctx, i, std::max<size_t>(16, esz * rows * P));
  if (!rc) rc = ensure_scratch(ctx, 5, std::max<size_t>(16, sizeof(int32_t) * rows));
  if (!rc) rc = ensure_pinned(ctx, 0, std::max<size_t>(16, sizeof(int32_t) * rows));
  if (rc) return rc;
  static const bool env_sequential = getenv("CA_RT_SEQUENTIAL") != nullptr;  // one stream group, for A/B
  if (rows && !host_pitch && !timing && !env_sequential) {
    bool overflowed = false;
    int64_t needed = 0;
    static const bool by_streams = [] { const char* e = getenv("CA_RT_PIPELINE"); return e && !strcmp(e, "streams"); }();  // A/B: the stream-group pipeline
    rc = by_streams ? roundtrip_device_pipelined(ctx, prm, h_x, pcm16, n_samples, n_streams, frames, h_out, h_stats, use64, &overflowed, &needed)
                    : roundtrip_device_chunked(ctx, prm, h_x, pcm16, n_samples, n_streams, frames, h_out, h_stats, use64, &overflowed, &needed);
    if (rc || !overflowed) return rc;
    rc = ensure_scratch(ctx, 6, sizeof(Segment) * static_cast<size_t>(needed));  // rare: redo in one group with the exact size
    if (rc) return rc;
  }
  int32_t* h_cnt = static_cast<int32_t*>(ctx->h_pinned[0]);
  int cmax = 0;
  if (rows) {
    float* d_x = static_cast<float*>(ctx->d_scratch[0]);
    rc = upload_samples(ctx, h_x, pcm16, 0, total, d_x, st);
    if (rc) return rc;
    if (use64)
      rc = stft64_peaks(ctx, ctx->plan_rt, d_x, n_samples, n_streams, n_samples, static_cast<int32_t*>(ctx->d_scratch[2]),
                        static_cast<double*>(ctx->d_scratch[3]), static_cast<double*>(ctx->d_scratch[4]),
                        static_cast<int32_t*>(ctx->d_scratch[5]), st);
    else
      rc = stft_analyze(ctx, ctx->plan_rt, d_x, n_samples, n_streams, n_samples, nullptr, 0, static_cast<int32_t*>(ctx->d_scratch[2]),
                        static_cast<float*>(ctx->d_scratch[3]), static_cast<float*>(ctx->d_scratch[4]),
                        static_cast<int32_t*>(ctx->d_scratch[5]), st);
    if (rc) return rc;
    if (!host_pitch) {
      if (timing) {
        CA_CUDA_OK(ctx, cudaStreamSynchronize(st));
        lap("analysis (H2D, K2)");
      }
      return roundtrip_device_tail(ctx, prm, n_samples, n_streams, frames, h_out, h_stats, timing, use64, lap);
    }
    CA_CUDA_OK(ctx, cudaMemcpyAsync(h_cnt, ctx->d_scratch[5], sizeof(int32_t) * rows, cudaMemcpyDeviceToHost, st));
    CA_CUDA_OK(ctx, cudaStreamSynchronize(st));
    for (size_t r = 0; r < rows; ++r) cmax = std::max(cmax, h_cnt[r]);
    cmax = std::min(cmax, P);
  }
  const int PS = std::max(cmax, 1);  // compacted entries per frame
  for (int i = 1; i <= 3 && !rc; ++i) rc = ensure_pinned(ctx, i, std::max<size_t>(16, (i == 1 ? sizeof(int32_t) : esz) * rows * PS));
  if (rc) return rc;
  if (rows && cmax > 0)
    for (int i = 0; i < 3; ++i) {
      const size_t e = i == 0 ? sizeof(int32_t) : esz;
      CA_CUDA_OK(ctx, cudaMemcpy2DAsync(ctx->h_pinned[1 + i], e * PS, ctx->d_scratch[2 + i], e * P, e * PS, rows, cudaMemcpyDeviceToHost, st));
    }
  CA_CUDA_OK(ctx, cudaMemsetAsync(ctx->d_scratch[7], 0, out_bytes, st));  // samples before the first frame are silent
  CA_CUDA_OK(ctx, cudaStreamSynchronize(st));
  const int32_t* pbin = static_cast<const int32_t*>(ctx->h_pinned[1]);
  lap("analysis (H2D, K2, D2H peaks)");

  // ---- 2. pitch pipeline on host threads, one stream at a time per thread ----
  PitchParams pp = make_pitch_params(prm, n_samples, frames, PS);
  pp.cap_raw = P;  // the lists were compacted to PS entries per frame; the analysis kept at most P
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
    const size_t o = static_cast<size_t>(s) * frames * PS;
    if (use64)
      run_stream(pp, pbin + o, static_cast<const double*>(ctx->h_pinned[2]) + o, static_cast<const double*>(ctx->h_pinned[3]) + o,
                 h_cnt + s * frames, results[s]);
    else
      run_stream(pp, pbin + o, static_cast<const float*>(ctx->h_pinned[2]) + o, static_cast<const float*>(ctx->h_pinned[3]) + o,
                 h_cnt + s * frames, results[s]);
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
  for (const auto& r : results) ctx->rt_truncated_frames += r.truncated;
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
