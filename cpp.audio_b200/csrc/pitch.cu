// pitch.cu — the pitch pipeline of the round trip on the device.
//
// Reference: RtResynth::step (source/rt.resynth.lib.cpp:1670-1759) runs, per analysis frame, the functions of
// source/rt.resynth.lib.algo.cpp on the peak list and synthesize_sounds (:264-382) turns the result into NoteOn / NoteChange /
// NoteOff for a pool of voices. The formulas live in pitch.cuh (shared with the host A/B path); this file maps them to the GPU:
//
//   pitch_prefix_kernel  one THREAD per frame: peaks -> tuned list. Stateless, so all frames of all streams run at once and the
//                        double-precision log2 / pow of every peak leave the sequential part.
//   pitch_track_kernel   one WARP per stream walks the frames in order (the played-notes state is carried from frame to frame,
//                        :1710-1758). Inside a frame lanes = tuned entries / notes: greedy monotone matching by bisection with an
//                        exact serial fallback, note-off / note-change / note-on, ordering by rank counting, one Segment per
//                        sounding note and render tile (the closed forms of tracks.cuh), pool compaction by ballot prefix sums.
//                        Segment storage comes from a global bump allocator in chunks (one atomic per ~500 tiles); when the
//                        capacity guess is too small the kernel keeps counting and the host re-runs it with the exact size.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "ctx.hpp"
#include "pitch.cuh"

namespace ca {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kSegChunk = 8192;  // segments per allocation chunk (>= kMaxNotes)
constexpr int kRing = 4;          // frames of tuned entries in flight

template <class R>
struct PrefixArgs {
  PitchParams p;
  const int32_t* peak_bin;   // [rows][P]
  const R* peak_off;         // R = float (float32 analysis kernels) or double (stft64.cu)
  const R* peak_db;
  const int32_t* peak_count; // [rows]
  Tuned* tuned;              // [rows][P]
  int32_t* tcount;           // [rows]
  int64_t rows;
};

template <class R>
__global__ void __launch_bounds__(128) pitch_prefix_kernel(const __grid_constant__ PrefixArgs<R> a) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= a.rows) return;
  const int P = a.p.P;
  const int cnt = min(a.peak_count[r], P);
  a.tcount[r] = pitch_prefix(a.p, a.peak_bin + r * P, a.peak_off + r * P, a.peak_db + r * P, cnt, a.tuned + r * P);
}

struct TrackArgs {
  PitchParams p;
  const Tuned* tuned;        // [streams*frames][P]
  const int32_t* tcount;     // [streams*frames]
  Segment* segs;
  unsigned long long* seg_counter;  // bump allocator (segments handed out so far, chunk granularity)
  int64_t seg_capacity;
  int64_t* tile_begin;       // [streams*tiles_per_stream]
  int64_t* tile_end;
  int64_t* tile_s0;
  int32_t* tile_len;
  int64_t tiles_per_stream;
  int64_t* stats;            // [streams][2]: notes started, notes dropped
  int32_t* overflow;         // set when a tile did not get its segments
};

// shared memory of one warp (= one stream)
struct TrackSmem {
  Note notes[kMaxNotes];
  Tuned ring[kRing][32];        // first 32 entries of the lists of the coming frames (bulk-copy staging)
  uint64_t bar[kRing];          // one mbarrier per ring slot
  double pl_pitch[2][kMaxNotes];   // pitch of played[j] (double buffered with the list)
  int32_t played[2][kMaxNotes];    // indices into notes of the tracked notes, ascending pitch
  int32_t remap[kMaxNotes];
  int32_t t_lo[kMaxTuned];
  int32_t t_match[kMaxTuned];   // index into played, or -1
  float t_fv[kMaxTuned];        // float(volume sum) of the candidates for a new note
  uint8_t t_new[kMaxTuned];
  uint8_t keep[kMaxNotes];
};

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// CA_PITCH_TIMELINE builds (profiles/build_pitch_timeline.sh): clocks per phase of stream 0, printed when the kernel ends
#ifdef CA_PITCH_TIMELINE
#define TL(i)                                   \
  do {                                          \
    const long long now_ = clock64();           \
    tl_acc[i] += now_ - tl_t;                   \
    tl_t = now_;                                \
  } while (0)
#else
#define TL(i) do { } while (0)
#endif

__global__ void __launch_bounds__(32) pitch_track_kernel(const __grid_constant__ TrackArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TrackSmem& sm = *reinterpret_cast<TrackSmem*>(smem_raw);
  const PitchParams& p = a.p;
  const int lane = threadIdx.x;
  const unsigned lt = lanemask_lt();
  const int64_t s = blockIdx.x;
  const int P = p.P;
  const double dtrack = p.max_track;

  int N = 0;    // sounding notes (tracked or releasing)
  int Np = 0;   // tracked notes
  int64_t started = 0, dropped = 0;
  int64_t tile = s * a.tiles_per_stream;
  int64_t chunk_next = 0, chunk_end = 0;  // this stream's current allocation chunk
  // the played list is double buffered: a frame builds the next list in the other buffer
  int32_t* played = sm.played[0];
  int32_t* next_played = sm.played[1];
  double* pl_pitch = sm.pl_pitch[0];
  double* np_pitch = sm.pl_pitch[1];

  const Tuned* tu = a.tuned + s * p.frames * P;
  const int32_t* tc = a.tcount + s * p.frames;
  // Global-memory latency stays off the critical path: the entry counts of 32 frames are fetched a block of frames ahead
  // (lane l holds the count of frame block + l), and the first 32 entries of every frame's list are staged in shared memory
  // kRing - 1 frames ahead by bulk asynchronous copies (completion on an mbarrier, not on the warp's scoreboard).
  auto load_counts = [&](int64_t k0) { return k0 + lane < p.frames ? tc[k0 + lane] : 0; };
  int cnt_cur = load_counts(0), cnt_next = load_counts(32);
  const int n_staged = min(P, 32);
  const uint32_t staged_bytes = static_cast<uint32_t>(n_staged * sizeof(Tuned));
  if (lane == 0) {
    for (int i = 0; i < kRing; ++i) mbar_init(&sm.bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
    for (int f = 0; f < kRing - 1 && f < p.frames; ++f) {
      mbar_expect_tx(&sm.bar[f], staged_bytes);
      bulk_g2s(sm.ring[f], tu + static_cast<int64_t>(f) * P, staged_bytes, &sm.bar[f]);
    }
  }
  __syncwarp();

#ifdef CA_PITCH_TIMELINE
  long long tl_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long tl_t = clock64();
#endif
  for (int64_t k = 0; k < p.frames; ++k, tu += P) {
    TL(7);
    if ((k & 31) == 0 && k > 0) {
      cnt_cur = cnt_next;
      cnt_next = load_counts(k + 32);
    }
    const int T = __shfl_sync(kFull, cnt_cur, static_cast<int>(k & 31));
    const int slot = static_cast<int>(k % kRing);
    mbar_wait(&sm.bar[slot], static_cast<uint32_t>((k / kRing) & 1));
    Tuned e0{};
    if (lane < n_staged) e0 = sm.ring[slot][lane];
    __syncwarp();
    if (lane == 0 && k + kRing - 1 < p.frames) {  // the slot read a frame ago is free: stage frame k + kRing - 1 into it
      const int nslot = static_cast<int>((k + kRing - 1) % kRing);
      // (no fence.proxy.async, a MEMBAR.ALL.CTA: the slot's last generic-proxy accesses were loads whose values were consumed a
      // whole frame ago, with warp barriers since)
      mbar_expect_tx(&sm.bar[nslot], staged_bytes);
      bulk_g2s(sm.ring[nslot], tu + static_cast<int64_t>(kRing - 1) * P, staged_bytes, &sm.bar[nslot]);
    }
    const int64_t n_k = static_cast<int64_t>(p.W) - 1 + k * p.hop;  // events of frame k act from this sample on
    const int len_frame = static_cast<int>(min(static_cast<int64_t>(p.hop), p.n_samples - n_k));
    auto entry = [&](int i0, int i) -> Tuned { return i0 == 0 ? e0 : tu[i]; };
    TL(8);

    // ---------------- tracking: greedy monotone matching of the tuned pitches against the played ones ----------------
    // serial form: it = 0; for each entry i (ascending pitch): skip played pitches below pitch_i - d; if the next one is
    // <= pitch_i + d it is entry i's note (it++). When no two entries compete for a played note the answer is the lower
    // bound of every entry; a competing pair is detected by comparing neighbours and handled serially (exact).
    int kept = 0;
    bool any_new = false;
    if (T <= 32 && Np <= 32) {
      // ---- register form: lane i = tuned entry i, lane j = played note j; flags travel as ballots ----
      const bool has = lane < T;
      int lo = 0, mt = -1;
      if (has) {
        const double lo_t = e0.pitch - dtrack, hi_t = e0.pitch + dtrack;
        for (int j = 0; j < Np; ++j) lo += pl_pitch[j] < lo_t ? 1 : 0;  // broadcast loads, independent of each other
        if (lo < Np && pl_pitch[lo] <= hi_t) mt = lo;
      }
      const int prev_lo = __shfl_up_sync(kFull, lo, 1), prev_mt = __shfl_up_sync(kFull, mt, 1);
      const bool bad = has && lane > 0 && lo < prev_lo + (prev_mt >= 0 ? 1 : 0);
      if (__any_sync(kFull, bad)) {
        if (lane == 0) {
          int it = 0;
          for (int i = 0; i < T; ++i) {
            const double tp = sm.ring[slot][i].pitch;
            sm.t_match[i] = -1;
            for (; it < Np; ++it) {
              const double pp = pl_pitch[it];
              if (pp < tp - dtrack) continue;
              if (pp <= tp + dtrack) {
                sm.t_match[i] = it;
                ++it;
              }
              break;
            }
          }
        }
        __syncwarp();
        if (has) mt = sm.t_match[lane];
      }
      TL(9);
      const unsigned keepmask = __reduce_or_sync(kFull, mt >= 0 ? 1u << mt : 0u);
      kept = __popc(keepmask);
      // NoteOff for the played notes nobody matched; the others stay, in order
      if (lane < Np) {
        const int note = played[lane];
        if ((keepmask >> lane) & 1) next_played[__popc(keepmask & lt)] = note;
        else note_off(p, sm.notes[note], n_k);
      }
      TL(10);
      // NoteChange for the matched entries; candidates for NoteOn
      if (has) {
        if (mt >= 0) {
          note_change(p, sm.notes[played[mt]], e0);
          np_pitch[__popc(keepmask & ((1u << mt) - 1u))] = e0.pitch;
        }
        const bool is_new = mt < 0 && e0.volume > 0;
        any_new = is_new;
        sm.t_new[lane] = is_new;
        sm.t_fv[lane] = static_cast<float>(e0.vol);
      }
      any_new = __any_sync(kFull, any_new);
      TL(0);
    } else {
      // ---- general form: loops over groups of 32 entries / notes, flags in shared memory ----
      bool bad = false;
      for (int i0 = 0; i0 < T; i0 += 32) {
        const int i = i0 + lane;
        if (i < T) {
          const double tp = entry(i0, i).pitch;
          const double lo_t = tp - dtrack, hi_t = tp + dtrack;
          int lo = 0, hi = Np;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (pl_pitch[mid] < lo_t) lo = mid + 1;
            else hi = mid;
          }
          sm.t_lo[i] = lo;
          sm.t_match[i] = (lo < Np && pl_pitch[lo] <= hi_t) ? lo : -1;
        }
      }
      for (int j = lane; j < Np; j += 32) sm.keep[j] = 0;
      __syncwarp();
      for (int i = lane; i < T; i += 32)
        if (i > 0 && sm.t_lo[i] < sm.t_lo[i - 1] + (sm.t_match[i - 1] >= 0 ? 1 : 0)) bad = true;
      if (__any_sync(kFull, bad)) {
        if (lane == 0) {
          int it = 0;
          for (int i = 0; i < T; ++i) {
            const double tp = tu[i].pitch;
            sm.t_match[i] = -1;
            for (; it < Np; ++it) {
              const double pp = pl_pitch[it];
              if (pp < tp - dtrack) continue;
              if (pp <= tp + dtrack) {
                sm.t_match[i] = it;
                ++it;
              }
              break;
            }
          }
        }
        __syncwarp();
      }
      for (int i = lane; i < T; i += 32)
        if (sm.t_match[i] >= 0) sm.keep[sm.t_match[i]] = 1;
      __syncwarp();
      // NoteOff for the played notes nobody matched; the others stay, in order (remap: position in the next list)
      for (int j0 = 0; j0 < Np; j0 += 32) {
        const int j = j0 + lane;
        const bool in = j < Np;
        const bool kp = in && sm.keep[j];
        if (in && !kp) note_off(p, sm.notes[played[j]], n_k);
        const unsigned m = __ballot_sync(kFull, kp);
        if (kp) {
          const int pos = kept + __popc(m & lt);
          next_played[pos] = played[j];
          sm.remap[j] = pos;
        }
        kept += __popc(m);
      }
      __syncwarp();
      // NoteChange for the matched entries; candidates for NoteOn
      for (int i0 = 0; i0 < T; i0 += 32) {
        const int i = i0 + lane;
        if (i < T) {
          const Tuned e = entry(i0, i);
          const int mt = sm.t_match[i];
          if (mt >= 0) {
            note_change(p, sm.notes[played[mt]], e);
            np_pitch[sm.remap[mt]] = e.pitch;
          }
          const bool is_new = mt < 0 && e.volume > 0;
          sm.t_new[i] = is_new;
          sm.t_fv[i] = static_cast<float>(e.vol);
          any_new |= is_new;
        }
      }
      any_new = __any_sync(kFull, any_new);
      TL(0);
    }
    __syncwarp();
    TL(2);
    int M = kept;
    if (any_new) {
      // new notes take channels in the order of decreasing volume (stable), :264-382; rank = position in that order
      int n_new = 0;
      for (int i0 = 0; i0 < T; i0 += 32) n_new += __popc(__ballot_sync(kFull, i0 + lane < T && sm.t_new[i0 + lane]));
      const int room = max(0, p.max_channels - N);
      const int accepted = min(n_new, room);
      for (int i0 = 0; i0 < T; i0 += 32) {
        const int i = i0 + lane;
        if (i < T && sm.t_new[i]) {
          const float fv = sm.t_fv[i];
          int rank = 0;
          for (int j = 0; j < T; ++j)
            if (sm.t_new[j] && (sm.t_fv[j] > fv || (sm.t_fv[j] == fv && j < i))) ++rank;
          if (rank < accepted) {
            const Tuned e = entry(i0, i);
            sm.notes[N + rank] = note_on(p, e, started + rank + 1, n_k);
            next_played[kept + rank] = N + rank;
            np_pitch[kept + rank] = e.pitch;
          }
        }
      }
      started += n_new;
      dropped += n_new - accepted;
      N += accepted;
      M += accepted;
      __syncwarp();
    }
    TL(3);
    // ---------------- played = next_played in ascending pitch (stable) ----------------
    bool unsorted = false;
    for (int j = lane; j < M; j += 32)
      if (j > 0 && np_pitch[j - 1] > np_pitch[j]) unsorted = true;
    if (__any_sync(kFull, unsorted)) {
      for (int j = lane; j < M; j += 32) {
        const double pj = np_pitch[j];
        int rank = 0;
        for (int l = 0; l < M; ++l) {
          const double pl = np_pitch[l];
          if (pl < pj || (pl == pj && l < j)) ++rank;
        }
        played[rank] = next_played[j];
        pl_pitch[rank] = pj;
      }
      __syncwarp();
    } else {  // already in order (the usual case): the buffers trade places
      int32_t* ti = played;
      played = next_played;
      next_played = ti;
      double* td = pl_pitch;
      pl_pitch = np_pitch;
      np_pitch = td;
    }
    Np = M;

    TL(4);
    // ---------------- one segment per sounding note and tile, note states advance ----------------
    for (int off = 0; off < max(len_frame, 0) || off == 0; off += kMaxTile) {
      const int len = min(kMaxTile, len_frame - off);
      const int64_t n_k_tile = n_k + off;
      int64_t base = 0;
      bool ok = true;
      if (len > 0 && N > 0) {
        if (chunk_next + N > chunk_end) {  // warp-uniform
          unsigned long long got = 0;
          if (lane == 0) got = atomicAdd(a.seg_counter, static_cast<unsigned long long>(kSegChunk));
          got = __shfl_sync(kFull, got, 0);
          chunk_next = static_cast<int64_t>(got);
          chunk_end = chunk_next + kSegChunk;
        }
        base = chunk_next;
        chunk_next += N;
        ok = base + N <= a.seg_capacity;
      }
      if (lane == 0) {
        const bool filled = ok && len > 0 && N > 0;
        a.tile_begin[tile] = filled ? base : 0;
        a.tile_end[tile] = filled ? base + N : 0;
        a.tile_s0[tile] = s * p.n_samples + n_k_tile;
        a.tile_len[tile] = max(len, 0);
        if (!ok) *a.overflow = 1;
      }
      ++tile;
      if (len <= 0) break;
      for (int idx = lane; idx < N; idx += 32) {
        const Segment g = emit_segment(p, sm.notes[idx], n_k_tile, len);
        if (ok) a.segs[base + idx] = g;
      }
      __syncwarp();
    }

    TL(5);
    // ---------------- released notes whose envelope is over leave the pool ----------------
    if (len_frame > 0) {
      const int64_t n_next = n_k + len_frame;
      bool dead = false;
      if (N > Np)  // only releasing notes can be done; every other note is in the played list
        for (int i = lane; i < N; i += 32) dead |= note_done(sm.notes[i], n_next);
      if (N > Np && __any_sync(kFull, dead)) {
        int w = 0;
        for (int i0 = 0; i0 < N; i0 += 32) {
          const int i = i0 + lane;
          Note tmp;
          bool alive = false;
          if (i < N) {
            tmp = sm.notes[i];
            alive = !note_done(tmp, n_next);
          }
          const unsigned m = __ballot_sync(kFull, alive);
          const int pos = w + __popc(m & lt);
          if (i < N) sm.remap[i] = alive ? pos : -1;
          __syncwarp();
          if (alive && pos != i) sm.notes[pos] = tmp;
          __syncwarp();
          w += __popc(m);
        }
        for (int j = lane; j < Np; j += 32) played[j] = sm.remap[played[j]];
        N = w;
        __syncwarp();
      }
    }
  }
#ifdef CA_PITCH_TIMELINE
  TL(6);
  if (s == 0 && lane == 0) {
    printf("front detail: stage %lld  match %lld  noteoff %lld  change %lld  sync %lld\n", tl_acc[8] / p.frames, tl_acc[9] / p.frames,
           tl_acc[10] / p.frames, tl_acc[0] / p.frames, tl_acc[2] / p.frames);
    // FP64 latencies of this GPU (dependent chains)
    double x = 1.0 + 1e-9 * static_cast<double>(tl_acc[1] & 7), y = 1.0000001;
    long long t0 = clock64();
    for (int i = 0; i < 256; ++i) x = fma(x, y, 1e-9);
    long long t1 = clock64();
    for (int i = 0; i < 64; ++i) x = 1.0 / (x + 1.0);
    long long t2 = clock64();
    for (int i = 0; i < 64; ++i) x = log(x + 2.0);
    long long t3 = clock64();
    for (int i = 0; i < 64; ++i) x = expm1(x * 0.5);
    long long t4 = clock64();
    printf("fp64 dependent latency: fma %lld  div %lld  log %lld  expm1 %lld clocks (%g)\n", (t1 - t0) / 256, (t2 - t1) / 64,
           (t3 - t2) / 64, (t4 - t3) / 64, x);
  }
  if (s == 0 && lane == 0)
    printf("pitch_track clocks/frame: noteon %lld  sort %lld  segments %lld  tail %lld  compaction+loop %lld\n", tl_acc[3] / p.frames,
           tl_acc[4] / p.frames, tl_acc[5] / p.frames, tl_acc[6] / p.frames, tl_acc[7] / p.frames);
#endif
  if (lane == 0) {
    a.stats[2 * s] = started;
    a.stats[2 * s + 1] = dropped;
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Two warps per stream (the default). Tracking needs nothing from the note arithmetic: the played list is keyed by the tuned
// pitches, and a released note leaves the pool after n_off + release + 17 samples, all integers known to the tracker. So warp 0
// (tracker) runs frame j while warp 1 (voices) applies frame j-1's events and writes its segments; they meet at one CTA barrier
// per frame and exchange a small record: the pool slot of every tuned entry (note-change / note-on) and the slots released.
// Both warps compact the pool by the same rule, so slot numbers stay in step without being exchanged. A frame then costs
// max(tracking, note arithmetic) instead of their sum.
// ---------------------------------------------------------------------------------------------------------------
struct GlideC {  // what Glide::start leaves behind for a changed note
  double from, to, K, L, c1, P_m1, f_m1;
  int32_t m;
};
struct FrameRec {
  int32_t T;          // tuned entries of the frame
  int32_t N_before;   // pool size before the frame's note-ons: slots >= N_before are new notes
  int32_t N_after;    // pool size after them (before compaction)
  int32_t n_off;      // released slots
  int32_t N_compact;  // pool size after the frame's compaction, -1: nothing left the pool
  int64_t id_base;    // notes started before this frame: note-on of slot s gets id id_base + (s - N_before) + 1
  int32_t slot[kMaxTuned];   // per tuned entry: pool slot, or -1
  int32_t off[kMaxNotes];    // released slots
  int32_t remap[kMaxNotes];  // compaction: new slot of every pool entry, -1: gone (valid when N_compact >= 0)
  Tuned entry[kMaxTuned];    // the tuned entries that carry an event (slot >= 0)
  GlideC glide[kMaxTuned];   // three-warp form: the glide of every changed note, computed by the glide warp
};

struct Track2Smem {
  Note notes[kMaxNotes];           // voices warp
  Tuned ring[kRing][32];
  uint64_t bar[kRing];
  FrameRec rec[3];
  double f_last[kMaxNotes];        // glide warp (three-warp form): last emitted increment of every slot
  // tracker warp
  double pl_pitch[2][kMaxNotes];
  int32_t played[2][kMaxNotes];
  int64_t tr_noff[kMaxNotes];      // sample of the first release step, -1 while tracked
  int32_t tr_release[kMaxNotes];
  int32_t tr_nmc[kMaxNotes];       // next_min_change of the slot
  int32_t remap[kMaxNotes];        // tracker scratch
  int32_t t_lo[kMaxTuned];
  int32_t t_match[kMaxTuned];
  float t_fv[kMaxTuned];
  uint8_t t_new[kMaxTuned];
  uint8_t keep[kMaxNotes];
};

template <int kWarps>
__global__ void __launch_bounds__(32 * kWarps) pitch_track2_kernel(const __grid_constant__ TrackArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Track2Smem& sm = *reinterpret_cast<Track2Smem*>(smem_raw);
  const PitchParams& p = a.p;
  const int lane = threadIdx.x & 31;
  const int role = threadIdx.x >> 5;  // 0: tracker; two warps: 1 = voices; three warps: 1 = glides, 2 = segments
  const bool is_tracker = role == 0;
  const unsigned lt = lanemask_lt();
  const int64_t s = blockIdx.x;
  const int P = p.P;
  const double dtrack = p.max_track;
  const Tuned* tu_base = a.tuned + s * p.frames * P;
  const int n_staged = min(P, 32);
  const uint32_t staged_bytes = static_cast<uint32_t>(n_staged * sizeof(Tuned));
  constexpr int kAhead = 2;  // frames staged ahead of the tracker (the voices warp still reads the slot of the frame before)

  // ---- tracker state ----
  int N = 0, Np = 0;
  int64_t started = 0, dropped = 0;
  int32_t* played = sm.played[0];
  int32_t* next_played = sm.played[1];
  double* pl_pitch = sm.pl_pitch[0];
  double* np_pitch = sm.pl_pitch[1];
  const int32_t* tc = a.tcount + s * p.frames;
  auto load_counts = [&](int64_t k0) { return k0 + lane < p.frames ? tc[k0 + lane] : 0; };
  int cnt_cur = 0, cnt_next = 0;
  // ---- voices state ----
  int64_t tile = s * a.tiles_per_stream;
  int64_t chunk_next = 0, chunk_end = 0;

  if (is_tracker) {
    cnt_cur = load_counts(0);
    cnt_next = load_counts(32);
    if (lane == 0) {
      for (int i = 0; i < kRing; ++i) mbar_init(&sm.bar[i], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      fence_proxy_async();
      for (int f = 0; f < kAhead && f < p.frames; ++f) {
        mbar_expect_tx(&sm.bar[f], staged_bytes);
        bulk_g2s(sm.ring[f], tu_base + static_cast<int64_t>(f) * P, staged_bytes, &sm.bar[f]);
      }
    }
  }
  __syncthreads();

#ifdef CA_PITCH_TIMELINE
  long long busy = 0;  // clocks this warp spent working (between leaving a barrier and arriving at the next)
#endif
  for (int64_t it = 0; it < p.frames + kWarps - 1; ++it) {
#ifdef CA_PITCH_TIMELINE
    const long long t_in = clock64();
#endif
    if (is_tracker) {
      // ================= tracker: frame k = it =================
      const int64_t k = it;
      if (k < p.frames) {
        const Tuned* tu = tu_base + k * P;
        FrameRec& rec = sm.rec[k % kWarps];
        if ((k & 31) == 0 && k > 0) {
          cnt_cur = cnt_next;
          cnt_next = load_counts(k + 32);
        }
        const int T = __shfl_sync(kFull, cnt_cur, static_cast<int>(k & 31));
        const int slot = static_cast<int>(k % kRing);
        mbar_wait(&sm.bar[slot], static_cast<uint32_t>((k / kRing) & 1));
        Tuned e0{};
        if (lane < n_staged) e0 = sm.ring[slot][lane];
        if (lane == 0 && k + kAhead < p.frames) {  // slot of frame k + kAhead: last read by the voices warp an iteration ago
          const int nslot = static_cast<int>((k + kAhead) % kRing);
          mbar_expect_tx(&sm.bar[nslot], staged_bytes);
          bulk_g2s(sm.ring[nslot], tu + static_cast<int64_t>(kAhead) * P, staged_bytes, &sm.bar[nslot]);
        }
        const int64_t n_k = static_cast<int64_t>(p.W) - 1 + k * p.hop;
        const int len_frame = static_cast<int>(min(static_cast<int64_t>(p.hop), p.n_samples - n_k));
        auto entry = [&](int i0, int i) -> Tuned { return i0 == 0 ? e0 : tu[i]; };
        int kept = 0, n_off = 0;
        bool any_new = false;
        if (T <= 32 && Np <= 32) {
          const bool has = lane < T;
          int lo = 0, mt = -1;
          // steady state of tracked tones: entry i continues played note i. Then the lower bound of entry i is i (the played
          // list is ascending: played[i-1] below the window, played[i] inside it) and no two entries compete.
          bool spec = true;
          if (has) {
            const double lo_t = e0.pitch - dtrack, hi_t = e0.pitch + dtrack;
            const double here = pl_pitch[min(lane, kMaxNotes - 1)];
            spec = lane < Np && here >= lo_t && here <= hi_t && (lane == 0 || pl_pitch[lane - 1] < lo_t);
          }
          const bool steady = T == Np && __all_sync(kFull, spec);
          if (steady) {
            if (has) lo = mt = lane;
          } else if (has) {
            const double lo_t = e0.pitch - dtrack, hi_t = e0.pitch + dtrack;
            for (int j = 0; j < Np; ++j) lo += pl_pitch[j] < lo_t ? 1 : 0;
            if (lo < Np && pl_pitch[lo] <= hi_t) mt = lo;
          }
          const int prev_lo = __shfl_up_sync(kFull, lo, 1), prev_mt = __shfl_up_sync(kFull, mt, 1);
          const bool bad = !steady && has && lane > 0 && lo < prev_lo + (prev_mt >= 0 ? 1 : 0);
          if (__any_sync(kFull, bad)) {
            if (lane == 0) {
              int itp = 0;
              for (int i = 0; i < T; ++i) {
                const double tp = sm.ring[slot][i].pitch;
                sm.t_match[i] = -1;
                for (; itp < Np; ++itp) {
                  const double pp = pl_pitch[itp];
                  if (pp < tp - dtrack) continue;
                  if (pp <= tp + dtrack) {
                    sm.t_match[i] = itp;
                    ++itp;
                  }
                  break;
                }
              }
            }
            __syncwarp();
            if (has) mt = sm.t_match[lane];
          }
          const unsigned keepmask = __reduce_or_sync(kFull, mt >= 0 ? 1u << mt : 0u);
          kept = __popc(keepmask);
          const unsigned playedmask = Np >= 32 ? kFull : ((1u << Np) - 1u);
          const unsigned offmask = playedmask & ~keepmask;
          n_off = __popc(offmask);
          if (lane < Np) {
            const int note = played[lane];
            if ((keepmask >> lane) & 1) next_played[__popc(keepmask & lt)] = note;
            else {  // NoteOff
              rec.off[__popc(offmask & lt)] = note;
              sm.tr_noff[note] = n_k;
              sm.tr_release[note] = max(sm.tr_nmc[note], p.R);
            }
          }
          if (has) {
            int sl = -1;
            if (mt >= 0) {
              sl = played[mt];
              sm.tr_nmc[sl] = e0.min_change;
              np_pitch[__popc(keepmask & ((1u << mt) - 1u))] = e0.pitch;
            }
            rec.slot[lane] = sl;
            if (sl >= 0) rec.entry[lane] = e0;
            const bool is_new = mt < 0 && e0.volume > 0;
            any_new = is_new;
            sm.t_new[lane] = is_new;
            sm.t_fv[lane] = static_cast<float>(e0.vol);
          }
          any_new = __any_sync(kFull, any_new);
        } else {
          bool bad = false;
          for (int i0 = 0; i0 < T; i0 += 32) {
            const int i = i0 + lane;
            if (i < T) {
              const double tp = entry(i0, i).pitch;
              const double lo_t = tp - dtrack, hi_t = tp + dtrack;
              int lo = 0, hi = Np;
              while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (pl_pitch[mid] < lo_t) lo = mid + 1;
                else hi = mid;
              }
              sm.t_lo[i] = lo;
              sm.t_match[i] = (lo < Np && pl_pitch[lo] <= hi_t) ? lo : -1;
            }
          }
          for (int j = lane; j < Np; j += 32) sm.keep[j] = 0;
          __syncwarp();
          for (int i = lane; i < T; i += 32)
            if (i > 0 && sm.t_lo[i] < sm.t_lo[i - 1] + (sm.t_match[i - 1] >= 0 ? 1 : 0)) bad = true;
          if (__any_sync(kFull, bad)) {
            if (lane == 0) {
              int itp = 0;
              for (int i = 0; i < T; ++i) {
                const double tp = tu[i].pitch;
                sm.t_match[i] = -1;
                for (; itp < Np; ++itp) {
                  const double pp = pl_pitch[itp];
                  if (pp < tp - dtrack) continue;
                  if (pp <= tp + dtrack) {
                    sm.t_match[i] = itp;
                    ++itp;
                  }
                  break;
                }
              }
            }
            __syncwarp();
          }
          for (int i = lane; i < T; i += 32)
            if (sm.t_match[i] >= 0) sm.keep[sm.t_match[i]] = 1;
          __syncwarp();
          for (int j0 = 0; j0 < Np; j0 += 32) {
            const int j = j0 + lane;
            const bool in = j < Np;
            const bool kp = in && sm.keep[j];
            const bool off = in && !kp;
            const unsigned m = __ballot_sync(kFull, kp), mo = __ballot_sync(kFull, off);
            if (kp) {
              const int pos = kept + __popc(m & lt);
              next_played[pos] = played[j];
              sm.remap[j] = pos;
            }
            if (off) {
              const int note = played[j];
              rec.off[n_off + __popc(mo & lt)] = note;
              sm.tr_noff[note] = n_k;
              sm.tr_release[note] = max(sm.tr_nmc[note], p.R);
            }
            kept += __popc(m);
            n_off += __popc(mo);
          }
          __syncwarp();
          for (int i0 = 0; i0 < T; i0 += 32) {
            const int i = i0 + lane;
            if (i < T) {
              const Tuned e = entry(i0, i);
              const int mt = sm.t_match[i];
              int sl = -1;
              if (mt >= 0) {
                sl = played[mt];
                sm.tr_nmc[sl] = e.min_change;
                np_pitch[sm.remap[mt]] = e.pitch;
              }
              rec.slot[i] = sl;
              if (sl >= 0) rec.entry[i] = e;
              const bool is_new = mt < 0 && e.volume > 0;
              sm.t_new[i] = is_new;
              sm.t_fv[i] = static_cast<float>(e.vol);
              any_new |= is_new;
            }
          }
          any_new = __any_sync(kFull, any_new);
        }
        __syncwarp();
        const int N_before = N;
        const int64_t id_base = started;
        int M = kept;
        if (any_new) {
          int n_new = 0;
          for (int i0 = 0; i0 < T; i0 += 32) n_new += __popc(__ballot_sync(kFull, i0 + lane < T && sm.t_new[i0 + lane]));
          const int room = max(0, p.max_channels - N);
          const int accepted = min(n_new, room);
          for (int i0 = 0; i0 < T; i0 += 32) {
            const int i = i0 + lane;
            if (i < T && sm.t_new[i]) {
              const float fv = sm.t_fv[i];
              int rank = 0;
              for (int j = 0; j < T; ++j)
                if (sm.t_new[j] && (sm.t_fv[j] > fv || (sm.t_fv[j] == fv && j < i))) ++rank;
              if (rank < accepted) {
                const Tuned e = entry(i0, i);
                const int sl = N + rank;
                rec.slot[i] = sl;
                rec.entry[i] = e;
                sm.tr_noff[sl] = -1;
                sm.tr_release[sl] = 0;
                sm.tr_nmc[sl] = e.min_change;
                next_played[kept + rank] = sl;
                np_pitch[kept + rank] = e.pitch;
              }
            }
          }
          started += n_new;
          dropped += n_new - accepted;
          N += accepted;
          M += accepted;
          __syncwarp();
        }
        bool unsorted = false;
        for (int j = lane; j < M; j += 32)
          if (j > 0 && np_pitch[j - 1] > np_pitch[j]) unsorted = true;
        if (__any_sync(kFull, unsorted)) {
          for (int j = lane; j < M; j += 32) {
            const double pj = np_pitch[j];
            int rank = 0;
            for (int l = 0; l < M; ++l) {
              const double pl = np_pitch[l];
              if (pl < pj || (pl == pj && l < j)) ++rank;
            }
            played[rank] = next_played[j];
            pl_pitch[rank] = pj;
          }
          __syncwarp();
        } else {
          int32_t* ti = played;
          played = next_played;
          next_played = ti;
          double* td = pl_pitch;
          pl_pitch = np_pitch;
          np_pitch = td;
        }
        Np = M;
        if (lane == 0) {
          rec.T = T;
          rec.N_before = N_before;
          rec.N_after = N;
          rec.n_off = n_off;
          rec.N_compact = -1;
          rec.id_base = id_base;
        }
        __syncwarp();
        // pool compaction, by the rule the voices warp applies to its own copy of the pool
        if (len_frame > 0 && N > Np) {
          const int64_t n_next = n_k + len_frame;
          auto done = [&](int i) { return sm.tr_noff[i] >= 0 && n_next >= sm.tr_noff[i] + sm.tr_release[i] + 17; };
          bool dead = false;
          for (int i = lane; i < N; i += 32) dead |= done(i);
          if (__any_sync(kFull, dead)) {
            int w = 0;
            for (int i0 = 0; i0 < N; i0 += 32) {
              const int i = i0 + lane;
              int64_t t_noff = 0;
              int32_t t_rel = 0, t_nmc = 0;
              bool alive = false;
              if (i < N) {
                alive = !done(i);
                t_noff = sm.tr_noff[i];
                t_rel = sm.tr_release[i];
                t_nmc = sm.tr_nmc[i];
              }
              const unsigned m = __ballot_sync(kFull, alive);
              const int pos = w + __popc(m & lt);
              if (i < N) {
                sm.remap[i] = alive ? pos : -1;
                rec.remap[i] = alive ? pos : -1;
              }
              __syncwarp();
              if (alive && pos != i) {
                sm.tr_noff[pos] = t_noff;
                sm.tr_release[pos] = t_rel;
                sm.tr_nmc[pos] = t_nmc;
              }
              __syncwarp();
              w += __popc(m);
            }
            for (int j = lane; j < Np; j += 32) played[j] = sm.remap[played[j]];
            N = w;
            if (lane == 0) rec.N_compact = w;
            __syncwarp();
          }
        }
      }
    } else if (kWarps == 3 && role == 1) {
      // ================= glides (three-warp form): frame k = it - 1 =================
      // The only thing a note's next glide needs from its past is the last emitted increment, and that follows from the glide
      // itself: f_at(len_frame - 1). So this warp owns that one value per slot, runs Glide::start for the changed notes and
      // leaves the constants in the record for the segments warp.
      const int64_t k = it - 1;
      if (k >= 0 && k < p.frames) {
        FrameRec& rec = sm.rec[k % kWarps];
        const int64_t n_k = static_cast<int64_t>(p.W) - 1 + k * p.hop;
        const int len_frame = static_cast<int>(min(static_cast<int64_t>(p.hop), p.n_samples - n_k));
        for (int i = lane; i < rec.T; i += 32) {
          const int sl = rec.slot[i];
          if (sl < 0) continue;
          const double ai = rec.entry[i].ai;
          if (sl >= rec.N_before) sm.f_last[sl] = ai;  // NoteOn: steady at ai
          else {
            Glide g;
            g.start(sm.f_last[sl], ai, p.hop);
            rec.glide[i] = GlideC{g.from, g.to, g.K, g.L, g.c1, g.P_m1, g.f_m1, g.m};
            if (len_frame > 0) sm.f_last[sl] = g.f_at(len_frame - 1);
          }
        }
        __syncwarp();
        if (rec.N_compact >= 0) {
          for (int i0 = 0; i0 < rec.N_after; i0 += 32) {
            const int i = i0 + lane;
            double v = 0;
            int pos = -1;
            if (i < rec.N_after) {
              v = sm.f_last[i];
              pos = rec.remap[i];
            }
            __syncwarp();
            if (pos >= 0 && pos != i) sm.f_last[pos] = v;
            __syncwarp();
          }
        }
      }
    } else {
      // ================= voices / segments: frame k = it - (kWarps - 1) =================
      const int64_t k = it - (kWarps - 1);
      if (k >= 0) {
        const FrameRec& rec = sm.rec[k % kWarps];
        const int T = rec.T, N_before = rec.N_before, NN = rec.N_after;
        const int64_t n_k = static_cast<int64_t>(p.W) - 1 + k * p.hop;
        const int len_frame = static_cast<int>(min(static_cast<int64_t>(p.hop), p.n_samples - n_k));
        for (int j = lane; j < rec.n_off; j += 32) note_off(p, sm.notes[rec.off[j]], n_k);
        for (int i = lane; i < T; i += 32) {
          const int sl = rec.slot[i];
          if (sl >= 0) {
            const Tuned e = rec.entry[i];
            if (sl >= N_before) sm.notes[sl] = note_on(p, e, rec.id_base + (sl - N_before) + 1, n_k);
            else if (kWarps == 2) note_change(p, sm.notes[sl], e);
            else {  // NoteChange with the glide the glide warp prepared (note_change, pitch.cuh)
              Note& n = sm.notes[sl];
              const GlideC& g = rec.glide[i];
              n.pitch = e.pitch;
              n.vol_x = static_cast<double>(0.1f * e.volume);
              n.next_min_change = e.min_change;
              n.glide.from = g.from;
              n.glide.to = g.to;
              n.glide.K = g.K;
              n.glide.L = g.L;
              n.glide.c1 = g.c1;
              n.glide.P_m1 = g.P_m1;
              n.glide.f_m1 = g.f_m1;
              n.glide.m = g.m;
              n.glide.i = 0;
            }
          }
        }
        __syncwarp();
        for (int off = 0; off < max(len_frame, 0) || off == 0; off += kMaxTile) {
          const int len = min(kMaxTile, len_frame - off);
          const int64_t n_k_tile = n_k + off;
          int64_t base = 0;
          bool ok = true;
          if (len > 0 && NN > 0) {
            if (chunk_next + NN > chunk_end) {
              unsigned long long got = 0;
              if (lane == 0) got = atomicAdd(a.seg_counter, static_cast<unsigned long long>(kSegChunk));
              got = __shfl_sync(kFull, got, 0);
              chunk_next = static_cast<int64_t>(got);
              chunk_end = chunk_next + kSegChunk;
            }
            base = chunk_next;
            chunk_next += NN;
            ok = base + NN <= a.seg_capacity;
          }
          if (lane == 0) {
            const bool filled = ok && len > 0 && NN > 0;
            a.tile_begin[tile] = filled ? base : 0;
            a.tile_end[tile] = filled ? base + NN : 0;
            a.tile_s0[tile] = s * p.n_samples + n_k_tile;
            a.tile_len[tile] = max(len, 0);
            if (!ok) *a.overflow = 1;
          }
          ++tile;
          if (len <= 0) break;
          for (int idx = lane; idx < NN; idx += 32) {
            const Segment g = emit_segment(p, sm.notes[idx], n_k_tile, len);
            if (ok) a.segs[base + idx] = g;
          }
          __syncwarp();
        }
        if (rec.N_compact >= 0) {  // the tracker's compaction, applied to this warp's copy of the pool
          for (int i0 = 0; i0 < NN; i0 += 32) {
            const int i = i0 + lane;
            Note tmp;
            int pos = -1;
            if (i < NN) {
              pos = rec.remap[i];
              if (pos >= 0 && pos != i) tmp = sm.notes[i];
            }
            __syncwarp();
            if (pos >= 0 && pos != i) sm.notes[pos] = tmp;
            __syncwarp();
          }
        }
      }
    }
#ifdef CA_PITCH_TIMELINE
    busy += clock64() - t_in;
#endif
    __syncthreads();
  }
#ifdef CA_PITCH_TIMELINE
  if (s == 0 && lane == 0) printf("pitch_track2<%d> role %d: %lld busy clocks per frame\n", kWarps, role, busy / p.frames);
#endif
  if (is_tracker && lane == 0) {
    a.stats[2 * s] = started;
    a.stats[2 * s + 1] = dropped;
  }
}

}  // namespace

// Tuned lists, then tracking and segments, for `n_streams` streams of `p.frames` frames whose peak lists are on the device.
// Outputs (device): segments, per-tile [begin, end) into them, first output sample and length of every tile, stats.
// `needed` receives the number of segments the allocator handed out; when it exceeds seg_capacity the tiles past the
// capacity carry no segments and the caller re-runs with a larger buffer.
int pitch_pipeline_device(ca_ctx* ctx, const PitchParams& p, int64_t n_streams, const int32_t* d_peak_bin, const void* d_peak_off,
                          const void* d_peak_db, bool peaks_f64, const int32_t* d_peak_count, Tuned* d_tuned, int32_t* d_tcount, Segment* d_segs,
                          int64_t seg_capacity, int64_t n_tiles_per_stream, int64_t* d_tile_begin, int64_t* d_tile_end,
                          int64_t* d_tile_s0, int32_t* d_tile_len, int64_t* d_stats, unsigned long long* d_counter,
                          int32_t* d_overflow, bool run_prefix, cudaStream_t st, bool reset_counters) {
  if (p.P > kMaxTuned) return fail(ctx, CA_ERR_UNSUPPORTED, "round trip: max_peaks above 256 is not supported by the device pitch pipeline");
  if (p.max_channels > kMaxNotes) return fail(ctx, CA_ERR_UNSUPPORTED, "round trip: max_channels above 256 is not supported");
  const int64_t rows = n_streams * p.frames;
  if (rows == 0) return CA_OK;
  if (run_prefix) {
    if (peaks_f64) {
      PrefixArgs<double> pa{p, d_peak_bin, static_cast<const double*>(d_peak_off), static_cast<const double*>(d_peak_db), d_peak_count, d_tuned, d_tcount, rows};
      pitch_prefix_kernel<double><<<static_cast<unsigned>((rows + 127) / 128), 128, 0, st>>>(pa);
    } else {
      PrefixArgs<float> pa{p, d_peak_bin, static_cast<const float*>(d_peak_off), static_cast<const float*>(d_peak_db), d_peak_count, d_tuned, d_tcount, rows};
      pitch_prefix_kernel<float><<<static_cast<unsigned>((rows + 127) / 128), 128, 0, st>>>(pa);
    }
    CA_CUDA_OK(ctx, cudaGetLastError());
    ctx->launches += 1;
  }
  if (reset_counters) {
    CA_CUDA_OK(ctx, cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), st));
    CA_CUDA_OK(ctx, cudaMemsetAsync(d_overflow, 0, sizeof(int32_t), st));
  }
  static const int n_warps = [] { const char* e = getenv("CA_PITCH_WARPS"); return e ? atoi(e) : 3; }();  // 1 / 2 / 3 warps per stream (A/B)
  TrackArgs ta{p, d_tuned, d_tcount, d_segs, d_counter, seg_capacity, d_tile_begin, d_tile_end, d_tile_s0, d_tile_len,
               n_tiles_per_stream, d_stats, d_overflow};
  if (n_warps == 1) {
    CA_CUDA_OK(ctx, cudaFuncSetAttribute(pitch_track_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(TrackSmem))));
    pitch_track_kernel<<<static_cast<unsigned>(n_streams), 32, sizeof(TrackSmem), st>>>(ta);
  } else if (n_warps == 2) {
    CA_CUDA_OK(ctx, cudaFuncSetAttribute(pitch_track2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(Track2Smem))));
    pitch_track2_kernel<2><<<static_cast<unsigned>(n_streams), 64, sizeof(Track2Smem), st>>>(ta);
  } else {
    CA_CUDA_OK(ctx, cudaFuncSetAttribute(pitch_track2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(Track2Smem))));
    pitch_track2_kernel<3><<<static_cast<unsigned>(n_streams), 96, sizeof(Track2Smem), st>>>(ta);
  }
  CA_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches += 1;
  return CA_OK;
}

}  // namespace ca
