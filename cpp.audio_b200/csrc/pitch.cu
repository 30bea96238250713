// pitch.cu — the pitch pipeline of the round trip on the device.
//
// Reference: RtResynth::step (source/rt.resynth.lib.cpp:1670-1759) runs, per analysis frame, the functions of
// source/rt.resynth.lib.algo.cpp on the peak list and synthesize_sounds (:264-382) turns the result into NoteOn / NoteChange /
// NoteOff for a pool of voices. The formulas live in pitch.cuh (shared with the host A/B path); this file maps them to the GPU:
//
//   pitch_prefix_kernel  one THREAD per frame: peaks -> tuned list. Stateless, so all frames of all streams run at once and the
//                        double-precision log2 / pow of every peak leave the sequential part.
//   pitch_track3_kernel  three WARPS per stream walk the frames in order (the played-notes state is carried from frame to
//                        frame, :1710-1758) as a pipeline: tracker, glides, segments (see the comment above the kernel). Inside a
//                        frame lanes = tuned entries / notes: greedy monotone matching with an exact serial fallback, note-off /
//                        note-change / note-on on the lowest free channel, ordering by rank counting, the start-phase rule of the
//                        voice pool, one Segment per sounding note and render tile (the closed forms of tracks.cuh), pool
//                        compaction by ballot prefix sums. Segment storage comes from a global bump allocator in chunks; when
//                        the capacity guess is too small the kernel keeps counting and the host re-runs it with the exact size.
//                        A launch takes a range of frames: the state of a stream rests in global memory in between, so the
//                        caller can render and ship the audio of the frames already tracked (roundtrip.cu).
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
  Tuned* tuned;              // [rows][PT]
  int32_t* tcount;           // [rows]
  int64_t rows;
  unsigned long long* truncated;  // frames whose raw or tuned list did not fit its capacity
};

template <class R>
__global__ void __launch_bounds__(128) pitch_prefix_kernel(const __grid_constant__ PrefixArgs<R> a) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= a.rows) return;
  const int P = a.p.P;
  const int found = a.peak_count[r];
  int cut = found > P ? 1 : 0;
  a.tcount[r] = pitch_prefix(a.p, a.peak_bin + r * P, a.peak_off + r * P, a.peak_db + r * P, min(found, P), a.tuned + r * a.p.PT, &cut);
  if (cut) atomicAdd(a.truncated, 1ull);
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
  // A launch walks the frames [k_begin, k_end) of every stream; tuned / tcount hold those frames only, [streams][chunk_frames].
  // Between launches the whole state of a stream (the shared-memory image and a few registers) rests in `state`, so that the
  // audio of the frames already tracked can be rendered and copied out while the tracker goes on (roundtrip.cu).
  int64_t k_begin, k_end, chunk_frames;
  unsigned char* state;      // [streams][state_stride]; null: a single launch over all frames
  int64_t state_stride;
  int32_t resume;            // take the state up from `state` instead of starting afresh
};

struct TrackRegs {  // what a stream keeps in registers between frames
  int32_t N, Np, played_sel, pad;
  int64_t started, dropped, tile, chunk_next, chunk_end;
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

// ---------------------------------------------------------------------------------------------------------------
// Three warps per stream, arranged as a pipeline over frames with one CTA barrier per frame.
//   warp 0, tracker (frame j): tracking needs nothing from the note arithmetic: the played list is keyed by the tuned pitches,
//     a released note leaves the pool after n_off + release + 17 samples, and a new note takes the lowest free channel of the
//     voice pool -- all integers the tracker knows.
//   warp 1, glides (frame j-1): owns the last emitted increment of every slot (what a note's next glide starts from, and what
//     the start-phase rule compares), runs Glide::start for the changed notes and decides which new notes take over the
//     phase of a sounding channel (setPhase<SynchronizePhase::Yes>, include/gen.crtp.h:127-156).
//   warp 2, segments (frame j-2): applies note-on / note-off, takes the prepared glides and phase sources, advances phase,
//     volume filter and envelope and writes the segments.
// The warps exchange a per-frame record and compact the pool by the same rule, so slot numbers stay in step without being
// exchanged. A frame costs the slowest stage instead of the sum. (The one- and two-warp forms this grew out of -- 31.1 and
// 24.1 ms against 17.6 ms for 32 streams x 11,235 frames -- are gone; the host form in roundtrip.cu is the A/B partner.)
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
  GlideC glide[kMaxTuned];   // the glide of every changed note, computed by the glide warp
  int16_t chan[kMaxTuned];   // note-on entries: the channel the tracker gave the note
  int32_t sync[kMaxTuned];   // note-on entries, by the glide warp: -1 random start phase; s >= 0: the phase of slot s;
                             // -2 - r: the random start phase of this frame's r-th new note (a chain of same-frame note-ons)
};

struct Track2Smem {
  Note notes[kMaxNotes];           // voices warp
  Tuned ring[kRing][32];
  uint64_t bar[kRing];
  FrameRec rec[3];
  // glide warp
  double f_last[kMaxNotes];        // last emitted increment of every slot
  double osc_ai_ch[kMaxNotes];     // per CHANNEL: the increment its oscillator last stepped with (survives the note, 0: never used)
  double new_ai[kMaxTuned];        // this frame's new notes in slot order: increment ...
  int16_t new_entry[kMaxTuned];    // ... and tuned entry
  int16_t gl_chan[kMaxNotes];      // channel of every slot
  // tracker warp
  double pl_pitch[2][kMaxNotes];
  int32_t played[2][kMaxNotes];
  int64_t tr_noff[kMaxNotes];      // sample of the first release step, -1 while tracked
  int32_t tr_release[kMaxNotes];
  int32_t tr_nmc[kMaxNotes];       // next_min_change of the slot
  int16_t tr_chan[kMaxNotes];      // channel of the slot
  int16_t new_chan[kMaxNotes];     // channels handed to this frame's new notes, by rank
  uint32_t free_mask[kMaxNotes / 32];  // free channels of the voice pool
  int32_t remap[kMaxNotes];        // tracker scratch
  int32_t t_lo[kMaxTuned];
  int32_t t_match[kMaxTuned];
  float t_fv[kMaxTuned];
  uint8_t t_new[kMaxTuned];
  uint8_t keep[kMaxNotes];
};

constexpr int kWarps = 3;
__global__ void __launch_bounds__(32 * kWarps) pitch_track3_kernel(const __grid_constant__ TrackArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Track2Smem& sm = *reinterpret_cast<Track2Smem*>(smem_raw);
  const PitchParams& p = a.p;
  const int lane = threadIdx.x & 31;
  const int role = threadIdx.x >> 5;  // 0: tracker, 1: glides, 2: segments
  const bool is_tracker = role == 0;
  const unsigned lt = lanemask_lt();
  const int64_t s = blockIdx.x;
  const int P = p.PT;  // entries per frame of the tuned lists
  const double dtrack = p.max_track;
  const int64_t n_frames = a.k_end - a.k_begin;  // frames of this launch; kr = k - k_begin indexes them
  const Tuned* tu_base = a.tuned + s * a.chunk_frames * P;
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
  const int32_t* tc = a.tcount + s * a.chunk_frames;
  auto load_counts = [&](int64_t k0) { return k0 + lane < n_frames ? tc[k0 + lane] : 0; };
  int cnt_cur = 0, cnt_next = 0;
  // ---- voices state ----
  int64_t tile = s * a.tiles_per_stream;
  int64_t chunk_next = 0, chunk_end = 0;

  uint4* st_img = a.state ? reinterpret_cast<uint4*>(a.state + s * a.state_stride) : nullptr;
  TrackRegs* st_regs = a.state ? reinterpret_cast<TrackRegs*>(a.state + s * a.state_stride + sizeof(Track2Smem)) : nullptr;
  static_assert(sizeof(Track2Smem) % 16 == 0, "the state image is copied in 16-byte words");
  if (a.resume) {
    for (int i = threadIdx.x; i < static_cast<int>(sizeof(Track2Smem) / 16); i += blockDim.x) reinterpret_cast<uint4*>(smem_raw)[i] = st_img[i];
    const TrackRegs r = *st_regs;
    N = r.N;
    Np = r.Np;
    started = r.started;
    dropped = r.dropped;
    tile = r.tile;
    chunk_next = r.chunk_next;
    chunk_end = r.chunk_end;
    if (r.played_sel) {
      played = sm.played[1];
      next_played = sm.played[0];
      pl_pitch = sm.pl_pitch[1];
      np_pitch = sm.pl_pitch[0];
    }
    __syncthreads();  // the image is in place before the barriers inside it are initialised again
  } else {
    for (int c = threadIdx.x; c < kMaxNotes; c += blockDim.x) sm.osc_ai_ch[c] = 0.0;
    if (threadIdx.x < kMaxNotes / 32) {
      const int lo = 32 * threadIdx.x;  // channels [0, max_channels) are free
      sm.free_mask[threadIdx.x] = p.max_channels >= lo + 32 ? kFull : p.max_channels > lo ? (1u << (p.max_channels - lo)) - 1u : 0u;
    }
  }
  if (is_tracker) {
    cnt_cur = load_counts(0);
    cnt_next = load_counts(32);
    if (lane == 0) {
      for (int i = 0; i < kRing; ++i) mbar_init(&sm.bar[i], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      fence_proxy_async();
      for (int f = 0; f < kAhead && f < n_frames; ++f) {
        mbar_expect_tx(&sm.bar[f], staged_bytes);
        bulk_g2s(sm.ring[f], tu_base + static_cast<int64_t>(f) * P, staged_bytes, &sm.bar[f]);
      }
    }
  }
  __syncthreads();

#ifdef CA_PITCH_TIMELINE
  long long busy = 0;  // clocks this warp spent working (between leaving a barrier and arriving at the next)
#endif
  for (int64_t it = 0; it < n_frames + kWarps - 1; ++it) {
#ifdef CA_PITCH_TIMELINE
    const long long t_in = clock64();
#endif
    if (is_tracker) {
      // ================= tracker: frame k = it =================
      const int64_t kr = it, k = a.k_begin + kr;
      if (kr < n_frames) {
        const Tuned* tu = tu_base + kr * P;
        FrameRec& rec = sm.rec[kr % kWarps];
        if ((kr & 31) == 0 && kr > 0) {
          cnt_cur = cnt_next;
          cnt_next = load_counts(kr + 32);
        }
        const int T = __shfl_sync(kFull, cnt_cur, static_cast<int>(kr & 31));
        const int slot = static_cast<int>(kr % kRing);
        mbar_wait(&sm.bar[slot], static_cast<uint32_t>((kr / kRing) & 1));
        Tuned e0{};
        if (lane < n_staged) e0 = sm.ring[slot][lane];
        if (lane == 0 && kr + kAhead < n_frames) {  // slot of frame kr + kAhead: last read by the voices warp an iteration ago
          const int nslot = static_cast<int>((kr + kAhead) % kRing);
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
          // note-ons are handled in the order of decreasing volume, and each takes the first free channel (gen.crtp.h:398-410)
          if (lane == 0) {
            int w = 0;
            for (int r = 0; r < accepted; ++r) {
              while (sm.free_mask[w] == 0u) ++w;
              const int b = __ffs(static_cast<int>(sm.free_mask[w])) - 1;
              sm.free_mask[w] &= ~(1u << b);
              sm.new_chan[r] = static_cast<int16_t>(32 * w + b);
            }
          }
          __syncwarp();
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
                rec.chan[i] = sm.new_chan[rank];
                sm.tr_chan[sl] = sm.new_chan[rank];
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
              int16_t t_chan = 0;
              bool alive = false;
              if (i < N) {
                alive = !done(i);
                t_noff = sm.tr_noff[i];
                t_rel = sm.tr_release[i];
                t_nmc = sm.tr_nmc[i];
                t_chan = sm.tr_chan[i];
                if (!alive) atomicOr(&sm.free_mask[t_chan >> 5], 1u << (t_chan & 31));  // the channel reached EnvelopeDone2
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
                sm.tr_chan[pos] = t_chan;
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
    } else if (role == 1) {
      // ================= glides (three-warp form): frame k = it - 1 =================
      // The only thing a note's next glide needs from its past is the last emitted increment, and that follows from the glide
      // itself: f_at(len_frame - 1). So this warp owns that one value per slot, runs Glide::start for the changed notes and
      // leaves the constants in the record for the segments warp.
      const int64_t kr = it - 1, k = a.k_begin + kr;
      if (kr >= 0 && kr < n_frames) {
        FrameRec& rec = sm.rec[kr % kWarps];
        const int64_t n_k = static_cast<int64_t>(p.W) - 1 + k * p.hop;
        const int len_frame = static_cast<int>(min(static_cast<int64_t>(p.hop), p.n_samples - n_k));
        // ---- start phase of the new notes (before this frame's glides move f_last) ----
        const int n_new = rec.N_after - rec.N_before;
        if (n_new > 0) {
          for (int i = lane; i < rec.T; i += 32) {
            const int sl = rec.slot[i];
            if (sl >= rec.N_before) {
              sm.new_ai[sl - rec.N_before] = rec.entry[i].ai;
              sm.new_entry[sl - rec.N_before] = static_cast<int16_t>(i);
              sm.gl_chan[sl] = rec.chan[i];
            }
          }
          __syncwarp();
          // in event order (= slot order): the candidates of new note r are the sounding slots and the new notes before it;
          // the first match in CHANNEL order wins
          for (int r = 0; r < n_new; ++r) {
            const int c = sm.gl_chan[rec.N_before + r];
            const double ai = chan_angle_increments(sm.osc_ai_ch[c], sm.new_ai[r]);
            int best = INT_MAX;  // (channel << 16) | slot of the matching candidate with the lowest channel
            for (int sl = lane; sl < rec.N_before + r; sl += 32) {
              const int oc = sm.gl_chan[sl];
              // sounding slot: its oscillator has stepped with f_last; new note: still the channel's previous increment
              const double oai = sl < rec.N_before ? sm.f_last[sl] : chan_angle_increments(sm.osc_ai_ch[oc], sm.new_ai[sl - rec.N_before]);
              if (oai == ai) best = min(best, (oc << 16) | sl);
            }
            best = __reduce_min_sync(kFull, best);
            if (lane == 0) {
              int src = -1;
              if (best != INT_MAX) {
                src = best & 0xffff;
                if (src >= rec.N_before) {  // a note of this frame: its phase is its own source's, or its random one
                  const int up = rec.sync[sm.new_entry[src - rec.N_before]];
                  src = up == -1 ? -2 - (src - rec.N_before) : up;
                }
              }
              rec.sync[sm.new_entry[r]] = src;
            }
            __syncwarp();
          }
        }
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
            int16_t ch = 0;
            if (i < rec.N_after) {
              v = sm.f_last[i];
              ch = sm.gl_chan[i];
              pos = rec.remap[i];
              if (pos < 0) sm.osc_ai_ch[ch] = v;  // the note is over: its channel keeps the oscillator's last increment
            }
            __syncwarp();
            if (pos >= 0 && pos != i) {
              sm.f_last[pos] = v;
              sm.gl_chan[pos] = ch;
            }
            __syncwarp();
          }
        }
      }
    } else {
      // ================= voices / segments: frame k = it - (kWarps - 1) =================
      const int64_t kr = it - (kWarps - 1), k = a.k_begin + kr;
      if (kr >= 0) {
        const FrameRec& rec = sm.rec[kr % kWarps];
        const int T = rec.T, N_before = rec.N_before, NN = rec.N_after;
        const int64_t n_k = static_cast<int64_t>(p.W) - 1 + k * p.hop;
        const int len_frame = static_cast<int>(min(static_cast<int64_t>(p.hop), p.n_samples - n_k));
        for (int j = lane; j < rec.n_off; j += 32) note_off(p, sm.notes[rec.off[j]], n_k);
        for (int i = lane; i < T; i += 32) {
          const int sl = rec.slot[i];
          if (sl >= 0) {
            const Tuned e = rec.entry[i];
            if (sl >= N_before) {
              Note nn = note_on(p, e, rec.id_base + (sl - N_before) + 1, n_k);
              const int src = rec.sync[i];  // start-phase rule, decided by the glide warp
              if (src >= 0) nn.phase = sm.notes[src].phase;  // (phases only move in emit_segment, below)
              else if (src <= -2) {
                double a0 = static_cast<double>(hash_unit(static_cast<uint64_t>(rec.id_base + (-2 - src) + 1), 2));
                nn.phase = a0 - 2.0 * floor(0.5 * a0);
              }
              sm.notes[sl] = nn;
            } else {  // NoteChange with the glide the glide warp prepared (note_change, pitch.cuh)
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
  if (s == 0 && lane == 0) printf("pitch_track3 role %d: %lld busy clocks per frame\n", role, busy / p.frames);
#endif
  if (is_tracker && lane == 0) {
    a.stats[2 * s] = started;
    a.stats[2 * s + 1] = dropped;
  }
  if (st_img) {  // the stream's state rests in global memory until the launch of the next frames
    __syncthreads();
    for (int i = threadIdx.x; i < static_cast<int>(sizeof(Track2Smem) / 16); i += blockDim.x) st_img[i] = reinterpret_cast<const uint4*>(smem_raw)[i];
    if (is_tracker && lane == 0) {
      st_regs->N = N;
      st_regs->Np = Np;
      st_regs->played_sel = played == sm.played[1] ? 1 : 0;
      st_regs->started = started;
      st_regs->dropped = dropped;
    }
    if (role == kWarps - 1 && lane == 0) {
      st_regs->tile = tile;
      st_regs->chunk_next = chunk_next;
      st_regs->chunk_end = chunk_end;
    }
  }
}

}  // namespace

// Tuned lists of `rows` frames (any frames of any streams: the stage is stateless) from their peak lists.
int pitch_prefix_device(ca_ctx* ctx, const PitchParams& p, int64_t rows, const int32_t* d_peak_bin, const void* d_peak_off, const void* d_peak_db,
                        bool peaks_f64, const int32_t* d_peak_count, unsigned long long* d_truncated, Tuned* d_tuned, int32_t* d_tcount,
                        cudaStream_t st) {
  if (p.PT > kMaxTuned) return fail(ctx, CA_ERR_UNSUPPORTED, "round trip: more than 256 tuned entries per frame are not supported by the device pitch pipeline");
  if (rows == 0) return CA_OK;
  if (peaks_f64) {
    PrefixArgs<double> pa{p, d_peak_bin, static_cast<const double*>(d_peak_off), static_cast<const double*>(d_peak_db), d_peak_count, d_tuned, d_tcount, rows, d_truncated};
    pitch_prefix_kernel<double><<<static_cast<unsigned>((rows + 127) / 128), 128, 0, st>>>(pa);
  } else {
    PrefixArgs<float> pa{p, d_peak_bin, static_cast<const float*>(d_peak_off), static_cast<const float*>(d_peak_db), d_peak_count, d_tuned, d_tcount, rows, d_truncated};
    pitch_prefix_kernel<float><<<static_cast<unsigned>((rows + 127) / 128), 128, 0, st>>>(pa);
  }
  CA_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches += 1;
  return CA_OK;
}

size_t pitch_track_state_bytes() { return sizeof(Track2Smem) + sizeof(TrackRegs); }

// Tracking and segments of the frames [k_begin, k_end) of `n_streams` streams; d_tuned / d_tcount hold those frames,
// [streams][k_end - k_begin]. d_state (nullable: one launch over all frames) carries the streams from launch to launch.
int pitch_track_device(ca_ctx* ctx, const PitchParams& p, int64_t n_streams, const Tuned* d_tuned, const int32_t* d_tcount, int64_t k_begin,
                       int64_t k_end, unsigned char* d_state, Segment* d_segs, int64_t seg_capacity, int64_t n_tiles_per_stream,
                       int64_t* d_tile_begin, int64_t* d_tile_end, int64_t* d_tile_s0, int32_t* d_tile_len, int64_t* d_stats,
                       unsigned long long* d_counter, int32_t* d_overflow, cudaStream_t st) {
  if (p.PT > kMaxTuned) return fail(ctx, CA_ERR_UNSUPPORTED, "round trip: more than 256 tuned entries per frame are not supported by the device pitch pipeline");
  if (p.max_channels > kMaxNotes) return fail(ctx, CA_ERR_UNSUPPORTED, "round trip: max_channels above 256 is not supported");
  if (n_streams == 0 || k_end <= k_begin) return CA_OK;
  TrackArgs ta{p, d_tuned, d_tcount, d_segs, d_counter, seg_capacity, d_tile_begin, d_tile_end, d_tile_s0, d_tile_len,
               n_tiles_per_stream, d_stats, d_overflow, k_begin, k_end, k_end - k_begin, d_state,
               static_cast<int64_t>((pitch_track_state_bytes() + 15) / 16 * 16), (d_state && k_begin > 0) ? 1 : 0};
  static bool attr_set = false;
  if (!attr_set) {
    CA_CUDA_OK(ctx, cudaFuncSetAttribute(pitch_track3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(Track2Smem))));
    attr_set = true;
  }
  pitch_track3_kernel<<<static_cast<unsigned>(n_streams), 32 * kWarps, sizeof(Track2Smem), st>>>(ta);
  CA_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches += 1;
  return CA_OK;
}

// Both stages over all frames of `n_streams` streams in one go.
// Outputs (device): segments, per-tile [begin, end) into them, first output sample and length of every tile, stats.
// When the allocator hands out more than seg_capacity segments the tiles past the capacity carry none, *d_overflow is set and
// the caller re-runs with a larger buffer.
int pitch_pipeline_device(ca_ctx* ctx, const PitchParams& p, int64_t n_streams, const int32_t* d_peak_bin, const void* d_peak_off,
                          const void* d_peak_db, bool peaks_f64, const int32_t* d_peak_count, unsigned long long* d_truncated, Tuned* d_tuned, int32_t* d_tcount, Segment* d_segs,
                          int64_t seg_capacity, int64_t n_tiles_per_stream, int64_t* d_tile_begin, int64_t* d_tile_end,
                          int64_t* d_tile_s0, int32_t* d_tile_len, int64_t* d_stats, unsigned long long* d_counter,
                          int32_t* d_overflow, bool run_prefix, cudaStream_t st, bool reset_counters) {
  const int64_t rows = n_streams * p.frames;
  if (rows == 0) return CA_OK;
  if (run_prefix)
    if (int rc = pitch_prefix_device(ctx, p, rows, d_peak_bin, d_peak_off, d_peak_db, peaks_f64, d_peak_count, d_truncated, d_tuned, d_tcount, st)) return rc;
  if (reset_counters) {
    CA_CUDA_OK(ctx, cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), st));
    CA_CUDA_OK(ctx, cudaMemsetAsync(d_overflow, 0, sizeof(int32_t), st));
  }
  return pitch_track_device(ctx, p, n_streams, d_tuned, d_tcount, 0, p.frames, nullptr, d_segs, seg_capacity, n_tiles_per_stream, d_tile_begin,
                            d_tile_end, d_tile_s0, d_tile_len, d_stats, d_counter, d_overflow, st);
}

}  // namespace ca
