// cuda_resynth_synth.hpp — C++ host adapter with the reference's voice-level SYNTHESIS interface, for the synth rt.resynth
// drives: sine::Synth<nOuts, ResynthElement<double>, TryAccountForTimeSourceJitter::No, SynchronizePhase::Yes,
// DefaultStartPhase::Random, HandleNoteOff::Yes, 127, ResynthElementInitializer<double>> (source/rt.resynth.lib.cpp:201-210).
//
// Reference interface (ImplCRTP, include/gen.crtp.h:243-348, :381-386):
//   onEventResult onEvent(int sample_rate, Event const&, Out&, Chans&, std::optional<TimestampAndSource> const&)
//   bool compute(double* buffer, int nFrames)            accumulate nFrames (<= 16) interleaved frames, false = unregister me
//   void allNotesOff();  void setSynchronousElementInitializer(Initializer const&)
// with Event / NoteId / mkNoteOn / mkNoteChange / mkNoteOff of include/events.h:12-107 and TimestampAndSource of :187-216.
// The names, argument order and meaning are kept; `Out` and `Chans` (the reference's lock and one-shot queue) are accepted and
// not used: events go to the GPU pool (ca_tracks_*, include/cpp_audio_b200.h) with the sample time they take effect at.
//
//   onEvent(..., std::nullopt)        the event acts at the next block boundary: the first sample compute() has not rendered yet
//                                     (the reference applies its one-shots at the start of the next audio callback,
//                                     include/simple_audio_context.h:69-76). With block_frames = 16 or less and events posted
//                                     between compute() calls, as the offline rt.resynth loop does (rt.resynth.lib.cpp:1204-1234),
//                                     timing is sample exact.
//   onEvent(..., TimestampAndSource)  the event acts at nanoseconds_to_frames(time) (include/sound.functions.h:21-26) on the
//                                     synth's timeline, or at the next block boundary when that is already past (gen.crtp.h:503-510)
//   compute(buffer, n)                renders ahead one block at a time on the GPU, so the per-call cost is a copy
// DROPPED_NOTE is reported, as in the reference, from the channels that are free when onEvent is called.
#pragma once

#include <algorithm>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/cpp_audio_b200.h"

namespace ca_host {
namespace resynth {

enum class onEventResult { OK = 0, DROPPED_NOTE };
enum class EventType : uint8_t { NoteOn, NoteOff, NoteChange };
struct NoteOnEvent {
  float frequency;
  float velocity;
};
struct NoteChangeEvent {
  float changed_frequency;
  float changed_velocity;
};
struct NoteOffEvent {
  float velocity;
};
struct NoteId {
  int64_t noteid;
  bool operator==(NoteId const& o) const { return noteid == o.noteid; }
  bool operator!=(NoteId const& o) const { return noteid != o.noteid; }
};
struct Event {
  Event(EventType t, NoteId const& n) : type(t), noteid(n) {}
  EventType type;
  NoteId noteid;
  void offsetNoteId(int64_t offset) { noteid.noteid += offset; }
  union {
    NoteOnEvent noteOn;
    NoteChangeEvent noteChange;
    NoteOffEvent noteOff;
  };
};
inline Event mkNoteOn(NoteId const& noteid, float frequency, float velocity) {
  Event e(EventType::NoteOn, noteid);
  e.noteOn.frequency = frequency;
  e.noteOn.velocity = velocity;
  return e;
}
inline Event mkNoteChange(NoteId const& noteid, float changed_velocity, float new_frequency) {
  Event e(EventType::NoteChange, noteid);
  e.noteChange.changed_frequency = new_frequency;
  e.noteChange.changed_velocity = changed_velocity;
  return e;
}
inline Event mkNoteOff(NoteId const& noteid) {
  Event e(EventType::NoteOff, noteid);
  e.noteOff.velocity = 0.f;
  return e;
}

struct TimestampAndSource {  // include/events.h:187-216, times in nanoseconds
  TimestampAndSource(uint64_t t, uint64_t sourceKey) : time(t), src_key(sourceKey) {}
  uint64_t getNanosTime() const { return time; }
  uint64_t getSourceKey() const { return src_key; }

 private:
  uint64_t time;
  uint64_t src_key;
};

inline int64_t nanoseconds_to_frames(uint64_t ns, int sample_rate) {  // include/sound.functions.h:21-26, in double for long timelines
  return static_cast<int64_t>(0.5 + static_cast<double>(sample_rate) * 1e-9 * static_cast<double>(ns));
}

// ResynthElementInitializer (source/rt.resynth.lib.cpp:96-118): what every new note is set up with
struct ResynthElementInitializer {
  int sample_rate = 48000;
  int stride = 256;  // frames between analysis frames: glide duration, volume filter limit 2/stride
  float stereo_spread = 1.f;
  ca_ahdsr ahdsr{0, CA_ITP_LINEAR, 0, 0, CA_ITP_LINEAR, 0, CA_ITP_LINEAR, 1.f};
};

struct NoOut {};    // stand-ins for the Out / Chans arguments of onEvent when the caller has none
struct NoChans {};

template <int nOuts>
class CudaResynthSynth {
 public:
  static constexpr int n_channels_per_note = 2;  // gen.crtp.h:221-233: n_channels = 2 * nVoices

  explicit CudaResynthSynth(ca_ctx* ctx, int n_voices = 127, int block_frames = 4096) : ctx_(ctx), n_voices_(n_voices), block_(block_frames) {
    static_assert(nOuts == 1 || nOuts == 2, "mono or stereo");
    rendered_.resize(static_cast<size_t>(block_) * nOuts);
  }
  ~CudaResynthSynth() { ca_tracks_close(pool_); }
  CudaResynthSynth(CudaResynthSynth const&) = delete;
  CudaResynthSynth& operator=(CudaResynthSynth const&) = delete;

  // gen.crtp.h:343-347. Takes effect for the pool opened by the next event; call it before the first one, as rt.resynth does.
  void setSynchronousElementInitializer(ResynthElementInitializer const& i) {
    if (pool_) throw std::logic_error("CudaResynthSynth: the element initializer cannot change once events were posted");
    init_ = i;
  }

  template <typename Out, typename Chans>
  onEventResult onEvent(int const sample_rate, Event const& e, Out&, Chans&, std::optional<TimestampAndSource> const& maybeTimeAndSource) {
    if (sample_rate != init_.sample_rate) throw std::invalid_argument("CudaResynthSynth::onEvent: sample rate differs from the initializer's");
    open();
    int64_t st[5];
    ca_tracks_state(pool_, st);
    ca_track_event ev{};
    ev.at_sample = maybeTimeAndSource ? std::max(st[0], nanoseconds_to_frames(maybeTimeAndSource->getNanosTime(), sample_rate)) : st[0];
    ev.at_sample = std::max(ev.at_sample, last_posted_);  // one-shots are applied in the order they were queued
    ev.noteid = e.noteid.noteid;
    switch (e.type) {
      case EventType::NoteOn:
        // channels are reserved when the event is posted (gen.crtp.h:398-414): the ones sounding plus the ones reserved since
        if (st[3] + reserved_ >= n_channels_per_note * n_voices_) {
          ++dropped_;
          return onEventResult::DROPPED_NOTE;
        }
        ++reserved_;
        ev.type = CA_NOTE_ON;
        ev.frequency = e.noteOn.frequency;
        ev.velocity = e.noteOn.velocity;
        on_.push_back(ev.noteid);
        break;
      case EventType::NoteChange:
        ev.type = CA_NOTE_CHANGE;
        ev.frequency = e.noteChange.changed_frequency;
        ev.velocity = e.noteChange.changed_velocity;
        break;
      case EventType::NoteOff:
        ev.type = CA_NOTE_OFF;
        on_.erase(std::remove(on_.begin(), on_.end(), ev.noteid), on_.end());
        break;
    }
    check(ca_tracks_post(pool_, &ev, 1), "ca_tracks_post");
    last_posted_ = ev.at_sample;
    return onEventResult::OK;
  }
  onEventResult onEvent(int const sample_rate, Event const& e, std::optional<TimestampAndSource> const& t = std::nullopt) {
    NoOut o;
    NoChans c;
    return onEvent(sample_rate, e, o, c, t);
  }

  // gen.crtp.h:301-319: release every sounding note
  void allNotesOff() {
    if (!pool_) return;
    std::vector<int64_t> const ids = on_;
    for (int64_t id : ids) onEvent(init_.sample_rate, mkNoteOff(NoteId{id}));
  }

  // ComputeFunc (include/simple_audio_context.h:20-26): accumulate n_frames interleaved frames; false when nothing sounds
  bool compute(double* buffer, int n_frames) {
    if (!pool_) return false;
    for (int i = 0; i < n_frames;) {
      if (cursor_ >= block_begin_ + block_len_) render_block();
      int64_t const off = cursor_ - block_begin_;
      int const take = static_cast<int>(std::min<int64_t>(n_frames - i, block_len_ - off));
      for (int j = 0; j < take * nOuts; ++j) buffer[static_cast<size_t>(i) * nOuts + j] += rendered_[static_cast<size_t>(off) * nOuts + j];
      i += take;
      cursor_ += take;
    }
    int64_t st[5];
    ca_tracks_state(pool_, st);
    return st[3] > 0 || st[4] > 0;
  }

  int64_t position() const { return cursor_; }
  int64_t dropped_notes() const { return dropped_; }

 private:
  void check(int rc, char const* what) {
    if (rc != CA_OK) throw std::runtime_error(std::string(what) + ": " + ca_last_error(ctx_));
  }
  void open() {
    if (pool_) return;
    ca_tracks_desc d{};
    d.sample_rate = init_.sample_rate;
    d.glide_samples = init_.stride;
    d.env = init_.ahdsr;
    d.max_channels = n_channels_per_note * n_voices_;
    d.stereo_spread = init_.stereo_spread;
    d.mix = nOuts == 2 ? CA_MIX_STEREO : CA_MIX_MONO;
    check(ca_tracks_open(ctx_, &d, &pool_), "ca_tracks_open");
  }
  void render_block() {
    block_begin_ = cursor_;
    block_len_ = block_;
    reserved_ = 0;  // the pool applies the pending note-ons of the block: its sounding count includes them from here on
    check(ca_tracks_pull_host(pool_, block_len_, rendered_.data()), "ca_tracks_pull_host");
  }

  ca_ctx* ctx_;
  int n_voices_;
  ResynthElementInitializer init_;
  ca_tracks* pool_ = nullptr;
  int64_t block_, block_begin_ = 0, block_len_ = 0, cursor_ = 0, last_posted_ = 0, reserved_ = 0, dropped_ = 0;
  std::vector<int64_t> on_;  // ids of the notes switched on and not yet off (allNotesOff)
  std::vector<float> rendered_;
};

}  // namespace resynth
}  // namespace ca_host
