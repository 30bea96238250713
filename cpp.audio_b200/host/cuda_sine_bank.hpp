// cuda_sine_bank.hpp — C++ host adapter that presents the reference's SYNTHESIS seam on top of the C ABI.
//
// Reference seam: SimpleAudioOutContext calls  ComputeFunc = bool(double* buffer, int nFrames)  which ACCUMULATES nFrames
// (<= 16) interleaved frames and returns false to be unregistered (include/simple_audio_context.h:20-26,69-112); above it
// ImplCRTP turns NoteOn/NoteOff events into element parameters (include/gen.crtp.h:382-622) and renders voice-major in
// compute() (:350-378). CudaSineBank keeps those names and meanings:
//
//   onEvent(mkNoteOn / mkNoteOff)   schedule a note of the harmonic bank at a sample time (events carry their time, as the
//                                   offline rt.resynth path does with MIDI timestamps, source/rt.resynth.lib.cpp:1601-1605)
//   compute(double* buffer, int n)  accumulate the next n frames; renders ahead on the GPU one block at a time
//                                   (ca_oscbank_render_host with first_sample), so the per-call cost is a memcpy
//   allNotesOff()                   release every note now
//
// Elements are VolumeAdjusted<MultiEnveloped<SineOscillatorAlgo,AHDSREnvelope>> (include/audioelement.h:486-657,1111-1265),
// base volume 0.1 (reduceUnadjustedVolumes*soundBaseVolume(SINE), :1270,2370), panned by StereoPanned when nOuts == 2.
// Limits: events must be posted before compute() reaches their sample time minus the render-ahead block (they are applied
// at the next block otherwise); note changes (glides) are not part of this adapter.
#pragma once

#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/cpp_audio_b200.h"

namespace ca_host {

struct NoteId {
  int64_t noteid;
};
enum class EventType { NoteOn, NoteOff };
struct Event {  // include/events.h:55-80, plus the sample time the offline path attaches
  EventType type;
  NoteId noteid;
  float frequency = 0.f;
  float velocity = 0.f;
  int64_t at_sample = 0;
};
inline Event mkNoteOn(NoteId id, float frequency, float velocity, int64_t at_sample) { return {EventType::NoteOn, id, frequency, velocity, at_sample}; }
inline Event mkNoteOff(NoteId id, int64_t at_sample) { return {EventType::NoteOff, id, 0.f, 0.f, at_sample}; }
enum class onEventResult { OK, DROPPED_NOTE };

struct harmonicProperties_t {  // built at source/main.tune.cpp:1998-2001
  float phase;
  float volume;
};

template <int nOuts>
class CudaSineBank {
 public:
  static constexpr float baseVolume = 0.1f;

  CudaSineBank(ca_ctx* ctx, int sample_rate, int n_max_voices, std::vector<harmonicProperties_t> harmonics, ca_ahdsr env,
               int block_frames = 1 << 15)
      : ctx_(ctx), sample_rate_(sample_rate), max_voices_(n_max_voices), harmonics_(std::move(harmonics)), env_(env), block_(block_frames) {
    static_assert(nOuts == 1 || nOuts == 2, "mono or stereo");
    if (ca_oscbank_plan(ctx_, max_voices_, static_cast<int32_t>(harmonics_.size())) != CA_OK)
      throw std::runtime_error(std::string("ca_oscbank_plan: ") + ca_last_error(ctx_));
    rendered_.resize(static_cast<size_t>(block_) * nOuts);
  }

  // pan in [-1,1] (include/stereo.h:18-27); start_angle in units of pi rad (DefaultStartPhase, gen.crtp.h:149-155)
  onEventResult onEvent(Event const& e, float pan = 0.f, float start_angle = 0.f) {
    if (e.type == EventType::NoteOn) {
      if (static_cast<int>(notes_.size()) >= max_voices_) {
        drop_finished();
        if (static_cast<int>(notes_.size()) >= max_voices_) return onEventResult::DROPPED_NOTE;  // gen.crtp.h:412-414
      }
      notes_.push_back({e.noteid.noteid, e.frequency, baseVolume * e.velocity, pan, start_angle, e.at_sample, -1});
      dirty_ = true;
      return onEventResult::OK;
    }
    for (auto& n : notes_)
      if (n.id == e.noteid.noteid && n.release_at < 0) {
        n.release_at = e.at_sample;
        dirty_ = true;
        return onEventResult::OK;
      }
    return onEventResult::OK;  // note off for a dropped note: ignored like the reference (gen.crtp.h:589-591)
  }

  void allNotesOff() {
    for (auto& n : notes_)
      if (n.release_at < 0) n.release_at = cursor_;
    dirty_ = true;
  }

  // ComputeFunc: accumulate n_frames interleaved frames into buffer; false when nothing is left to play
  bool compute(double* buffer, int n_frames) {
    bool silence = notes_.empty();
    for (int i = 0; i < n_frames;) {
      if (cursor_ >= block_begin_ + block_len_ || dirty_) render_block();
      int64_t const off = cursor_ - block_begin_;
      int const take = static_cast<int>(std::min<int64_t>(n_frames - i, block_len_ - off));
      for (int j = 0; j < take * nOuts; ++j) buffer[static_cast<size_t>(i) * nOuts + j] += rendered_[static_cast<size_t>(off) * nOuts + j];
      i += take;
      cursor_ += take;
    }
    return !silence;
  }

  int64_t position() const { return cursor_; }

 private:
  struct Note {
    int64_t id;
    float freq, volume, pan, start_angle;
    int64_t start, release_at;
  };

  void drop_finished() {
    // a note is finished once its longest release has elapsed: 2.5 periods of the fundamental or the envelope's release
    notes_.erase(std::remove_if(notes_.begin(), notes_.end(), [this](Note const& n) {
                   if (n.release_at < 0) return false;
                   double const period = sample_rate_ / std::max(1.f, n.freq);
                   int64_t const rel = std::max<int64_t>({static_cast<int64_t>(env_.release), static_cast<int64_t>(sample_rate_ / 1000), static_cast<int64_t>(0.5 + 2.5 * period)});
                   return cursor_ > n.release_at + rel + 16;
                 }),
                 notes_.end());
  }

  void render_block() {
    drop_finished();
    block_begin_ = cursor_;
    block_len_ = block_;
    dirty_ = false;
    std::fill(rendered_.begin(), rendered_.end(), 0.f);
    size_t const nv = notes_.size();
    if (nv == 0) return;
    size_t const np = harmonics_.size();
    f0_.resize(nv); vol_.resize(nv); sa_.resize(nv); pan_.resize(nv); kd_.resize(nv); ra_.resize(nv);
    pv_.resize(nv * np); pp_.resize(nv * np);
    // one global timeline: oscillator phases count from sample 0 of the bank (a voice's oscillators rotate during its key
    // delay, as MultiEnveloped::step does, include/audioelement.h:529-545), so key delays are absolute sample times
    int64_t const origin = 0;
    for (size_t v = 0; v < nv; ++v) {
      Note const& n = notes_[v];
      f0_[v] = n.freq; vol_[v] = n.volume; sa_[v] = n.start_angle; pan_[v] = n.pan;
      if (n.start - origin > 0x7fffffff) throw std::runtime_error("CudaSineBank: note start beyond 2^31 samples");
      kd_[v] = static_cast<int32_t>(n.start - origin);
      ra_[v] = n.release_at < 0 ? -1 : n.release_at - origin;
      for (size_t p = 0; p < np; ++p) { pv_[v * np + p] = harmonics_[p].volume; pp_[v * np + p] = harmonics_[p].phase; }
    }
    ca_oscbank_desc d{};
    d.n_voices = static_cast<int64_t>(nv); d.n_partials = static_cast<int32_t>(np); d.n_samples = block_len_; d.sample_rate = sample_rate_;
    d.f0_hz = f0_.data(); d.volume = vol_.data(); d.start_angle = sa_.data(); d.key_delay = kd_.data(); d.release_at = ra_.data();
    d.partial_volume = pv_.data(); d.partial_phase = pp_.data(); d.env = &env_; d.n_env = 1; d.pan = pan_.data();
    d.mix = nOuts == 2 ? CA_MIX_STEREO : CA_MIX_MONO;
    d.first_sample = cursor_ - origin;
    if (ca_oscbank_render_host(ctx_, &d, rendered_.data()) != CA_OK) throw std::runtime_error(std::string("ca_oscbank_render_host: ") + ca_last_error(ctx_));
  }

  ca_ctx* ctx_;
  int sample_rate_, max_voices_;
  std::vector<harmonicProperties_t> harmonics_;
  ca_ahdsr env_;
  int64_t block_, block_begin_ = 0, block_len_ = 0, cursor_ = 0;
  bool dirty_ = true;
  std::vector<Note> notes_;
  std::vector<float> rendered_, f0_, vol_, sa_, pan_, pv_, pp_;
  std::vector<int32_t> kd_;
  std::vector<int64_t> ra_;
};

}  // namespace ca_host
