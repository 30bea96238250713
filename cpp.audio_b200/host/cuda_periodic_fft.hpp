// cuda_periodic_fft.hpp — C++ host adapters that present the reference's ANALYSIS seam on top of the C ABI.
//
// The reference composes its analysis by C++ concepts (no FFI): PeriodicFFT<FftOperation> drives an operator that has
//   using FPT; using Result; Op(int reserve_for_max_fft_size, ...); bool init_data(int window_size);
//   template<It> void operator()(It begin, It end); Result const& get_results() const;
// (source/rt.resynth.lib.periodicfft.cpp:13-30,111,122-134,146-154). The types below have the same names, members and
// argument meaning, so host code written against the reference shapes compiles against them:
//
//   ca_host::SqMagFftOperation<Window::Gaussian, double>   ~ rtresynth::SqMagFftOperation  (periodicfft.cpp:251-325)
//   ca_host::FftOperation<double>                          ~ rtresynth::FftOperation       (periodicfft.cpp:183-244; consumed as
//                                                            PeriodicFFT<FftOperation<double>> by the vocoder's carrier,
//                                                            source/rt.resynth.lib.vocoder.cpp:544)
//   ca_host::PeriodicFFT<Op>                               ~ rtresynth::PeriodicFFT<Op>    (periodicfft.cpp:13-181)
//   ca_host::FrequenciesSqMag<T>, FreqMag<T>               ~ the cpp.algorithms result types used at
//                                                            source/rt.resynth.lib.cpp:1579-1595
//
// Difference that matters: PeriodicFFT here is BATCHED. feed() buffers samples and hands `batch_frames` frames at a time to
// the GPU (one ca_stft_analyze_host call), then fires on_fft_result for each frame in order; flush() forces the pending
// frames out. Results are identical to the per-frame reference; the callbacks are late by at most batch_frames frames.
// The peak list of every frame (extractLocalMaxFreqsMags, source/rt.resynth.lib.cpp:1592-1595) is computed on the GPU in
// the same pass and is available through last_peaks() inside the callback.
//
// Everything here is header-only C++17 and needs only include/cpp_audio_b200.h + libcpp_audio_b200.so (no CUDA headers).
#pragma once

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/cpp_audio_b200.h"

namespace ca_host {

enum class Window { Rectangular, Gaussian };

template <typename T>
struct FrequenciesSqMag {
  std::vector<T> frequencies_sqmag;
  int fft_length = 0;
  int get_fft_length() const { return fft_length; }
  double bin_index_to_Hz(int sample_rate) const { return sample_rate / static_cast<double>(fft_length); }
};

template <typename T>
struct FreqMag {
  T freq;
  T mag_db;
};

inline void check(ca_ctx* ctx, int rc, const char* what) {
  if (rc != CA_OK) throw std::runtime_error(std::string(what) + ": " + ca_last_error(ctx));
}

struct AnalysisConfig {
  int sample_rate = 48000;
  float min_db = -80.f;  // the reference drops reduced volumes < 1e-4 downstream (rt.resynth.lib.cpp:947)
  int max_peaks = 200;   // rt.resynth.lib.cpp:1102
  int hop = 256;         // needed at plan time; PeriodicFFT overrides it from its stride functor
};

// One frame at a time, exactly the reference operator concept. Slow by construction (one launch per frame); it exists so
// that code templated on the operator concept keeps compiling. Use PeriodicFFT below for throughput.
template <Window window_type, typename T>
struct SqMagFftOperation {
  using FPT = T;
  using Result = FrequenciesSqMag<T>;
  static int constexpr zero_padding_factor = 1;
  static int constexpr windowed_signal_stride = 1;

  SqMagFftOperation(int reserve_for_max_fft_size, ca_ctx* ctx, AnalysisConfig cfg = {}) : ctx_(ctx), cfg_(cfg) {
    int const max_fft_length = ca_fft_length_for(reserve_for_max_fft_size, zero_padding_factor);
    result_.frequencies_sqmag.reserve(max_fft_length / 2 + 1);
    staging_.reserve(reserve_for_max_fft_size);
    sq_.reserve(max_fft_length / 2 + 1);
  }

  bool init_data(int const window_size) {
    if (window_size % 2) throw std::logic_error("window size must be even");  // Assert at periodicfft.cpp:274
    if (window_size == window_) return false;
    plan(window_size, cfg_.hop);
    return true;  // the caller zeroes its samples (periodicfft.cpp:296)
  }

  template <typename It>
  void operator()(It const& it, It const& end) {
    replan_if_stale();
    staging_.assign(it, end);  // T -> float: the device path is float32
    int const n_bins = fft_length_ / 2 + 1;
    sq_.resize(n_bins);
    check(ctx_, ca_stft_analyze_host(ctx_, staging_.data(), static_cast<int64_t>(staging_.size()), 1, static_cast<int64_t>(staging_.size()),
                                     sq_.data(), n_bins, nullptr, nullptr, nullptr, nullptr),
          "ca_stft_analyze_host");
    result_.frequencies_sqmag.assign(sq_.begin(), sq_.end());
    result_.fft_length = fft_length_;
  }

  Result const& get_results() const { return result_; }

  // ---- used by PeriodicFFT: `frames` consecutive frames of `samples` (hop apart) in one GPU call ----
  template <typename OnFrame>
  void process_frames(float const* samples, int64_t n, int frames, OnFrame&& on_frame) {
    replan_if_stale();
    int const n_bins = fft_length_ / 2 + 1;
    int const P = cfg_.max_peaks;
    bsq_.resize(static_cast<size_t>(frames) * n_bins);
    pb_.resize(static_cast<size_t>(frames) * P);
    pf_.resize(pb_.size());
    pd_.resize(pb_.size());
    pc_.resize(frames);
    check(ctx_, ca_stft_analyze_host(ctx_, samples, n, 1, n, bsq_.data(), n_bins, pb_.data(), pf_.data(), pd_.data(), pc_.data()),
          "ca_stft_analyze_host");
    batch_result_.fft_length = fft_length_;
    for (int f = 0; f < frames; ++f) {
      batch_result_.frequencies_sqmag.assign(bsq_.begin() + static_cast<size_t>(f) * n_bins, bsq_.begin() + static_cast<size_t>(f + 1) * n_bins);
      found_ = pc_[f];
      int const stored = found_ < P ? found_ : P;
      peaks_.resize(stored);
      for (int i = 0; i < stored; ++i) peaks_[i] = {static_cast<T>(pf_[static_cast<size_t>(f) * P + i]), static_cast<T>(pd_[static_cast<size_t>(f) * P + i])};
      on_frame(batch_result_);
    }
  }
  std::vector<FreqMag<T>> const& last_peaks() const { return peaks_; }
  int last_peak_count_found() const { return found_; }

  void plan(int window_size, int hop) {
    ca_stft_params p{};
    p.window = window_size;
    p.hop = hop;
    p.window_type = window_type == Window::Gaussian ? CA_WINDOW_GAUSSIAN : CA_WINDOW_RECTANGULAR;
    p.sigma_factor = 4.f;  // periodicfft.cpp:288
    p.sample_rate = cfg_.sample_rate;
    p.min_db = cfg_.min_db;
    p.max_peaks = cfg_.max_peaks;
    check(ctx_, ca_stft_plan(ctx_, &p), "ca_stft_plan");
    generation_ = ca_stft_plan_generation(ctx_);
    window_ = window_size;
    hop_ = hop;
    fft_length_ = ca_fft_length_for(window_size, zero_padding_factor);
  }
  ca_ctx* ctx() const { return ctx_; }
  AnalysisConfig const& config() const { return cfg_; }
  int planned_hop() const { return hop_; }
  int fft_length() const { return fft_length_; }

 private:
  ca_ctx* ctx_;
  AnalysisConfig cfg_;
  // a ctx holds one analysis plan: when another operator planned on it since, plan again
  void replan_if_stale() {
    if (window_ > 0 && ca_stft_plan_generation(ctx_) != generation_) plan(window_, hop_);
  }
  int window_ = 0, hop_ = 0, fft_length_ = 0, found_ = 0;
  uint64_t generation_ = 0;
  std::vector<float> staging_, sq_, bsq_, pf_, pd_;
  std::vector<int32_t> pb_, pc_;
  std::vector<FreqMag<T>> peaks_;
  Result result_, batch_result_;
};

// The complex half spectrum of a frame: what the reference keeps in fft::RealFBins_<Tag, T, Alloc>::type, whose packing depends on
// the FFT backend; here plainly bins 0 .. N/2.
template <typename T>
struct RealFBins {
  std::vector<std::complex<T>> bins;
  int fft_length = 0;
  int get_fft_length() const { return fft_length; }
};

// FftOperation<T> (source/rt.resynth.lib.periodicfft.cpp:183-244): rectangular window (:228), zero padding to the FFT length
// (:231), forward real FFT, unscaled (:234; fft::Algo_::forward) on K1 (ca_rfft_batched_host).
template <typename T>
struct FftOperation {
  using FPT = T;
  using Result = RealFBins<T>;
  static int constexpr zero_padding_factor = 1;

  FftOperation(int reserve_for_max_fft_size, ca_ctx* ctx) : ctx_(ctx) {
    int const max_fft_length = ca_fft_length_for(reserve_for_max_fft_size, zero_padding_factor);
    result_.bins.reserve(max_fft_length / 2 + 1);
    frames_.reserve(max_fft_length);
    out_.reserve(2 * (max_fft_length / 2 + 1));
  }

  bool init_data(int const window_size) {  // :192-208: nothing depends on the window but the FFT length
    int const n = ca_fft_length_for(window_size, zero_padding_factor);
    if (n < 512 || n > (1 << 20)) throw std::logic_error("FftOperation: fft length must be within 2^9 .. 2^20");
    window_ = window_size;
    fft_length_ = n;
    return false;
  }

  template <typename It>
  void operator()(It const& it, It const& end) {
    frames_.assign(static_cast<size_t>(fft_length_), 0.f);
    size_t i = 0;
    for (It p = it; p != end && i < frames_.size(); ++p, ++i) frames_[i] = static_cast<float>(*p);
    transform(1);
    take(0, result_);
  }
  Result const& get_results() const { return result_; }

  template <typename OnFrame>
  void process_frames(float const* samples, int64_t n, int frames, OnFrame&& on_frame) {
    (void)n;
    frames_.assign(static_cast<size_t>(frames) * fft_length_, 0.f);
    for (int f = 0; f < frames; ++f)
      std::copy(samples + static_cast<int64_t>(f) * hop_, samples + static_cast<int64_t>(f) * hop_ + window_, frames_.begin() + static_cast<size_t>(f) * fft_length_);
    transform(frames);
    for (int f = 0; f < frames; ++f) {
      take(f, batch_result_);
      on_frame(batch_result_);
    }
  }
  void plan(int window_size, int hop) {
    init_data(window_size);
    hop_ = hop;
  }
  ca_ctx* ctx() const { return ctx_; }
  int fft_length() const { return fft_length_; }

 private:
  void transform(int frames) {
    out_.resize(static_cast<size_t>(frames) * 2 * (fft_length_ / 2 + 1));
    check(ctx_, ca_rfft_batched_host(ctx_, frames_.data(), fft_length_, frames, out_.data()), "ca_rfft_batched_host");
  }
  void take(int f, Result& r) const {
    int const nb = fft_length_ / 2 + 1;
    r.bins.resize(nb);
    r.fft_length = fft_length_;
    float const* o = out_.data() + static_cast<size_t>(f) * 2 * nb;
    for (int k = 0; k < nb; ++k) r.bins[k] = {static_cast<T>(o[2 * k]), static_cast<T>(o[2 * k + 1])};
  }
  ca_ctx* ctx_;
  int window_ = 0, hop_ = 1, fft_length_ = 0;
  std::vector<float> frames_, out_;
  Result result_, batch_result_;
};

template <typename FftOperation>
struct PeriodicFFT {
  using FftOp = FftOperation;
  using FPT = typename FftOperation::FPT;
  using Result = typename FftOperation::Result;
  using OnFftResult = std::function<void(int const, Result const&)>;
  using GetWindowSizeFunc = std::function<int()>;
  using GetWindowCenterStrideFramesFunc = std::function<int()>;

  template <typename... Args>
  PeriodicFFT(int reserve_for_max_fft_size, Args&&... args) : fft_op(reserve_for_max_fft_size, std::forward<Args>(args)...) {}

  // frames handed to the GPU per call; 1 reproduces the reference's per-frame latency
  void set_batch_frames(int n) { batch_frames_ = n < 1 ? 1 : n; }

  void setLambdas(GetWindowSizeFunc&& get_window_size, GetWindowCenterStrideFramesFunc&& get_window_center_stride_frames,
                  OnFftResult&& on_fft_result) {
    this->get_window_size = std::move(get_window_size);
    this->get_window_center_stride_frames = std::move(get_window_center_stride_frames);
    this->on_fft_result = std::move(on_fft_result);
    window_ = this->get_window_size();
    hop_ = this->get_window_center_stride_frames();
    fft_op.plan(window_, hop_);
    reset_samples();
  }

  // The window size and the stride are read when setLambdas runs and after every flush (the reference re-reads them after
  // every frame, periodicfft.cpp:138,156; a change takes effect at the next batch boundary here).
  void feed(float const sample) {
    pending_.push_back(sample);
    if (static_cast<int64_t>(pending_.size()) >= needed_for(batch_frames_)) process(batch_frames_);
  }

  // run whatever complete frames are pending
  void flush() {
    int64_t const n = static_cast<int64_t>(pending_.size());
    if (n < window_) return;
    process(static_cast<int>(1 + (n - window_) / hop_));
  }

  void reset_samples() { pending_.clear(); }

  // the peak list of the frame the callback is being called for (SqMagFftOperation only)
  auto const& last_peaks() const { return fft_op.last_peaks(); }
  int last_peak_count_found() const { return fft_op.last_peak_count_found(); }

 private:
  int64_t needed_for(int frames) const { return static_cast<int64_t>(window_) + static_cast<int64_t>(frames - 1) * hop_; }

  void process(int frames) {
    fft_op.process_frames(pending_.data(), needed_for(frames), frames, [this](Result const& r) { on_fft_result(hop_, r); });
    // keep the overlap: the next frame starts `frames * hop` samples later (periodicfft.cpp:167-172)
    int64_t const consumed = static_cast<int64_t>(frames) * hop_;
    if (consumed >= static_cast<int64_t>(pending_.size())) pending_.clear();
    else pending_.erase(pending_.begin(), pending_.begin() + consumed);
    int const w = get_window_size(), h = get_window_center_stride_frames();
    if (w != window_ || h != hop_) {  // periodicfft.cpp:122-134,156-158: a new window size zeroes the samples
      if (w != window_) pending_.clear();
      window_ = w;
      hop_ = h;
      fft_op.plan(window_, hop_);
    }
  }

  FftOperation fft_op;
  OnFftResult on_fft_result;
  GetWindowSizeFunc get_window_size;
  GetWindowCenterStrideFramesFunc get_window_center_stride_frames;
  int window_ = 0, hop_ = 1, batch_frames_ = 64;
  std::vector<float> pending_;
};

}  // namespace ca_host
