// oracle/oracle.hpp — CPU restatement of cpp.audio's rt.resynth analysis + sine-bank synthesis hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under cpp.audio_b200/ (the product) includes, links or runs this file.
// It is used by tests/, by __graft_entry__.smoke() and by the cpu_baseline / --impl reference legs of bench.py.
//
// PARITY STATUS: the synthesis half restates in-tree reference code and is pinned by it; the FFT / window /
// peak-interpolation arithmetic lives in the un-vendored sibling repo cpp.algorithms (include/public.h:13) and is
// therefore "PARITY UNPINNED" — see oracle/SPEC.md (decisions S1..S9).
//
// All citations are path:line under /root/reference/.
#pragma once

#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <complex>
#include <cstdint>
#include <functional>
#include <limits>
#include <optional>
#include <stdexcept>
#include <utility>
#include <vector>

namespace orc {

constexpr double kPi = 3.141592653589793238462643383279502884;

// ------------------------------------------------------------------------------------------------------------
// FFT (the reference delegates to fft::Fastest in cpp.algorithms; call sites
// source/rt.resynth.lib.periodicfft.cpp:214-234, :304-311; source/main.test_fft.cpp:78).
// Forward transform, unscaled: X[k] = sum_n x[n] exp(-2 pi i k n / N).
// ------------------------------------------------------------------------------------------------------------

inline bool is_pow2(int64_t n) { return n > 0 && (n & (n - 1)) == 0; }

template <class T>
class ComplexFft {
public:
  explicit ComplexFft(int n) : n_(n) {
    if (!is_pow2(n)) throw std::invalid_argument("fft size must be a power of two");
    tw_.resize(std::max(1, n / 2));
    for (int k = 0; k < n / 2; ++k) {
      long double a = -2.0L * static_cast<long double>(kPi) * k / n;
      tw_[k] = {static_cast<T>(std::cos(a)), static_cast<T>(std::sin(a))};
    }
    rev_.resize(n);
    int bits = 0;
    while ((1 << bits) < n) ++bits;
    for (int i = 0; i < n; ++i) {
      uint32_t r = 0;
      for (int b = 0; b < bits; ++b)
        if (i & (1 << b)) r |= 1u << (bits - 1 - b);
      rev_[i] = r;
    }
  }
  int size() const { return n_; }

  // in-place, iterative radix-2 decimation in time
  void forward(std::complex<T>* a) const {
    for (int i = 0; i < n_; ++i) {
      int j = static_cast<int>(rev_[i]);
      if (i < j) std::swap(a[i], a[j]);
    }
    for (int len = 2; len <= n_; len <<= 1) {
      int half = len >> 1;
      int step = n_ / len;
      for (int i = 0; i < n_; i += len) {
        for (int j = 0; j < half; ++j) {
          std::complex<T> w = tw_[j * step];
          T xr = a[i + j + half].real(), xi = a[i + j + half].imag();
          T tr = xr * w.real() - xi * w.imag();
          T ti = xr * w.imag() + xi * w.real();
          T ur = a[i + j].real(), ui = a[i + j].imag();
          a[i + j] = {ur + tr, ui + ti};
          a[i + j + half] = {ur - tr, ui - ti};
        }
      }
    }
  }

private:
  int n_;
  std::vector<std::complex<T>> tw_;
  std::vector<uint32_t> rev_;
};

// N real samples -> N/2+1 complex bins, through one N/2-point complex FFT and the usual untangling step.
template <class T>
class RealFft {
public:
  explicit RealFft(int n) : n_(n), half_(std::max(1, n / 2)), work_(std::max(1, n / 2)) {
    if (!is_pow2(n) || n < 2) throw std::invalid_argument("real fft size must be a power of two >= 2");
    tw_.resize(n / 2 + 1);
    for (int k = 0; k <= n / 2; ++k) {
      long double a = -2.0L * static_cast<long double>(kPi) * k / n;
      tw_[k] = {static_cast<T>(std::cos(a)), static_cast<T>(std::sin(a))};
    }
  }
  int size() const { return n_; }
  int bins() const { return n_ / 2 + 1; }

  void forward(const T* x, std::complex<T>* X) {
    int h = n_ / 2;
    for (int i = 0; i < h; ++i) work_[i] = {x[2 * i], x[2 * i + 1]};
    half_.forward(work_.data());
    for (int k = 0; k <= h; ++k) {
      std::complex<T> zk = work_[k % h];
      std::complex<T> zm = std::conj(work_[(h - k) % h]);
      std::complex<T> even = (zk + zm) * static_cast<T>(0.5);
      std::complex<T> odd = (zk - zm) * std::complex<T>(0, static_cast<T>(-0.5));
      X[k] = even + tw_[k] * odd;
    }
  }

  // unscaled inverse: x[n] = sum over the Hermitian-extended spectrum of X[k] exp(+2 pi i k n / N); inverse(forward(x)) = N x
  // (the reference's fft.inverse + extractRealOutput, source/rt.resynth.lib.vocoder.cpp:452-467). Plain full-size complex FFT.
  void inverse(const std::complex<T>* X, T* x) {
    if (!full_) full_.emplace(n_);
    full_work_.resize(n_);
    for (int k = 0; k <= n_ / 2; ++k) full_work_[k] = std::conj(X[k]);
    for (int k = n_ / 2 + 1; k < n_; ++k) full_work_[k] = X[n_ - k];
    full_->forward(full_work_.data());  // conj(sum X e^{+}) = sum conj(X) e^{-}
    for (int i = 0; i < n_; ++i) x[i] = full_work_[i].real();
  }

private:
  int n_;
  ComplexFft<T> half_;
  std::vector<std::complex<T>> work_;
  std::vector<std::complex<T>> tw_;
  std::optional<ComplexFft<T>> full_;
  std::vector<std::complex<T>> full_work_;
};

// ------------------------------------------------------------------------------------------------------------
// Windows and spectrum (SPEC S1..S5; call sites source/rt.resynth.lib.periodicfft.cpp:273-311)
// ------------------------------------------------------------------------------------------------------------

enum class Window : int { Rectangular = 0, Gaussian = 1 };

inline int fft_length_for(int n_samples, int zero_padding_factor) {  // S1
  int64_t want = static_cast<int64_t>(n_samples) * zero_padding_factor;
  int n = 1;
  while (n < want) n <<= 1;
  return n;
}

// S2 + S3: half window (index = distance from the centre), normalised so that the full symmetric window has mean 1.
template <class T>
std::vector<T> make_half_window(Window type, double sigma_factor, int half_size) {
  std::vector<T> hw(half_size);
  long double sum = 0;
  for (int j = 0; j < half_size; ++j) {
    long double v = 1;
    if (type == Window::Gaussian) {
      long double u = static_cast<long double>(sigma_factor) * (j + 0.5L) / half_size;
      v = std::exp(-0.5L * u * u);
    }
    hw[j] = static_cast<T>(v);
    sum += v;
  }
  long double scale = half_size / sum;  // mean over the full window == mean over the half window
  for (auto& v : hw) v = static_cast<T>(v * scale);
  return hw;
}

template <class T>
inline T window_at(const std::vector<T>& hw, int n) {  // full window of length 2*hw.size()
  int H = static_cast<int>(hw.size());
  return n < H ? hw[H - 1 - n] : hw[n - H];
}

// Restates SqMagFftOperation<window,T> (source/rt.resynth.lib.periodicfft.cpp:251-325):
// init_data(W) builds the normalised half window; operator() = findFrequenciesSqMag (window, zero pad, real FFT, |X|^2).
template <class T>
class SqMagOp {
public:
  using FPT = T;
  using Result = std::vector<T>;
  SqMagOp(Window type, double sigma_factor) : type_(type), sigma_(sigma_factor) {}

  bool init_data(int window_size) {
    if (window_size % 2) throw std::invalid_argument("window size must be even");  // Assert at :274
    int n = fft_length_for(window_size, 1);
    if (!fft_ || fft_->size() != n) {
      fft_.emplace(n);
      signal_.assign(n, T(0));
      bins_.assign(n / 2 + 1, {});
      sqmag_.assign(n / 2 + 1, T(0));
    }
    if (static_cast<int>(half_window_.size()) != window_size / 2) {
      half_window_ = make_half_window<T>(type_, sigma_, window_size / 2);
      return true;  // :296 — caller zeroes its samples
    }
    return false;
  }

  template <class It>
  void operator()(It begin, It end) {
    int w = static_cast<int>(end - begin);
    int n = fft_->size();
    for (int i = 0; i < w; ++i) signal_[i] = static_cast<T>(begin[i]) * window_at(half_window_, i);
    for (int i = w; i < n; ++i) signal_[i] = T(0);
    fft_->forward(signal_.data(), bins_.data());
    T scale = T(1) / static_cast<T>(w);  // S4
    for (int k = 0; k <= n / 2; ++k) {   // S5
      T re = bins_[k].real() * scale, im = bins_[k].imag() * scale;
      sqmag_[k] = re * re + im * im;
    }
  }

  const std::vector<T>& get_results() const { return sqmag_; }
  int fft_length() const { return fft_ ? fft_->size() : 0; }

private:
  Window type_;
  double sigma_;
  std::optional<RealFft<T>> fft_;
  std::vector<T> half_window_, signal_, sqmag_;
  std::vector<std::complex<T>> bins_;
};

// Restates FftOperation<T> (source/rt.resynth.lib.periodicfft.cpp:183-244): rectangular window (:228), zero padding to the
// FFT length (:231), forward real FFT (:234), unscaled; the result is the complex half spectrum (N/2+1 bins).
template <class T>
class FftOp {
public:
  using FPT = T;
  using Result = std::vector<std::complex<T>>;

  bool init_data(int window_size) {
    int n = fft_length_for(window_size, 1);
    if (!fft_ || fft_->size() != n) {
      fft_.emplace(n);
      signal_.assign(n, T(0));
      bins_.assign(n / 2 + 1, {});
    }
    return false;
  }

  template <class It>
  void operator()(It begin, It end) {
    int w = static_cast<int>(end - begin);
    int n = fft_->size();
    for (int i = 0; i < w; ++i) signal_[i] = static_cast<T>(begin[i]);
    for (int i = w; i < n; ++i) signal_[i] = T(0);
    fft_->forward(signal_.data(), bins_.data());
  }

  const Result& get_results() const { return bins_; }
  int fft_length() const { return fft_ ? fft_->size() : 0; }
  RealFft<T>& fft() { return *fft_; }

private:
  std::optional<RealFft<T>> fft_;
  std::vector<T> signal_;
  Result bins_;
};

// Restates PeriodicFFT<Op> (source/rt.resynth.lib.periodicfft.cpp:13-181): sample-at-a-time feed into a window
// buffer, operator fired when the buffer is full, then the overlap is shifted left by the stride.
template <class Op>
class PeriodicFft {
public:
  using FPT = typename Op::FPT;
  using OnResult = std::function<void(int, const typename Op::Result&)>;

  template <class... Args>
  explicit PeriodicFft(Args&&... args) : op_(std::forward<Args>(args)...) {}

  void set_lambdas(std::function<int()> get_window_size, std::function<int()> get_stride, OnResult on_result) {
    get_window_size_ = std::move(get_window_size);
    get_stride_ = std::move(get_stride);
    on_result_ = std::move(on_result);
    ignore_ = 0;
    init_data();
    reset_samples();
  }

  void feed(float sample) {  // :55-67
    if (ignore_ > 0) {
      --ignore_;
      return;
    }
    samples_[end_] = sample;
    ++end_;
    if (end_ == window_) {
      end_ = 0;
      on_full_buffer();
    }
  }

  void reset_samples() { samples_.assign(window_, FPT(0)); }
  Op& op() { return op_; }

private:
  bool init_data() {  // :122-134
    bool changed = false;
    int w = get_window_size_();
    if (w != window_) {
      window_ = w;
      changed = true;
    }
    return op_.init_data(window_) || changed;
  }

  void on_full_buffer() {  // :136-180
    int stride = get_stride_();
    op_(samples_.begin(), samples_.begin() + window_);
    on_result_(stride, op_.get_results());
    if (init_data()) reset_samples();
    int overlap = window_ - stride;
    if (overlap >= 0) {
      for (; end_ < overlap; ++end_) samples_[end_] = samples_[end_ + stride];
      ignore_ = 0;
    } else {
      ignore_ = -overlap;
    }
  }

  Op op_;
  std::vector<FPT> samples_;
  int window_ = 0, end_ = 0, ignore_ = 0;
  std::function<int()> get_window_size_, get_stride_;
  OnResult on_result_;
};

inline int64_t num_frames(int64_t n_samples, int window, int hop) {
  if (n_samples < window) return 0;
  if (hop >= window) return 1 + (n_samples - window) / hop;  // ignore_frames path, same count
  return 1 + (n_samples - window) / hop;
}

// ------------------------------------------------------------------------------------------------------------
// Peak extraction (SPEC S6; call site source/rt.resynth.lib.cpp:1592-1595)
// ------------------------------------------------------------------------------------------------------------

template <class T>
struct FreqMag {
  T freq;
  T mag_db;
  int bin;
};

template <class T>
inline T sqmag_to_db(T s) {
  T floor_v = std::is_same<T, float>::value ? T(1e-37) : T(1e-300);
  return T(10) * std::log10(std::max(s, floor_v));
}
template <class T>
inline T db_to_mag(T db) { return std::pow(T(10), db / T(20)); }

template <class T>
void extract_local_max_freqs_mags(T sample_rate, const T* sqmag, int fft_length, T min_db,
                                  std::vector<FreqMag<T>>& out) {
  out.clear();
  int half = fft_length / 2;
  T bin_to_hz = sample_rate / static_cast<T>(fft_length);
  for (int k = 1; k <= half - 1; ++k) {
    T c = sqmag[k];
    if (!(c > sqmag[k - 1] && c > sqmag[k + 1])) continue;
    T beta = sqmag_to_db(c);
    if (!(beta >= min_db)) continue;
    T alpha = sqmag_to_db(sqmag[k - 1]);
    T gamma = sqmag_to_db(sqmag[k + 1]);
    T denom = alpha - T(2) * beta + gamma;
    T p = denom != T(0) ? T(0.5) * (alpha - gamma) / denom : T(0);
    out.push_back({(static_cast<T>(k) + p) * bin_to_hz, beta - T(0.25) * (alpha - gamma) * p, k});
  }
}

// ------------------------------------------------------------------------------------------------------------
// Interpolation (SPEC S7; call sites include/audioelement.h:733-737,994-998,2774)
// ------------------------------------------------------------------------------------------------------------

enum Itp : int {
  LINEAR = 0,
  EASE_IN_QUAD = 1, EASE_OUT_QUAD = 2, EASE_INOUT_QUAD = 3,
  EASE_IN_CUBIC = 4, EASE_OUT_CUBIC = 5, EASE_INOUT_CUBIC = 6,
  EASE_IN_QUART = 7, EASE_OUT_QUART = 8, EASE_INOUT_QUART = 9,
  EASE_IN_QUINT = 10, EASE_OUT_QUINT = 11, EASE_INOUT_QUINT = 12,
  EASE_IN_SINE = 13, EASE_OUT_SINE = 14, EASE_INOUT_SINE = 15,
  EASE_IN_EXPO = 16, EASE_OUT_EXPO = 17, EASE_INOUT_EXPO = 18,
  EASE_IN_CIRC = 19, EASE_OUT_CIRC = 20, EASE_INOUT_CIRC = 21,
};

// Penner's easing set, written as b + c*e(t/d) with e(0) = 0 and e(1) = 1 (the exponential family jumps at its
// end points exactly as the published equations do: e(0+) = 2^-10 for "in", e(1-) = 1 - 2^-10 for "out").
template <class T>
T interpolate(int type, T t, T b, T c, T d) {
  T u = t / d;
  auto in_pow = [](T x, int p) { T r = 1; for (int i = 0; i < p; ++i) r *= x; return r; };
  if (type == LINEAR) return b + c * u;
  const T kHalfPi = T(1.5707963267948966192313216916398);
  int fam = (type - 1) / 3;      // 0 quad .. 3 quint, 4 sine, 5 expo, 6 circ
  int mode = (type - 1) % 3;     // 0 in, 1 out, 2 inout
  T e;
  if (fam <= 3) {
    int p = fam + 2;
    if (mode == 0) e = in_pow(u, p);
    else if (mode == 1) e = T(1) - in_pow(T(1) - u, p);
    else {
      if (u < T(0.5)) e = T(0.5) * in_pow(T(2) * u, p);
      else e = T(1) - T(0.5) * in_pow(T(2) - T(2) * u, p);
    }
  } else if (fam == 4) {
    if (mode == 0) e = T(1) - std::cos(u * kHalfPi);
    else if (mode == 1) e = std::sin(u * kHalfPi);
    else e = T(0.5) * (T(1) - std::cos(T(2) * kHalfPi * u));
  } else if (fam == 5) {
    if (mode == 0) e = (t == T(0)) ? T(0) : std::exp2(T(10) * (u - T(1)));
    else if (mode == 1) e = (t == d) ? T(1) : T(1) - std::exp2(T(-10) * u);
    else {
      if (t == T(0)) e = T(0);
      else if (t == d) e = T(1);
      else if (u < T(0.5)) e = T(0.5) * std::exp2(T(10) * (T(2) * u - T(1)));
      else e = T(0.5) * (T(2) - std::exp2(T(-10) * (T(2) * u - T(1))));
    }
  } else {
    auto root = [](T x) { return std::sqrt(x > T(0) ? x : T(0)); };
    if (mode == 0) e = T(1) - root(T(1) - u * u);
    else if (mode == 1) e = root(T(1) - (u - T(1)) * (u - T(1)));
    else {
      T w = T(2) * u;
      if (w < T(1)) e = T(0.5) * (T(1) - root(T(1) - w * w));
      else e = T(0.5) * (root(T(1) - (w - T(2)) * (w - T(2))) + T(1));
    }
  }
  return b + c * e;
}

// ------------------------------------------------------------------------------------------------------------
// Oscillator (include/audioelement.h:2366-2452) and aliasing gain (:466-483)
// ------------------------------------------------------------------------------------------------------------

template <class T>
T freq_aliasing_multiplicator(T angle_increment) {
  if (angle_increment == T(0)) return T(1);
  T half_spp = T(1) / angle_increment;
  if (half_spp < T(1)) return T(0);
  if (half_spp > T(4)) return T(1);
  return (half_spp - T(1)) * (T(1) / T(3));
}

template <class T>
T freq_to_angle_increment(T freq, int sample_rate) { return 2 * freq / sample_rate; }  // include/sound.functions.h:58-62

template <class T>
T angle_increment_to_period_in_continuous_samples(T i) {  // include/sound.functions.h:70-76
  if (i == T(0)) return T(1);
  return T(2) / std::abs(i);
}

template <class T>
class SineOsc {
public:
  void set_start_angle(T a) { cre_ = std::cos(static_cast<T>(kPi) * a); cim_ = std::sin(static_cast<T>(kPi) * a); }  // :2401
  void set_angle_increments(T f) {                                                                                  // :2404
    if (ai_ && *ai_ == f) return;
    ai_ = f;
    mre_ = std::cos(static_cast<T>(kPi) * f);
    mim_ = std::sin(static_cast<T>(kPi) * f);
    alias_ = freq_aliasing_multiplicator(f);
  }
  void synchronize_angles(const SineOsc& o) { cre_ = o.cre_; cim_ = o.cim_; }
  void step() {                                                                                                     // :2413, :2443
    T r = cre_ * mre_ - cim_ * mim_;
    T i = cre_ * mim_ + cim_ * mre_;
    T k = T(0.5) * (T(3) - (r * r + i * i));
    cre_ = r * k;
    cim_ = i * k;
  }
  T imag() const { return alias_ * cim_; }
  T real() const { return alias_ * cre_; }
  T angle_increments() const { return ai_ ? *ai_ : T(0); }

private:
  T cre_ = 1, cim_ = 0, mre_ = 1, mim_ = 0, alias_ = 1;
  std::optional<T> ai_;
};

// ------------------------------------------------------------------------------------------------------------
// AHDSR envelope (include/audioelement.h:314-335, :666-842, :855-1099), WaitForKeyRelease / AllowZeroAttack::No
// ------------------------------------------------------------------------------------------------------------

struct Ahdsr {
  int32_t attack; int32_t attack_itp;
  int32_t hold;
  int32_t decay; int32_t decay_itp;
  int32_t release; int32_t release_itp;
  float sustain;
};

enum class EnvState : uint8_t { SoonKeyPressed, KeyPressed, KeyReleased, Done1, Done2 };
constexpr int kFramesPerBuffer = 16;  // include/audioelement.h:3

template <class T>
class Envelope {
public:
  void set_ahdsr(const Ahdsr& s, int sample_rate) {  // :923-952
    int32_t min_dt = sample_rate / 1000;            // :863-872
    bool has_decay = s.sustain < 0.999999;
    s_minus_one_ = has_decay ? std::min(std::max(s.sustain, 0.f), 1.f) - 1.f : 0.f;
    A_ = std::max(s.attack, min_dt);
    H_ = std::max(s.hold, 0);
    D_ = has_decay ? std::max(s.decay, min_dt) : 0;
    R_ = std::max(s.release, min_dt);
    attack_itp_ = s.attack_itp; decay_itp_ = s.decay_itp; release_itp_ = s.release_itp;
  }
  void set_min_change_duration_samples(int n) { next_min_change_ = n; }  // :857-861
  void forget_past_signals() { counter_ = 0; min_change_ = next_min_change_; }  // :694-700
  bool try_acquire() { if (state_ != EnvState::Done2) return false; state_ = EnvState::SoonKeyPressed; return true; }

  void on_key_pressed(int32_t delay) {  // :762-767
    state_ = EnvState::KeyPressed;
    counter_ = -delay;
    release_delay_ = -1;
    value_ = 0;
  }
  void on_key_released(int32_t delay) {  // :769-781
    if (state_ != EnvState::KeyPressed) return;
    if (counter_ + delay <= 0) state_ = EnvState::Done1;
    else release_delay_ = delay;
  }

  void step() {  // :702-756
    if (release_delay_ >= 0) {
      if (release_delay_ == 0) start_released();
      --release_delay_;
    }
    ++counter_;
    switch (state_) {
      case EnvState::KeyPressed:
        if (counter_ < 1) return;
        if (counter_ == 1) { ahd_ = 0; on_ahd_change(); }
        value_ = step_pressed();
        break;
      case EnvState::KeyReleased:
        value_ = interpolate<T>(release_itp_, static_cast<T>(counter_), top_, -top_, static_cast<T>(release_time()));
        if (counter_ >= release_time()) { state_ = EnvState::Done1; counter_ = 0; value_ = 0; }
        break;
      case EnvState::Done1:
        if (counter_ > kFramesPerBuffer) { state_ = EnvState::Done2; counter_ = 0; }
        value_ = 0;
        break;
      case EnvState::Done2:
      case EnvState::SoonKeyPressed:
        value_ = 0;
        break;
    }
  }

  T value() const { return value_; }
  EnvState state() const { return state_; }
  bool rt_active() const { return state_ != EnvState::SoonKeyPressed && state_ != EnvState::Done2; }
  bool finished() const { return state_ == EnvState::Done2; }
  int count_steps_to_release() const { return release_delay_; }  // :783-785

private:
  // ahd_: 0 attacking, 1 holding, 2 decaying, 3 none (sustain)
  int32_t max_counter() const {  // :1066-1084
    switch (ahd_) {
      case 0: return std::max(min_change_, A_);
      case 1: return H_;
      default: return D_ == 0 ? 0 : std::max(min_change_, D_);
    }
  }
  void on_ahd_change() {  // :1040-1053
    ahd_counter_ = 0;
    min_change_ = next_min_change_;
    if (ahd_ == 3) return;
    if (max_counter() == 0) step_ahd();
  }
  void step_ahd() {  // :1055-1064
    if (ahd_counter_ < max_counter()) return;
    ++ahd_;
    on_ahd_change();
  }
  T step_pressed() {  // :960-1000
    if (ahd_ != 3) { step_ahd(); ++ahd_counter_; }
    if (ahd_ == 3) return T(1) + static_cast<T>(s_minus_one_);
    T from, delta; int itp;
    if (ahd_ == 0) { from = 0; delta = 1; itp = attack_itp_; }
    else if (ahd_ == 1) { from = 1; delta = 0; itp = LINEAR; }
    else { from = 1; delta = static_cast<T>(s_minus_one_); itp = decay_itp_; }
    return interpolate<T>(itp, static_cast<T>(ahd_counter_), from, delta, static_cast<T>(max_counter()));
  }
  int32_t release_time() const { return std::max(min_change_, R_); }  // :1002-1004
  void start_released() {  // :836-841
    min_change_ = next_min_change_;
    state_ = EnvState::KeyReleased;
    top_ = value_;
    counter_ = 0;
  }

  T value_ = 0, top_ = 0;
  EnvState state_ = EnvState::Done2;
  int32_t counter_ = 0, release_delay_ = -1;
  int32_t min_change_ = 0, next_min_change_ = 0;
  int32_t ahd_counter_ = 0, A_ = 100000, H_ = 0, D_ = 100000, R_ = 100000;
  float s_minus_one_ = -0.5f;
  int ahd_ = 3;
  int attack_itp_ = LINEAR, decay_itp_ = LINEAR, release_itp_ = LINEAR;
};

// Enveloped<SineOscillatorAlgo, AHDSREnvelope> (include/audioelement.h:182-291)
template <class T>
class EnvelopedOsc {
public:
  void forget_past_signals() { env.forget_past_signals(); }
  void set_angle_increments(T a) {  // :218-225
    T period = angle_increment_to_period_in_continuous_samples(a);
    env.set_min_change_duration_samples(static_cast<int>(0.5f + T(2.5) * period));
    osc.set_angle_increments(a);
  }
  void step() { env.step(); osc.step(); }      // :197-200
  T imag() const { return osc.imag() * env.value(); }  // :231
  Envelope<T> env;
  SineOsc<T> osc;
};

// ------------------------------------------------------------------------------------------------------------
// Harmonic bank = MultiEnveloped (include/audioelement.h:486-657) inside VolumeAdjusted (:1111-1265)
// ------------------------------------------------------------------------------------------------------------

struct HarmonicProps { float phase; float volume; };  // harmonicProperties_t (ext.; built at source/main.tune.cpp:1998-2001)

// SPEC S8 one-pole low pass standing in for Filter<T,1,LOW_PASS,1>
template <class T>
struct OnePole {
  T a = 1, y = 0;
  void init_with_angle_increment(T inc) { a = T(1) - std::exp(-static_cast<T>(kPi) * inc); }
  void set_initial_value(T v) { y = v; }
  T feed(T x) { y += a * (x - y); return y; }
};

template <class T>
class BankVoice {
public:
  void set_harmonics(const HarmonicProps* props, int n) {  // :507-521 (trailing inaudible harmonics dropped, :401-415)
    int last = -1;
    for (int i = n - 1; i >= 0; --i)
      if (std::abs(props[i].volume) > 0.000001f) { last = i; break; }
    harmonics_.assign(last + 1, {});
    props_.assign(props, props + last + 1);
  }
  void set_ahdsr(const Ahdsr& s, int sr) { for (auto& h : harmonics_) h.env.set_ahdsr(s, sr); }
  void forget_past_signals() { for (auto& h : harmonics_) h.forget_past_signals(); volume_.reset(); target_.reset(); }
  void set_volume_target(T v) { target_ = v; }
  void set_angle_increments(T a) {  // :614-619
    for (size_t i = 0; i < harmonics_.size(); ++i) harmonics_[i].set_angle_increments(static_cast<T>(i + 1) * a);
  }
  void set_start_angle(T a) {  // :590-594
    for (size_t i = 0; i < harmonics_.size(); ++i)
      harmonics_[i].osc.set_start_angle(props_[i].phase + static_cast<T>(i + 1) * a);
  }
  void on_key_pressed(int32_t delay) { for (auto& h : harmonics_) h.env.on_key_pressed(delay); state_ = EnvState::KeyPressed; }
  void on_key_released(int32_t delay) { for (auto& h : harmonics_) h.env.on_key_released(delay); }
  bool try_acquire() { if (state_ != EnvState::Done2) return false; state_ = EnvState::SoonKeyPressed; return true; }
  bool rt_active() const { return state_ != EnvState::SoonKeyPressed && state_ != EnvState::Done2; }

  void step() {  // VolumeAdjusted::step :1195-1216 around MultiEnveloped::step :529-545
    imag_ = 0;
    bool go_on = false;
    for (size_t i = 0; i < harmonics_.size(); ++i) {
      auto& h = harmonics_[i];
      h.step();
      if (!h.env.finished()) go_on = true;
      imag_ += h.imag() * props_[i].volume;
    }
    if (!go_on) state_ = EnvState::Done2;
    T inc = std::min(max_filter_increment, harmonics_.empty() ? T(0) : std::abs(harmonics_[0].osc.angle_increments()));
    if (!volume_) {
      filter_inc_ = inc;
      filter_.init_with_angle_increment(inc);
      filter_.set_initial_value(*target_);
    } else if (*volume_ != *target_ && filter_inc_ != inc) {
      filter_inc_ = inc;
      filter_.init_with_angle_increment(inc);
    }
    volume_ = filter_.feed(*target_);
  }
  T imag() const { return *volume_ * imag_; }  // :1126-1129

  T max_filter_increment = std::numeric_limits<T>::max();

private:
  std::vector<EnvelopedOsc<T>> harmonics_;
  std::vector<HarmonicProps> props_;
  EnvState state_ = EnvState::Done2;
  T imag_ = 0;
  std::optional<T> volume_, target_;
  T filter_inc_ = 0;
  OnePole<T> filter_;
};

// stereo pan law (include/stereo.h:18-27)
inline void stereo_gains(double pan, double& left, double& right) {
  pan = std::max(-1., std::min(1., pan));
  double angle = 0.25 * kPi * (pan + 1.);
  left = std::cos(angle);
  right = std::sin(angle);
}

// ------------------------------------------------------------------------------------------------------------
// Frequency glide (InterpolatedFreq, include/audioelement.h:2706-2817) + FreqCtrl_ (:3056-3155)
// and the rt.resynth element StereoPanned<VolumeAdjusted<Enveloped<FreqCtrl_<Sine,InterpolatedFreq>,AHDSR>>>
// (source/rt.resynth.lib.cpp:5-18)
// ------------------------------------------------------------------------------------------------------------

template <class T>
class InterpolatedFreq {
public:
  void forget_past_signals() { f_result_.reset(); }
  void setup(T duration, int itp) {  // :2739-2750
    duration_ = duration; itp_ = itp; cur_ = 0;
    if (f_result_) { from_ = *f_result_; C_ = prop_const(); }
  }
  void set_angle_increments(T inc) {  // :2752-2762
    cur_ = 0; to_ = inc;
    from_ = f_result_ ? *f_result_ : inc;
    C_ = prop_const();
  }
  T step() {  // :2764-2788
    if (f_result_ && *f_result_ == to_) return to_;
    if (cur_ + T(.5f) > duration_) cur_ = duration_;
    f_result_ = interpolate<T>(itp_, cur_, from_, to_ - from_, duration_);
    if (cur_ < duration_) {
      T f = from_ + (to_ - from_) * (cur_ + T(.5f)) / duration_;
      cur_ += C_ * f;
    }
    return *f_result_;
  }
  T from() const { return from_; }

private:
  T prop_const() const { return (to_ == from_) ? T(1) : -std::log(from_ / to_) / (to_ - from_); }  // :2804-2816
  T from_ = 0, to_ = 0, cur_ = 0, duration_ = 0, C_ = 1;
  int itp_ = LINEAR;
  std::optional<T> f_result_;
};

template <class T>
class ResynthVoice {
public:
  // ResynthElementInitializer (source/rt.resynth.lib.cpp:96-118)
  void initialize(const Ahdsr& a, int sample_rate, int stride, double pan) {
    env_.set_ahdsr(a, sample_rate);
    max_filter_increment_ = T(2) / static_cast<T>(stride);
    ctrl_.setup(static_cast<T>(stride), LINEAR);
    stereo_gains(pan, gl_, gr_);
  }
  void reinitialize(const Ahdsr& a, int sample_rate, int stride) {  // InitializationType::ExistingNote
    env_.set_ahdsr(a, sample_rate);
    max_filter_increment_ = T(2) / static_cast<T>(stride);
    ctrl_.setup(static_cast<T>(stride), LINEAR);
  }
  void forget_past_signals() { env_.forget_past_signals(); ctrl_.forget_past_signals(); volume_.reset(); target_.reset(); }
  void set_volume_target(T v) { target_ = v; }
  void set_angle_increments(T a) {  // Enveloped :218-225 -> FreqCtrl_ :3103-3108
    T period = angle_increment_to_period_in_continuous_samples(a);
    env_.set_min_change_duration_samples(static_cast<int>(0.5f + T(2.5) * period));
    ctrl_.set_angle_increments(a);
  }
  void set_start_angle(T a) { osc_.set_start_angle(a); }
  void synchronize_angles(const ResynthVoice& o) { osc_.synchronize_angles(o.osc_); }  // :209-211 -> :3080-3082 -> :2397-2399
  void on_key_pressed(int32_t d) { env_.on_key_pressed(d); }
  void on_key_released(int32_t d) { env_.on_key_released(d); }
  bool try_acquire() { return env_.try_acquire(); }
  bool rt_active() const { return env_.rt_active(); }
  // FreqCtrl_::angleIncrements :3117-3124. The oscillator's own increment wins when it is non-zero, and nothing resets it
  // when a channel is reused (SineOscillatorAlgo::forgetPastSignals is empty, :2387-2388): on a reused channel this is the
  // last increment of the channel's PREVIOUS note until the first step of the new one.
  T angle_increments() const {
    T ai = osc_.angle_increments();
    return ai != T(0) ? ai : ctrl_.from();
  }
  void step() {  // StereoPanned -> VolumeAdjusted::step -> Enveloped::step -> FreqCtrl_::step
    env_.step();
    osc_.set_angle_increments(ctrl_.step());
    osc_.step();
    T inc = std::min(max_filter_increment_, std::abs(angle_increments()));
    if (!volume_) {
      filter_inc_ = inc;
      filter_.init_with_angle_increment(inc);
      filter_.set_initial_value(*target_);
    } else if (*volume_ != *target_ && filter_inc_ != inc) {
      filter_inc_ = inc;
      filter_.init_with_angle_increment(inc);
    }
    volume_ = filter_.feed(*target_);
  }
  T imag(int ch) const { return (ch == 0 ? gl_ : gr_) * (*volume_ * (osc_.imag() * env_.value())); }
  T imag_mono() const { return *volume_ * (osc_.imag() * env_.value()); }

private:
  Envelope<T> env_;
  SineOsc<T> osc_;
  InterpolatedFreq<T> ctrl_;
  std::optional<T> volume_, target_;
  T filter_inc_ = 0, max_filter_increment_ = std::numeric_limits<T>::max();
  OnePole<T> filter_;
  double gl_ = 1, gr_ = 1;
};

// ------------------------------------------------------------------------------------------------------------
// Frequency ramp (LogRamp, include/audioelement.h:2464-2572) and the gen.voice ramp element
// VolumeAdjusted<Enveloped<FreqCtrl_<Sine, LogRamp>, AHDSR>> (include/soundengine.h:245-257) with a constant volume.
// The loudness adjustment of the full element (Mix<LoudnessAdjusted<Sine>>, include/loudness.h) is not restated.
// ------------------------------------------------------------------------------------------------------------

template <class T>
class LogRamp {
public:
  void setup(T from, T to, T duration, T start_sample, int itp) {  // :2491-2520 (start_sample >= 0 branch)
    cur_ = start_sample;
    from_ = from; to_ = to; duration_ = duration; itp_ = itp;
    C_ = (to_ == from_) ? T(1) : -std::log(from_ / to_) / (to_ - from_);  // :2560-2571
  }
  T step() {  // do_step, :2522-2543
    if (cur_ + T(.5f) > duration_) { cur_ = 0; std::swap(from_, to_); }
    T f_result = interpolate<T>(itp_, cur_, from_, to_ - from_, duration_);
    T f = from_ + (to_ - from_) * (cur_ + T(.5f)) / duration_;
    cur_ += C_ * f;
    return f_result;
  }
  T from() const { return from_; }

private:
  T from_ = 0, to_ = 0, cur_ = 0, duration_ = 1, C_ = 1;
  int itp_ = LINEAR;
};

template <class T>
class RampVoice {
public:
  void initialize(const Ahdsr& a, int sample_rate, T volume) { env_.set_ahdsr(a, sample_rate); volume_ = volume; }
  void forget_past_signals() { env_.forget_past_signals(); }
  void setup_ramp(T from, T to, T duration, T start_sample) { ctrl_.setup(from, to, duration, start_sample, LINEAR); }
  void set_angle_increments(T a) {  // Enveloped :218-225 (the control ignores it, :2489)
    T period = angle_increment_to_period_in_continuous_samples(a);
    env_.set_min_change_duration_samples(static_cast<int>(0.5f + T(2.5) * period));
  }
  void set_start_angle(T a) { osc_.set_start_angle(a); }
  void on_key_pressed(int32_t d) { env_.on_key_pressed(d); }
  void on_key_released(int32_t d) { env_.on_key_released(d); }
  bool try_acquire() { return env_.try_acquire(); }
  bool rt_active() const { return env_.rt_active(); }
  void step() {  // Enveloped::step -> FreqCtrl_::step (:3126-3129)
    env_.step();
    osc_.set_angle_increments(ctrl_.step());
    osc_.step();
  }
  T imag() const { return volume_ * (osc_.imag() * env_.value()); }

private:
  Envelope<T> env_;
  SineOsc<T> osc_;
  LogRamp<T> ctrl_;
  T volume_ = 1;
};

// ImplCRTP::compute (include/gen.crtp.h:350-378): voice-major accumulation of n_frames <= 16 frames into an interleaved,
// pre-zeroed double buffer (caller: SimpleAudioOutContext::step, include/simple_audio_context.h:69-112).
template <class Voice>
bool pool_compute_mono(std::vector<Voice>& voices, double* buffer, int n_frames, int n_out) {
  bool silence = true;
  for (auto& v : voices) {
    if (!v.rt_active()) continue;
    silence = false;
    for (int i = 0; i < n_frames; ++i) {
      v.step();
      if (!v.rt_active()) break;
      for (int j = 0; j < n_out; ++j) buffer[n_out * i + j] += v.imag();
    }
  }
  return !silence;
}

// nOuts == 1 with an element that is not StereoPanned: imag_mono() of the round-trip element
template <class Voice>
bool pool_compute_mono_unpanned(std::vector<Voice>& voices, double* buffer, int n_frames) {
  bool silence = true;
  for (auto& v : voices) {
    if (!v.rt_active()) continue;
    silence = false;
    for (int i = 0; i < n_frames; ++i) {
      v.step();
      if (!v.rt_active()) break;
      buffer[i] += v.imag_mono();
    }
  }
  return !silence;
}

template <class Voice>
bool pool_compute_stereo(std::vector<Voice>& voices, double* buffer, int n_frames) {
  bool silence = true;
  for (auto& v : voices) {
    if (!v.rt_active()) continue;
    silence = false;
    for (int i = 0; i < n_frames; ++i) {
      v.step();
      if (!v.rt_active()) break;
      buffer[2 * i + 0] += v.imag(0);
      buffer[2 * i + 1] += v.imag(1);
    }
  }
  return !silence;
}

// ------------------------------------------------------------------------------------------------------------
// Pitch pipeline (source/rt.resynth.lib.algo.cpp:10-440) and the analysis -> resynthesis round trip of one stream as
// the offline constructor runs it (source/rt.resynth.lib.cpp:1185-1235, :1572-1759, :264-382): one sample at a time,
// a frame fires when the window is full, its events are applied before the output sample of the same index.
// ------------------------------------------------------------------------------------------------------------

inline double frequency_to_midi_pitch(double freq) { return 69. + 12. * std::log2(freq / 440.f); }           // include/midi.h:226-231
inline double midi_pitch_to_freq(double pitch) { return 440.f * std::pow(std::pow(2., 1. / 12.), pitch - 69.); }  // :258-264

struct PitchVolume {
  double midipitch;
  double volume;
};

// Equal-loudness contours (include/loudness.h): the ISO 226:2003 tables (:9-149), compute_equal_loudness_volume (:188-197, float
// arithmetic; the reference evaluates it at compile time with sprout::pow / sprout::log), the table over midi pitch (:43-54),
// closest() (:155-186) and equal_loudness_volume_db (:232-244).
struct EqualLoudness {
  static constexpr int n_freq = 29;
  std::array<double, n_freq> pitches;
  std::array<float, n_freq> freqs_hz;
  std::array<float, n_freq> elv;  // contour of one loudness level, dB
  explicit EqualLoudness(float phons) {
    static const float freqs[n_freq] = {20.f,  25.f,  31.5f, 40.f,   50.f,   63.f,   80.f,   100.f,  125.f,  160.f,
                                        200.f, 250.f, 315.f, 400.f,  500.f,  630.f,  800.f,  1000.f, 1250.f, 1600.f,
                                        2000.f, 2500.f, 3150.f, 4000.f, 5000.f, 6300.f, 8000.f, 10000.f, 12500.f};
    static const float alpha_f_[n_freq] = {.532f, .506f, .480f, .455f, .432f, .409f, .387f, .367f, .349f, .330f, .315f, .301f, .288f, .276f, .267f,
                                           .259f, .253f, .250f, .246f, .244f, .243f, .243f, .243f, .242f, .242f, .245f, .254f, .271f, .301f};
    static const float Lu_[n_freq] = {-31.6f, -27.2f, -23.0f, -19.1f, -15.9f, -13.0f, -10.3f, -8.1f, -6.2f, -4.5f, -3.1f, -2.0f, -1.1f, -0.4f, 0.f,
                                      .3f,    .5f,    0.f,    -2.7f,  -4.1f,  -1.0f,  1.7f,   2.5f,  1.2f,  -2.1f, -7.1f, -11.2f, -10.7f, -3.1f};
    static const float Tf_[n_freq] = {78.5f, 68.7f, 59.5f, 51.1f, 44.0f, 37.5f, 31.5f, 26.5f, 22.1f, 17.9f, 14.4f, 11.4f, 8.6f, 6.2f, 4.4f,
                                      3.0f,  2.2f,  2.4f,  3.5f,  1.7f,  -1.3f, -4.2f, -6.0f, -5.4f, -1.5f, 6.0f,  12.6f, 13.9f, 12.3f};
    // phons_to_index (:220-229) picks one of the contours 20, 30, .. 100 phons
    int level = static_cast<int>(phons * .1f) - 2;
    level = std::max(0, std::min(8, level));
    const float LN = (level + 2) * 10.f;
    for (int i = 0; i < n_freq; ++i) {
      pitches[i] = frequency_to_midi_pitch(freqs[i]);
      freqs_hz[i] = freqs[i];
      const float Af = 4.47e-3f * (std::pow(10.f, .025f * LN) - 1.14f) + std::pow(.4f * std::pow(10.f, ((Tf_[i] + Lu_[i]) * .1f) - 9.f), alpha_f_[i]);
      elv[i] = 94.f - Lu_[i] + (10.f / alpha_f_[i]) * std::log(Af) / std::log(10.f);
    }
  }
  float volume_db(double pitch) const {  // equal_loudness_volume_db(pitches, pitch, level)
    int min_ = 0, max_ = n_freq - 1;
    do {
      const int center = (min_ + max_) / 2;
      if (pitch < pitches[center]) max_ = center;
      else min_ = center;
    } while (max_ - min_ >= 2);
    if (pitch <= pitches[min_]) return elv[min_];
    if (pitch >= pitches[max_]) return elv[max_];
    const float ratio = static_cast<float>((pitch - pitches[min_]) / (pitches[max_] - pitches[min_]));
    return ratio * elv[max_] + (1.f - ratio) * elv[max_ - 1];
  }
  // equal_loudness_volume_from_freq (:258-275): the amplitude that equalises the loudness of a sine of `freq` with the one at
  // freqs[index_freq_ref]; log_ratio in [0,1] scales the compensation
  float volume_from_freq(float freq, int index_freq_ref, float log_ratio) const {
    int min_ = 0, max_ = n_freq - 1;
    do {
      const int center = (min_ + max_) / 2;
      if (freq < freqs_hz[center]) max_ = center;
      else min_ = center;
    } while (max_ - min_ >= 2);
    float db;
    if (freq <= freqs_hz[min_]) db = elv[min_];
    else if (freq >= freqs_hz[max_]) db = elv[max_];
    else {
      const float ratio = (freq - freqs_hz[min_]) / (freqs_hz[max_] - freqs_hz[min_]);
      db = ratio * elv[max_] + (1.f - ratio) * elv[max_ - 1];
    }
    const float max_db = elv[index_freq_ref];
    if (db > max_db) return 1.f;                                   // db_to_amplitude, :246-255
    return std::pow(10.f, log_ratio * (db - max_db) / 20.f);
  }
};

// Optional stages of RtResynth::step between reduce_pitches and track_pitches (source/rt.resynth.lib.cpp:1687-1707).
// All zero = what rt.resynth does out of the box (no shift, no harmonisation, use_autotune == false).
// Same layout as ca_pitch_options (include/cpp_audio_b200.h) so that a test hands one block to both sides.
struct PitchOptions {
  float shift_pre_autotune = 0;       // pitch_shift_pre_autotune
  float harmonize_pre_autotune = 0;   // pitch_harmonize_pre_autotune
  float shift_post_autotune = 0;
  float harmonize_post_autotune = 0;
  int32_t autotune = 0;               // 0: identity (:1763-1765); 1: closest of `table` = allowed_pitches (Chord and
                                      // FixedSizeIntervals, :1767-1858); 2: musical scale, `table` = scale pitches 0 .. 12 (:1859-1868)
  float autotune_max_pitch = 0;       // pitches above are played but not tuned
  float autotune_tolerance_pitches = 0;
  int32_t n_table = 0;
  const double* table = nullptr;
  double scale_root_pitch = 0;        // A_pitch + transpose + half_tones_distance(La, root note), :1862-1865
};

constexpr float kPitchEpsilon = 0.0001f;  // source/rt.resynth.lib.algo.cpp:3

// find_closest_pitch, source/rt.resynth.lib.autotune.cpp:189-218 (precondition: ordered; harmonize_pitches calls it on a
// vector whose tail is not, so the bisection of std::lower_bound is part of the behaviour)
template <class V, class F>
int find_closest_pitch(double pitch, const V& pitches, int n, F&& get) {
  if (n == 0) return -1;
  int lb = static_cast<int>(std::lower_bound(pitches, pitches + n, pitch, [&](const auto& a, double c) { return get(a) < c; }) - pitches);
  if (lb == n) return lb - 1;
  if (lb == 0) return 0;
  const float dist_lb = static_cast<float>(std::abs(pitch - get(pitches[lb])));
  const float dist_prev = static_cast<float>(std::abs(pitch - get(pitches[lb - 1])));
  return dist_lb < dist_prev ? lb : lb - 1;
}

inline void shift_pitches(float amount, std::vector<PitchVolume>& v) {  // algo.cpp:307-316
  if (!amount) return;
  for (auto& q : v) q.midipitch += amount;
}

inline void harmonize_pitches(float amount, std::vector<PitchVolume>& work, std::vector<PitchVolume>& v) {  // algo.cpp:318-371
  if (!amount) return;
  const int sz = static_cast<int>(v.size());
  for (int i = 0; i < sz; ++i) {
    const double harmonized = v[i].midipitch + amount;
    const int c = find_closest_pitch(harmonized, v.data(), static_cast<int>(v.size()), [](const PitchVolume& q) { return q.midipitch; });
    const float dist = static_cast<float>(std::abs(v[c].midipitch - harmonized));
    if (dist <= kPitchEpsilon) v[c].volume += v[i].volume;
    else v.push_back({harmonized, v[i].volume});
  }
  // merge_sort_step (cpp.algorithms): the two sorted runs [0, sz) and [sz, end) merged by pitch; ties keep the first run first
  work.resize(v.size());
  std::merge(v.begin(), v.begin() + sz, v.begin() + sz, v.end(), work.begin(),
             [](const PitchVolume& a, const PitchVolume& b) { return a.midipitch < b.midipitch; });
  work.swap(v);
}

// MusicalScalePitches::closest_pitch / distance_to, source/rt.resynth.lib.autotune.cpp:96-146
inline double scale_closest_pitch(const double* scale, int n_scale, double root_pitch, double pitch) {
  const double octave_dist = (pitch - root_pitch) / 12;
  const int octaves = octave_dist >= 0. ? static_cast<int>(octave_dist) : static_cast<int>(octave_dist) - 1;
  const double rel = (pitch - octaves * 12) - root_pitch;
  int i = 0;
  for (; i < n_scale - 1; ++i)
    if (rel < 0.5 * (scale[i] + scale[i + 1])) break;
  return pitch - (rel - scale[i]);
}

// autotune_pitches, algo.cpp:186-229, with the transforms of mkAutotuneFunction (rt.resynth.lib.cpp:1762-1869)
inline void autotune_pitches(const PitchOptions& o, const std::vector<PitchVolume>& in, std::vector<PitchVolume>& out) {
  out.clear();
  for (const auto& pv : in) {
    double t = pv.midipitch;
    bool has = true;
    if (o.autotune != 0) {
      if (pv.midipitch <= o.autotune_max_pitch) {
        if (o.autotune == 1) {
          const int c = find_closest_pitch(pv.midipitch, o.table, o.n_table, [](double q) { return q; });
          has = c >= 0;
          if (has) t = o.table[c];
        } else {
          t = scale_closest_pitch(o.table, o.n_table, o.scale_root_pitch, pv.midipitch);
        }
        if (has && t > o.autotune_max_pitch) t = pv.midipitch;  // keeps the output sorted
      }
      if (!has || !(std::abs(t - pv.midipitch) <= o.autotune_tolerance_pitches)) continue;
    }
    if (!out.empty() && std::abs(out.back().midipitch - t) < kPitchEpsilon) out.back().volume += pv.volume;
    else out.push_back({t, pv.volume});
  }
}

// SPEC S10: the reference draws start phases and pans from mersenne<> (external); the oracle and the product use this hash.
inline float hash_unit(uint64_t id, uint64_t salt) {
  uint64_t z = id * 0x9E3779B97F4A7C15ull + salt * 0xBF58476D1CE4E5B9ull + 0x94D049BB133111EBull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return static_cast<float>(static_cast<double>(z >> 11) * (2.0 / 9007199254740992.0) - 1.0);
}

// Test hooks: SynchronizePhase::Yes is what rt.resynth instantiates; a test switches it off to show that a case depends on it,
// and reads how often the rule fired.
inline std::atomic<bool>& sync_phase() { static std::atomic<bool> v{true}; return v; }
inline std::atomic<int64_t>& sync_count() { static std::atomic<int64_t> v{0}; return v; }

struct RoundTripParams {
  int sample_rate = 48000;
  int window = 4096;
  int hop = 256;
  double sigma_factor = 4.0;
  double min_db = -80.0;                // SPEC S6 threshold on the peak list
  double nearby_distance_tones = 0.4;   // rt.resynth.lib.cpp:948
  double min_volume = 0.0001;           // :947
  double max_track_pitches = 1.0;       // :949
  float gain = 1.f;                     // analysis_volume (:995 defaults to 0, i.e. muted; 1 = unity)
  float stereo_spread = 1.f;            // :955
  Ahdsr env{0, LINEAR, 0, 0, LINEAR, 0, LINEAR, 1.f};  // :957-961
  int max_channels = 254;               // 2 * 127 voices, :201-210, gen.crtp.h:225
  PitchOptions opts;                    // shift / harmonize / autotune (:1687-1707), off by default
};

// ------------------------------------------------------------------------------------------------------------
// SoundEngine (include/soundengine.h:240-468, :738-803): a voice of gen.voice. Up to three ramp elements
//   VolumeAdjusted<Enveloped<FreqCtrl_<Mix<LoudnessVolumeAdjusted<Sine>>, LogRamp>, AHDSR>>   (:245-257, MixOf :209-216)
// play a list of ramp specs one after the other; a new ramp fades in (attack = xfade_len) while the previous one fades out
// (release = xfade_len). The specs come from a Markov chain at note-on (:565-736, RNG and chain are external): here they
// are an input.
// ------------------------------------------------------------------------------------------------------------
struct RampSpec {  // same layout as ca_ramp_spec
  float from_hz, to_hz;   // LogRamp::setup(freq_to_angle_increment(from), freq_to_angle_increment(to), ...)
  float duration;         // samples
  float start_sample;
  int32_t itp;            // interpolation of the emitted frequency
  float volume;           // RampSpecs::Ctrl::setVolume
  int32_t silence_follows;
};

struct SoundEngineParams {
  int sample_rate = 48000;
  int xfade_len = 480;            // setEnvelopeCharacTime
  int articulative_pause = 0;     // articulative_pause_length
  int loudness_on = 1;            // the Mix element is LoudnessVolumeAdjusted<Sine> (0: plain sine, for tests of the ramps alone)
  int loudness_ref_index = 0;     // setLoudnessParams(sample_rate, low_index, log_ratio, loudness_level), gen.voice.h:934-937
  float loudness_compensation = 1.f;
  float loudness_level = 40.f;
  float sine_gain = 1.f;          // Mix gain of the sine (SetGains, gen.voice.h:951-956)
};

template <class T>
class RampElement {
public:
  void set_sample_rate(int sr) { sample_rate_ = sr; }
  void set_loudness(const EqualLoudness* contour, int ref_index, float log_ratio) { contour_ = contour; ref_index_ = ref_index; log_ratio_ = log_ratio; }
  void set_gain(float g) { gain_ = g; }
  void set_stereo(double gl, double gr) { gl_ = gl; gr_ = gr; }
  LogRamp<T>& ctrl() { return ctrl_; }
  Envelope<T>& env() { return env_; }
  const Envelope<T>& env() const { return env_; }
  void forget_past_signals() { env_.forget_past_signals(); lvol_.reset(); ltarget_.reset(); }  // Enveloped :192-195, BaseVolumeAdjusted :1168-1172
  void set_start_angle(T a) { osc_.set_start_angle(a); }
  void set_volume_target(T v) { volume_ = v; }  // the outer VolumeAdjusted: constant target, so its filter is the identity
  void step() {
    env_.step();                                  // Enveloped::step :197-200
    const T f = ctrl_.step();                     // FreqCtrl_::step :3126-3129 -> Mix::setAngleIncrements :1843-1847
    if (contour_) {                               // LoudnessVolumeAdjusted::setAngleIncrements :1291-1298
      const float freq = static_cast<float>(f * sample_rate_ / 2);  // angle_increment_to_freq
      ltarget_ = static_cast<T>(contour_->volume_from_freq(freq, ref_index_, log_ratio_));
    }
    osc_.set_angle_increments(f);
    osc_.step();                                  // BaseVolumeAdjusted::step :1195-1216 (max_filter_increment at its default)
    if (contour_) {
      const T inc = std::abs(f);
      if (!lvol_) {
        lfilter_inc_ = inc;
        lfilter_.init_with_angle_increment(inc);
        lfilter_.set_initial_value(*ltarget_);
      } else if (*lvol_ != *ltarget_ && lfilter_inc_ != inc) {
        lfilter_inc_ = inc;
        lfilter_.init_with_angle_increment(inc);
      }
      lvol_ = lfilter_.feed(*ltarget_);
    }
  }
  T imag_mono() const {  // VolumeAdjusted( Enveloped( Mix( Loudness( Sine ) ) ) )
    const T tone = contour_ ? *lvol_ * osc_.imag() : osc_.imag();
    return volume_ * ((static_cast<T>(gain_) * tone) * env_.value());
  }
  T imag(int ch) const { return (ch == 0 ? gl_ : gr_) * imag_mono(); }  // StereoPanned :1376-1380

private:
  Envelope<T> env_;
  SineOsc<T> osc_;
  LogRamp<T> ctrl_;
  const EqualLoudness* contour_ = nullptr;
  int ref_index_ = 0, sample_rate_ = 48000;
  float log_ratio_ = 1.f, gain_ = 1.f;
  std::optional<T> lvol_, ltarget_;
  T lfilter_inc_ = 0, volume_ = 1;
  OnePole<T> lfilter_;
  double gl_ = 1, gr_ = 1;
};

template <class T>
class SoundEngineVoice {
public:
  SoundEngineVoice(const SoundEngineParams& p, const RampSpec* specs, int n_specs, bool stereo, double pan)
      : prm_(p), specs_(specs, specs + n_specs), contour_(p.loudness_level), stereo_(stereo) {
    double gl = 1, gr = 1;
    if (stereo) stereo_gains(pan, gl, gr);
    for (auto& r : ramps_) {
      r.set_sample_rate(p.sample_rate);
      if (p.loudness_on) r.set_loudness(&contour_, p.loudness_ref_index, p.loudness_compensation);
      r.set_gain(p.sine_gain);
      r.set_stereo(gl, gr);
    }
  }
  void on_key_pressed() { go_on_ = true; }       // tryAcquire + onKeyPressed :297-312, :349-353
  void on_key_released(int32_t delay) {           // :357-365
    if (!go_on_) return;
    go_on_ = false;
    for (auto& r : ramps_) r.env().on_key_released(delay);
  }
  void step() {  // :738-741
    orchestrate();
    for (auto& r : ramps_)  // step_algos :412-426
      if (r.env().state() != EnvState::Done2) r.step();
  }
  T imag(int ch) const {  // :428-464
    T v = 0;
    for (auto& r : ramps_)
      if (r.env().state() == EnvState::KeyReleased || r.env().state() == EnvState::KeyPressed) v += stereo_ ? r.imag(ch) : r.imag_mono();
    return v;
  }

private:
  void orchestrate() {  // :744-772
    if (it_ == static_cast<int>(specs_.size())) return;  // ramp_specs.done()
    const RampSpec* cur = it_ >= 0 && it_ < static_cast<int>(specs_.size()) ? &specs_[it_] : nullptr;
    RampElement<T>* pressed = nullptr;
    RampElement<T>* done = nullptr;
    for (auto& r : ramps_) {  // get_ramps :277-295: the last ramp of each kind wins
      if (r.env().state() == EnvState::Done2) {
        if (go_on_) done = &r;
      } else if (r.env().state() == EnvState::KeyPressed) {
        pressed = &r;
      }
    }
    if (pressed) {
      const int remaining = pressed->env().count_steps_to_release();
      if (remaining > 0) return;
      if (remaining == 0 && cur && cur->silence_follows) remaining_silence_steps_ = prm_.articulative_pause;
    }
    if (cur && cur->silence_follows && remaining_silence_steps_ > 0) {
      --remaining_silence_steps_;
      return;
    }
    // playNextSpec :774-803
    if (!done) return;
    ++it_;
    if (it_ == static_cast<int>(specs_.size())) return;
    const RampSpec& sp = specs_[it_];
    const T from = static_cast<T>(freq_to_angle_increment<float>(sp.from_hz, prm_.sample_rate));
    const T to = static_cast<T>(freq_to_angle_increment<float>(sp.to_hz, prm_.sample_rate));
    done->ctrl().setup(from, to, static_cast<T>(sp.duration), static_cast<T>(sp.start_sample), sp.itp);
    done->forget_past_signals();
    done->set_start_angle(0);  // start_angle: forgetPastSignals zeroes it (:320-323), DefaultStartPhase::Zero (gen.voice.h:1100)
    done->set_volume_target(static_cast<T>(sp.volume));
    done->env().set_ahdsr(Ahdsr{prm_.xfade_len, LINEAR, 0, 0, LINEAR, prm_.xfade_len, LINEAR, 1.f}, prm_.sample_rate);
    done->env().try_acquire();
    const int time_to_release = static_cast<int>(.5f + sp.duration) - prm_.xfade_len;
    done->env().on_key_pressed(0);
    done->env().on_key_released(std::max(time_to_release, 0));
  }

  SoundEngineParams prm_;
  std::vector<RampSpec> specs_;
  EqualLoudness contour_;
  bool stereo_;
  std::array<RampElement<T>, 3> ramps_;
  int it_ = -1, remaining_silence_steps_ = 0;
  bool go_on_ = false;
};

// ------------------------------------------------------------------------------------------------------------
// The voice pool of sine::Synth<..., ResynthElement, ..., SynchronizePhase::Yes, DefaultStartPhase::Random, HandleNoteOff::Yes,
// 127, ResynthElementInitializer> (source/rt.resynth.lib.cpp:201-210): ImplCRTP::onEvent (include/gen.crtp.h:382-622) over
// 2 * 127 channels (:221-233) of StereoPanned<VolumeAdjusted<Enveloped<FreqCtrl_<Sine, InterpolatedFreq>, AHDSR>>>.
// ------------------------------------------------------------------------------------------------------------
template <class T>
class VoicePool {
public:
  VoicePool(const Ahdsr& env, int sample_rate, int max_channels, float stereo_spread)
      : env_(env), sample_rate_(sample_rate), stereo_spread_(stereo_spread), channels_(max_channels), note_of_(max_channels, -1) {}

  // NoteOn (:389-531) with ResynthElementInitializer (rt.resynth.lib.cpp:96-118); false: no free channel (DROPPED_NOTE, :412-414)
  bool note_on(int64_t id, float freq, float velocity, int stride) {
    for (size_t c = 0; c < channels_.size(); ++c) {
      if (!channels_[c].try_acquire()) continue;
      auto& e = channels_[c];
      note_of_[c] = id;
      e.forget_past_signals();
      e.set_volume_target(static_cast<T>(0.1f * velocity));
      e.initialize(env_, sample_rate_, stride, stereo_spread_ * hash_unit(static_cast<uint64_t>(id), 1));
      e.set_angle_increments(static_cast<T>(freq_to_angle_increment<float>(freq, sample_rate_)));
      // setPhase<SynchronizePhase::Yes, DefaultStartPhase::Random> (include/gen.crtp.h:127-156; Synth instantiated with these
      // at source/rt.resynth.lib.cpp:201-210): the first OTHER channel, in channel order, whose envelope is RT-active
      // (releasing channels included) and whose angleIncrements() equals this element's hands over its oscillator phase;
      // otherwise the start phase is random (SPEC S10: a hash of the note id).
      bool synced = false;
      if (sync_phase()) {
        const T ai = e.angle_increments();
        for (size_t o = 0; o < channels_.size(); ++o) {
          if (o == c || !channels_[o].rt_active()) continue;
          if (channels_[o].angle_increments() == ai) {
            e.synchronize_angles(channels_[o]);
            synced = true;
            ++sync_count();
            break;
          }
        }
      }
      if (!synced) e.set_start_angle(static_cast<T>(hash_unit(static_cast<uint64_t>(id), 2)));
      e.on_key_pressed(0);
      return true;
    }
    return false;
  }
  // NoteChange (:595-618): every channel that holds the note id
  void note_change(int64_t id, float velocity, float freq) {
    for (size_t c = 0; c < channels_.size(); ++c)
      if (note_of_[c] == id) {
        channels_[c].set_volume_target(static_cast<T>(0.1f * velocity));
        channels_[c].set_angle_increments(static_cast<T>(freq_to_angle_increment<float>(freq, sample_rate_)));
      }
  }
  // NoteOff (:532-593): the first channel that holds the note id, released if its key is pressed (:802-811)
  void note_off(int64_t id) {
    for (size_t c = 0; c < channels_.size(); ++c)
      if (note_of_[c] == id) {
        channels_[c].on_key_released(0);
        return;
      }
  }
  std::vector<ResynthVoice<T>>& channels() { return channels_; }

private:
  Ahdsr env_;
  int sample_rate_;
  float stereo_spread_;
  std::vector<ResynthVoice<T>> channels_;
  std::vector<int64_t> note_of_;
};

// A note schedule played on the pool: the events of sample n are applied, in order, before the pool steps to produce output
// sample n (what SimpleAudioOutContext::step does with its one-shots, include/simple_audio_context.h:69-112).
struct TrackEvent {  // same layout as ca_track_event (include/cpp_audio_b200.h)
  int64_t at_sample;
  int64_t noteid;
  int32_t type;  // EventType, include/events.h:12-17: 0 NoteOn, 1 NoteOff, 2 NoteChange
  float frequency;
  float velocity;
  int32_t reserved;
};

template <class T>
void render_tracks(const TrackEvent* ev, int64_t n_events, int64_t n_samples, int sample_rate, int glide_samples, const Ahdsr& env,
                   int max_channels, float stereo_spread, bool mono, double* out, int64_t* stats) {
  VoicePool<T> pool(env, sample_rate, max_channels, stereo_spread);
  int64_t started = 0, dropped = 0, e = 0;
  for (int64_t i = 0; i < n_samples; ++i) {
    for (; e < n_events && ev[e].at_sample <= i; ++e) {
      if (ev[e].type == 0) {
        ++started;
        if (!pool.note_on(ev[e].noteid, ev[e].frequency, ev[e].velocity, glide_samples)) ++dropped;
      } else if (ev[e].type == 1) {
        pool.note_off(ev[e].noteid);
      } else {
        pool.note_change(ev[e].noteid, ev[e].velocity, ev[e].frequency);
      }
    }
    double buf[2] = {0, 0};
    if (mono) {
      pool_compute_mono_unpanned(pool.channels(), buf, 1);
      out[i] = buf[0];
    } else {
      pool_compute_stereo(pool.channels(), buf, 1);
      out[2 * i] = buf[0];
      out[2 * i + 1] = buf[1];
    }
  }
  if (stats) {
    stats[0] = started;
    stats[1] = dropped;
  }
}

template <class T>
class RoundTrip {
public:
  explicit RoundTrip(const RoundTripParams& p)
      : prm_(p), pfft_(Window::Gaussian, p.sigma_factor), pool_(p.env, p.sample_rate, p.max_channels, p.stereo_spread) {
    pfft_.set_lambdas([this] { return prm_.window; }, [this] { return prm_.hop; },
                      [this](int stride, const std::vector<double>& sq) { on_frame(stride, sq); });
  }

  // out: interleaved stereo [n][2]
  void run(const float* x, int64_t n, double* out) {
    for (int64_t i = 0; i < n; ++i) {
      pfft_.feed(x[i]);
      double buf[2] = {0, 0};
      pool_compute_stereo(pool_.channels(), buf, 1);
      out[2 * i] = buf[0];
      out[2 * i + 1] = buf[1];
    }
  }
  int64_t dropped_notes() const { return dropped_; }
  int64_t notes_started() const { return next_noteid_; }

private:
  struct PlayedNote {
    int64_t noteid;
    double midi_pitch;
  };

  void on_frame(int stride, const std::vector<double>& sq) {
    extract_local_max_freqs_mags<double>(prm_.sample_rate, sq.data(), pfft_.op().fft_length(), prm_.min_db, freqmags_);
    // frequencies_to_pitches (:10-29)
    pv_.clear();
    for (auto& f : freqmags_)
      if (f.freq > 0) pv_.push_back({frequency_to_midi_pitch(f.freq), db_to_mag(f.mag_db)});
    // aggregate_pitches (:124-161) + reduce_pitches (:166-184, PonderateByVolume / SumVolumes, rt.resynth.lib.cpp:400-403)
    reduced_.clear();
    bool open = false;
    double mn = 0, mx = 0, sum_v = 0, sum_pv = 0;
    auto close = [&] {
      if (open && !(sum_v < prm_.min_volume)) reduced_.push_back({sum_pv / sum_v, sum_v});
      open = false;
    };
    for (auto& q : pv_) {
      if (open && std::max({mn, mx, q.midipitch}) - std::min({mn, mx, q.midipitch}) > prm_.nearby_distance_tones) close();
      if (!open) {
        open = true;
        mn = mx = q.midipitch;
        sum_v = 0;
        sum_pv = 0;
      } else {
        mn = std::min(mn, q.midipitch);
        mx = std::max(mx, q.midipitch);
      }
      sum_v += q.volume;
      sum_pv += q.midipitch * q.volume;
    }
    close();
    // shift / harmonize / autotune / shift / harmonize (rt.resynth.lib.cpp:1687-1707); with the default options only the
    // merge of pitches closer than pitch_epsilon inside autotune_pitches (identity transform) acts
    shift_pitches(prm_.opts.shift_pre_autotune, reduced_);
    harmonize_pitches(prm_.opts.harmonize_pre_autotune, pitches_tmp_, reduced_);
    autotune_pitches(prm_.opts, reduced_, tuned_);
    shift_pitches(prm_.opts.shift_post_autotune, tuned_);
    harmonize_pitches(prm_.opts.harmonize_post_autotune, pitches_tmp_, tuned_);
    // track_pitches (:256-305)
    std::vector<int> change(tuned_.size(), -1);
    std::vector<char> keep(played_.size(), 0);
    {
      size_t it = 0;
      for (size_t i = 0; i < tuned_.size(); ++i) {
        for (; it < played_.size(); ++it) {
          if (played_[it].midi_pitch < tuned_[i].midipitch - prm_.max_track_pitches) continue;
          if (played_[it].midi_pitch <= tuned_[i].midipitch + prm_.max_track_pitches) {
            change[i] = static_cast<int>(it);
            keep[it] = 1;
            ++it;
          }
          break;
        }
      }
    }
    // order_pitches_by_perceived_loudness (:377-411) with the key of rt.resynth.lib.cpp:1716-1724: volume over the 60-phon
    // equal-loudness contour at the pitch, kept as float. (The reference uses std::sort; ties, which need equal float keys, keep
    // the pitch order here.)
    static const EqualLoudness contour(60.f);
    std::vector<float> loud(tuned_.size());
    for (size_t i = 0; i < loud.size(); ++i) loud[i] = static_cast<float>(tuned_[i].volume / contour.volume_db(tuned_[i].midipitch));
    std::vector<int> order(tuned_.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = static_cast<int>(i);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return loud[a] > loud[b]; });
    // synthesize_sounds (rt.resynth.lib.cpp:264-382)
    for (size_t i = 0; i < played_.size(); ++i)
      if (!keep[i]) note_off(played_[i].noteid);
    size_t n_before = played_.size();
    for (int idx : order) {
      float new_freq = static_cast<float>(midi_pitch_to_freq(tuned_[idx].midipitch));
      float volume = static_cast<float>(prm_.gain * tuned_[idx].volume / 0.1f);
      if (change[idx] >= 0) {
        note_change(played_[change[idx]].noteid, volume, new_freq);
        played_[change[idx]].midi_pitch = tuned_[idx].midipitch;
      } else {
        if (volume <= 0) continue;
        ++next_noteid_;
        if (note_on(next_noteid_, new_freq, volume, stride)) played_.push_back({next_noteid_, tuned_[idx].midipitch});
        else ++dropped_;
      }
    }
    // remove_dead_notes (:415-432) + sort_by_current_pitch (:434-440)
    std::vector<PlayedNote> alive;
    for (size_t i = 0; i < played_.size(); ++i)
      if (i >= n_before || keep[i]) alive.push_back(played_[i]);
    std::stable_sort(alive.begin(), alive.end(), [](const PlayedNote& a, const PlayedNote& b) { return a.midi_pitch < b.midi_pitch; });
    played_.swap(alive);
  }

  bool note_on(int64_t id, float freq, float velocity, int stride) { return pool_.note_on(id, freq, velocity, stride); }
  void note_change(int64_t id, float velocity, float freq) { pool_.note_change(id, velocity, freq); }
  void note_off(int64_t id) { pool_.note_off(id); }

  RoundTripParams prm_;
  PeriodicFft<SqMagOp<double>> pfft_;
  VoicePool<T> pool_;
  std::vector<FreqMag<double>> freqmags_;
  std::vector<PitchVolume> pv_, reduced_, tuned_, pitches_tmp_;
  std::vector<PlayedNote> played_;
  int64_t next_noteid_ = 0, dropped_ = 0;
};


// ------------------------------------------------------------------------------------------------------------
// Vocoder (source/rt.resynth.lib.vocoder.cpp): FFTModulator (:100-257) measures band amplitudes of the modulator signal,
// FFTCarrier (:396-551) scales the carrier's spectrum band by band, transforms back and cross-fades consecutive frames.
// Unpinned (cpp.algorithms): RealFBins::modulate_bands -> SPEC S12, EqualGainXFade -> SPEC S13.
// ------------------------------------------------------------------------------------------------------------

struct VocoderSetup {  // SetupParams, :5-44
  float modulator_window_size_seconds = 0.04f;
  float stride_seconds = 0.005f;
  int count_bands = 20;
  float min_freq = 100.f;
  float max_freq = 5000.f;
};

inline void vocoder_fill_freqs(const VocoderSetup& p, std::vector<double>& freqs) {  // :15-26
  freqs.clear();
  auto log_min = std::log(p.min_freq);
  auto log_max = std::log(p.max_freq);
  for (int i = 0; i <= p.count_bands; ++i) {
    double const ratio = static_cast<double>(i) / p.count_bands;
    freqs.push_back(std::exp(log_min + ratio * (log_max - log_min)));
  }
}

inline int vocoder_good_stride(float seconds, int sample_rate) {  // :84-92
  int res = std::max(1, static_cast<int>(0.5f + seconds * sample_rate));
  if (0 == res % 2) ++res;
  return res;
}

inline int vocoder_modulator_window(float seconds, int sample_rate) {  // :110-118
  int sz = std::max(1, static_cast<int>(0.5f + sample_rate * seconds));
  if (1 == sz % 2) ++sz;
  return sz;
}

// band of a frequency: freqs[b] < Hz <= freqs[b+1] (the membership rule of the modulator's loop, :131-156), -1 outside
inline int vocoder_band_of(const std::vector<double>& freqs, double hz) {
  int nb = static_cast<int>(freqs.size()) - 1;
  if (!(hz > freqs[0])) return -1;
  for (int b = 0; b < nb; ++b)
    if (hz <= freqs[b + 1]) return b;
  return -1;
}

// SPEC S13: equal-gain sinusoidal cross-fade over `stride` steps: step p in [1, stride] gives the new signal the weight
// (1 - cos(pi p / stride)) / 2 and the old one the rest (weights sum to 1; p = stride: new signal only).
template <class T>
inline void vocoder_xfade(int p, int stride, T& new_mult, T& old_mult) {
  new_mult = T(0.5) * (T(1) - std::cos(static_cast<T>(kPi) * static_cast<T>(p) / static_cast<T>(stride)));
  old_mult = T(1) - new_mult;
}

template <class T>
class Vocoder {
public:
  Vocoder(int sample_rate, const VocoderSetup& setup, float vol_modulator, float vol_carrier, float vol_vocoder)
      : sr_(sample_rate), vm_(vol_modulator), vc_(vol_carrier), vv_(vol_vocoder), modulator_(Window::Gaussian, 4.0) {
    vocoder_fill_freqs(setup, freqs_);
    nb_ = static_cast<int>(freqs_.size()) - 1;
    amps_.assign(nb_, T(0));
    stride_ = vocoder_good_stride(setup.stride_seconds, sr_);
    int wm = vocoder_modulator_window(setup.modulator_window_size_seconds, sr_);
    // FFTModulator::init_dynamic_allocs, :108-165
    modulator_.set_lambdas([wm] { return wm; }, [this] { return stride_; }, [this](int, const std::vector<T>& sqmag) {
      std::fill(amps_.begin(), amps_.end(), T(0));
      int n_fft = modulator_.op().fft_length();
      double const bin_to_Hz = static_cast<double>(sr_) / n_fft;
      size_t lo = 0, hi = 1;
      int band = 0;
      for (int i = 0, sz = static_cast<int>(sqmag.size()); i < sz; ++i) {
        double const Hz = bin_to_Hz * i;
        if (Hz <= freqs_[lo]) continue;
        while (Hz > freqs_[hi]) {
          ++band;
          ++hi;
          ++lo;
          if (hi == freqs_.size()) break;
        }
        if (hi == freqs_.size()) break;
        amps_[band] += sqmag[i];
      }
      for (auto& a : amps_) a = std::sqrt(a);
    });
    // FFTCarrier::init_dynamic_allocs, :413-476
    carrier_.set_lambdas([this] { return 2 * stride_; }, [this] { return stride_; },
                         [this](int, const std::vector<std::complex<T>>& bins) {
      int n_fft = carrier_.op().fft_length();
      if (static_cast<int>(sig_new_.size()) != n_fft) {
        sig_new_.assign(n_fft, T(0));
        sig_old_.assign(n_fft, T(0));
      }
      sig_old_.swap(sig_new_);
      tmp_ = bins;
      for (int i = 0; i < static_cast<int>(tmp_.size()); ++i) {  // SPEC S12: modulate_bands
        int b = vocoder_band_of(freqs_, static_cast<double>(sr_) / n_fft * i);
        tmp_[i] *= b < 0 ? T(0) : amps_[b];
      }
      carrier_.op().fft().inverse(tmp_.data(), sig_new_.data());
      T const scale = T(1) / static_cast<T>(n_fft);  // 1 / (Algo::scale * fft_length): forward o inverse = identity
      for (auto& v : sig_new_) v *= scale;
      idx_ = 0;
    });
  }

  // one sample of each input -> one output sample (Vocoder::initialize's compute lambda, :777-812)
  T feed(T modulator_sample, T carrier_sample) {
    modulator_.feed(static_cast<float>(modulator_sample));
    carrier_.feed(static_cast<float>(carrier_sample));
    T vocoded = 0;
    if (!sig_new_.empty() && idx_ < stride_) {  // FFTCarrier::feed, :501-531
      T a, b;
      vocoder_xfade<T>(idx_ + 1, stride_, a, b);
      vocoded = sig_new_[idx_] * a + sig_old_[idx_ + stride_] * b;
    }
    ++idx_;
    return vm_ * modulator_sample + vc_ * carrier_sample + vv_ * vocoded;
  }

  int stride() const { return stride_; }
  const std::vector<T>& bands_amplitudes() const { return amps_; }

private:
  int sr_;
  float vm_, vc_, vv_;
  std::vector<double> freqs_;
  int nb_ = 0, stride_ = 1, idx_ = 0;
  std::vector<T> amps_;
  PeriodicFft<SqMagOp<T>> modulator_;
  PeriodicFft<FftOp<T>> carrier_;
  std::vector<std::complex<T>> tmp_;
  std::vector<T> sig_new_, sig_old_;
};

}  // namespace orc
