// tracks.cuh — closed-form rendering of gliding sine voices, shared by the round trip (roundtrip.cu) and the ramp voices
// (ramp.cu). A SEGMENT describes one voice over one tile of consecutive samples: every per-sample recurrence of the
// reference element (oscillator with a per-sample frequency control, include/audioelement.h:3056-3155; time-warped
// frequency interpolation :2706-2817 and :2464-2572; volume low-pass :1195-1216; AHDSR envelope :666-1099) has a closed
// form in the step index, so tiles are independent and one warp renders a tile with lanes = consecutive samples.
#pragma once

#include <math.h>
#include <stdint.h>

#include <climits>

#include "common.cuh"

namespace ca {

struct __align__(8) Segment {
  double phase_start;  // oscillator angle (units of pi rad, [0,2)) before the first step of the tile
  double from, to;     // glide end points (angle increments)
  double K, L;         // f_i = from + K*(exp(i*L) - 1) for glide steps i < m, f_i = to afterwards
  double c1;           // K / lambda:  P_i = sum_{l<=i} f_l = (i+1)*(from-K) + c1*(exp((i+1)*L) - 1)
  double P_before;     // P_{i0-1} (0 when i0 == 0)
  double P_m1;         // P_{m-1}
  int32_t i0;          // glide step index of the tile's first sample
  int32_t m;           // first clamped glide step; 0 = steady at `to`
  float vol_x, vol_dy; // volume = x + dy*exp(-pi*Mv)
  float cap;           // max filter increment (2/stride)
  int32_t regime;      // 0: Mv = cap*(j+1); 1: Mv = P-sum since tile start; 2: f-limited until istar then cap; 3: cap then f
  int32_t istar;       // tile-relative index where a mixed regime switches
  float m_star;        // Mv accumulated over [0, istar)
  int32_t c0;          // envelope step counter at the tile's first sample
  int32_t attack, hold, decay, release;
  float sustain;
  int32_t rel_base;    // tile-relative index of the first release step (may be negative); INT_MAX = not released
  float top;           // envelope value when the release started
  float gl, gr;
};

struct TileArgs {
  const Segment* segs;
  const int64_t* tile_seg_begin;  // [tiles+1] CSR into segs ([tiles] when tile_seg_end is given)
  const int64_t* tile_sample0;    // [tiles] first output sample (absolute index into out / 2)
  const int32_t* tile_len;        // [tiles]
  int64_t n_tiles;
  float* out;                     // interleaved stereo (mono for tracks_kernel<true>)
  int32_t accumulate;             // mono only: atomicAdd into out (tiles of different voices overlap) instead of storing
  const int64_t* tile_seg_end = nullptr;  // [tiles] one past the last segment of each tile; null: tile_seg_begin[tile + 1]
  // a launch over a tile range of every row of a [rows][row_tiles] tile table: flat index i -> row i / range_tiles, tile
  // range_first + i % range_tiles of that row (row_tiles == 0: the flat index is the tile)
  int64_t row_tiles = 0, range_first = 0, range_tiles = 0;
};

__device__ __forceinline__ float alias_gain(float ai) {  // freqAliasingMultiplicator, include/audioelement.h:466-483
  if (ai == 0.f) return 1.f;
  const float h = 1.f / ai;
  if (h < 1.f) return 0.f;
  if (h > 4.f) return 1.f;
  return (h - 1.f) * (1.f / 3.f);
}

constexpr int kMaxTile = 1024;  // samples per tile (hop); 32 steps of 32 lanes

// kMono: one float per sample (the left gain only) instead of an interleaved stereo pair
template <bool kMono>
__global__ void __launch_bounds__(128) tracks_kernel(const __grid_constant__ TileArgs a) {
  const int lane = threadIdx.x & 31;
  int64_t tile = static_cast<int64_t>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (tile >= a.n_tiles) return;
  if (a.row_tiles) tile = (tile / a.range_tiles) * a.row_tiles + a.range_first + tile % a.range_tiles;
  const int len = a.tile_len[tile];
  const int64_t s_begin = a.tile_seg_begin[tile], s_end = a.tile_seg_end ? a.tile_seg_end[tile] : a.tile_seg_begin[tile + 1];
  float2* out = reinterpret_cast<float2*>(a.out) + (kMono ? 0 : a.tile_sample0[tile]);
  float* out1 = a.out + (kMono ? a.tile_sample0[tile] : 0);
  for (int j0 = 0; j0 < len; j0 += 32) {
    const int j = j0 + lane;  // tile-relative sample index
    float accl = 0.f, accr = 0.f;
    for (int64_t si = s_begin; si < s_end; ++si) {
      const Segment& g = a.segs[si];
      // ---- frequency glide and phase (FP64) ----
      const int i = g.i0 + j;  // glide step index
      double f, P;             // emitted increment at this step, sum of increments over glide steps 0..i
      if (i < g.m) {
        // expm1, not exp()-1: when `to` is within 1e-10 of `from`, c1 = K/lambda is ~1e14 and the rounding of exp()-1
        // (1e-16) would reach the phase
        P = static_cast<double>(i + 1) * (g.from - g.K) + g.c1 * expm1(g.L * static_cast<double>(i + 1));
        f = g.from + g.K * expm1(g.L * static_cast<double>(i));
      } else {
        P = g.P_m1 + static_cast<double>(i - g.m + 1) * g.to;
        f = g.to;
      }
      const double adv = P - g.P_before;  // phase advance since the tile start, this step included
      double ph = g.phase_start + adv;
      ph -= 2.0 * floor(0.5 * ph);
      const float s = sinpif(static_cast<float>(ph));
      const float ff = static_cast<float>(f);
      // ---- volume low-pass (SPEC S8): y = x + dy * exp(-pi * sum of min(cap, f)) ----
      float mv;
      if (g.regime == 0) mv = g.cap * static_cast<float>(j + 1);
      else if (g.regime == 1) mv = static_cast<float>(adv);
      else if (g.regime == 2) {  // f-limited first, cap from istar on
        if (j < g.istar) mv = static_cast<float>(adv);
        else mv = g.m_star + g.cap * static_cast<float>(j - g.istar + 1);
      } else {  // cap first, f-limited from istar on: m_star holds cap*istar - (advance before istar)
        if (j < g.istar) mv = g.cap * static_cast<float>(j + 1);
        else mv = g.m_star + static_cast<float>(adv);
      }
      const float vol = g.vol_x + g.vol_dy * __expf(-3.14159265358979f * mv);
      // ---- envelope (include/audioelement.h:960-1000, :733-742) ----
      float env;
      if (j >= g.rel_base) {
        const int jr = j - g.rel_base + 1;
        env = jr >= g.release ? 0.f : g.top - g.top * (static_cast<float>(jr) / static_cast<float>(g.release));
      } else {
        int c = g.c0 + j;
        if (c < 1) env = 0.f;
        else if (c <= g.attack) env = static_cast<float>(c) / static_cast<float>(g.attack);
        else if ((c -= g.attack) <= g.hold) env = 1.f;
        else if ((c -= g.hold) <= g.decay) env = 1.f + (g.sustain - 1.f) * (static_cast<float>(c) / static_cast<float>(g.decay));
        else env = g.sustain;
      }
      const float y = vol * (alias_gain(ff) * s * env);
      accl = fmaf(g.gl, y, accl);
      accr = fmaf(g.gr, y, accr);
    }
    if (j < len) {
      if (kMono && a.accumulate) atomicAdd(out1 + j, accl);
      else if (kMono) out1[j] = accl;
      else out[j] = make_float2(accl, accr);
    }
  }
}

// InterpolatedFreq / LogRamp in closed form: the control's position obeys cur' = (1+lambda)*cur + beta (it advances by
// C*f with f linear in cur), so cur_i = beta*(alpha^i - 1)/lambda and the emitted value and its running sum are
// exponentials of the step index.
// 1/x to the last bit or two: hardware seed + two Newton steps on the device (~100 clocks instead of ~320)
__host__ __device__ inline double rcp(double x) {
#ifdef __CUDA_ARCH__
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
#else
  return 1.0 / x;
#endif
}

// LogRamp (include/audioelement.h:2464-2572) in closed form, one leg at a time (the ramp turns around when its position
// reaches the duration): the position advances by C*f_lin(position) per step whatever the interpolation of the EMITTED
// frequency is (:2537-2540), so it is an exponential of the step index, and so is the emitted frequency when that
// interpolation is LINEAR.
struct Leg {  // closed form of one ramp leg that starts at position c_start
  double from1 = 0, K1 = 0, L = 0, c1 = 0;  // f_i = from1 + K1*expm1(L*i),  P_i = (i+1)*(from1-K1) + c1*expm1(L*(i+1))
  double u0 = 0, Ku = 0;                    // ramp position over duration: u_i = u0 + Ku*expm1(L*i) (the argument of the easings)
  int64_t m = 0;                            // steps before the ramp turns around
  bool steady = false;                      // from == to
  __host__ __device__ void start(double from, double to, double dur, double c_start) {
    steady = (to == from);
    u0 = c_start / dur;
    Ku = 0;
    if (steady) {
      from1 = from;
      K1 = L = c1 = 0;
      m = LLONG_MAX / 4;
      return;
    }
    const double delta = to - from;
    const double C = -log(from / to) / delta;            // get_linear_proportionality_constant, :2560-2571
    const double lambda = C * delta / dur;
    const double beta = C * (from + 0.5 * delta / dur);  // do_step, :2537-2540
    L = log1p(lambda);
    K1 = (delta / dur) * (c_start + beta / lambda);
    from1 = from + delta * c_start / dur;
    c1 = K1 / lambda;
    Ku = (c_start + beta / lambda) / dur;
    // first step i with cur_i + .5 > dur (:2523): cur_i = (c_start + beta/lambda)*alpha^i - beta/lambda
    const double x = (log1p(lambda * (dur - 0.5) / beta) - log1p(lambda * c_start / beta)) / L;
    double mf = floor(x) + 1.0;
    if (!(mf >= 0.0)) mf = 0.0;
    if (mf > 4e18) mf = 4e18;
    m = static_cast<int64_t>(mf);
  }
  __host__ __device__ double P(int64_t k) const {  // sum of the emitted increments over steps 0..k
    if (k < 0) return 0.0;
    if (steady) return static_cast<double>(k + 1) * from1;
    return static_cast<double>(k + 1) * (from1 - K1) + c1 * expm1(L * static_cast<double>(k + 1));
  }
};

struct Glide {  // InterpolatedFreq in closed form
  double from = 0, to = 0, K = 0, L = 0, c1 = 0, P_m1 = 0;
  double f_m1 = 0;  // f_at(m - 1), the last glide step before the clamp
  int32_t m = 0;   // first clamped step
  int64_t i = 0;   // next glide step index
  __host__ __device__ void start(double from_, double to_, double dur) {
    from = from_;
    to = to_;
    i = 0;
    if (to == from) {  // C = 1: the first step emits `to`, after which the steady-state shortcut holds (:2765-2767)
      m = 0;
      K = L = c1 = P_m1 = 0;
      f_m1 = to;
      return;
    }
    // Reference constants (:2783-2784, :2815): C = -log(from/to)/delta, cur' = (1+lambda)*cur + beta with lambda = C*delta/dur,
    // beta = C*(from + delta/(2 dur)); the first clamped step is the first i with cur_i + .5 > dur, cur_i = beta*(alpha^i-1)/lambda.
    // Arranged for the device: a correctly rounded FP64 division costs as much as a log() there (~320 clocks, measured), so
    // quotients become products with reciprocals, and lambda*(dur-.5)/beta = (dur-.5)*delta/(dur*from + delta/2) is formed
    // without the logarithm so that the two logarithm chains are independent.
    const double delta = to - from;
    const double inv_dur = rcp(dur);
    const double lg = log1p(-delta * rcp(to));  // log(from/to); never 0 for from != to, however close
    const double t = (dur - 0.5) * delta * rcp(dur * from + 0.5 * delta);
    const double lt = log1p(t);
    const double C = -lg * rcp(delta);
    const double lambda = -lg * inv_dur;
    const double beta = C * (from + 0.5 * delta * inv_dur);
    const double inv_lambda = rcp(lambda);
    K = delta * beta * inv_dur * inv_lambda;
    L = log1p(lambda);
    c1 = K * inv_lambda;
    const double x = lt * rcp(L);
    double mf = floor(x) + 1.0;
    if (mf < 1.0) mf = 1.0;
    if (mf > 1e9) mf = 1e9;
    m = static_cast<int32_t>(mf);
    // one expm1 serves both f_at(m - 1) and P(m - 1): exp(L) = 1 + lambda, so expm1(L m) = e + lambda e + lambda with
    // e = expm1(L (m - 1))
    const double e = m > 1 ? expm1(L * static_cast<double>(m - 1)) : 0.0;
    f_m1 = from + K * e;
    P_m1 = static_cast<double>(m) * (from - K) + c1 * (e + lambda * e + lambda);  // P(m - 1)
  }
  __host__ __device__ double f_at(int64_t k) const {
    if (k >= m) return to;
    if (k == m - 1) return f_m1;
    return k == 0 ? from : from + K * expm1(L * static_cast<double>(k));  // expm1(0) == 0
  }
  __host__ __device__ double P(int64_t k) const {  // sum of f over steps 0..k
    if (k < 0) return 0.0;
    if (k == m - 1) return P_m1;  // the same expression, evaluated by start()
    if (k < m) return static_cast<double>(k + 1) * (from - K) + c1 * expm1(L * static_cast<double>(k + 1));
    return P_m1 + static_cast<double>(k - m + 1) * to;
  }
};


}  // namespace ca
