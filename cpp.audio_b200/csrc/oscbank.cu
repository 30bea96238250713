// oscbank.cu — K3: sine oscillator bank (harmonic voices) rendered for whole notes at once.
//
// Reference semantics (include/audioelement.h:486-657 MultiEnveloped, :2366-2452 SineOscillatorAlgo, :666-1099 AHDSR
// envelope, :1111-1265 VolumeAdjusted; include/gen.crtp.h:350-378 compute): sample n of a voice is
//     volume * sum_p vol_p * env_p(n) * g_p * sin(pi * (phase_p + p*a0 + (n+1) * p*ai))
// with one envelope per partial (attack/decay/release floors of 2.5 periods of THAT partial, :216-223).
//
// B200 design (not a translation of the per-sample phasor loop):
//   * lanes = 32 consecutive samples; a warp renders a chunk of R*32 samples of one voice.
//   * in the sustain regime a partial is advanced 32 samples at a time by the Reinsch form of the two-term
//     recurrence  d -= kappa*y ; y += d  (1 FFMA + 1 FADD, + 1 FADD to accumulate) instead of a 4-instruction complex
//     rotation; the 32-sample rotation angle is folded into (-pi/2, pi/2] (alternating sign for the other half) so
//     the recurrence stays as accurate as the rotation (measured: 2.5e-6 max abs after 64 steps, tests/test_oscbank.py).
//   * every chunk is re-seeded from an EXACT phase: phases are 64-bit fixed-point turns (phase0 + (n+1)*inc mod 2^64),
//     so fp32 never accumulates phase over the 480,000-sample note (the reference runs in double).
//   * chunks that contain an envelope transition (attack, decay, release) take a general path: complex rotation per
//     32 samples + the envelope evaluated per sample (piecewise linear, or eased: itp::interpolate with the ca_itp set).
//   * the kernel is instantiated twice, for voices with LINEAR phases only and for voices with an eased phase; a render
//     launches the pass(es) its envelopes need and each pass skips the voices of the other kind.
//   * partial parameters are prepared once per render by a small kernel, in double where the reference is in double.
#include <math.h>

#include "common.cuh"
#include "ctx.hpp"
#include "packed.cuh"

namespace ca {

namespace {

constexpr int kNP = 32;       // partials per pass (register resident)
constexpr int kR = 64;        // 32-sample steps per chunk
constexpr int kChunk = 32 * kR;
constexpr int kMaxPartials = 128;

struct __align__(16) PartialParams {
  uint64_t inc;   // phase increment per sample, turns * 2^64
  uint64_t ph0;   // phase before the first step
  float amp;      // volume * partial volume * aliasing gain (0 for padding / dropped harmonics)
  float kappa;    // 4 sin^2(Of/2), Of = 32-sample rotation folded into (-pi/2, pi/2]
  float so;       // sin(Of)
  float sgn;      // +1, or -1 when Of = O -+ pi (output alternates sign every step)
  float rot_re;   // exp(i*O), unfolded 32-sample rotation
  float rot_im;
  int32_t attack;   // effective durations in samples (floors applied)
  int32_t decay;
  int32_t release;
  float inv_attack;
  float inv_decay;
  float inv_release;
};
static_assert(sizeof(PartialParams) == 64, "PartialParams must be 64 bytes");

struct __align__(16) VoiceParams {
  int64_t release_at;  // <0: never
  int32_t key_delay;
  int32_t hold;
  float sustain;  // 1 + SMinusOne
  float gain_l;
  float gain_r;
  int32_t itp;  // ca_itp of attack | decay << 8 | release << 16
};
static_assert(sizeof(VoiceParams) == 32, "VoiceParams must be 32 bytes");

struct PrepArgs {
  ca_oscbank_desc d;
  PartialParams* partials;
  VoiceParams* voices;
  int32_t ppad;
};

__global__ void osc_prep_kernel(const __grid_constant__ PrepArgs a) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t v = idx / a.ppad;
  const int p = static_cast<int>(idx - v * a.ppad);
  if (v >= a.d.n_voices) return;
  const ca_ahdsr env = a.d.env[a.d.n_env == 1 ? 0 : v];
  const int sr = a.d.sample_rate;

  // AHDSREnvelopeBase::setAHDSR, include/audioelement.h:923-952
  const int32_t min_dt = sr / 1000;
  const bool has_decay = env.sustain < 0.999999;
  const float s_minus_one = has_decay ? fminf(fmaxf(env.sustain, 0.f), 1.f) - 1.f : 0.f;
  const int32_t A = max(env.attack, min_dt);
  const int32_t H = max(env.hold, 0);
  const int32_t D = has_decay ? max(env.decay, min_dt) : 0;
  const int32_t R = max(env.release, min_dt);

  if (p == 0) {
    VoiceParams vp;
    vp.release_at = a.d.release_at ? a.d.release_at[v] : -1;
    vp.key_delay = a.d.key_delay ? a.d.key_delay[v] : 0;
    vp.hold = H;
    vp.sustain = 1.f + s_minus_one;
    double gl = 1.0, gr = 1.0;
    if (a.d.mix == CA_MIX_STEREO) {  // include/stereo.h:18-27
      double pan = a.d.pan ? static_cast<double>(a.d.pan[v]) : 0.0;
      pan = fmin(fmax(pan, -1.0), 1.0);
      sincospi(0.25 * (pan + 1.0), &gr, &gl);
    }
    vp.gain_l = static_cast<float>(gl);
    vp.gain_r = static_cast<float>(gr);
    vp.itp = (env.attack_itp & 0xff) | ((env.decay_itp & 0xff) << 8) | ((env.release_itp & 0xff) << 16);
    a.voices[v] = vp;
  }

  PartialParams pp;
  // trailing inaudible harmonics are dropped (include/audioelement.h:401-415, :512)
  bool audible_tail = false;
  if (p < a.d.n_partials)
    for (int q = p; q < a.d.n_partials; ++q)
      if (fabsf(a.d.partial_volume[v * a.d.n_partials + q]) > 0.000001f) {
        audible_tail = true;
        break;
      }

  // freq_to_angle_increment is evaluated in float (include/sound.functions.h:58-62 with the float of Event::noteOn)
  const float ai_f = 2 * a.d.f0_hz[v] / sr;
  const double ai = static_cast<double>(p + 1) * static_cast<double>(ai_f);  // harmonic_angle, :456-459
  // freqAliasingMultiplicator, :466-483
  double g = 1.0;
  if (ai != 0.0) {
    const double h = 1.0 / ai;
    if (h < 1.0) g = 0.0;
    else if (h > 4.0) g = 1.0;
    else g = (h - 1.0) * (1.0 / 3.0);
  }
  // Enveloped::setAngleIncrements, :218-225
  const double period = (ai == 0.0) ? 1.0 : 2.0 / fabs(ai);
  const int32_t min_change = static_cast<int32_t>(0.5f + 2.5 * period);
  pp.attack = max(min_change, A);                 // :1066-1084
  pp.decay = (D == 0) ? 0 : max(min_change, D);
  pp.release = max(min_change, R);                // :1002-1004
  pp.inv_attack = 1.f / static_cast<float>(pp.attack);
  pp.inv_decay = pp.decay ? 1.f / static_cast<float>(pp.decay) : 0.f;
  pp.inv_release = 1.f / static_cast<float>(pp.release);

  const double vol = audible_tail ? static_cast<double>(a.d.volume[v]) *
                                        static_cast<double>(a.d.partial_volume[v * a.d.n_partials + p]) * g
                                  : 0.0;
  pp.amp = static_cast<float>(vol);

  const double two64 = 18446744073709551616.0;
  double t = ai * 0.5;
  t -= floor(t);
  pp.inc = __double2ull_rn(t * two64);
  const double a0 = (p < a.d.n_partials ? static_cast<double>(a.d.partial_phase[v * a.d.n_partials + p]) : 0.0) +
                    static_cast<double>(p + 1) * static_cast<double>(a.d.start_angle ? a.d.start_angle[v] : 0.f);
  double t0 = a0 * 0.5;
  t0 -= floor(t0);
  pp.ph0 = __double2ull_rn(t0 * two64);

  const uint64_t step = pp.inc * 32ull;
  const int64_t s_unf = static_cast<int64_t>(step);
  double sn, cs;
  sincospi(2.0 * (static_cast<double>(s_unf) / two64), &sn, &cs);
  pp.rot_re = static_cast<float>(cs);
  pp.rot_im = static_cast<float>(sn);
  int64_t s_f = s_unf;
  pp.sgn = 1.f;
  if (s_f >= (1ll << 62) || s_f < -(1ll << 62)) {
    s_f = static_cast<int64_t>(step ^ 0x8000000000000000ull);
    pp.sgn = -1.f;
  }
  const double of_turns = static_cast<double>(s_f) / two64;  // in [-1/4, 1/4]
  const double sh = sinpi(of_turns);                          // sin(Of/2)
  pp.kappa = static_cast<float>(4.0 * sh * sh);
  pp.so = static_cast<float>(sinpi(2.0 * of_turns));
  a.partials[v * a.ppad + p] = pp;
}

struct OscArgs {
  const PartialParams* partials;
  const VoiceParams* voices;
  int64_t n_voices;
  int64_t n_samples;
  int32_t ppad;
  int32_t mix;
  int32_t voices_per_group;
  int64_t n_chunks;
  int64_t first_sample;  // global index of output sample 0 (phases and envelopes are closed-form in the global index)
  float* out;
};

__device__ __forceinline__ float2 cis_turns_u64(uint64_t ph) {  // exp(2 pi i * ph / 2^64)
  const int32_t top = static_cast<int32_t>(ph >> 32);
  float s, c;
  sincospif(static_cast<float>(top) * (1.f / 2147483648.f), &s, &c);  // top/2^31 = 2*turns
  return make_float2(c, s);
}

// itp::interpolate as b + c*e(u), u = t/d in (0, 1]: Penner's easing set in the ca_itp numbering (oracle/SPEC.md S7;
// call sites include/audioelement.h:733-737, 994-998). Only the chunks that contain an envelope transition get here.
__device__ __forceinline__ float ease_in(int fam, float x) {
  switch (fam) {
    case 0: return x * x;
    case 1: return x * x * x;
    case 2: { const float x2 = x * x; return x2 * x2; }
    case 3: { const float x2 = x * x; return x2 * x2 * x; }
    case 4: return 1.f - cospif(0.5f * x);
    case 5: return x == 0.f ? 0.f : exp2f(10.f * (x - 1.f));
    default: return 1.f - sqrtf(fmaxf(0.f, 1.f - x * x));
  }
}
__device__ __forceinline__ float ease(int type, float u) {
  if (type == CA_ITP_LINEAR) return u;
  const int fam = (type - 1) / 3, mode = (type - 1) - 3 * fam;
  if (mode == 0) return ease_in(fam, u);
  if (mode == 1) return 1.f - ease_in(fam, 1.f - u);
  if (fam == 4) return 0.5f * (1.f - cospif(u));
  return u < 0.5f ? 0.5f * ease_in(fam, 2.f * u) : 1.f - 0.5f * ease_in(fam, 2.f - 2.f * u);
}

// value of the envelope while the key is pressed, c = step counter (>=1 once audible), :960-1000
template <bool EASED>
__device__ __forceinline__ float env_pressed(int64_t c, const PartialParams& q, const VoiceParams& vp) {
  if (c < 1) return 0.f;
  if (EASED) {
    if (c <= q.attack) return c == q.attack ? 1.f : ease(vp.itp & 0xff, static_cast<float>(c) * q.inv_attack);
    c -= q.attack;
    if (c <= vp.hold) return 1.f;
    c -= vp.hold;
    if (c < q.decay) return 1.f + (vp.sustain - 1.f) * ease((vp.itp >> 8) & 0xff, static_cast<float>(c) * q.inv_decay);
    return vp.sustain;
  }
  if (c <= q.attack) return static_cast<float>(c) * q.inv_attack;
  c -= q.attack;
  if (c <= vp.hold) return 1.f;
  c -= vp.hold;
  if (c <= q.decay) return 1.f + (vp.sustain - 1.f) * (static_cast<float>(c) * q.inv_decay);
  return vp.sustain;
}

template <bool EASED>
__device__ __forceinline__ float env_value_t(int64_t n, const PartialParams& q, const VoiceParams& vp) {
  const int64_t rel = vp.release_at;
  if (rel >= 0 && n >= rel) {
    const int64_t crel = rel - vp.key_delay;  // counter when onKeyReleased(0) is called, :769-781
    if (crel <= 0) return 0.f;
    const int64_t j = n - rel + 1;
    if (j >= q.release) return 0.f;
    const float top = env_pressed<EASED>(crel, q, vp);
    const float u = static_cast<float>(j) * q.inv_release;
    return top - top * (EASED ? ease((vp.itp >> 16) & 0xff, u) : u);  // :733-737
  }
  return env_pressed<EASED>(n + 1 - vp.key_delay, q, vp);
}

// EASED = false renders the voices whose three envelope phases are LINEAR (rt.resynth, every benchmark configuration) with the
// code the kernel had before easings existed; EASED = true renders the other voices (and only those) with the easing switch
// inlined in the transition chunks. Each launch skips the voices of the other kind, so a render is the sum of the two.
template <int WARPS, bool EASED>
__global__ void __launch_bounds__(WARPS * 32, 2) oscbank_kernel(const __grid_constant__ OscArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PartialParams* sp = reinterpret_cast<PartialParams*>(smem_raw);                 // [ppad]
  float2* E = reinterpret_cast<float2*>(sp + a.ppad);                             // [ppad][32] lane offsets
  float4* sgn4 = reinterpret_cast<float4*>(E + a.ppad * 32);                      // [ppad/4]
  float2* base_all = reinterpret_cast<float2*>(sgn4 + a.ppad / 4);                // [WARPS][kNP]
  int* kend_all = reinterpret_cast<int*>(base_all + WARPS * kNP);                 // [WARPS][kNP]
  float* cval_all = reinterpret_cast<float*>(kend_all + WARPS * kNP);             // [WARPS][kNP]
  float* outbuf_all = cval_all + WARPS * kNP;                                     // [WARPS][kChunk]
  float* bus_all = outbuf_all + WARPS * kChunk;                                   // [WARPS][kChunk*2] (mix only)
  __shared__ VoiceParams s_vp;
  __shared__ int s_max_press, s_max_release;

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float2* base = base_all + warp * kNP;
  int* kend = kend_all + warp * kNP;
  float* cval = cval_all + warp * kNP;
  float* outbuf = outbuf_all + warp * kChunk;
  float* bus = bus_all + warp * (kChunk * 2);

  const int64_t chunk = static_cast<int64_t>(blockIdx.x) * WARPS + warp;
  const bool has_chunk = chunk < a.n_chunks;
  const int64_t l0 = chunk * kChunk;         // local (output) index of the chunk
  const int64_t n0 = a.first_sample + l0;    // global sample index of the chunk
  const int64_t n_end = a.first_sample + a.n_samples;
  const int n_out = a.mix == CA_MIX_STEREO ? 2 : 1;

  if (a.mix != CA_MIX_NONE)
    for (int i = lane; i < kChunk * 2; i += 32) bus[i] = 0.f;

  const int64_t v_begin = static_cast<int64_t>(blockIdx.y) * a.voices_per_group;
  const int64_t v_end = min(a.n_voices, v_begin + a.voices_per_group);
  bool any_voice = false;
  for (int64_t v = v_begin; v < v_end; ++v) {
    if ((a.voices[v].itp != 0) != EASED) continue;  // uniform over the CTA
    any_voice = true;
    __syncthreads();
    // ---- per-voice tables ----
    if (threadIdx.x == 0) {
      s_vp = a.voices[v];
      s_max_press = 0;
      s_max_release = 0;
    }
    {
      const uint4* src = reinterpret_cast<const uint4*>(a.partials + v * a.ppad);
      uint4* dst = reinterpret_cast<uint4*>(sp);
      for (int i = threadIdx.x; i < a.ppad * 4; i += WARPS * 32) dst[i] = src[i];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < a.ppad * 32; i += WARPS * 32) {
      const int p = i >> 5, l = i & 31;
      E[i] = cis_turns_u64(sp[p].inc * static_cast<uint64_t>(l));
    }
    for (int p = threadIdx.x; p < a.ppad; p += WARPS * 32) {
      reinterpret_cast<float*>(sgn4)[p] = sp[p].sgn;
      if (sp[p].amp != 0.f) {
        atomicMax(&s_max_press, sp[p].attack + sp[p].decay);
        atomicMax(&s_max_release, sp[p].release);
      }
    }
    __syncthreads();
    const VoiceParams vp = s_vp;
    if (!has_chunk) continue;

    const int64_t n1 = min(n0 + kChunk, n_end);
    const int64_t press_end = static_cast<int64_t>(vp.key_delay) + s_max_press + vp.hold;  // last non-sustain counter
    const bool released = vp.release_at >= 0;
    const bool dead_note = released && (vp.release_at - vp.key_delay <= 0);
    bool silent = (n1 - vp.key_delay < 1);  // chunk entirely before the attack starts
    if (released) {
      if (dead_note) silent = true;  // released before the key delay elapsed: the note is skipped (:774-777)
      else if (n0 >= vp.release_at + s_max_release - 1) silent = true;  // every partial has finished its release
    }
    const bool steady = !silent && (n0 + 1 > press_end) && (!released || n1 <= vp.release_at);

    if (silent) {
#pragma unroll 4
      for (int k = 0; k < kR; ++k) outbuf[32 * k + lane] = 0.f;
    } else if (steady) {
      // ================= fast path: Reinsch recurrence, constant envelope =================
      for (int p0 = 0; p0 < a.ppad; p0 += kNP) {
        // exact seeds: lane p computes amp*S*exp(i*theta(n0)) for partial p0+lane
        {
          const PartialParams& q = sp[p0 + lane];
          const float2 z = cis_turns_u64(q.ph0 + q.inc * static_cast<uint64_t>(n0 + 1));
          const float am = q.amp * vp.sustain;
          base[lane] = make_float2(am * z.x, am * z.y);
        }
        __syncwarp();
        // two partials per register pair: the recurrence and the accumulation run on FFMA2 / FADD2 (packed.cuh), half
        // the issue slots of the scalar form for the same FP32 work
        c2 y[kNP / 2], d[kNP / 2], nkap[kNP / 2];
#pragma unroll
        for (int p = 0; p < kNP; p += 2) {
          const float2 z0 = cmul(base[p], E[(p0 + p) * 32 + lane]);
          const float2 z1 = cmul(base[p + 1], E[(p0 + p + 1) * 32 + lane]);
          const float k0 = sp[p0 + p].kappa, k1 = sp[p0 + p + 1].kappa;
          nkap[p / 2] = pk(-k0, -k1);
          y[p / 2] = pk(z0.y, z1.y);
          // y0 - y(-1) = Im(z * (1 - exp(-i Of)))
          d[p / 2] = pk(fmaf(z0.x, sp[p0 + p].so, 0.5f * k0 * z0.y), fmaf(z1.x, sp[p0 + p + 1].so, 0.5f * k1 * z1.y));
        }
        __syncwarp();
        const c2* sgn2 = reinterpret_cast<const c2*>(sgn4) + (p0 >> 1);
#pragma unroll 1
        for (int k = 0; k < kR; k += 2) {
          // even step: every partial enters with a plus sign
          c2 acc0 = y[0], acc1 = y[1], acc2 = y[2], acc3 = y[3];
#pragma unroll
          for (int p = 4; p < kNP / 2; p += 4) {
            acc0 = add2(acc0, y[p]);
            acc1 = add2(acc1, y[p + 1]);
            acc2 = add2(acc2, y[p + 2]);
            acc3 = add2(acc3, y[p + 3]);
          }
          {
            const float2 t = upk(add2(add2(acc0, acc1), add2(acc2, acc3)));
            const float s = t.x + t.y;
            if (p0 == 0) outbuf[32 * k + lane] = s;
            else outbuf[32 * k + lane] += s;
          }
#pragma unroll
          for (int p = 0; p < kNP / 2; ++p) {
            d[p] = fma2(nkap[p], y[p], d[p]);
            y[p] = add2(y[p], d[p]);
          }
          // odd step: folded partials enter with a minus sign
          acc0 = mul2(sgn2[0], y[0]);
          acc1 = mul2(sgn2[1], y[1]);
          acc2 = mul2(sgn2[2], y[2]);
          acc3 = mul2(sgn2[3], y[3]);
#pragma unroll
          for (int p = 4; p < kNP / 2; p += 4) {
            acc0 = fma2(sgn2[p], y[p], acc0);
            acc1 = fma2(sgn2[p + 1], y[p + 1], acc1);
            acc2 = fma2(sgn2[p + 2], y[p + 2], acc2);
            acc3 = fma2(sgn2[p + 3], y[p + 3], acc3);
          }
          {
            const float2 t = upk(add2(add2(acc0, acc1), add2(acc2, acc3)));
            const float s = t.x + t.y;
            if (p0 == 0) outbuf[32 * (k + 1) + lane] = s;
            else outbuf[32 * (k + 1) + lane] += s;
          }
#pragma unroll
          for (int p = 0; p < kNP / 2; ++p) {
            d[p] = fma2(nkap[p], y[p], d[p]);
            y[p] = add2(y[p], d[p]);
          }
        }
      }
    } else {
      // ================= general path: complex rotation + envelope =================
      // The envelope of a partial is piecewise linear and becomes constant (sustain level, or 0 after the release) at a
      // sample T_p; only the steps of the chunk before T_p evaluate it per sample, the others use the constant.
      for (int p0 = 0; p0 < a.ppad; p0 += kNP) {
        {
          const PartialParams& q = sp[p0 + lane];
          const float2 z = cis_turns_u64(q.ph0 + q.inc * static_cast<uint64_t>(n0 + 1));
          base[lane] = make_float2(q.amp * z.x, q.amp * z.y);
          const bool rel_seen = released && vp.release_at < n1;
          const int64_t T = rel_seen ? vp.release_at + q.release - 1
                                     : static_cast<int64_t>(vp.key_delay) + q.attack + vp.hold + q.decay;
          const int64_t ke = (T - n0 + 31) >> 5;
          kend[lane] = static_cast<int>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(kR), ke)));
          cval[lane] = rel_seen ? 0.f : vp.sustain;
        }
        __syncwarp();
        float2 z[kNP];
#pragma unroll
        for (int p = 0; p < kNP; ++p) z[p] = cmul(base[p], E[(p0 + p) * 32 + lane]);
#pragma unroll 1
        for (int k = 0; k < kR; ++k) {
          const int64_t n = n0 + 32 * k + lane;
          float acc = 0.f;
#pragma unroll
          for (int p = 0; p < kNP; ++p) {
            const PartialParams& q = sp[p0 + p];
            float e = cval[p];
            if (k < kend[p]) e = env_value_t<EASED>(n, q, vp);
            acc = fmaf(z[p].y, e, acc);
            z[p] = cmul(z[p], make_float2(q.rot_re, q.rot_im));
          }
          if (p0 == 0) outbuf[32 * k + lane] = acc;
          else outbuf[32 * k + lane] += acc;
        }
        __syncwarp();
      }
    }
    __syncwarp();

    // ---- emit ----
    if (a.mix == CA_MIX_NONE) {
      float* o = a.out + v * a.n_samples + l0;
#pragma unroll 4
      for (int k = 0; k < kR; ++k) {
        const int i = 32 * k + lane;
        if (l0 + i < a.n_samples) o[i] = outbuf[i];
      }
    } else if (!silent) {
      if (a.mix == CA_MIX_STEREO) {
        float2* b2 = reinterpret_cast<float2*>(bus);
        for (int k = 0; k < kR; ++k) {
          const int i = 32 * k + lane;
          const float s = outbuf[i];
          float2 acc = b2[i];
          acc.x = fmaf(vp.gain_l, s, acc.x);
          acc.y = fmaf(vp.gain_r, s, acc.y);
          b2[i] = acc;
        }
      } else {
        for (int k = 0; k < kR; ++k) {
          const int i = 32 * k + lane;
          bus[i] += outbuf[i];
        }
      }
    }
    __syncwarp();
  }

  if (a.mix != CA_MIX_NONE && has_chunk && any_voice) {
    float* o = a.out + l0 * n_out;
    const int64_t lim = (min(l0 + kChunk, a.n_samples) - l0) * n_out;
    for (int i = lane; i < lim; i += 32) atomicAdd(o + i, bus[i]);
  }
}

__global__ void fp32_peak_kernel(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f,
        a7 = a0 + 7.f;
  const float m = 0.999999f, c = 1e-7f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
      a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
    }
  }
  const float s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 12345.678f) out[0] = s;
}

}  // namespace

int oscbank_plan(ca_ctx* ctx, int64_t max_voices, int32_t max_partials) {
  if (max_voices < 1 || max_partials < 1) return fail(ctx, CA_ERR_INVALID, "ca_oscbank_plan: sizes must be positive");
  if (max_partials > kMaxPartials)
    return fail(ctx, CA_ERR_UNSUPPORTED, "ca_oscbank_plan: more than 128 partials per voice is not built");
  const int32_t ppad = (max_partials + kNP - 1) / kNP * kNP;
  oscbank_release(ctx);
  CA_CUDA_OK(ctx, cudaMalloc(&ctx->d_osc_partials, sizeof(PartialParams) * static_cast<size_t>(max_voices) * ppad));
  CA_CUDA_OK(ctx, cudaMalloc(&ctx->d_osc_voices, sizeof(VoiceParams) * static_cast<size_t>(max_voices)));
  ctx->osc_max_voices = max_voices;
  ctx->osc_max_partials = max_partials;
  return CA_OK;
}

void oscbank_release(ca_ctx* ctx) {
  if (ctx->d_osc_partials) cudaFree(ctx->d_osc_partials);
  if (ctx->d_osc_voices) cudaFree(ctx->d_osc_voices);
  ctx->d_osc_partials = ctx->d_osc_voices = nullptr;
  ctx->osc_max_voices = 0;
  ctx->osc_max_partials = 0;
}

// eased: what the caller knows about the envelopes of this call. 0: every phase of every envelope is LINEAR (only the LINEAR
// pass is launched); 2: every envelope has an eased phase (only the eased pass); 1: both kinds; -1: unknown, the envelopes live
// in device memory (both passes are launched, each skips the voices of the other kind).
int oscbank_render(ca_ctx* ctx, const ca_oscbank_desc* desc, float* d_out, cudaStream_t stream, int eased) {
  if (!desc || !d_out) return fail(ctx, CA_ERR_INVALID, "ca_oscbank_render: null argument");
  const ca_oscbank_desc& d = *desc;
  if (d.n_voices < 0 || d.n_partials < 1 || d.n_samples < 0 || d.first_sample < 0 || d.sample_rate <= 0 || !d.f0_hz || !d.volume ||
      !d.partial_volume || !d.partial_phase || !d.env || (d.n_env != 1 && d.n_env != d.n_voices))
    return fail(ctx, CA_ERR_INVALID, "ca_oscbank_render: bad descriptor");
  if (d.mix != CA_MIX_NONE && d.mix != CA_MIX_STEREO && d.mix != CA_MIX_MONO)
    return fail(ctx, CA_ERR_INVALID, "ca_oscbank_render: bad mix mode");
  if (d.n_voices > ctx->osc_max_voices || d.n_partials > ctx->osc_max_partials)
    return fail(ctx, CA_ERR_NOT_PLANNED, "ca_oscbank_render: call ca_oscbank_plan with sizes covering this render");
  if (d.n_voices == 0 || d.n_samples == 0) return CA_OK;
  const int32_t ppad = (d.n_partials + kNP - 1) / kNP * kNP;

  PrepArgs pa{d, static_cast<PartialParams*>(ctx->d_osc_partials), static_cast<VoiceParams*>(ctx->d_osc_voices), ppad};
  const int64_t n_prep = d.n_voices * ppad;
  osc_prep_kernel<<<static_cast<unsigned>((n_prep + 127) / 128), 128, 0, stream>>>(pa);
  CA_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches += 1;

  OscArgs a{};
  a.partials = pa.partials;
  a.voices = pa.voices;
  a.n_voices = d.n_voices;
  a.n_samples = d.n_samples;
  a.ppad = ppad;
  a.mix = d.mix;
  a.n_chunks = (d.n_samples + kChunk - 1) / kChunk;
  a.first_sample = d.first_sample;
  a.out = d_out;

  const size_t common = sizeof(PartialParams) * ppad + sizeof(float2) * ppad * 32 + sizeof(float) * ppad;
  if (d.mix == CA_MIX_NONE) {
    constexpr int WARPS = 8;
    a.voices_per_group = 1;
    const size_t smem = common + WARPS * (sizeof(float2) * kNP * 2 + sizeof(float) * kChunk);
    CA_CUDA_OK(ctx, cudaFuncSetAttribute(oscbank_kernel<WARPS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
    CA_CUDA_OK(ctx, cudaFuncSetAttribute(oscbank_kernel<WARPS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
    dim3 grid(static_cast<unsigned>((a.n_chunks + WARPS - 1) / WARPS), static_cast<unsigned>(d.n_voices));
    if (d.n_voices > 65535) return fail(ctx, CA_ERR_UNSUPPORTED, "ca_oscbank_render: more than 65535 voices per call");
    if (eased != 2) oscbank_kernel<WARPS, false><<<grid, WARPS * 32, smem, stream>>>(a);
    if (eased != 0) oscbank_kernel<WARPS, true><<<grid, WARPS * 32, smem, stream>>>(a);
  } else {
    constexpr int WARPS = 4;
    const int n_out = d.mix == CA_MIX_STEREO ? 2 : 1;
    CA_CUDA_OK(ctx, cudaMemsetAsync(d_out, 0, sizeof(float) * d.n_samples * n_out, stream));
    const size_t smem = common + WARPS * (sizeof(float2) * kNP * 2 + sizeof(float) * kChunk * 3);
    CA_CUDA_OK(ctx, cudaFuncSetAttribute(oscbank_kernel<WARPS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
    CA_CUDA_OK(ctx, cudaFuncSetAttribute(oscbank_kernel<WARPS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
    const int64_t tiles = (a.n_chunks + WARPS - 1) / WARPS;
    // enough voice groups to fill the GPU a few times over, few enough that the final atomics stay negligible
    int64_t groups = (static_cast<int64_t>(ctx->sm_count) * 8 + tiles - 1) / tiles;
    if (groups < 1) groups = 1;
    if (groups > d.n_voices) groups = d.n_voices;
    a.voices_per_group = static_cast<int32_t>((d.n_voices + groups - 1) / groups);
    groups = (d.n_voices + a.voices_per_group - 1) / a.voices_per_group;
    dim3 grid(static_cast<unsigned>(tiles), static_cast<unsigned>(groups));
    if (eased != 2) oscbank_kernel<WARPS, false><<<grid, WARPS * 32, smem, stream>>>(a);
    if (eased != 0) oscbank_kernel<WARPS, true><<<grid, WARPS * 32, smem, stream>>>(a);
  }
  CA_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches += (eased == -1 || eased == 1) ? 2 : 1;
  return CA_OK;
}

int measure_fp32_peak(ca_ctx* ctx, double* tflops) {
  if (!tflops) return fail(ctx, CA_ERR_INVALID, "ca_measure_fp32_peak: null output");
  float* d = nullptr;
  CA_CUDA_OK(ctx, cudaMalloc(&d, sizeof(float)));
  cudaEvent_t e0, e1;
  CA_CUDA_OK(ctx, cudaEventCreate(&e0));
  CA_CUDA_OK(ctx, cudaEventCreate(&e1));
  const int iters = 4096, threads = 256, blocks = ctx->sm_count * 8;
  fp32_peak_kernel<<<blocks, threads>>>(d, 64);
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    fp32_peak_kernel<<<blocks, threads>>>(d, iters);
    cudaEventRecord(e1);
    CA_CUDA_OK(ctx, cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 16 * 8 * static_cast<double>(iters) * threads * blocks;
    best = fmax(best, flops / (ms * 1e-3) / 1e12);
  }
  ctx->launches += 6;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return CA_OK;
}

}  // namespace ca
