// oscbank.cu — K3: sine oscillator bank (harmonic voices) rendered for whole notes at once.
//
// Reference semantics (include/audioelement.h:486-657 MultiEnveloped, :2366-2452 SineOscillatorAlgo, :666-1099 AHDSR
// envelope, :1111-1265 VolumeAdjusted; include/gen.crtp.h:350-378 compute): sample n of a voice is
//     volume * sum_p vol_p * env_p(n) * g_p * sin(pi * (phase_p + p*a0 + (n+1) * p*ai))
// with one envelope per partial (attack/decay/release floors of 2.5 periods of THAT partial, :216-223).
//
// B200 design (not a translation of the per-sample phasor loop):
//   * lanes = 32 consecutive samples; a warp renders a chunk of R*32 samples of one voice.
//   * in the sustain regime the chunk is a small matrix product on the FP32 pipe: sample 32k+l of the chunk is
//         sum_p Im( e_p[l] * F_p[k] ),   e_p[l] = amp_p exp(i theta_p(n0+l))  (32 values per partial, per chunk),
//                                        F_p[k] = exp(i 32 k w_p)             (64 values per partial, per voice),
//     i.e. TWO fused multiply-adds per partial-sample and no serial dependence between steps. A thread owns 8 l x 8 k
//     outputs (32 packed accumulators) and reads 8+8 complex operands per partial, so one partial costs 64 FFMA2 for
//     8 LDS.128; the k operand enters FFMA2 as a 32-bit broadcast register. (Round 1's Reinsch recurrence
//     d -= kappa*y ; y += d ; acc += y was 3 operations per partial-sample in a dependent chain that the register-file
//     ports held to 2.2-2.6 of 4 warp-instructions per clock, profiles/microbench/fp32_ports_b200.txt.)
//   * every chunk is re-seeded from an EXACT phase: phases are 64-bit fixed-point turns (phase0 + (n+1)*inc mod 2^64),
//     so fp32 never accumulates phase over the 480,000-sample note (the reference runs in double).
//   * chunks that contain an envelope transition (attack, decay, release) take a general path: complex rotation per
//     32 samples + the envelope evaluated per sample (piecewise linear, or eased: itp::interpolate with the ca_itp set).
//   * the kernel is instantiated twice, for voices with LINEAR phases only and for voices with an eased phase; a render
//     launches the pass(es) its envelopes need and each pass skips the voices of the other kind.
//   * partial parameters are prepared once per render by a small kernel, in double where the reference is in double.
#include <cuda_fp16.h>

#include <algorithm>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "ctx.hpp"
#include "packed.cuh"

namespace ca {

namespace {

constexpr int kNP = 32;       // partials per pass (register resident)
constexpr int kR = 64;        // 32-sample steps per chunk
constexpr int kChunk = 32 * kR;
constexpr int kMaxPartials = 128;
constexpr int kUmmaChunk = 128 * 128;  // samples per chunk of the tcgen05 kernel (osc_umma_kernel)
constexpr int kEP = 36;       // row pitch of the lane-offset table in float2: rows p .. p+3 start 8 banks apart

struct __align__(16) PartialParams {
  uint64_t inc;   // phase increment per sample, turns * 2^64
  uint64_t ph0;   // phase before the first step
  float amp;      // volume * partial volume * aliasing gain (0 for padding / dropped harmonics)
  float pad0[3];
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
  pp.pad0[0] = pp.pad0[1] = pp.pad0[2] = 0.f;
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
  int32_t chunks_per_warp;
  int32_t umma_skip;  // 1: the steady 16,384-sample chunks are rendered by osc_umma_kernel, this launch fills in the rest
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

// The tensor-core form of the sustain-regime product: 3 x FP16. An operand pair (x, y) is split into hi = fp16(x, y) and
// lo = fp16((x, y) - hi); hi*hi' + lo*hi' + hi*lo' reproduces the FP32 product to 2^-22 relative (the dropped lo*lo' term) as long
// as the lo parts are normal FP16 numbers, which the power-of-two operand scales (kScaleB, the per-voice scale of A) arrange; the
// tensor core accumulates in FP32. mma.m16n8k16 on FP16 issues at the rate of mma.m16n8k8 on TF32 and does twice the work
// (profiles/microbench/mma_tf32_b200.txt), which is why this is not 3 x TF32.
__device__ __forceinline__ void f16_split(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x, y);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ float2 f16_join(uint32_t hi, uint32_t lo) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hi)), b = __half22float2(*reinterpret_cast<const __half2*>(&lo));
  return make_float2(a.x + b.x, a.y + b.y);
}
// D(16x8) += A(16x16, row) * B(16x8, col), FP16 operands, FP32 accumulation (SASS HMMA.16816.F32)
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
constexpr float kScaleB = 256.f;  // exp(i x) * 2^8: lo parts are normal down to |value| = 5e-4, below that exact to 1e-10

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
// TENSOR = true runs the sustain-regime product on the tensor cores (3 x FP16, above), TENSOR = false on FFMA2.
template <int WARPS, int MINB, bool EASED, bool TENSOR>
__global__ void __launch_bounds__(WARPS * 32, MINB) oscbank_kernel(const __grid_constant__ OscArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PartialParams* sp = reinterpret_cast<PartialParams*>(smem_raw);                 // [ppad]
  float2* E = reinterpret_cast<float2*>(sp + a.ppad);                             // [ppad][kEP] exp(i l w_p), l < 32
  float* Fx = reinterpret_cast<float*>(E + a.ppad * kEP);                         // [ppad][kR] Re exp(i 32 k w_p)
  float* Fy = Fx + a.ppad * kR;                                                   // [ppad][kR] Im
  uint4* Ft = reinterpret_cast<uint4*>(Fx);                                       // TENSOR: [ppad/8][8][32] B fragments
  float2* base_all = reinterpret_cast<float2*>(Fx + a.ppad * kR * 2);             // [WARPS][kNP]
  int* kend_all = reinterpret_cast<int*>(base_all + WARPS * kNP);                 // [WARPS][kNP]
  float* cval_all = reinterpret_cast<float*>(kend_all + WARPS * kNP);             // [WARPS][kNP]
  float* outbuf_all = cval_all + WARPS * kNP;                                     // [WARPS][kChunk]
  float* bus_all = outbuf_all + WARPS * kChunk;                                   // [WARPS][kChunk*2] (mix only)
  __shared__ VoiceParams s_vp;
  __shared__ int s_max_press, s_max_release;
  __shared__ float s_amp_max;

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float2* base = base_all + warp * kNP;
  int* kend = kend_all + warp * kNP;
  float* cval = cval_all + warp * kNP;
  float* outbuf = outbuf_all + warp * kChunk;
  float* bus = bus_all + warp * (kChunk * 2);

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
      s_amp_max = 0.f;
    }
    {
      const uint4* src = reinterpret_cast<const uint4*>(a.partials + v * a.ppad);
      uint4* dst = reinterpret_cast<uint4*>(sp);
      for (int i = threadIdx.x; i < a.ppad * 4; i += WARPS * 32) dst[i] = src[i];
    }
    __syncthreads();
    for (int p = threadIdx.x; p < a.ppad; p += WARPS * 32) {
      if (sp[p].amp != 0.f) {
        atomicMax(reinterpret_cast<int*>(&s_amp_max), __float_as_int(fabsf(sp[p].amp)));  // non-negative floats order as ints
        atomicMax(&s_max_press, sp[p].attack + sp[p].decay);
        atomicMax(&s_max_release, sp[p].release);
      }
    }
    __syncthreads();
    const VoiceParams vp = s_vp;
    const int64_t press_end = static_cast<int64_t>(vp.key_delay) + s_max_press + vp.hold;  // last non-sustain counter
    const bool released = vp.release_at >= 0;
    const bool dead_note = released && (vp.release_at - vp.key_delay <= 0);
    // Which 2048-sample chunks does the tcgen05 kernel render? The steady ones, m_lo .. m_hi: first sample at or after the end
    // of the pressed transitions, last sample before the release (the same closed form as in osc_umma_kernel; it also writes
    // sustain-level signal over the other chunks of its 16,384-sample chunks, which this kernel, launched after it, renders).
    int64_t m_lo = 1, m_hi = 0;
    if (a.umma_skip && vp.itp == 0) {
      const int64_t d0 = press_end - a.first_sample;
      m_lo = d0 <= 0 ? 0 : (d0 + kChunk - 1) / kChunk;
      m_hi = a.n_chunks - 1;
      if (released && vp.release_at < n_end) {
        const int64_t d1 = vp.release_at - a.first_sample;
        m_hi = d1 < kChunk ? -1 : min(m_hi, d1 / kChunk - 1);
      }
    }
    auto umma_renders = [&](int64_t chunk) { return chunk >= m_lo && chunk <= m_hi; };
    if (a.umma_skip) {  // nothing to do for this voice in this CTA: skip the tables too (uniform)
      const int64_t first = static_cast<int64_t>(blockIdx.x) * a.chunks_per_warp * WARPS;
      const int64_t last = min(a.n_chunks, first + static_cast<int64_t>(a.chunks_per_warp) * WARPS);
      if (first >= last || (first >= m_lo && last - 1 <= m_hi)) continue;
    }
    for (int i = threadIdx.x; i < a.ppad * 32; i += WARPS * 32) {
      const int p = i >> 5, l = i & 31;
      E[p * kEP + l] = cis_turns_u64(sp[p].inc * static_cast<uint64_t>(l));
    }
    if (TENSOR) {
      // B fragments of mma.m16n8k16, ready to load: entry (s, i, ln) serves lane ln = 4 g + t in step s (partials 8s .. 8s+7) for
      // the n tile i, k = 8 i + g: rows 2t, 2t+1 = (Im, Re) of partial 8s + t, rows 2t+8, 2t+9 = the same of partial 8s + 4 + t,
      // as (hi, hi', lo, lo'): two register pairs, as the instruction wants them. A lane reads 16 bytes next to its neighbours'.
      for (int i = threadIdx.x; i < a.ppad * 32; i += WARPS * 32) {  // (ppad / 8) steps x 8 tiles x 32 lanes
        const int ln = i & 31, nt = (i >> 5) & 7, st = i >> 8;
        const int p = 8 * st + (ln & 3), k = 8 * nt + (ln >> 2);
        const float2 f0 = cis_turns_u64(sp[p].inc * static_cast<uint64_t>(32 * k));
        const float2 f1 = cis_turns_u64(sp[p + 4].inc * static_cast<uint64_t>(32 * k));
        uint4 e;
        f16_split(kScaleB * f0.y, kScaleB * f0.x, e.x, e.z);
        f16_split(kScaleB * f1.y, kScaleB * f1.x, e.y, e.w);
        Ft[i] = e;
      }
    }
    for (int i = threadIdx.x; i < (TENSOR ? 0 : a.ppad * kR); i += WARPS * 32) {
      // row p holds k = 8 tk + 4 h + j at position 32 h + 4 tk + j: the two 16-byte loads of a thread (h = 0, 1) are then
      // contiguous across the eight tk of a quarter warp (no bank conflict)
      const int p = i / kR, k = i - p * kR;
      const float2 f = cis_turns_u64(sp[p].inc * static_cast<uint64_t>(32 * k));
      const int pos = p * kR + ((k >> 2) & 1) * 32 + (k >> 3) * 4 + (k & 3);
      Fx[pos] = f.x;
      Fy[pos] = f.y;
    }
    __syncthreads();
#pragma unroll 1
    for (int ci = 0; ci < a.chunks_per_warp; ++ci) {
      const int64_t chunk = (static_cast<int64_t>(blockIdx.x) * a.chunks_per_warp + ci) * WARPS + warp;
      if (chunk >= a.n_chunks) break;
      if (umma_renders(chunk)) continue;
      const int64_t l0 = chunk * kChunk;         // local (output) index of the chunk
      const int64_t n0 = a.first_sample + l0;    // global sample index of the chunk
      const int64_t n1 = min(n0 + kChunk, n_end);
      bool silent = (n1 - vp.key_delay < 1);  // chunk entirely before the attack starts
      if (released) {
        if (dead_note) silent = true;  // released before the key delay elapsed: the note is skipped (:774-777)
        else if (n0 >= vp.release_at + s_max_release - 1) silent = true;  // every partial has finished its release
      }
      const bool steady = !silent && (n0 + 1 > press_end) && (!released || n1 <= vp.release_at);

      // the constant part of the chunk: every partial at the level its envelope settles on (sustain; 0 once the release
      // is in sight). A steady chunk is nothing else; a chunk with a transition adds corrections below.
      const bool rel_seen = released && vp.release_at < n1;
      const float cv = rel_seen ? 0.f : vp.sustain;
      // a whole steady chunk of a per-voice render goes from the accumulators to HBM (8 lanes = one 32-byte sector)
      const bool direct = TENSOR && a.mix == CA_MIX_NONE && steady && l0 + kChunk <= a.n_samples;
      // scale of the A operand (TENSOR): the power of two that brings the loudest partial of the voice just below 2^15, so that
      // the lo halves of all but the faintest partials are normal FP16 numbers
      int amp_exp;
      frexpf(s_amp_max * cv, &amp_exp);
      const float scale_a = exp2f(static_cast<float>(min(100, 15 - amp_exp)));
      const float unscale = 1.f / (scale_a * kScaleB);
      if (silent || cv == 0.f) {
#pragma unroll 4
        for (int k = 0; k < kR; ++k) outbuf[32 * k + lane] = 0.f;
      } else if (TENSOR) {
        // ================= fast path on the tensor cores: chunk(32 l x 64 k) = A(32 x 2P) B(2P x 64), 3 x FP16 =================
        // A[l][2p], A[l][2p+1] = Re, Im of e_p[l] = amp_p exp(i theta_p(n0 + l));  B[2p][k], B[2p+1][k] = Im, Re of exp(i 32 k w_p).
        // One mma step takes eight partials (sixteen columns of A); lane (g, t) of the warp holds partials 8s + t and 8s + 4 + t:
        // its A values are e_p at l = g, g+8, g+16, g+24 (computed here from the seed and the lane-offset table, split hi/lo in
        // registers), its B values come from the per-voice fragment table.
        const int g = lane >> 2, t = lane & 3;
        float acc[2][8][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[mt][nt][c] = 0.f;
        // exact seeds of every partial, amp * scale * exp(i theta(n0)), in the (still unused) output staging buffer
        float2* seed = reinterpret_cast<float2*>(outbuf);
        for (int p = lane; p < a.ppad; p += 32) {
          const PartialParams& q = sp[p];
          const float2 z = cis_turns_u64(q.ph0 + q.inc * static_cast<uint64_t>(n0 + 1));
          const float am = q.amp * cv * scale_a;
          seed[p] = make_float2(am * z.x, am * z.y);
        }
        __syncwarp();
        // A fragments of step st (partials 8 st + t and 8 st + 4 + t)
        auto gen_a = [&](int st, uint32_t (&hi)[2][4], uint32_t (&lo)[2][4]) {
          const float2 b0 = seed[8 * st + t], b1 = seed[8 * st + 4 + t];
          const float2* er0 = E + (8 * st + t) * kEP + g;
          const float2* er1 = er0 + 4 * kEP;
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const float2 e0 = cmul(b0, er0[16 * mt]), e1 = cmul(b0, er0[16 * mt + 8]);
            const float2 e2 = cmul(b1, er1[16 * mt]), e3 = cmul(b1, er1[16 * mt + 8]);
            f16_split(e0.x, e0.y, hi[mt][0], lo[mt][0]);  // row g,     columns 2t, 2t+1
            f16_split(e1.x, e1.y, hi[mt][1], lo[mt][1]);  // row g + 8
            f16_split(e2.x, e2.y, hi[mt][2], lo[mt][2]);  // row g,     columns 2t+8, 2t+9
            f16_split(e3.x, e3.y, hi[mt][3], lo[mt][3]);  // row g + 8
          }
        };
        // one step: three sweeps over the sixteen accumulator tiles, so that consecutive tensor instructions never wait for each
        // other. (Preparing the A operands of the next step between the tensor instructions of this one, by hand or by ptxas,
        // measured the same: the other warps of the scheduler already fill those slots.)
        const int n_steps = a.ppad >> 3;
#pragma unroll 1
        for (int st = 0; st < n_steps; ++st) {
          uint32_t hi[2][4], lo[2][4];
          gen_a(st, hi, lo);
          const uint4* fq = Ft + st * 256 + lane;
          uint4 f[8];
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) f[nt] = fq[nt * 32];
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) {
            mma_f16(acc[0][nt], lo[0], f[nt].x, f[nt].y);
            mma_f16(acc[1][nt], lo[1], f[nt].x, f[nt].y);
          }
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) {
            mma_f16(acc[0][nt], hi[0], f[nt].z, f[nt].w);
            mma_f16(acc[1][nt], hi[1], f[nt].z, f[nt].w);
          }
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) {
            mma_f16(acc[0][nt], hi[0], f[nt].x, f[nt].y);
            mma_f16(acc[1][nt], hi[1], f[nt].x, f[nt].y);
          }
        }
        __syncwarp();
        // accumulator (mt, nt): c0, c1 = (l = 16 mt + g, k = 8 nt + 2t, +1), c2, c3 = (l + 8, same k)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) {
            float* o = (direct ? a.out + v * a.n_samples + l0 : outbuf) + 32 * (8 * nt + 2 * t) + 16 * mt + g;
            o[0] = acc[mt][nt][0] * unscale;
            o[32] = acc[mt][nt][1] * unscale;
            o[8] = acc[mt][nt][2] * unscale;
            o[40] = acc[mt][nt][3] * unscale;
          }
      } else {
        // ================= fast path: constant envelope, chunk = Im(e F) on FFMA2 =================
        const int tl = lane >> 3, tk = lane & 7;  // the thread owns l = 8 tl .. 8 tl + 7 and k = 8 tk .. 8 tk + 7
        c2 acc[4][8];                             // [pair of l][k]
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = 0ull;
        float* ex = outbuf;              // [kNP][32] Re e_p[l]   (the operand planes live in the output staging buffer,
        float* ey = outbuf + kNP * 32;   // [kNP][32] Im e_p[l]    which is written after the last pass)
        for (int p0 = 0; p0 < a.ppad; p0 += kNP) {
          // exact seeds: lane p computes amp*S*exp(i*theta(n0)) for partial p0+lane
          {
            const PartialParams& q = sp[p0 + lane];
            const float2 z = cis_turns_u64(q.ph0 + q.inc * static_cast<uint64_t>(n0 + 1));
            const float am = q.amp * cv;
            base[lane] = make_float2(am * z.x, am * z.y);
          }
          __syncwarp();
#pragma unroll 8
          for (int p = 0; p < kNP; ++p) {
            const float2 z = cmul(base[p], E[(p0 + p) * kEP + lane]);
            ex[p * 32 + lane] = z.x;
            ey[p * 32 + lane] = z.y;
          }
          __syncwarp();
          const float4* ex4 = reinterpret_cast<const float4*>(ex + 8 * tl);
          const float4* ey4 = reinterpret_cast<const float4*>(ey + 8 * tl);
          const float4* fx4 = reinterpret_cast<const float4*>(Fx + p0 * kR + 4 * tk);
          const float4* fy4 = reinterpret_cast<const float4*>(Fy + p0 * kR + 4 * tk);
#pragma unroll 2
          for (int p = 0; p < kNP; ++p) {
            const float4 a0 = ex4[p * 8], a1 = ex4[p * 8 + 1];
            const float4 b0 = ey4[p * 8], b1 = ey4[p * 8 + 1];
            const float4 f0 = fx4[p * (kR / 4)], f1 = fx4[p * (kR / 4) + 8];
            const float4 g0 = fy4[p * (kR / 4)], g1 = fy4[p * (kR / 4) + 8];
            const c2 xr[4] = {pk(a0.x, a0.y), pk(a0.z, a0.w), pk(a1.x, a1.y), pk(a1.z, a1.w)};
            const c2 xi[4] = {pk(b0.x, b0.y), pk(b0.z, b0.w), pk(b1.x, b1.y), pk(b1.z, b1.w)};
            const float fr[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
            const float fi[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[i][j] = fma2(xi[i], bc(fr[j]), fma2(xr[i], bc(fi[j]), acc[i][j]));
          }
          __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float* o = outbuf + 32 * (8 * tk + j) + 8 * tl;
          const float2 v0 = upk(acc[0][j]), v1 = upk(acc[1][j]), v2 = upk(acc[2][j]), v3 = upk(acc[3][j]);
          *reinterpret_cast<float4*>(o) = make_float4(v0.x, v0.y, v1.x, v1.y);
          *reinterpret_cast<float4*>(o + 4) = make_float4(v2.x, v2.y, v3.x, v3.y);
        }
      }
      __syncwarp();
      if (!silent && !steady) {
        // ================= envelope transitions =================
        // The envelope of a partial is piecewise and settles (sustain level, or 0 after the release) at a sample T_p; the
        // steps of the chunk before T_p get (env(n) - settled level) * amp * sin(theta) added, one partial and one step at a
        // time: sin(theta(n0 + 32k + l)) = Im(e[l] F[k]) from the same tables as the product.
        const float lane_f = static_cast<float>(lane);
        for (int p0 = 0; p0 < a.ppad; p0 += kNP) {
          {
            const PartialParams& q = sp[p0 + lane];
            const float2 z = cis_turns_u64(q.ph0 + q.inc * static_cast<uint64_t>(n0 + 1));
            base[lane] = make_float2(q.amp * z.x, q.amp * z.y);
            const int64_t T = rel_seen ? vp.release_at + q.release - 1
                                       : static_cast<int64_t>(vp.key_delay) + q.attack + vp.hold + q.decay;
            const int64_t ke = (T - n0 + 31) >> 5;
            kend[lane] = q.amp == 0.f ? 0 : static_cast<int>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(kR), ke)));
            // value the release starts from (:769-781), per partial
            cval[lane] = rel_seen ? env_pressed<EASED>(vp.release_at - vp.key_delay, q, vp) : 0.f;
          }
          __syncwarp();
          // step counter of sample n0 (+1: the envelope steps before it is read), clamped where float holds it exactly enough
          const float c0 = static_cast<float>(max(static_cast<int64_t>(-(1 << 30)), min(static_cast<int64_t>(1 << 30), n0 + 1 - vp.key_delay)));
          const float crel = static_cast<float>(max(static_cast<int64_t>(-(1 << 30)), min(static_cast<int64_t>(1 << 30), vp.release_at - vp.key_delay)));
#pragma unroll 1
          for (int p = 0; p < kNP; ++p) {
            const int ke = kend[p];
            if (ke == 0) continue;
            const int P = p0 + p;
            const PartialParams& q = sp[P];
            const float2 e = cmul(base[p], E[P * kEP + lane]);
            const float inv_a = q.inv_attack, ah = static_cast<float>(q.attack + vp.hold), slope_d = (vp.sustain - 1.f) * q.inv_decay;
            const float inv_r = q.inv_release, top = cval[p];
#pragma unroll 2
            for (int k = 0; k < ke; ++k) {
              float2 f;
              if (TENSOR) {
                const uint4 t4 = Ft[((P >> 3) * 8 + (k >> 3)) * 32 + 4 * (k & 7) + (P & 3)];
                const float2 yx = (P & 4) ? f16_join(t4.y, t4.w) : f16_join(t4.x, t4.z);  // (Im, Re) * kScaleB
                f = make_float2(yx.y * (1.f / kScaleB), yx.x * (1.f / kScaleB));
              } else {
                const int pos = P * kR + ((k >> 2) & 1) * 32 + (k >> 3) * 4 + (k & 3);
                f = make_float2(Fx[pos], Fy[pos]);
              }
              const float im = fmaf(e.x, f.y, e.y * f.x);
              float env;
              if (EASED) {
                env = env_value_t<true>(n0 + 32 * k + lane, q, vp);
              } else {
                // LINEAR phases, branch free: attack ramp, hold, decay ramp, sustain; the release ramp from `top`
                const float c = c0 + static_cast<float>(32 * k) + lane_f;
                const float up = c * inv_a;
                const float dn = fmaxf(fmaf(c - ah, slope_d, 1.f), vp.sustain);
                env = fmaxf(fminf(fminf(up, 1.f), dn), 0.f);
                const float j = c - crel;
                if (rel_seen && j >= 1.f) env = top * fmaxf(1.f - j * inv_r, 0.f);
              }
              outbuf[32 * k + lane] = fmaf(env - cv, im, outbuf[32 * k + lane]);
            }
          }
          __syncwarp();
        }
      }
      __syncwarp();

      // ---- emit ----
      if (direct) {
      } else if (a.mix == CA_MIX_NONE) {
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
  }

  // mix modes: one chunk per warp (chunks_per_warp = 1), the voices of the group were accumulated on the warp's bus
  const int64_t chunk = static_cast<int64_t>(blockIdx.x) * WARPS + warp;
  if (a.mix != CA_MIX_NONE && chunk < a.n_chunks && any_voice) {
    const int64_t l0 = chunk * kChunk;
    float* o = a.out + l0 * n_out;
    const int64_t lim = (min(l0 + kChunk, a.n_samples) - l0) * n_out;
    for (int i = lane; i < lim; i += 32) atomicAdd(o + i, bus[i]);
  }
}

// =====================================================================================================================
// tcgen05 form of the sustain regime (per-voice renders of up to 64 partials): D[l][k] = sum_q A[l][q] B[k][q] on the
// 5th-generation tensor cores, accumulator in tensor memory.
//   l = sample inside a block of 128 (M = 128: the 128 TMEM lanes), k = block of the chunk (N = 128 columns): a chunk is
//   16,384 samples; q = (partial, Re/Im), K = 2 * ppad.
//   A[l][2p], A[l][2p+1] = Re, Im of e_p[l] = amp_p exp(i theta_p(n0 + l))      per chunk: seed_p (exact 64-bit phase) times the
//                                                                                thread's own exp(i l w_p), kept in registers
//   B[k][2p], B[k][2p+1] = Im, Re of exp(i 128 k w_p)                           per voice: built once by osc_umma_tables_kernel,
//                                                                                bulk-copied (TMA) into shared memory
// Both operands are FP16 pairs hi + lo (f16_split above; three MMAs per K slice: lo*hi, hi*lo, hi*hi), stored K-major without
// swizzle in the canonical core-matrix layout the shared-memory descriptors describe: 8 rows x 16 bytes contiguous (128 B),
// the next 8 rows 128 B further (SBO), the next 16 bytes of K 128 rows * 16 B = 2048 B further (LBO).
// One CTA (8 warps) per SM: every warp prepares A for chunk c+1 and drains the accumulator of chunk c-1 while the tensor core
// works on chunk c (A and D are double-buffered; one elected thread issues the MMAs and commits them to an mbarrier).
// Lane = sample makes the epilogue stores coalesced: a warp writes 32 consecutive samples per accumulator column.
constexpr int kUM = 128, kUN = 128;
constexpr int kUChunk = kUmmaChunk;                   // samples per chunk
static_assert(kUM * kUN == kUmmaChunk, "chunk of the tcgen05 kernel");
constexpr uint32_t kUOperand = 128 * 128 * 2;         // bytes of one 128 x 128 FP16 operand (ppad = 64)
constexpr uint32_t kUTableBytes = 2 * kUOperand + 64 * 128 * sizeof(float2);  // per voice: B hi, B lo, exp(i l w_p) for l < 128

struct UmmaArgs {
  const PartialParams* partials;
  const VoiceParams* voices;
  unsigned char* tables;
  int64_t n_voices, n_samples, first_sample;
  int32_t ppad;
  int32_t chunks_per_cta;
  int32_t ablate;  // profiling only (CA_OSC_ABLATE): 1 no MMA, 2 no operand preparation, 4 no stores, 8 no seeds
  float* out;
};

__device__ __forceinline__ uint32_t canon_off(int row, int q) {  // byte offset of element (row, q) of a K-major operand
  return static_cast<uint32_t>((q >> 3) * 2048 + (row >> 3) * 128 + (row & 7) * 16 + (q & 7) * 2);
}

__global__ void osc_umma_tables_kernel(const __grid_constant__ UmmaArgs a) {
  const int64_t v = blockIdx.x;
  unsigned char* base = a.tables + v * kUTableBytes;
  const PartialParams* pp = a.partials + v * a.ppad;
  for (int idx = threadIdx.x; idx < 64 * 128; idx += blockDim.x) {
    const int k = idx & 127, p = idx >> 7;
    uint32_t hi = 0, lo = 0;
    float2 e = make_float2(0.f, 0.f);
    if (p < a.ppad) {
      const uint64_t inc = pp[p].inc;
      const float2 f = cis_turns_u64(inc * static_cast<uint64_t>(128 * k));
      f16_split(kScaleB * f.y, kScaleB * f.x, hi, lo);
      e = cis_turns_u64(inc * static_cast<uint64_t>(k));
    }
    const uint32_t off = canon_off(k, 2 * p);
    *reinterpret_cast<uint32_t*>(base + off) = hi;
    *reinterpret_cast<uint32_t*>(base + kUOperand + off) = lo;
    reinterpret_cast<float2*>(base + 2 * kUOperand)[p * 128 + k] = e;
  }
}

__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* bar, uint32_t parity) {  // a lost arrival must not hang the GPU
  for (int spin = 0; spin < (1 << 24); ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
// 1-D bulk copy shared -> global through the TMA engine (the LSU sees none of it), tracked in bulk groups of the issuing thread
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// shared-memory matrix descriptor (tcgen05): start address, leading / stride byte offsets in 16-byte units, version 1, no swizzle
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return static_cast<uint64_t>((smem_addr >> 4) & 0x3fff) | (static_cast<uint64_t>(2048 >> 4) << 16) |
         (static_cast<uint64_t>(128 >> 4) << 32) | (1ull << 46);
}
// instruction descriptor, kind::f16: D = F32 (bits 4-5 = 1), A = B = F16 (0), both K-major (0), N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kUmmaIdesc = (1u << 4) | (static_cast<uint32_t>(kUN >> 3) << 17) | (static_cast<uint32_t>(kUM >> 4) << 24);
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(da), "l"(db), "r"(kUmmaIdesc), "r"(accumulate)
      : "memory");
}

// Two teams of eight warps per CTA, each a pipeline of its own over every other chunk: prepare A (256 threads: four... two per
// row, 32 partials each) -> team barrier -> one thread issues the 24 MMAs and commits them to the team's mbarrier -> all wait
// for it and drain the accumulator (128 lanes x 128 columns, 32 x 64 per warp). A team's waits (barrier, tensor core, stores)
// are the other team's issue slots; nothing is shared between the teams but the B tables and the tensor core itself.
constexpr int kUWarps = 16;

__global__ void __launch_bounds__(kUWarps * 32, 1) osc_umma_kernel(const __grid_constant__ UmmaArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* sA = smem_raw;                          // [2 teams][hi, lo][32 KB]
  unsigned char* sB = smem_raw + 4 * kUOperand;          // [hi, lo][32 KB]
  float2* seed = reinterpret_cast<float2*>(sB + 2 * kUOperand);  // [2 teams][2][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(seed + 256);      // [0]: table copy, [1], [2]: the teams' accumulators
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  __shared__ float s_amp_max;
  __shared__ int s_max_press;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int team = warp >> 3, tt = tid & 255, tw = warp & 7;  // thread and warp inside the team
  const int64_t v = blockIdx.y;
  const VoiceParams vp = a.voices[v];
  if (vp.itp != 0) return;  // eased envelopes: everything is left to the other kernel (uniform over the CTA)
  const PartialParams* pp = a.partials + v * a.ppad;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_amp_max = 0.f;
    s_max_press = 0;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot + static_cast<uint32_t>(team * kUN);

  const unsigned char* tab = a.tables + v * kUTableBytes;
  if (tid == 0) {
    mbar_expect_tx(&bars[0], 2 * kUOperand);
    bulk_g2s(sB, tab, kUOperand, &bars[0]);
    bulk_g2s(sB + kUOperand, tab + kUOperand, kUOperand, &bars[0]);
  }
  // the thread's part of A: sample l of every block, partials 32 h .. 32 h + 31, as packed pairs (Re, Im) of exp(i l w_p)
  const int l = tt & 127, h = tt >> 7;
  c2 eg[32];
  {
    const c2* e128 = reinterpret_cast<const c2*>(tab + 2 * kUOperand);
#pragma unroll
    for (int i = 0; i < 32; ++i) eg[i] = e128[(32 * h + i) * 128 + l];
  }
  if (tid < a.ppad && pp[tid].amp != 0.f) {
    atomicMax(reinterpret_cast<int*>(&s_amp_max), __float_as_int(fabsf(pp[tid].amp)));
    atomicMax(&s_max_press, pp[tid].attack + pp[tid].decay);
  }
  __syncthreads();
  int amp_exp;
  frexpf(s_amp_max * vp.sustain, &amp_exp);
  const float scale_a = exp2f(static_cast<float>(min(100, 15 - amp_exp)));
  const float unscale = 1.f / (scale_a * kScaleB);
  const int64_t press_end = static_cast<int64_t>(vp.key_delay) + s_max_press + vp.hold;
  const bool released = vp.release_at >= 0;
  const int n_slices = a.ppad >> 3;  // K slices of 16 = eight partials
  const int64_t n_end = a.first_sample + a.n_samples;
  unsigned char* myA = sA + (2 * team) * kUOperand;
  float2* myseed = seed + team * 128;

  auto make_seed = [&](int64_t c, int b) {  // threads 0 .. ppad-1 of the team: exact phase of the chunk's first sample
    const PartialParams q = pp[tt];
    const float2 z = cis_turns_u64(q.ph0 + q.inc * static_cast<uint64_t>(a.first_sample + c * kUChunk + 1));
    const float am = q.amp * vp.sustain * scale_a;
    myseed[b * 64 + tt] = make_float2(am * z.x, am * z.y);
  };
  auto team_barrier = [&]() { asm volatile("bar.sync %0, 256;" ::"r"(1 + team) : "memory"); };

  // Epilogue: 128 lanes (samples) x 128 columns (blocks of 128 samples); the warp takes the lanes it may read
  // (32 (warp % 4) ..) and 64 columns, 16 at a time. The 64 KB of a whole chunk are staged in the team's A operand (the tensor
  // core is done with it: same size, and in the order of the output) and handed to the TMA engine as ONE bulk copy: the output
  // never enters the LSU's global queue, which 32 stores per thread kept full (lg_throttle) while the other team's shared-
  // memory traffic queued behind them. The copy must have read the buffer before the team prepares its next A.
  const bool tma_ok = (reinterpret_cast<uintptr_t>(a.out + v * a.n_samples) & 15) == 0 && !(a.ablate & 16);  // uniform
  bool staged = false;
  auto drain = [&](int64_t chunk) {
    const int smp = 32 * (tw & 3) + lane;  // sample inside the block
    const int col0 = 64 * (tw >> 2);
    float* o = a.out + v * a.n_samples + chunk * kUChunk;
    const bool whole = (chunk + 1) * kUChunk <= a.n_samples;
    float* stage = reinterpret_cast<float*>(myA);
#pragma unroll
    for (int part = 0; part < 4; ++part) {
      uint32_t r[16];
      const uint32_t taddr = tmem_d + (static_cast<uint32_t>(32 * (tw & 3)) << 16) + static_cast<uint32_t>(col0 + 16 * part);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
            "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if ((a.ablate & 4) && r[0] != 0x12345678u) continue;
      if (whole && tma_ok) {
#pragma unroll
        for (int j = 0; j < 16; ++j) stage[(col0 + 16 * part + j) * kUM + smp] = __uint_as_float(r[j]) * unscale;
      } else if (whole) {
#pragma unroll
        for (int j = 0; j < 16; ++j) o[(col0 + 16 * part + j) * kUM + smp] = __uint_as_float(r[j]) * unscale;
      } else {  // the last, partial chunk of the render
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (chunk * kUChunk + (col0 + 16 * part + j) * kUM + smp < a.n_samples)
            o[(col0 + 16 * part + j) * kUM + smp] = __uint_as_float(r[j]) * unscale;
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    staged = whole && tma_ok;
    if (staged) {
      fence_proxy_async();
      team_barrier();
      if (tt == 0) {
        bulk_s2g(o, stage, kUChunk * sizeof(float));
        bulk_commit();
      }
    }
  };

  // which chunks of this CTA are rendered here: at least one of the eight 2048-sample chunks inside is steady in the sense of
  // oscbank_kernel (the envelope sits at the sustain level); that kernel, launched afterwards, renders the ones that are not.
  // The teams take them alternately.
  const int64_t n_chunks = (a.n_samples + kUChunk - 1) / kUChunk;
  const int64_t c_begin = static_cast<int64_t>(blockIdx.x) * a.chunks_per_cta;
  const int n_mine = static_cast<int>(min(n_chunks, c_begin + a.chunks_per_cta) - c_begin);  // <= 64
  // steady 2048-sample chunks of the render: m_lo .. m_hi (first sample first_sample + 2048 m at or after the end of the
  // pressed transitions, last sample before the release)
  const int64_t n_small = (a.n_samples + kChunk - 1) / kChunk;
  const int64_t d0 = press_end - a.first_sample;
  const int64_t m_lo = d0 <= 0 ? 0 : (d0 + kChunk - 1) / kChunk;
  int64_t m_hi = n_small - 1;
  if (released && vp.release_at < n_end) {
    const int64_t d1 = vp.release_at - a.first_sample;
    m_hi = d1 < kChunk ? -1 : min(m_hi, d1 / kChunk - 1);
  }
  uint64_t todo = 0;
  int taken = 0;
  for (int i = 0; i < n_mine; ++i) {
    const int64_t s0 = (c_begin + i) * (kUChunk / kChunk);
    const bool any = max(s0, m_lo) <= min(s0 + kUChunk / kChunk - 1, m_hi);
    if (any && (taken++ & 1) == team) todo |= 1ull << i;
  }
  if (todo && tt < a.ppad) make_seed(c_begin + __ffsll(static_cast<long long>(todo)) - 1, 0);
  mbar_wait_or_trap(&bars[0], 0);
  __syncthreads();
  int it = 0;
  while (todo) {  // every condition in this loop is uniform over the team
    const int b = it & 1;
    const int64_t c = c_begin + __ffsll(static_cast<long long>(todo)) - 1;
    todo &= todo - 1;
    if (staged) {  // the bulk copy of the previous chunk must have read the buffer
      if (tt == 0) bulk_wait_read();
      team_barrier();
    }
    if (32 * h < a.ppad && !(a.ablate & 2)) {
      unsigned char* dst_hi = myA + canon_off(l, 64 * h);
      unsigned char* dst_lo = dst_hi + kUOperand;
      const float4* sd = reinterpret_cast<const float4*>(myseed + b * 64 + 32 * h);
#pragma unroll
      for (int gq = 0; gq < 8; ++gq) {  // four partials = eight columns of K = one 16-byte store per half
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; i += 2) {
          const float4 s2 = sd[2 * gq + (i >> 1)];  // two seeds
          // e = seed * exp(i l w): (sx gx - sy gy, sx gy + sy gx) = sx * (gx, gy) + sy * (-gy, gx), packed
          const c2 e0 = fma2(bc(s2.y), cmuli(eg[4 * gq + i]), mul2(bc(s2.x), eg[4 * gq + i]));
          const c2 e1 = fma2(bc(s2.w), cmuli(eg[4 * gq + i + 1]), mul2(bc(s2.z), eg[4 * gq + i + 1]));
          const float2 f0 = upk(e0), f1 = upk(e1);
          f16_split(f0.x, f0.y, hi[i], lo[i]);
          f16_split(f1.x, f1.y, hi[i + 1], lo[i + 1]);
        }
        *reinterpret_cast<uint4*>(dst_hi + gq * 2048) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(dst_lo + gq * 2048) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
    if (todo && tt < a.ppad && !(a.ablate & 8)) make_seed(c_begin + __ffsll(static_cast<long long>(todo)) - 1, b ^ 1);  // published by the barrier
    fence_proxy_async();  // the tensor core reads shared memory through the async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    team_barrier();
    if (tt == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = smem_u32(myA), a_lo = a_hi + kUOperand;
      const uint32_t b_hi = smem_u32(sB), b_lo = b_hi + kUOperand;
      for (int ks = 0; ks < ((a.ablate & 1) ? 0 : n_slices); ++ks) {
        const uint32_t o = static_cast<uint32_t>(ks) * 4096;  // two 16-byte K columns of 128 rows
        umma_f16(tmem_d, umma_desc(a_lo + o), umma_desc(b_hi + o), ks > 0);
        umma_f16(tmem_d, umma_desc(a_hi + o), umma_desc(b_lo + o), 1);
        umma_f16(tmem_d, umma_desc(a_hi + o), umma_desc(b_hi + o), 1);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[1 + team])) : "memory");
    }
    mbar_wait_or_trap(&bars[1 + team], static_cast<uint32_t>(it & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    drain(c);
    ++it;
  }
  if (tt == 0) bulk_wait_all();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"(256));
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

namespace {
// bufs = staging buffers of kChunk floats per warp (1: output; 3: output + stereo bus)
template <int WARPS, int MINB, bool TENSOR>
int launch_osc(ca_ctx* ctx, const OscArgs& a, int bufs, int64_t grid_y, int eased, cudaStream_t stream) {
  const size_t smem = sizeof(PartialParams) * a.ppad + sizeof(float2) * a.ppad * kEP + sizeof(float) * a.ppad * kR * 2 +
                      WARPS * (sizeof(float2) * kNP * 2 + sizeof(float) * kChunk * bufs);
  const int64_t per_cta = static_cast<int64_t>(WARPS) * a.chunks_per_warp;
  dim3 grid(static_cast<unsigned>((a.n_chunks + per_cta - 1) / per_cta), static_cast<unsigned>(grid_y));
  if (eased != 2) {
    CA_CUDA_OK(ctx, cudaFuncSetAttribute(oscbank_kernel<WARPS, MINB, false, TENSOR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
    oscbank_kernel<WARPS, MINB, false, TENSOR><<<grid, WARPS * 32, smem, stream>>>(a);
  }
  if (eased != 0) {
    CA_CUDA_OK(ctx, cudaFuncSetAttribute(oscbank_kernel<WARPS, MINB, true, TENSOR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
    oscbank_kernel<WARPS, MINB, true, TENSOR><<<grid, WARPS * 32, smem, stream>>>(a);
  }
  return CA_OK;
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
  if (ctx->d_osc_umma) cudaFree(ctx->d_osc_umma);
  ctx->d_osc_partials = ctx->d_osc_voices = ctx->d_osc_umma = nullptr;
  ctx->osc_umma_bytes = 0;
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

  if (d.n_voices > 65535) return fail(ctx, CA_ERR_UNSUPPORTED, "ca_oscbank_render: more than 65535 voices per call");
  // CA_OSC_PATH=fp32 keeps the sustain-regime product on the FP32 pipe (FFMA2); the default runs it on the tensor cores (3 x FP16)
  // (CA_OSC_PATH=mma: tensor cores through mma.sync only, without the tcgen05 kernel for the long steady stretches)
  const char* path = getenv("CA_OSC_PATH");
  const bool tensor = !(path && !strcmp(path, "fp32"));
  int rc;
  if (d.mix == CA_MIX_NONE) {
    a.voices_per_group = 1;
    // Per-voice renders of up to 64 partials: the steady stretches go to the tcgen05 kernel in chunks of 16,384 samples (operand
    // tables of 128 KB per voice in device memory, at most 2 GB of them), the 2048-sample chunks with an envelope transition
    // (and eased voices) to the kernel below, which skips what the other one renders.
    const int64_t n_full = (d.n_samples + kUChunk - 1) / kUChunk;
    const size_t table_bytes = static_cast<size_t>(d.n_voices) * kUTableBytes;
    if (tensor && !(path && !strcmp(path, "mma")) && ppad <= 64 && d.n_samples >= kUChunk / 2 && eased != 2 && table_bytes <= (size_t{2} << 30)) {
      if (ctx->osc_umma_bytes < table_bytes) {
        if (ctx->d_osc_umma) cudaFree(ctx->d_osc_umma);
        ctx->d_osc_umma = nullptr;
        ctx->osc_umma_bytes = 0;
        CA_CUDA_OK(ctx, cudaMalloc(&ctx->d_osc_umma, table_bytes));
        ctx->osc_umma_bytes = table_bytes;
      }
      UmmaArgs u{};
      u.partials = pa.partials;
      u.voices = pa.voices;
      u.tables = static_cast<unsigned char*>(ctx->d_osc_umma);
      u.n_voices = d.n_voices;
      u.n_samples = d.n_samples;
      u.first_sample = d.first_sample;
      u.ppad = ppad;
      u.out = d_out;
      u.ablate = getenv("CA_OSC_ABLATE") ? atoi(getenv("CA_OSC_ABLATE")) : 0;
      // one CTA per SM; enough CTAs per voice to fill the GPU twice when there are few voices
      const int64_t per_voice = std::min<int64_t>(n_full, std::max<int64_t>(1, (2 * ctx->sm_count + d.n_voices - 1) / d.n_voices));
      u.chunks_per_cta = static_cast<int32_t>((n_full + per_voice - 1) / per_voice);
      osc_umma_tables_kernel<<<static_cast<unsigned>(d.n_voices), 256, 0, stream>>>(u);
      const size_t usmem = 6 * kUOperand + 256 * sizeof(float2) + 64;
      CA_CUDA_OK(ctx, cudaFuncSetAttribute(osc_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(usmem)));
      dim3 ugrid(static_cast<unsigned>((n_full + u.chunks_per_cta - 1) / u.chunks_per_cta), static_cast<unsigned>(d.n_voices));
      osc_umma_kernel<<<ugrid, kUWarps * 32, usmem, stream>>>(u);
      CA_CUDA_OK(ctx, cudaGetLastError());
      ctx->launches += 2;
      a.umma_skip = 1;
    }
    // shape of the per-voice render: warps per CTA (the per-voice tables are built once per CTA, so a CTA takes chunks_per_warp
    // chunks per warp). Measured on config 3 (profiles/oscbank_shapes_r02.txt); CA_OSC_SHAPE / CA_OSC_CPW are the A/B switches.
    const char* shape = getenv("CA_OSC_SHAPE");
    const char* cpw = getenv("CA_OSC_CPW");
    const int sh = shape ? atoi(shape) : 6;
    a.chunks_per_warp = cpw ? max(1, atoi(cpw)) : 8;
    if (tensor) {
      if (sh == 8) rc = launch_osc<8, 1, true>(ctx, a, 1, d.n_voices, eased, stream);
      else if (sh == 6 && ppad <= 64) rc = launch_osc<6, 2, true>(ctx, a, 1, d.n_voices, eased, stream);
      else if (sh == 16 && ppad <= 64) rc = launch_osc<16, 1, true>(ctx, a, 1, d.n_voices, eased, stream);
      else rc = launch_osc<12, 1, true>(ctx, a, 1, d.n_voices, eased, stream);
    } else {
      if (sh == 16) rc = launch_osc<16, 1, false>(ctx, a, 1, d.n_voices, eased, stream);
      else if (sh == 12) rc = launch_osc<12, 1, false>(ctx, a, 1, d.n_voices, eased, stream);
      else rc = launch_osc<6, 2, false>(ctx, a, 1, d.n_voices, eased, stream);
    }
  } else {
    constexpr int WARPS = 4;
    const int n_out = d.mix == CA_MIX_STEREO ? 2 : 1;
    CA_CUDA_OK(ctx, cudaMemsetAsync(d_out, 0, sizeof(float) * d.n_samples * n_out, stream));
    const int64_t tiles = (a.n_chunks + WARPS - 1) / WARPS;
    // enough voice groups to fill the GPU a few times over, few enough that the final atomics stay negligible
    int64_t groups = (static_cast<int64_t>(ctx->sm_count) * 8 + tiles - 1) / tiles;
    if (groups < 1) groups = 1;
    if (groups > d.n_voices) groups = d.n_voices;
    a.voices_per_group = static_cast<int32_t>((d.n_voices + groups - 1) / groups);
    a.chunks_per_warp = 1;
    groups = (d.n_voices + a.voices_per_group - 1) / a.voices_per_group;
    rc = tensor ? launch_osc<WARPS, 1, true>(ctx, a, 3, groups, eased, stream) : launch_osc<WARPS, 2, false>(ctx, a, 3, groups, eased, stream);
  }
  if (rc != CA_OK) return rc;
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
