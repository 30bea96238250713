// stft64.cu — the periodic-FFT analysis in FP64, peaks only: the form the analysis -> resynthesis round trip uses.
//
// Why it exists. The reference analyses in double (PeriodicFFT<SqMagFftOperation<Window::Gaussian, double>>,
// source/rt.resynth.lib.cpp:1010) and turns every frame's peak frequencies into note frequencies that are rounded to
// float (Event::noteOn.frequency, source/rt.resynth.lib.cpp:322-352). A float32 spectrum carries ~1e-7 of relative noise;
// that moves a peak by ~1e-7 bins, which now and then flips the rounded note frequency by one float ulp (6e-8 relative),
// and the resynthesised phase integrates those flips over the whole stream: a random walk that reaches 1e-4 after 60 s
// (profiles/roundtrip_time_r01.txt). With the spectrum in double the peaks agree with the reference's to ~1e-13 bins, a
// flip happens about once per million note-frames, and the 60 s round trip stays within 1e-5 of the double oracle.
// The float32 kernels (stft.cu, rfftw.cu) remain the analysis product (spectra and peak lists as float32, BASELINE
// configs 1, 2, 5); this kernel feeds the pitch pipeline only.
//
// Restates SqMagFftOperation<Window,double>::operator() (source/rt.resynth.lib.periodicfft.cpp:302-313) +
// extractLocalMaxFreqsMags (call site source/rt.resynth.lib.cpp:1592-1595) per frame, arithmetic per oracle/SPEC.md S1-S6:
//   one CTA per frame; the N real samples are packed as M = N/2 complex points in shared memory (double2, padded),
//   in-place radix-4 decimation in time (bit-reversed on the way in), real-FFT untangling, |X/W|^2 in double,
//   strict local maxima with 10 log10 >= min_db, parabolic interpolation on dB, a block prefix sum orders the peaks.
#include <math.h>

#include <vector>

#include "common.cuh"
#include "ctx.hpp"

namespace ca {

namespace {

struct Stft64Args {
  const float* x;
  int64_t stream_stride;
  int64_t frames_per_stream;
  int64_t total_frames;
  int32_t hop, window;
  const double* win;   // [N] normalised window (mean 1), zero beyond `window`
  const double2* tw;   // [N/2 + 1] exp(-2 pi i k / N): untangling
  const double2* tab;  // per-pass twiddle tables of the N/2-point complex FFT (fft_pass64)
  int32_t* peak_bin;   // [frames][max_peaks]
  double* peak_off;    // sub-bin offset p of the parabola's vertex
  double* peak_db;
  int32_t* peak_count; // peaks found (may exceed max_peaks)
  int32_t max_peaks;
  double min_db;       // -inf: none
  double thr_lo;       // a little below 10^(min_db/10): bins under it cannot reach min_db
  double inv_w;        // 1 / window (SPEC S4)
};

constexpr int kThreads64 = 256;

// shared-memory slot of complex element i: one slot of padding per 8 keeps the strided accesses of the first pass and the
// unit-stride accesses of the later ones spread over the 16-byte bank groups
__device__ __forceinline__ int phys64(int i) { return i + (i >> 3); }

__device__ __forceinline__ double2 cmul64(double2 a, double2 b) {
  return make_double2(fma(-a.y, b.y, a.x * b.x), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cadd64(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub64(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// Passes of the M-point decimation-in-time FFT: a pass fuses K consecutive radix-2 stages (half lengths h, 2h, .., 2^(K-1) h)
// in registers, so the data crosses shared memory LOGM/3 (+1) times instead of LOGM times. The first pass takes the
// LOGM % 3 left-over stages. Twiddles W_{2 h 2^s}^j, j < h, come from one table per pass and stage, contiguous in j
// (stft64_plan), so that a warp's twiddle loads are as compact as its data accesses.
constexpr int first_k64(int logm) { return logm % 3 == 0 ? 3 : logm % 3; }
constexpr int n_passes64(int logm) { return (logm + 2) / 3; }
constexpr int pass_k64(int logm, int p) { return p == 0 ? first_k64(logm) : 3; }
constexpr int pass_logh64(int logm, int p) { return p == 0 ? 0 : first_k64(logm) + 3 * (p - 1); }
constexpr int pass_off64(int logm, int p) {  // first table entry of pass p
  int off = 0;
  for (int q = 0; q < p; ++q) off += pass_k64(logm, q) << pass_logh64(logm, q);
  return off;
}
constexpr int table_len64(int logm) { return pass_off64(logm, n_passes64(logm)); }

// K fused DIT stages on the 2^K values v[e] = z[b + j + e h]. Stage s pairs (e, e + 2^s) with the twiddle
// W_{2 h 2^s}^(j + (e mod 2^s) h) = tw_s * W_{2^(s+1)}^(e mod 2^s): a table twiddle times a power of W_8.
template <int K>
__device__ __forceinline__ void fused_stages64(double2 (&v)[1 << K], const double2 (&tw)[K]) {
  constexpr double kR = 0.70710678118654752440;
#pragma unroll
  for (int s = 0; s < K; ++s) {
#pragma unroll
    for (int m = 0; m < (1 << s); ++m) {  // m = e mod 2^s: twiddle tw[s] * W_{2^(s+1)}^m
      double2 w = tw[s];
      const int m8 = m << (2 - s);  // the power of W_8 (s <= 2)
      if (m8 == 2) w = make_double2(w.y, -w.x);                             // * (-i)
      else if (m8 == 1) w = make_double2(kR * (w.x + w.y), kR * (w.y - w.x));   // * (1 - i)/sqrt 2
      else if (m8 == 3) w = make_double2(kR * (w.y - w.x), -kR * (w.x + w.y));  // * (-1 - i)/sqrt 2
#pragma unroll
      for (int g = 0; g < (1 << (K - 1 - s)); ++g) {
        const int lo = (g << (s + 1)) + m, hi = lo + (1 << s);
        const double2 t = cmul64(w, v[hi]);
        v[hi] = csub64(v[lo], t);
        v[lo] = cadd64(v[lo], t);
      }
    }
  }
}

template <int LOGM, int P>
__device__ __forceinline__ void fft_pass64(double2* __restrict__ z, const double2* __restrict__ tab, int tid) {
  constexpr int M = 1 << LOGM, K = pass_k64(LOGM, P), LH = pass_logh64(LOGM, P), H = 1 << LH, R = 1 << K;
  constexpr int OFF = pass_off64(LOGM, P);
  for (int q = tid; q < M / R; q += kThreads64) {
    const int j = q & (H - 1);
    const int b = (q >> LH) << (LH + K);
    double2 tw[K];
#pragma unroll
    for (int s = 0; s < K; ++s) tw[s] = __ldg(tab + OFF + s * H + j);
    double2 v[R];
#pragma unroll
    for (int e = 0; e < R; ++e) v[e] = z[phys64(b + j + e * H)];
    fused_stages64<K>(v, tw);
#pragma unroll
    for (int e = 0; e < R; ++e) z[phys64(b + j + e * H)] = v[e];
  }
  __syncthreads();
  if constexpr (P + 1 < n_passes64(LOGM)) fft_pass64<LOGM, P + 1>(z, tab, tid);
}

template <int LOGM>
__global__ void __launch_bounds__(kThreads64) stft64_peaks_kernel(const __grid_constant__ Stft64Args a) {
  constexpr int M = 1 << LOGM;
  constexpr int B = (M + kThreads64 - 1) / kThreads64;  // bins per thread of the peak scan
  constexpr int LB = ilog2(B);
  extern __shared__ __align__(16) unsigned char smem64[];
  double2* z = reinterpret_cast<double2*>(smem64);          // [M + M/8]
  double* sq = reinterpret_cast<double*>(z + (M + M / 8));   // |X|^2 of bins 0 .. M, one slot of padding per B bins
  auto physq = [](int k) { return B > 1 ? k + (k >> LB) : k; };
  __shared__ int warp_tot[kThreads64 / 32];

  const int tid = threadIdx.x;
  const int64_t f = blockIdx.x;
  const int64_t stream = f / a.frames_per_stream;
  const int64_t fr = f - stream * a.frames_per_stream;
  const float* __restrict__ xs = a.x + stream * a.stream_stride + fr * a.hop;

  // ---- windowed samples, packed two per complex point, stored bit-reversed (decimation in time) ----
  // lane = b + 8a takes point i = a + 4c + b*(M/8): the four a of a quarter warp read one 32-byte sector, and the eight b land
  // in the low three bits of the bit-reversed position, i.e. in eight different 16-byte bank groups
  {
    const int la = (tid & 31) >> 3, lb = tid & 7;
    constexpr int kIter = M / 32 / (kThreads64 / 32);  // points per thread; every load is issued before the first use
    float2 sv[kIter];
#pragma unroll
    for (int it = 0; it < kIter; ++it) {
      const int n = 2 * (la + 4 * ((tid >> 5) + (kThreads64 / 32) * it) + lb * (M / 8));
      // 8-byte aligned: even hop and stream stride (stft64_supported); window is even, so n + 1 < window as well
      sv[it] = n < a.window ? *reinterpret_cast<const float2*>(xs + n) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int it = 0; it < kIter; ++it) {
      const int i = la + 4 * ((tid >> 5) + (kThreads64 / 32) * it) + lb * (M / 8);
      const double2 w = __ldg(reinterpret_cast<const double2*>(a.win) + i);  // zero beyond the window
      z[phys64(static_cast<int>(__brev(static_cast<unsigned>(i)) >> (32 - LOGM)))] =
          make_double2(static_cast<double>(sv[it].x) * w.x, static_cast<double>(sv[it].y) * w.y);
    }
  }
  __syncthreads();

  // ---- M-point complex FFT in place ----
  fft_pass64<LOGM, 0>(z, a.tab, tid);

  // ---- untangle the packed transform into X[k], k = 0 .. M, and form |X / W|^2 (SPEC S4, S5) ----
  for (int k = tid; k <= M / 2; k += kThreads64) {
    const int km = M - k;
    const double2 zk = z[phys64(k & (M - 1))];
    const double2 zc = z[phys64(km & (M - 1))];
    // bins k and M-k share the two inputs
    {
      const double2 zm = make_double2(zc.x, -zc.y);
      const double2 even = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y + zm.y));
      const double2 d = csub64(zk, zm);
      const double2 odd = make_double2(0.5 * d.y, -0.5 * d.x);  // (zk - zm) * (-i/2)
      const double2 t = cmul64(__ldg(a.tw + k), odd);
      const double re = (even.x + t.x) * a.inv_w, im = (even.y + t.y) * a.inv_w;
      sq[physq(k)] = re * re + im * im;
    }
    if (km != k) {
      const double2 zm = make_double2(zk.x, -zk.y);
      const double2 even = make_double2(0.5 * (zc.x + zm.x), 0.5 * (zc.y + zm.y));
      const double2 d = csub64(zc, zm);
      const double2 odd = make_double2(0.5 * d.y, -0.5 * d.x);
      const double2 t = cmul64(__ldg(a.tw + km), odd);
      const double re = (even.x + t.x) * a.inv_w, im = (even.y + t.y) * a.inv_w;
      sq[physq(km)] = re * re + im * im;
    }
  }
  __syncthreads();

  // ---- peaks: thread owns B consecutive bins; candidates are bins 1 .. M-1 (SPEC S6) ----
  const int k0 = tid * B;
  auto to_db = [](double s) { return 10.0 * log10(fmax(s, 1e-300)); };
  int cnt = 0;
  unsigned mask = 0;  // B <= 32
  if (k0 < M) {
    double l = k0 > 0 ? sq[physq(k0 - 1)] : INFINITY;  // bin 0 is never a peak
    double c = sq[physq(k0)];
#pragma unroll
    for (int j = 0; j < B; ++j) {
      const double r = sq[physq(k0 + j + 1)];
      // the threshold is defined on dB (SPEC S6); the comparison on |X|^2 with a margin keeps log10 off the bins far below it
      if (c > l && c > r && c >= a.thr_lo && to_db(c) >= a.min_db) {
        mask |= 1u << j;
        ++cnt;
      }
      l = c;
      c = r;
    }
  }
  int incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, incl, d);
    if ((tid & 31) >= d) incl += o;
  }
  if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
  __syncthreads();
  int base = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kThreads64 / 32; ++w) {
    const int t = warp_tot[w];
    if (w < (tid >> 5)) base += t;
    total += t;
  }
  if (tid == 0) a.peak_count[f] = total;
  int slot = base + incl - cnt;
  const int64_t out_base = f * static_cast<int64_t>(a.max_peaks);
  while (mask && slot < a.max_peaks) {
    const int j = __ffs(static_cast<int>(mask)) - 1;
    mask &= mask - 1;
    const int k = k0 + j;
    const double alpha = to_db(sq[physq(k - 1)]), beta = to_db(sq[physq(k)]), gamma = to_db(sq[physq(k + 1)]);
    const double denom = alpha - 2.0 * beta + gamma;
    const double p = denom != 0.0 ? 0.5 * (alpha - gamma) / denom : 0.0;
    a.peak_bin[out_base + slot] = k;
    a.peak_off[out_base + slot] = p;
    a.peak_db[out_base + slot] = beta - 0.25 * (alpha - gamma) * p;
    ++slot;
  }
}

template <int LOGM>
size_t smem_bytes64() {
  constexpr int M = 1 << LOGM;
  constexpr int B = (M + kThreads64 - 1) / kThreads64;
  return sizeof(double2) * (M + M / 8) + sizeof(double) * (M + 1 + (M + 1) / B + 1);
}

// per-pass twiddle tables of the M-point FFT (layout: fft_pass64)
template <int LOGM>
void build_table64(std::vector<double2>& tab) {
  tab.assign(table_len64(LOGM), make_double2(1.0, 0.0));
  const long double pi = static_cast<long double>(3.14159265358979323846);  // the oracle's kPi (a double constant)
  int off = 0;
  for (int p = 0; p < n_passes64(LOGM); ++p) {
    const int K = pass_k64(LOGM, p), H = 1 << pass_logh64(LOGM, p);
    for (int s = 0; s < K; ++s)
      for (int j = 0; j < H; ++j) {
        const long double ang = -2.0L * pi * j / (2.0L * H * (1 << s));
        tab[off + s * H + j] = make_double2(static_cast<double>(cosl(ang)), static_cast<double>(sinl(ang)));
      }
    off += K * H;
  }
}

template <int LOGM>
int launch64(ca_ctx* ctx, const Stft64Args& a, cudaStream_t stream) {
  static bool attr_set[16] = {};
  const size_t smem = smem_bytes64<LOGM>();
  if (!attr_set[ctx->device & 15]) {
    CA_CUDA_OK(ctx, cudaFuncSetAttribute(stft64_peaks_kernel<LOGM>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_set[ctx->device & 15] = true;
  }
  stft64_peaks_kernel<LOGM><<<static_cast<unsigned>(a.total_frames), kThreads64, smem, stream>>>(a);
  CA_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches += 1;
  return CA_OK;
}

}  // namespace

bool stft64_supported(const StftPlan& plan) {
  return plan.planned && plan.n_fft >= 1024 && plan.n_fft <= 16384 && plan.prm.hop % 2 == 0;
}

// window (SPEC S2, S3, exactly as the double oracle rounds it) and twiddles, built in long double on the host
int stft64_plan(ca_ctx* ctx, StftPlan& plan) {
  if (!plan.planned) return fail(ctx, CA_ERR_NOT_PLANNED, "stft64_plan: plan the float32 analysis first");
  const int N = plan.n_fft, W = plan.prm.window, H = W / 2;
  std::vector<double> hw(H);
  long double sum = 0;
  for (int j = 0; j < H; ++j) {
    long double v = 1;
    if (plan.prm.window_type == CA_WINDOW_GAUSSIAN) {
      const long double u = static_cast<long double>(static_cast<double>(plan.prm.sigma_factor)) * (j + 0.5L) / H;
      v = expl(-0.5L * u * u);
    }
    hw[j] = static_cast<double>(v);
    sum += v;
  }
  const long double scale = H / sum;
  for (auto& v : hw) v = static_cast<double>(v * scale);
  std::vector<double> full(N, 0.0);
  for (int n = 0; n < W; ++n) full[n] = n < H ? hw[H - 1 - n] : hw[n - H];
  std::vector<double2> tw(N / 2 + 1), tab;
  const long double pi = static_cast<long double>(3.14159265358979323846);  // the oracle's kPi (a double constant)
  for (int k = 0; k <= N / 2; ++k) {
    const long double ang = -2.0L * pi * k / N;
    tw[k] = make_double2(static_cast<double>(cosl(ang)), static_cast<double>(sinl(ang)));
  }
  switch (N) {
    case 1024: build_table64<9>(tab); break;
    case 2048: build_table64<10>(tab); break;
    case 4096: build_table64<11>(tab); break;
    case 8192: build_table64<12>(tab); break;
    default: build_table64<13>(tab); break;
  }
  const size_t tw_total = tw.size() + tab.size();
  tw.insert(tw.end(), tab.begin(), tab.end());  // one allocation: untangling twiddles, then the pass tables
  if (plan.n64 != N) {
    if (plan.d_window64) cudaFree(plan.d_window64);
    if (plan.d_tw64) cudaFree(plan.d_tw64);
    plan.d_window64 = nullptr;
    plan.d_tw64 = nullptr;
    plan.n64 = 0;
    CA_CUDA_OK(ctx, cudaMalloc(&plan.d_window64, sizeof(double) * N));
    CA_CUDA_OK(ctx, cudaMalloc(&plan.d_tw64, sizeof(double2) * tw_total));
    plan.n64 = N;
  }
  CA_CUDA_OK(ctx, cudaMemcpy(plan.d_window64, full.data(), sizeof(double) * N, cudaMemcpyHostToDevice));
  CA_CUDA_OK(ctx, cudaMemcpy(plan.d_tw64, tw.data(), sizeof(double2) * tw_total, cudaMemcpyHostToDevice));
  return CA_OK;
}

int stft64_peaks(ca_ctx* ctx, const StftPlan& plan, const float* d_x, int64_t n_samples, int64_t n_streams, int64_t stream_stride,
                 int32_t* d_peak_bin, double* d_peak_off, double* d_peak_db, int32_t* d_peak_count, cudaStream_t stream) {
  if (!stft64_supported(plan) || plan.n64 != plan.n_fft)
    return fail(ctx, CA_ERR_NOT_PLANNED, "stft64_peaks: not planned for this fft length");
  if (stream_stride % 2 != 0 || reinterpret_cast<uintptr_t>(d_x) % 8 != 0)
    return fail(ctx, CA_ERR_UNSUPPORTED, "stft64_peaks: streams must start on 8-byte boundaries");
  const int64_t frames = ca_stft_num_frames(n_samples, plan.prm.window, plan.prm.hop);
  const int64_t total = frames * n_streams;
  if (total == 0) return CA_OK;
  if (total >= (1ll << 31)) return fail(ctx, CA_ERR_UNSUPPORTED, "stft64_peaks: more than 2^31 frames per call");
  Stft64Args a{};
  a.x = d_x;
  a.stream_stride = stream_stride;
  a.frames_per_stream = frames;
  a.total_frames = total;
  a.hop = plan.prm.hop;
  a.window = plan.prm.window;
  a.win = plan.d_window64;
  a.tw = plan.d_tw64;
  a.tab = plan.d_tw64 + (plan.n_fft / 2 + 1);
  a.peak_bin = d_peak_bin;
  a.peak_off = d_peak_off;
  a.peak_db = d_peak_db;
  a.peak_count = d_peak_count;
  a.max_peaks = plan.prm.max_peaks;
  a.min_db = static_cast<double>(plan.prm.min_db);
  a.inv_w = 1.0 / plan.prm.window;
  a.thr_lo = isinf(a.min_db) && a.min_db < 0 ? 0.0 : pow(10.0, a.min_db / 10.0) * (1.0 - 1e-9);
  switch (plan.n_fft) {
    case 1024: return launch64<9>(ctx, a, stream);
    case 2048: return launch64<10>(ctx, a, stream);
    case 4096: return launch64<11>(ctx, a, stream);
    case 8192: return launch64<12>(ctx, a, stream);
    case 16384: return launch64<13>(ctx, a, stream);
  }
  return fail(ctx, CA_ERR_UNSUPPORTED, "stft64_peaks: fft length");
}

}  // namespace ca
