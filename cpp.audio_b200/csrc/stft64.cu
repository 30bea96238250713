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
  const double2* tw;   // [N/2 + 1] exp(-2 pi i k / N)
  int32_t* peak_bin;   // [frames][max_peaks]
  double* peak_off;    // sub-bin offset p of the parabola's vertex
  double* peak_db;
  int32_t* peak_count; // peaks found (may exceed max_peaks)
  int32_t max_peaks;
  double min_db;       // -inf: none
  double inv_w;        // 1 / window (SPEC S4)
};

constexpr int kThreads64 = 256;

// shared-memory slot of complex element i: one slot of padding per 8 keeps the stride-4 accesses of the first radix-4 pass
// and the unit-stride accesses of the later ones spread over the 16-byte bank groups
__device__ __forceinline__ int phys64(int i) { return i + (i >> 3); }

__device__ __forceinline__ double2 cmul64(double2 a, double2 b) {
  return make_double2(fma(-a.y, b.y, a.x * b.x), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cadd64(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub64(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

template <int LOGM>
__global__ void __launch_bounds__(kThreads64) stft64_peaks_kernel(const __grid_constant__ Stft64Args a) {
  constexpr int M = 1 << LOGM, N = 2 * M;
  extern __shared__ __align__(16) unsigned char smem64[];
  double2* z = reinterpret_cast<double2*>(smem64);               // [M + M/8]
  double* sq = reinterpret_cast<double*>(z + (M + M / 8));        // [M + 1]
  __shared__ int warp_tot[kThreads64 / 32];

  const int tid = threadIdx.x;
  const int64_t f = blockIdx.x;
  const int64_t stream = f / a.frames_per_stream;
  const int64_t fr = f - stream * a.frames_per_stream;
  const float* __restrict__ xs = a.x + stream * a.stream_stride + fr * a.hop;

  // ---- windowed samples, packed two per complex point, stored bit-reversed (decimation in time) ----
  for (int i = tid; i < M; i += kThreads64) {
    const int n = 2 * i;
    double2 v = make_double2(0.0, 0.0);
    if (n < a.window) {  // window is even: n + 1 < window as well
      const float2 s = *reinterpret_cast<const float2*>(xs + n);  // 8-byte aligned: even hop and stream stride (stft64_supported)
      v.x = static_cast<double>(s.x) * __ldg(a.win + n);
      v.y = static_cast<double>(s.y) * __ldg(a.win + n + 1);
    }
    z[phys64(static_cast<int>(__brev(static_cast<unsigned>(i)) >> (32 - LOGM)))] = v;
  }
  __syncthreads();

  // ---- M-point complex FFT in place. Stage with half length h uses W_{2h}^j = tw[j * M / h] ----
  int h = 1;
  if (LOGM & 1) {  // odd number of stages: one radix-2 stage first (twiddle 1)
    for (int q = tid; q < M / 2; q += kThreads64) {
      const int i0 = phys64(2 * q), i1 = phys64(2 * q + 1);
      const double2 u = z[i0], v = z[i1];
      z[i0] = cadd64(u, v);
      z[i1] = csub64(u, v);
    }
    __syncthreads();
    h = 2;
  }
  for (; h < M; h <<= 2) {  // radix-4: stages of half lengths h and 2h fused
    const int lh = 31 - __clz(h);
    for (int q = tid; q < M / 4; q += kThreads64) {
      const int j = q & (h - 1);
      const int b = (q >> lh) << (lh + 2);
      const int i0 = phys64(b + j), i1 = phys64(b + j + h), i2 = phys64(b + j + 2 * h), i3 = phys64(b + j + 3 * h);
      const double2 w1 = __ldg(a.tw + (static_cast<int64_t>(j) << (LOGM - lh)));       // W_{2h}^j
      const double2 w2 = __ldg(a.tw + (static_cast<int64_t>(j) << (LOGM - lh - 1)));   // W_{4h}^j
      const double2 w3 = make_double2(w2.y, -w2.x);                                    // W_{4h}^{j+h} = -i W_{4h}^j
      const double2 a0 = z[i0], a1 = z[i1], a2 = z[i2], a3 = z[i3];
      const double2 t1 = cmul64(w1, a1), t3 = cmul64(w1, a3);
      const double2 b0 = cadd64(a0, t1), b1 = csub64(a0, t1), b2 = cadd64(a2, t3), b3 = csub64(a2, t3);
      const double2 u = cmul64(w2, b2), v = cmul64(w3, b3);
      z[i0] = cadd64(b0, u);
      z[i2] = csub64(b0, u);
      z[i1] = cadd64(b1, v);
      z[i3] = csub64(b1, v);
    }
    __syncthreads();
  }

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
      sq[k] = re * re + im * im;
    }
    if (km != k) {
      const double2 zm = make_double2(zk.x, -zk.y);
      const double2 even = make_double2(0.5 * (zc.x + zm.x), 0.5 * (zc.y + zm.y));
      const double2 d = csub64(zc, zm);
      const double2 odd = make_double2(0.5 * d.y, -0.5 * d.x);
      const double2 t = cmul64(__ldg(a.tw + km), odd);
      const double re = (even.x + t.x) * a.inv_w, im = (even.y + t.y) * a.inv_w;
      sq[km] = re * re + im * im;
    }
  }
  __syncthreads();

  // ---- peaks: thread owns B consecutive bins; candidates are bins 1 .. M-1 (SPEC S6) ----
  constexpr int B = (M + kThreads64 - 1) / kThreads64;
  const int k0 = tid * B;
  auto to_db = [](double s) { return 10.0 * log10(fmax(s, 1e-300)); };
  int cnt = 0;
  unsigned mask = 0;  // B <= 32
  if (k0 < M) {
    double l = k0 > 0 ? sq[k0 - 1] : INFINITY;  // bin 0 is never a peak
    double c = sq[k0];
#pragma unroll 4
    for (int j = 0; j < B; ++j) {
      const double r = sq[k0 + j + 1];
      if (c > l && c > r && to_db(c) >= a.min_db) {
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
    const double alpha = to_db(sq[k - 1]), beta = to_db(sq[k]), gamma = to_db(sq[k + 1]);
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
  return sizeof(double2) * (M + M / 8) + sizeof(double) * (M + 1);
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
  std::vector<double2> tw(N / 2 + 1);
  const long double pi = 3.14159265358979323846L;  // the oracle's kPi is this double constant
  for (int k = 0; k <= N / 2; ++k) {
    const long double ang = -2.0L * static_cast<long double>(static_cast<double>(pi)) * k / N;
    tw[k] = make_double2(static_cast<double>(cosl(ang)), static_cast<double>(sinl(ang)));
  }
  if (plan.n64 != N) {
    if (plan.d_window64) cudaFree(plan.d_window64);
    if (plan.d_tw64) cudaFree(plan.d_tw64);
    plan.d_window64 = nullptr;
    plan.d_tw64 = nullptr;
    plan.n64 = 0;
    CA_CUDA_OK(ctx, cudaMalloc(&plan.d_window64, sizeof(double) * N));
    CA_CUDA_OK(ctx, cudaMalloc(&plan.d_tw64, sizeof(double2) * (N / 2 + 1)));
    plan.n64 = N;
  }
  CA_CUDA_OK(ctx, cudaMemcpy(plan.d_window64, full.data(), sizeof(double) * N, cudaMemcpyHostToDevice));
  CA_CUDA_OK(ctx, cudaMemcpy(plan.d_tw64, tw.data(), sizeof(double2) * (N / 2 + 1), cudaMemcpyHostToDevice));
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
  a.peak_bin = d_peak_bin;
  a.peak_off = d_peak_off;
  a.peak_db = d_peak_db;
  a.peak_count = d_peak_count;
  a.max_peaks = plan.prm.max_peaks;
  a.min_db = static_cast<double>(plan.prm.min_db);
  a.inv_w = 1.0 / plan.prm.window;
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
