// vocoder.cu — the FFT vocoder of rt.resynth on the device.
//
// Reference: source/rt.resynth.lib.vocoder.cpp. FFTModulator (:100-257) runs PeriodicFFT<SqMagFftOperation<Gaussian>> on the
// modulator signal (window modulator_window_size_seconds, hop = stride) and keeps, per frame, the amplitude of every band:
// sqrt of the sum of |X|^2 over the bins whose frequency lies in (freqs[b], freqs[b+1]] (:131-160). FFTCarrier (:396-551)
// runs PeriodicFFT<FftOperation> on the carrier (window 2*stride, hop stride, rectangular, zero padded to a power of two),
// scales every bin by the latest amplitude of its band (modulate_bands), transforms back, scales by 1/fft_length, and
// emits  new[i]*xf.new(i+1) + old[i+stride]*xf.old(i+1)  for the next `stride` samples (:501-531). Both PeriodicFFTs are
// fed sample by sample (modulator first, :786-797), so carrier frame j (fired by sample 2S-1+jS) sees modulator frame
// floor((2S + jS - Wm)/S), and its samples leave with a latency of 2S-1.
//
// Here whole streams are known, so every stage is one batched launch over all frames of all streams:
//   K2 (stft.cu / rfftw.cu)  |X|^2 of the modulator frames          band_amps_kernel  sums per band -> amplitudes
//   voc_frames_kernel        carrier frames, zero padded            K1 forward (rfftw.cu)
//   voc_modulate_kernel      bins *= amplitude of (frame, band)     K1 inverse (rfftw.cu), scale 1/N
//   voc_mix_kernel           cross-fade of consecutive frames + the dry signals
// Unpinned pieces (cpp.algorithms is absent): modulate_bands (oracle/SPEC.md S12), EqualGainXFade (S13).
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "ctx.hpp"

namespace ca {

namespace {

// one warp per modulator frame: amplitude of band b = sqrt(sum of sqmag over bins [begin[b], begin[b+1]))
__global__ void __launch_bounds__(128) band_amps_kernel(const float* __restrict__ sqmag, int64_t row_stride, int64_t rows,
                                                        const int32_t* __restrict__ band_begin, int n_bands, float* __restrict__ amps) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* row = sqmag + r * row_stride;
  for (int b = 0; b < n_bands; ++b) {
    const int lo = band_begin[b], hi = band_begin[b + 1];
    float acc = 0.f;
    for (int i = lo + lane; i < hi; i += 32) acc += row[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) amps[r * n_bands + b] = sqrtf(acc);
  }
}

// carrier frame j of stream s: samples [jS, jS + 2S), zero padded to n_fft
__global__ void __launch_bounds__(256) voc_frames_kernel(const float* __restrict__ x, int64_t n_samples, int64_t frames_per_stream,
                                                         int stride, int n_fft, float* __restrict__ out) {
  const int64_t f = blockIdx.x;
  const int64_t s = f / frames_per_stream, j = f - s * frames_per_stream;
  const float* xs = x + s * n_samples + j * stride;
  float* o = out + f * n_fft;
  for (int i = threadIdx.x; i < n_fft; i += 256) o[i] = i < 2 * stride ? __ldg(xs + i) : 0.f;
}

// bins[f][i] *= amplitude of the band of bin i in the modulator frame the carrier frame sees (0 outside the bands / before the
// first modulator frame)
__global__ void __launch_bounds__(256) voc_modulate_kernel(float2* __restrict__ bins, int n_bins, int64_t frames_c, int64_t frames_m,
                                                           int first_m /* modulator frame seen by carrier frame 0 (may be < 0) */,
                                                           const int8_t* __restrict__ band_of_bin, const float* __restrict__ amps,
                                                           int n_bands) {
  const int64_t f = blockIdx.x;
  const int64_t s = f / frames_c, j = f - s * frames_c;
  const int64_t m = j + first_m;
  const float* a = (m >= 0 && m < frames_m) ? amps + (s * frames_m + m) * n_bands : nullptr;
  float2* row = bins + f * n_bins;
  for (int i = threadIdx.x; i < n_bins; i += 256) {
    const int b = band_of_bin[i];
    const float g = (a && b >= 0) ? a[b] : 0.f;
    float2 v = row[i];
    v.x *= g;
    v.y *= g;
    row[i] = v;
  }
}

// out[n] = vm*xm[n] + vc*xc[n] + vv*(new_j[i]*w(i+1) + new_{j-1}[i+S]*(1-w(i+1))), n = 2S-1 + jS + i
__global__ void __launch_bounds__(256) voc_mix_kernel(const float* __restrict__ xm, const float* __restrict__ xc,
                                                      const float* __restrict__ sig, int64_t n_samples, int64_t n_streams,
                                                      int64_t frames_c, int stride, int n_fft, float vm, float vc, float vv,
                                                      float* __restrict__ out) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= n_samples * n_streams) return;
  const int64_t s = idx / n_samples, n = idx - s * n_samples;
  float vocoded = 0.f;
  const int64_t q = n - (2 * static_cast<int64_t>(stride) - 1);
  if (q >= 0) {
    const int64_t j = q / stride;
    const int i = static_cast<int>(q - j * stride);
    if (j < frames_c) {
      const float* cur = sig + (s * frames_c + j) * n_fft;
      const float w = 0.5f * (1.f - cospif(static_cast<float>(i + 1) / static_cast<float>(stride)));  // SPEC S13
      const float old = j > 0 ? cur[i + stride - n_fft] : 0.f;                                      // previous frame, second half
      vocoded = cur[i] * w + old * (1.f - w);
    }
  }
  out[idx] = vm * xm[idx] + vc * xc[idx] + vv * vocoded;
}

}  // namespace

int vocoder_host(ca_ctx* ctx, const ca_vocoder_params* prm, const float* h_modulator, const float* h_carrier, int64_t n_samples,
                 int64_t n_streams, float* h_out) {
  if (!prm || !h_modulator || !h_carrier || !h_out || n_samples < 0 || n_streams < 0)
    return fail(ctx, CA_ERR_INVALID, "ca_vocoder_host: null argument or negative size");
  if (prm->count_bands < 1 || prm->count_bands > 50 || !(prm->min_freq > 0) || !(prm->max_freq > prm->min_freq))
    return fail(ctx, CA_ERR_INVALID, "ca_vocoder_host: 1..50 bands between 0 < min_freq < max_freq (max_count_bands, vocoder.cpp:98)");
  const int sr = prm->sample_rate;
  // good_stride (:84-92) and the modulator window (:110-118)
  int S = std::max(1, static_cast<int>(0.5f + prm->stride_seconds * sr));
  if (S % 2 == 0) ++S;
  int Wm = std::max(1, static_cast<int>(0.5f + sr * prm->modulator_window_size_seconds));
  if (Wm % 2 == 1) ++Wm;
  const int Nm = ca_fft_length_for(Wm, 1), Nc = ca_fft_length_for(2 * S, 1);
  if (Nc < 512 || Nc > 32768) return fail(ctx, CA_ERR_UNSUPPORTED, "ca_vocoder_host: carrier FFT length (2*stride rounded up) must be in [512, 32768]");
  if (Nm < 1024 || Nm > 32768) return fail(ctx, CA_ERR_UNSUPPORTED, "ca_vocoder_host: modulator FFT length must be in [1024, 32768]");
  const size_t total = static_cast<size_t>(n_samples) * n_streams;
  if (total == 0) return CA_OK;
  const int64_t frames_m = ca_stft_num_frames(n_samples, Wm, S), frames_c = ca_stft_num_frames(n_samples, 2 * S, S);
  const int nb = prm->count_bands;

  // band edges (SetupParams::fill_freqs, :15-26) and the bins of each band, in double as the reference
  std::vector<double> freqs(nb + 1);
  {
    const auto log_min = std::log(prm->min_freq), log_max = std::log(prm->max_freq);
    for (int i = 0; i <= nb; ++i) freqs[i] = std::exp(log_min + static_cast<double>(i) / nb * (log_max - log_min));
  }
  auto band_of = [&](double hz) {
    if (!(hz > freqs[0])) return -1;
    for (int b = 0; b < nb; ++b)
      if (hz <= freqs[b + 1]) return b;
    return -1;
  };
  // bins of band b of the modulator spectrum: [begin[b], begin[b+1]) with begin[b] = first bin above freqs[b]
  std::vector<int32_t> band_begin(nb + 1, Nm / 2 + 1);
  for (int b = 0, i = 0; b <= nb; ++b) {
    while (i <= Nm / 2 && !(static_cast<double>(sr) / Nm * i > freqs[b])) ++i;
    band_begin[b] = i;
  }
  std::vector<int8_t> band_c(Nc / 2 + 1);
  for (int i = 0; i <= Nc / 2; ++i) band_c[i] = static_cast<int8_t>(band_of(static_cast<double>(sr) / Nc * i));

  cudaStream_t st = ctx->own_stream;
  const int64_t row_stride = (Nm / 2 + 1 + 3) & ~3ll;
  const size_t rows_m = static_cast<size_t>(frames_m) * n_streams, rows_c = static_cast<size_t>(frames_c) * n_streams;
  const size_t b_tab = sizeof(int32_t) * (nb + 1) + band_c.size();
  // FFT lengths >= 8192 (the default window of 0.10 s is 4800 samples -> 8192 points): the band sums come out of the analysis kernel and
  // the |X|^2 rows never reach HBM (CA_VOC_FUSED_BANDS=0: the two-kernel form, A/B)
  static const bool fused_env = [] { const char* e = getenv("CA_VOC_FUSED_BANDS"); return !e || atoi(e) != 0; }();
  const bool fused_bands = fused_env && Nm >= 8192;
  int rc = ensure_scratch(ctx, 0, sizeof(float) * total);
  if (!rc) rc = ensure_scratch(ctx, 1, sizeof(float) * total);
  if (!rc) rc = ensure_scratch(ctx, 2, std::max<size_t>(16, fused_bands ? 16 : sizeof(float) * rows_m * row_stride));
  if (!rc) rc = ensure_scratch(ctx, 3, std::max<size_t>(16, sizeof(float) * rows_m * nb));
  if (!rc) rc = ensure_scratch(ctx, 4, std::max<size_t>(16, sizeof(float) * rows_c * Nc));
  if (!rc) rc = ensure_scratch(ctx, 5, std::max<size_t>(16, sizeof(float2) * rows_c * (Nc / 2 + 1)));
  if (!rc) rc = ensure_scratch(ctx, 6, sizeof(float) * total);
  if (!rc) rc = ensure_scratch(ctx, 7, b_tab);
  if (rc) return rc;
  float* d_xm = static_cast<float*>(ctx->d_scratch[0]);
  float* d_xc = static_cast<float*>(ctx->d_scratch[1]);
  float* d_sq = static_cast<float*>(ctx->d_scratch[2]);
  float* d_amps = static_cast<float*>(ctx->d_scratch[3]);
  float* d_sig = static_cast<float*>(ctx->d_scratch[4]);
  float* d_bins = static_cast<float*>(ctx->d_scratch[5]);
  float* d_out = static_cast<float*>(ctx->d_scratch[6]);
  int32_t* d_band_begin = static_cast<int32_t*>(ctx->d_scratch[7]);
  int8_t* d_band_c = reinterpret_cast<int8_t*>(d_band_begin + nb + 1);
  CA_CUDA_OK(ctx, cudaMemcpyAsync(d_band_begin, band_begin.data(), sizeof(int32_t) * (nb + 1), cudaMemcpyHostToDevice, st));
  CA_CUDA_OK(ctx, cudaMemcpyAsync(d_band_c, band_c.data(), band_c.size(), cudaMemcpyHostToDevice, st));
  if (rows_m) {
    ca_stft_params sp{};
    sp.window = Wm;
    sp.hop = S;
    sp.window_type = CA_WINDOW_GAUSSIAN;
    sp.sigma_factor = 4.f;
    sp.sample_rate = sr;
    sp.min_db = -INFINITY;
    sp.max_peaks = 1;
    rc = stft_plan(ctx, ctx->plan_voc, &sp);
    if (rc) return rc;
  }
  CA_CUDA_OK(ctx, cudaStreamSynchronize(st));  // tables and window are in place before the stream groups start
  // carrier frame j is fired by sample 2S-1+jS and sees the modulator frames fired up to that sample (Wm-1+mS <= 2S-1+jS)
  const int64_t num = 2ll * S - Wm;  // m(j) = floor((num + jS)/S) = j + floor(num/S)
  const int first_m = static_cast<int>(num >= 0 ? num / S : -((-num + S - 1) / S));
  // Streams are independent: they go through the stages in up to four groups, each on its own CUDA stream, so that the copies
  // of one group overlap the kernels of the others (the call is PCIe-bound: 12 bytes per sample cross it).
  const int groups = static_cast<int>(std::min<int64_t>(4, n_streams));
  for (int g = 0; g < groups; ++g)
    if (!ctx->pipe_streams[g]) CA_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->pipe_streams[g], cudaStreamNonBlocking));
  for (int g = 0; g < groups && !rc; ++g) {
    cudaStream_t sg = ctx->pipe_streams[g];
    const int64_t a0 = n_streams * g / groups, a1 = n_streams * (g + 1) / groups, ns = a1 - a0;
    const size_t x0 = static_cast<size_t>(a0) * n_samples, nx = static_cast<size_t>(ns) * n_samples;
    const size_t m0 = static_cast<size_t>(a0) * frames_m, nm = static_cast<size_t>(ns) * frames_m;
    const size_t c0 = static_cast<size_t>(a0) * frames_c, nc = static_cast<size_t>(ns) * frames_c;
    cudaMemcpyAsync(d_xm + x0, h_modulator + x0, sizeof(float) * nx, cudaMemcpyHostToDevice, sg);
    cudaMemcpyAsync(d_xc + x0, h_carrier + x0, sizeof(float) * nx, cudaMemcpyHostToDevice, sg);
    if (nm && fused_bands) {  // modulator: the analysis kernel sums the bands of every frame from its shared-memory spectrum
      ctx->plan_voc.band_begin = d_band_begin;
      ctx->plan_voc.n_bands = nb;
      ctx->plan_voc.band_amps = d_amps + m0 * nb;
      rc = stft_analyze(ctx, ctx->plan_voc, d_xm + x0, n_samples, ns, n_samples, nullptr, 0, nullptr, nullptr, nullptr, nullptr, sg);
      ctx->plan_voc.band_amps = nullptr;
      if (rc) break;
    } else if (nm) {  // modulator: |X|^2 of every frame, then the band amplitudes
      rc = stft_analyze(ctx, ctx->plan_voc, d_xm + x0, n_samples, ns, n_samples, d_sq + m0 * row_stride, row_stride, nullptr, nullptr, nullptr, nullptr, sg);
      if (rc) break;
      band_amps_kernel<<<static_cast<unsigned>((nm + 3) / 4), 128, 0, sg>>>(d_sq + m0 * row_stride, row_stride, static_cast<int64_t>(nm),
                                                                           d_band_begin, nb, d_amps + m0 * nb);
      ctx->launches += 1;
    }
    float* sig_g = d_sig + c0 * Nc;
    if (nc) {  // carrier: frames -> spectrum -> band gains -> signal
      float* bins_g = d_bins + c0 * 2 * (Nc / 2 + 1);
      voc_frames_kernel<<<static_cast<unsigned>(nc), 256, 0, sg>>>(d_xc + x0, n_samples, frames_c, S, Nc, sig_g);
      rc = rfft_batched(ctx, sig_g, Nc, static_cast<int64_t>(nc), bins_g, sg);
      if (rc) break;
      voc_modulate_kernel<<<static_cast<unsigned>(nc), 256, 0, sg>>>(reinterpret_cast<float2*>(bins_g), Nc / 2 + 1, frames_c, frames_m, first_m,
                                                                    d_band_c, d_amps + m0 * nb, nb);
      rc = irfft_batched(ctx, bins_g, Nc, static_cast<int64_t>(nc), sig_g, 1.f / Nc, sg);
      if (rc) break;
      ctx->launches += 2;
    }
    voc_mix_kernel<<<static_cast<unsigned>((nx + 255) / 256), 256, 0, sg>>>(d_xm + x0, d_xc + x0, sig_g, n_samples, ns, frames_c, S, Nc,
                                                                           prm->vol_modulator, prm->vol_carrier, prm->vol_vocoder, d_out + x0);
    ctx->launches += 1;
    cudaMemcpyAsync(h_out + x0, d_out + x0, sizeof(float) * nx, cudaMemcpyDeviceToHost, sg);
  }
  cudaError_t err = cudaGetLastError();
  for (int g = 0; g < groups; ++g) {
    const cudaError_t e = cudaStreamSynchronize(ctx->pipe_streams[g]);
    if (err == cudaSuccess) err = e;
  }
  if (rc) return rc;
  if (err != cudaSuccess) return fail(ctx, CA_ERR_CUDA, std::string("ca_vocoder_host: ") + cudaGetErrorString(err));
  return CA_OK;
}

}  // namespace ca
