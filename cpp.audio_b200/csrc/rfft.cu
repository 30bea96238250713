// rfft.cu — K1: batched forward real FFT, N = 2^10 .. 2^20, float32, unscaled.
//
// Stands in for fft::Algo_<Tag,T>::forward as FftOperation<T> uses it (source/rt.resynth.lib.periodicfft.cpp:210-235) and
// forward_fft (source/main.test_fft.cpp:78). An N-point real transform is an N/2-point complex transform of the packed
// signal z[n] = x[2n] + i x[2n+1] followed by the untangling step; every frame is read from HBM once and written once.
//
//   N <= 32768 : one CTA runs a Stockham autosort FFT (radix 8, last pass radix 2/4/8) entirely in shared memory:
//                float2-coalesced loads straight into registers for the first pass, one shared-memory exchange per pass,
//                untangling fused into the store.
//   N >= 65536 : four-step through HBM (column FFTs with twiddle, row FFTs, untangle) — three passes over the data.
#include <math.h>

#include "common.cuh"
#include "ctx.hpp"

namespace ca {

namespace {

// T[i] = exp(-2 pi i * i / n), i < n
__global__ void twiddle_table_kernel(float2* tw, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s, c;
  sincospi(2.0 * static_cast<double>(i) / n, &s, &c);
  tw[i] = make_float2(static_cast<float>(c), static_cast<float>(-s));
}

template <int R>
__device__ __forceinline__ void radix_pass_load(const float2* __restrict__ buf, const float2* __restrict__ tw, int tw_stride,
                                                int M, int s, int j, float2 (&a)[R]) {
  // butterfly j of a radix-R Stockham pass with accumulated stride s: inputs buf[j + q*M/R], twiddles W_(s*R)^(k*q)
  const int k = j & (s - 1);
#pragma unroll
  for (int q = 0; q < R; ++q) a[q] = buf[j + q * (M / R)];
  if (s > 1) {
    const int unit = (M / (s * R)) * tw_stride;  // W_(sR)^1 expressed in table steps
#pragma unroll
    for (int q = 1; q < R; ++q) a[q] = cmul(a[q], __ldg(tw + k * q * unit));
  }
  Dif<R, R, 0>::run(a);
}

template <int R>
__device__ __forceinline__ void radix_pass_store(float2* __restrict__ buf, int s, int j, const float2 (&a)[R]) {
  const int k = j & (s - 1);
  const int o = (j - k) * R + k;
#pragma unroll
  for (int q = 0; q < R; ++q) buf[o + q * s] = a[bitrev(q, ilog2(R))];
}

// One frame per group of T threads, FPB frames per CTA. M = N/2 complex points; B = M/(8T) radix-8 butterflies/thread.
template <int M, int T, int FPB>
__global__ void __launch_bounds__(T * FPB) rfft_smem_kernel(const float* __restrict__ in, float2* __restrict__ out,
                                                             int64_t batch, const float2* __restrict__ tw /* W_(2M) */) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int B = M / (8 * T);
  static_assert(B >= 1, "too many threads per frame");
  const int fi = threadIdx.x / T;
  const int t = threadIdx.x - fi * T;
  const int64_t frame = static_cast<int64_t>(blockIdx.x) * FPB + fi;
  const bool live = frame < batch;
  float2* buf = reinterpret_cast<float2*>(smem_raw) + fi * M;
  const float2* zin = reinterpret_cast<const float2*>(in) + frame * M;

  constexpr int LOGM = ilog2(M);
  constexpr int N8 = LOGM / 3;        // radix-8 passes
  constexpr int REM = LOGM - 3 * N8;  // 0: none, 1: radix 2, 2: radix 4

  // ---- pass 0 (radix 8, s = 1): straight from HBM ----
  {
    float2 a[B][8];
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const int j = t + b * T;
#pragma unroll
      for (int q = 0; q < 8; ++q) a[b][q] = live ? __ldg(zin + j + q * (M / 8)) : make_float2(0.f, 0.f);
      Dif<8, 8, 0>::run(a[b]);
    }
#pragma unroll
    for (int b = 0; b < B; ++b) radix_pass_store<8>(buf, 1, t + b * T, a[b]);
  }
  __syncthreads();

  int s = 8;
#pragma unroll
  for (int pass = 1; pass < N8; ++pass) {
    float2 a[B][8];
#pragma unroll
    for (int b = 0; b < B; ++b) radix_pass_load<8>(buf, tw, 2, M, s, t + b * T, a[b]);
    __syncthreads();
#pragma unroll
    for (int b = 0; b < B; ++b) radix_pass_store<8>(buf, s, t + b * T, a[b]);
    __syncthreads();
    s *= 8;
  }
  if constexpr (REM == 2) {
    float2 a[2 * B][4];
#pragma unroll
    for (int b = 0; b < 2 * B; ++b) radix_pass_load<4>(buf, tw, 2, M, s, t + b * T, a[b]);
    __syncthreads();
#pragma unroll
    for (int b = 0; b < 2 * B; ++b) radix_pass_store<4>(buf, s, t + b * T, a[b]);
    __syncthreads();
  } else if constexpr (REM == 1) {
    float2 a[4 * B][2];
#pragma unroll
    for (int b = 0; b < 4 * B; ++b) radix_pass_load<2>(buf, tw, 2, M, s, t + b * T, a[b]);
    __syncthreads();
#pragma unroll
    for (int b = 0; b < 4 * B; ++b) radix_pass_store<2>(buf, s, t + b * T, a[b]);
    __syncthreads();
  }

  // ---- untangle + store: X[k] = (E + T)/2, X[M-k] = conj(E - T)/2, T = -i W_N^k (Z[k] - conj Z[M-k]) ----
  if (!live) return;
  float2* o = out + frame * (M + 1);
  for (int k = t; k <= M / 2; k += T) {
    const float2 zk = buf[k];
    const float2 zm = buf[(M - k) & (M - 1)];
    const float2 w = __ldg(tw + k);             // W_(2M)^k
    const float2 wp = make_float2(w.y, -w.x);   // -i * w
    const float ex = zk.x + zm.x, ey = zk.y - zm.y;
    const float dx = zk.x - zm.x, dy = zk.y + zm.y;
    const float tx = fmaf(-dy, wp.y, dx * wp.x), ty = fmaf(dx, wp.y, dy * wp.x);
    o[k] = make_float2(0.5f * (ex + tx), 0.5f * (ey + ty));
    o[M - k] = make_float2(0.5f * (ex - tx), -0.5f * (ey - ty));
  }
}

template <int M, int T, int FPB>
int launch_smem(ca_ctx* ctx, const float* d_in, int64_t batch, float* d_out, cudaStream_t stream) {
  const size_t smem = sizeof(float2) * M * FPB;
  auto kern = rfft_smem_kernel<M, T, FPB>;
  CA_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int64_t grid = (batch + FPB - 1) / FPB;
  kern<<<static_cast<unsigned>(grid), T * FPB, smem, stream>>>(d_in, reinterpret_cast<float2*>(d_out), batch, ctx->d_fft_tw);
  CA_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches += 1;
  return CA_OK;
}

int ensure_twiddles(ca_ctx* ctx, int n, cudaStream_t stream) {
  if (ctx->fft_tw_n == n) return CA_OK;
  if (ctx->d_fft_tw) cudaFree(ctx->d_fft_tw);
  ctx->d_fft_tw = nullptr;
  ctx->fft_tw_n = 0;
  CA_CUDA_OK(ctx, cudaMalloc(&ctx->d_fft_tw, sizeof(float2) * n));
  twiddle_table_kernel<<<(n + 255) / 256, 256, 0, stream>>>(ctx->d_fft_tw, n);
  CA_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches += 1;
  ctx->fft_tw_n = n;
  return CA_OK;
}

}  // namespace

void rfft_release(ca_ctx* ctx) {
  if (ctx->d_fft_tw) cudaFree(ctx->d_fft_tw);
  if (ctx->d_fft_work) cudaFree(ctx->d_fft_work);
  ctx->d_fft_tw = nullptr;
  ctx->d_fft_work = nullptr;
  ctx->fft_tw_n = 0;
  ctx->fft_work_bytes = 0;
}

int rfft_batched(ca_ctx* ctx, const float* d_in, int n_fft, int64_t batch, float* d_out, cudaStream_t stream) {
  if (!d_in || !d_out || batch < 0) return fail(ctx, CA_ERR_INVALID, "ca_rfft_batched: null pointer or negative batch");
  if (n_fft < 1024 || n_fft > (1 << 20) || (n_fft & (n_fft - 1)))
    return fail(ctx, CA_ERR_UNSUPPORTED, "ca_rfft_batched: n_fft must be a power of two in [2^10, 2^20]");
  if (batch == 0) return CA_OK;
  if (reinterpret_cast<uintptr_t>(d_in) % 8 || reinterpret_cast<uintptr_t>(d_out) % 8)
    return fail(ctx, CA_ERR_INVALID, "ca_rfft_batched: buffers must be 8-byte aligned");
  int rc = ensure_twiddles(ctx, n_fft, stream);
  if (rc) return rc;
  switch (n_fft) {
    case 1024: return launch_smem<512, 64, 4>(ctx, d_in, batch, d_out, stream);
    case 2048: return launch_smem<1024, 128, 2>(ctx, d_in, batch, d_out, stream);
    case 4096: return launch_smem<2048, 256, 1>(ctx, d_in, batch, d_out, stream);
    case 8192: return launch_smem<4096, 512, 1>(ctx, d_in, batch, d_out, stream);
    case 16384: return launch_smem<8192, 512, 1>(ctx, d_in, batch, d_out, stream);
    case 32768: return launch_smem<16384, 1024, 1>(ctx, d_in, batch, d_out, stream);
    default: break;
  }
  return fail(ctx, CA_ERR_UNSUPPORTED, "ca_rfft_batched: n_fft >= 65536 (four-step path) is not built yet");
}

}  // namespace ca
