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

#include <algorithm>

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
// kComplex: the rows are already complex (M float2 in, M float2 out, no untangling) — the row pass of the four-step path;
// tw_stride is the step of W_M^1 in the twiddle table (2 when the table is W_(2M), larger for the four-step path).
template <int M, int T, int FPB, bool kComplex>
__global__ void __launch_bounds__(T * FPB) rfft_smem_kernel(const float* __restrict__ in, float2* __restrict__ out,
                                                             int64_t batch, const float2* __restrict__ tw, int tw_stride) {
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
    for (int b = 0; b < B; ++b) radix_pass_load<8>(buf, tw, tw_stride, M, s, t + b * T, a[b]);
    __syncthreads();
#pragma unroll
    for (int b = 0; b < B; ++b) radix_pass_store<8>(buf, s, t + b * T, a[b]);
    __syncthreads();
    s *= 8;
  }
  if constexpr (REM == 2) {
    float2 a[2 * B][4];
#pragma unroll
    for (int b = 0; b < 2 * B; ++b) radix_pass_load<4>(buf, tw, tw_stride, M, s, t + b * T, a[b]);
    __syncthreads();
#pragma unroll
    for (int b = 0; b < 2 * B; ++b) radix_pass_store<4>(buf, s, t + b * T, a[b]);
    __syncthreads();
  } else if constexpr (REM == 1) {
    float2 a[4 * B][2];
#pragma unroll
    for (int b = 0; b < 4 * B; ++b) radix_pass_load<2>(buf, tw, tw_stride, M, s, t + b * T, a[b]);
    __syncthreads();
#pragma unroll
    for (int b = 0; b < 4 * B; ++b) radix_pass_store<2>(buf, s, t + b * T, a[b]);
    __syncthreads();
  }

  // ---- untangle + store: X[k] = (E + T)/2, X[M-k] = conj(E - T)/2, T = -i W_N^k (Z[k] - conj Z[M-k]) ----
  if (!live) return;
  if constexpr (kComplex) {
    float2* o = out + frame * M;
    for (int k = t; k < M; k += T) o[k] = buf[k];
    return;
  }
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
  auto kern = rfft_smem_kernel<M, T, FPB, false>;
  CA_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int64_t grid = (batch + FPB - 1) / FPB;
  kern<<<static_cast<unsigned>(grid), T * FPB, smem, stream>>>(d_in, reinterpret_cast<float2*>(d_out), batch, ctx->d_fft_tw, 2);
  CA_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches += 1;
  return CA_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// Four-step path for N >= 65536. M = N/2 complex points viewed as [M1 rows][M2 = 1024 columns] (n = r*M2 + c):
//   Z[k1 + M1*k2] = sum_c W_M2^(c k2) * ( W_M^(c k1) * sum_r z[r*M2 + c] W_M1^(r k1) )
//   pass A: column FFTs of length M1 (+ twiddle W_M^(c k1)), tiles of 16 columns staged in shared memory
//   pass B: row FFTs of length 1024 (the shared-memory kernel above in complex mode), in place
//   pass C: untangle to the real spectrum, transposing [k1][k2] -> k = k1 + M1*k2 through shared-memory tiles
// ---------------------------------------------------------------------------------------------------------------
constexpr int kM2 = 1024;
constexpr int kColTile = 16;

template <int M1>
__global__ void __launch_bounds__(256) fourstep_cols_kernel(const float2* __restrict__ zin, float2* __restrict__ work,
                                                             const float2* __restrict__ tw /* W_(2M), 2M entries */) {
  // grid: (kM2 / kColTile, frames). smem: [M1][kColTile+1] float2
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* buf = reinterpret_cast<float2*>(smem_raw);
  constexpr int LD = kColTile + 1;
  constexpr int M = M1 * kM2;
  const int c0 = blockIdx.x * kColTile;
  const float2* src = zin + static_cast<int64_t>(blockIdx.y) * M;
  float2* dst = work + static_cast<int64_t>(blockIdx.y) * M;
  // load in bit-reversed row order so that the in-place DIT passes below end in natural order
  constexpr int LOG = ilog2(M1);
  for (int i = threadIdx.x; i < M1 * kColTile; i += 256) {
    const int r = i / kColTile, c = i % kColTile;
    const int rr = static_cast<int>(__brev(static_cast<unsigned>(r)) >> (32 - LOG));
    buf[rr * LD + c] = __ldg(src + static_cast<int64_t>(r) * kM2 + c0 + c);
  }
  __syncthreads();
  for (int half = 1; half < M1; half <<= 1) {  // radix-2 DIT, twiddle W_(2*half)^j = W_M1^(j*M1/(2*half))
    const int tstep = (M1 / (2 * half)) * (2 * kM2);  // table is W_(2M): W_M1^1 = W_(2M)^(2*M2)
    for (int i = threadIdx.x; i < (M1 / 2) * kColTile; i += 256) {
      const int b = i / kColTile, c = i % kColTile;
      const int j = b & (half - 1);
      const int lo = ((b - j) << 1) + j;
      const float2 w = __ldg(tw + j * tstep);
      const float2 x0 = buf[lo * LD + c];
      const float2 x1 = cmul(buf[(lo + half) * LD + c], w);
      buf[lo * LD + c] = make_float2(x0.x + x1.x, x0.y + x1.y);
      buf[(lo + half) * LD + c] = make_float2(x0.x - x1.x, x0.y - x1.y);
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < M1 * kColTile; i += 256) {
    const int k1 = i / kColTile, c = i % kColTile;
    const float2 w = __ldg(tw + 2 * ((static_cast<int64_t>(k1) * (c0 + c)) % M));  // W_M^(c*k1)
    dst[static_cast<int64_t>(k1) * kM2 + c0 + c] = cmul(buf[k1 * LD + c], w);
  }
}

// pass C: work holds Z[k1 + M1*k2] at [k1][k2]. Each CTA produces X[k] for a 32(k1) x 32(k2) tile and its mirror.
template <int M1>
__global__ void __launch_bounds__(256) fourstep_untangle_kernel(const float2* __restrict__ work, float2* __restrict__ out,
                                                                 const float2* __restrict__ tw /* W_(2M) */) {
  constexpr int M = M1 * kM2;
  __shared__ float2 ta[32][33], tb[32][33];
  const float2* z = work + static_cast<int64_t>(blockIdx.z) * M;
  float2* o = out + static_cast<int64_t>(blockIdx.z) * (M + 1);
  const int k1_0 = blockIdx.x * 32, k2_0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  // tile A: Z[k1][k2] read along k2 (coalesced); tile B: the mirrors Z[M-k]
  for (int i = ty; i < 32; i += 8) {
    const int k1 = k1_0 + i, k2 = k2_0 + tx;
    ta[i][tx] = __ldg(z + static_cast<int64_t>(k1) * kM2 + k2);
    // M - k with k = k1 + M1*k2:  k1 > 0: (M1-k1) + M1*(M2-1-k2);  k1 == 0: M1*((M2-k2) % M2)
    const int m1 = k1 ? M1 - k1 : 0;
    const int m2 = k1 ? kM2 - 1 - k2 : (kM2 - k2) & (kM2 - 1);
    tb[i][tx] = __ldg(z + static_cast<int64_t>(m1) * kM2 + m2);
  }
  __syncthreads();
  // write along k1 (coalesced in k): thread tx -> k1 = k1_0 + tx, rows over k2
  for (int i = ty; i < 32; i += 8) {
    const int k1 = k1_0 + tx, k2 = k2_0 + i;
    const int64_t k = k1 + static_cast<int64_t>(M1) * k2;
    const float2 zk = ta[tx][i], zm = tb[tx][i];
    const float2 w = __ldg(tw + k);            // W_(2M)^k
    const float2 wp = make_float2(w.y, -w.x);  // -i*w
    const float ex = zk.x + zm.x, ey = zk.y - zm.y;
    const float dx = zk.x - zm.x, dy = zk.y + zm.y;
    const float tx_ = fmaf(-dy, wp.y, dx * wp.x), ty_ = fmaf(dx, wp.y, dy * wp.x);
    o[k] = make_float2(0.5f * (ex + tx_), 0.5f * (ey + ty_));
    if (k == 0) o[M] = make_float2(0.5f * (ex - tx_), -0.5f * (ey - ty_));
  }
}

template <int M1>
int launch_fourstep(ca_ctx* ctx, const float* d_in, int64_t batch, float* d_out, cudaStream_t stream) {
  constexpr int M = M1 * kM2;
  // bounded work buffer: frames are processed in groups
  const size_t want = static_cast<size_t>(256) << 20;
  const int64_t group = std::max<int64_t>(1, std::min<int64_t>(batch, static_cast<int64_t>(want / (sizeof(float2) * M))));
  const size_t need = sizeof(float2) * M * static_cast<size_t>(group);
  if (ctx->fft_work_bytes < need) {
    if (ctx->d_fft_work) cudaFree(ctx->d_fft_work);
    ctx->d_fft_work = nullptr;
    ctx->fft_work_bytes = 0;
    CA_CUDA_OK(ctx, cudaMalloc(&ctx->d_fft_work, need));
    ctx->fft_work_bytes = need;
  }
  const size_t smem_cols = sizeof(float2) * M1 * (kColTile + 1);
  CA_CUDA_OK(ctx, cudaFuncSetAttribute(fourstep_cols_kernel<M1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem_cols)));
  auto rows = rfft_smem_kernel<kM2, 128, 2, true>;
  const size_t smem_rows = sizeof(float2) * kM2 * 2;
  CA_CUDA_OK(ctx, cudaFuncSetAttribute(rows, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_rows)));
  for (int64_t b0 = 0; b0 < batch; b0 += group) {
    const int64_t nb = std::min(group, batch - b0);
    const float2* zin = reinterpret_cast<const float2*>(d_in) + b0 * M;
    float2* out = reinterpret_cast<float2*>(d_out) + b0 * (M + 1);
    fourstep_cols_kernel<M1><<<dim3(kM2 / kColTile, static_cast<unsigned>(nb)), 256, smem_cols, stream>>>(zin, ctx->d_fft_work, ctx->d_fft_tw);
    CA_CUDA_OK(ctx, cudaGetLastError());
    const int64_t n_rows = nb * M1;
    // W_1024^1 = W_(2M)^(2*M1)
    rows<<<static_cast<unsigned>((n_rows + 1) / 2), 256, smem_rows, stream>>>(reinterpret_cast<const float*>(ctx->d_fft_work), ctx->d_fft_work,
                                                                             n_rows, ctx->d_fft_tw, 2 * M1);
    CA_CUDA_OK(ctx, cudaGetLastError());
    fourstep_untangle_kernel<M1><<<dim3(M1 / 32, kM2 / 32, static_cast<unsigned>(nb)), 256, 0, stream>>>(ctx->d_fft_work, out, ctx->d_fft_tw);
    CA_CUDA_OK(ctx, cudaGetLastError());
    ctx->launches += 3;
  }
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
    case 1024:
    case 2048:
    case 4096: return rfft_warp(ctx, d_in, n_fft, batch, d_out, stream);  // one warp per frame (rfftw.cu)
    case 8192: return launch_smem<4096, 512, 1>(ctx, d_in, batch, d_out, stream);
    case 16384: return launch_smem<8192, 512, 1>(ctx, d_in, batch, d_out, stream);
    case 32768: return launch_smem<16384, 1024, 1>(ctx, d_in, batch, d_out, stream);
    case 1 << 16: return launch_fourstep<32>(ctx, d_in, batch, d_out, stream);
    case 1 << 17: return launch_fourstep<64>(ctx, d_in, batch, d_out, stream);
    case 1 << 18: return launch_fourstep<128>(ctx, d_in, batch, d_out, stream);
    case 1 << 19: return launch_fourstep<256>(ctx, d_in, batch, d_out, stream);
    case 1 << 20: return launch_fourstep<512>(ctx, d_in, batch, d_out, stream);
    default: break;
  }
  return fail(ctx, CA_ERR_UNSUPPORTED, "ca_rfft_batched: unsupported n_fft");
}

}  // namespace ca
