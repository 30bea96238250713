// rfftw.cu — K1 for N = 1024 / 2048 / 4096: batched forward real FFT with one warp per frame (two frames per warp at
// N = 1024), packed FP32 arithmetic (packed.cuh), every sample read from HBM once and every bin written once.
//
// Replaces fft::Algo_<Tag,T>::forward as FftOperation<T> uses it (source/rt.resynth.lib.periodicfft.cpp:210-235) and
// forward_fft (source/main.test_fft.cpp:78) for these lengths. The N real points are packed as M = N/2 complex points
// z[n] = x[2n] + i x[2n+1], M = R1 x 32:
//   pass 1  lane n2 transforms its R1 points z[32*n1 + n2] in registers (radix-2 DIT, immediate twiddles),
//   twiddle W_M^(n2*k1), k1 = B*a + b, applied as two table twiddles; 128-bit transposed stores to shared memory,
//   pass 2  each lane transforms whole columns k1 (32 points over n2) and writes Z[k1 + R1*k2] back to shared memory,
//   untangle X[k] = (E + T)/2, X[M-k] = conj(E - T)/2 with E = Z[k] + conj Z[M-k], T = -i W_N^k (Z[k] - conj Z[M-k]);
//            lane l takes k = l + 32*i, so the stores to HBM are coalesced and -i W_N^k = wpbase(l) * W_N^(32 i) needs one
//            per-lane table entry and an immediate.
// The kernel is HBM-bound by design: 4N bytes in, 8(N/2+1) bytes out per frame, ~0.06 instructions per byte.
#include <algorithm>

#include "common.cuh"
#include "ctx.hpp"
#include "packed.cuh"

namespace ca {

namespace {

constexpr int kXStride = 66;  // see stft.cu: element (k1, n2) at (k1 >> 1)*66 + 2*n2 + (k1 & 1), conflict-free both ways

template <int R1>
struct WarpFftCfg {
  static constexpr int M = 32 * R1;                  // complex points per frame
  static constexpr int FPW = R1 >= 32 ? 1 : 32 / R1;  // frames per warp and iteration
  static constexpr int CPL = R1 * FPW / 32;          // columns per lane in pass 2
  static constexpr int B = R1 >= 32 ? 8 : 4;         // k1 = B*a + b
  static constexpr int A = R1 / B;
  static constexpr int kRows = (B - 1) + (A - 1) + 1;  // table rows: wb[1..B-1], wa[1..A-1], wpbase
  static constexpr int kBufC2 = FPW * (R1 / 2) * kXStride > FPW * M ? FPW * (R1 / 2) * kXStride : FPW * M;  // per warp
};

__device__ __forceinline__ float2 cis_neg_w(float turns_times_two) {  // exp(-i*pi*arg)
  float s, c;
  sincospif(turns_times_two, &s, &c);
  return make_float2(c, -s);
}

template <int R1, int ROW, int... E>
__device__ __forceinline__ void tww_row(c2 (&v)[R1], const TwP& w, std::integer_sequence<int, E...>) {
  using C = WarpFftCfg<R1>;
  if constexpr (ROW < C::B - 1) ((v[C::B * E + ROW + 1] = cmul_tw(v[C::B * E + ROW + 1], w)), ...);  // wb[ROW+1], all a
  else ((v[C::B * (ROW - (C::B - 2)) + E] = cmul_tw(v[C::B * (ROW - (C::B - 2)) + E], w)), ...);      // wa[ROW-B+2], all b
}
template <int R1, int ROW>
__device__ __forceinline__ void tww_chain(c2 (&v)[R1], const float4* __restrict__ tab4, int lane, const TwP& w) {
  using C = WarpFftCfg<R1>;
  constexpr int kLast = (C::B - 1) + (C::A - 1);
  TwP wn = w;
  if constexpr (ROW + 1 < kLast) wn = make_twp(tab4[(ROW + 1) * 32 + lane]);
  if constexpr (ROW < C::B - 1) tww_row<R1, ROW>(v, w, std::make_integer_sequence<int, C::A>{});
  else tww_row<R1, ROW>(v, w, std::make_integer_sequence<int, C::B>{});
  if constexpr (ROW + 1 < kLast) tww_chain<R1, ROW + 1>(v, tab4, lane, wn);
}

template <int I, int R1>
__device__ __forceinline__ void untangle_store(const c2* __restrict__ Z, float2* __restrict__ o, const TwP& wpb, int lane) {
  using C = WarpFftCfg<R1>;
  constexpr int M = C::M;
  const int k = lane + 32 * I;
  const c2 zk = Z[k], zm = Z[(M - k) & (M - 1)];
  const c2 E = add2(zk, cconj(zm));
  c2 D = csub(zk, cconj(zm));
  if constexpr (I != 0) D = cmul_imm<Tw<2 * M, 32 * I>>(D);  // W_N^(32 i)
  const c2 T = cmul_tw(D, wpb);                               // wpb = 0.5 * (-i) * W_N^lane
  const c2 P = fma2(bc(0.5f), E, T);
  const c2 Q = cconj(fma2(bc(0.5f), E, cneg(T)));
  o[k] = upk(P);
  o[M - k] = upk(Q);
}
template <int R1, int... I>
__device__ __forceinline__ void untangle_all(const c2* Z, float2* o, const TwP& wpb, int lane, std::integer_sequence<int, I...>) {
  (untangle_store<I, R1>(Z, o, wpb, lane), ...);
}

template <int R1, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) rfft_warp_kernel(const float* __restrict__ in, float2* __restrict__ out,
                                                                  const int64_t batch) {
  using C = WarpFftCfg<R1>;
  constexpr int M = C::M, FPW = C::FPW, CPL = C::CPL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* tab4 = reinterpret_cast<float4*>(smem_raw);               // [kRows][32]
  c2* buf_all = reinterpret_cast<c2*>(tab4 + C::kRows * 32);        // [WARPS][kBufC2]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  for (int i = threadIdx.x; i < C::kRows * 32; i += WARPS * 32) {
    const int r = i >> 5, l = i & 31;
    float2 w;
    if (r < C::B - 1) w = cis_neg_w(2.f * static_cast<float>(l * (r + 1)) / M);                          // W_M^(l*b)
    else if (r < C::B - 1 + C::A - 1) w = cis_neg_w(2.f * static_cast<float>(l * C::B * (r - (C::B - 2))) / M);  // W_M^(l*B*a)
    else {
      float s, c;
      sincospif(static_cast<float>(l) / M, &s, &c);  // W_N^l = exp(-2 pi i l / 2M)
      w = make_float2(-0.5f * s, -0.5f * c);         // 0.5 * (-i) * W_N^l
    }
    tab4[i] = make_float4(w.x, w.y, -w.y, w.y);
  }
  __syncthreads();
  c2* buf = buf_all + warp * C::kBufC2;
  const TwP wpb = make_twp(tab4[(C::kRows - 1) * 32 + lane]);

  const int64_t groups = (batch + FPW - 1) / FPW;
  for (int64_t g = static_cast<int64_t>(blockIdx.x) * WARPS + warp; g < groups; g += static_cast<int64_t>(gridDim.x) * WARPS) {
    // ---- pass 1 per frame of the group: load, R1-point FFT, twiddle, transposed store ----
#pragma unroll
    for (int fw = 0; fw < FPW; ++fw) {
      const int64_t frame = g * FPW + fw;
      c2 v[R1];
      if (frame < batch) {
        const float* x = in + frame * (2 * M);
#pragma unroll
        for (int n1 = 0; n1 < R1; ++n1) v[bitrev(n1, ilog2(R1))] = ldg_c2(x + 64 * n1 + 2 * lane);
      } else {
#pragma unroll
        for (int n1 = 0; n1 < R1; ++n1) v[n1] = pk(0.f, 0.f);
      }
      DitP<R1, R1, 0>::run(v);
      tww_chain<R1, 0>(v, tab4, lane, make_twp(tab4[lane]));
      c2* xb = buf + fw * (R1 / 2) * kXStride;
#pragma unroll
      for (int m = 0; m < R1 / 2; ++m)
        *reinterpret_cast<ulonglong2*>(xb + m * kXStride + 2 * lane) = make_ulonglong2(v[2 * m], v[2 * m + 1]);
    }
    __syncwarp();
    // ---- pass 2: whole columns ----
    c2 A[CPL][32];
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int col = FPW > 1 ? (lane & (R1 - 1)) : lane + 32 * c;
      const c2* xb = buf + (FPW > 1 ? (lane / R1) * (R1 / 2) * kXStride : 0);
#pragma unroll
      for (int n2 = 0; n2 < 32; ++n2) A[c][bitrev(n2, 5)] = xb[(col >> 1) * kXStride + 2 * n2 + (col & 1)];
    }
    __syncwarp();  // the buffer is rewritten in Z order
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      DitP<32, 32, 0>::run(A[c]);
      const int col = FPW > 1 ? (lane & (R1 - 1)) : lane + 32 * c;
      c2* Z = buf + (FPW > 1 ? (lane / R1) * M : 0);
#pragma unroll
      for (int k2 = 0; k2 < 32; ++k2) Z[col + R1 * k2] = A[c][k2];
    }
    __syncwarp();
    // ---- untangle + coalesced stores ----
#pragma unroll
    for (int fw = 0; fw < FPW; ++fw) {
      const int64_t frame = g * FPW + fw;
      if (frame >= batch) break;
      const c2* Z = buf + fw * M;
      float2* o = out + frame * (M + 1);
      untangle_all<R1>(Z, o, wpb, lane, std::make_integer_sequence<int, M / 64>{});
      if (lane == 0) {  // k = M/2 pairs with itself: X[M/2] = conj(Z[M/2])
        const float2 z = upk(Z[M / 2]);
        o[M / 2] = make_float2(z.x, -z.y);
      }
    }
    __syncwarp();
  }
}

template <int R1, int WARPS>
int launch_warp(ca_ctx* ctx, const float* d_in, int64_t batch, float* d_out, cudaStream_t stream) {
  using C = WarpFftCfg<R1>;
  auto kern = rfft_warp_kernel<R1, WARPS>;
  const size_t smem = sizeof(float4) * C::kRows * 32 + sizeof(c2) * C::kBufC2 * WARPS;
  CA_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int per_sm = 0;
  CA_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
  if (per_sm < 1) per_sm = 1;
  const int64_t groups = (batch + C::FPW - 1) / C::FPW;
  int64_t grid = static_cast<int64_t>(ctx->sm_count) * per_sm;
  grid = std::min<int64_t>(grid, (groups + WARPS - 1) / WARPS);
  kern<<<static_cast<unsigned>(grid), WARPS * 32, smem, stream>>>(d_in, reinterpret_cast<float2*>(d_out), batch);
  CA_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches += 1;
  return CA_OK;
}

}  // namespace

int rfft_warp(ca_ctx* ctx, const float* d_in, int n_fft, int64_t batch, float* d_out, cudaStream_t stream) {
  switch (n_fft) {
    case 1024: return launch_warp<16, 8>(ctx, d_in, batch, d_out, stream);
    case 2048: return launch_warp<32, 8>(ctx, d_in, batch, d_out, stream);
    case 4096: return launch_warp<64, 12>(ctx, d_in, batch, d_out, stream);
    default: break;
  }
  return fail(ctx, CA_ERR_UNSUPPORTED, "rfft_warp: unsupported n_fft");
}

}  // namespace ca
