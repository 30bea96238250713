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
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "ctx.hpp"
#include "packed.cuh"

namespace ca {

namespace {

// Shared-memory opt-in and occupancy of a kernel, resolved once per launch function (= kernel instantiation) and device: the
// two runtime calls cost more than the launch itself, which matters for L2-resident batches (BASELINE config 1: 33.6 MB).
#define CA_KERNEL_SETUP(ctx, kern, threads, smem, per_sm)                                                                        \
  do {                                                                                                                          \
    static int s_dev_ = -1, s_per_sm_ = 0;                                                                                      \
    if (s_dev_ != (ctx)->device) {                                                                                              \
      CA_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));         \
      CA_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s_per_sm_, kern, threads, smem));                          \
      s_dev_ = (ctx)->device;                                                                                                   \
    }                                                                                                                           \
    per_sm = s_per_sm_;                                                                                                         \
  } while (0)

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

  // The samples of the warp's NEXT group are requested before the untangling / store phase of the current one (which
  // needs few registers) and land while it runs: more bytes in flight per SM, which is what an HBM-bound kernel with 8-12
  // warps per SM lacks.
  const int64_t groups = (batch + FPW - 1) / FPW;
  const int64_t gstep = static_cast<int64_t>(gridDim.x) * WARPS;
  c2 v[FPW][R1];
  // (not at R1 = 64: the 128 data registers do not fit next to the untangling phase — measured: spills, 70 % instead of 74 %)
  constexpr int kPre = R1 > 32 ? 0 : R1;
  auto load_group = [&](int64_t g, int n1_begin, int n1_end) {
#pragma unroll
    for (int fw = 0; fw < FPW; ++fw) {
      const int64_t frame = g * FPW + fw;
      if (frame < batch) {
        const float* x = in + frame * (2 * M);
#pragma unroll
        for (int n1 = 0; n1 < R1; ++n1)
          if (n1 >= n1_begin && n1 < n1_end) v[fw][bitrev(n1, ilog2(R1))] = ldg_c2(x + 64 * n1 + 2 * lane);
      } else {
#pragma unroll
        for (int n1 = 0; n1 < R1; ++n1)
          if (n1 >= n1_begin && n1 < n1_end) v[fw][bitrev(n1, ilog2(R1))] = pk(0.f, 0.f);
      }
    }
  };
  int64_t g = static_cast<int64_t>(blockIdx.x) * WARPS + warp;
  if (g < groups) load_group(g, 0, kPre);
  for (; g < groups; g += gstep) {
    if (kPre < R1) load_group(g, kPre, R1);
    // ---- pass 1 per frame of the group: R1-point FFT, twiddle, transposed store ----
#pragma unroll
    for (int fw = 0; fw < FPW; ++fw) {
      DitP<R1, R1, 0>::run(v[fw]);
      tww_chain<R1, 0>(v[fw], tab4, lane, make_twp(tab4[lane]));
      c2* xb = buf + fw * (R1 / 2) * kXStride;
#pragma unroll
      for (int m = 0; m < R1 / 2; ++m)
        *reinterpret_cast<ulonglong2*>(xb + m * kXStride + 2 * lane) = make_ulonglong2(v[fw][2 * m], v[fw][2 * m + 1]);
    }
    __syncwarp();
    // ---- pass 2: whole columns ----
    {
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
    }
    __syncwarp();
    if (g + gstep < groups) load_group(g + gstep, 0, kPre);  // in flight during the untangling
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
  int per_sm = 0;
  CA_KERNEL_SETUP(ctx, kern, WARPS * 32, smem, per_sm);
  if (per_sm < 1) per_sm = 1;
  const int64_t groups = (batch + C::FPW - 1) / C::FPW;
  int64_t grid = static_cast<int64_t>(ctx->sm_count) * per_sm;
  grid = std::min<int64_t>(grid, (groups + WARPS - 1) / WARPS);
  kern<<<static_cast<unsigned>(grid), WARPS * 32, smem, stream>>>(d_in, reinterpret_cast<float2*>(d_out), batch);
  CA_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches += 1;
  return CA_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// N = 8192 / 16384 / 32768: one CTA per frame, M = N/2 = M1 x 1024 complex points (M1 = 4, 8, 16), n = 1024*n1 + n2,
// k = k1 + M1*k2. The whole frame stays in shared memory between its single read and its single write of HBM:
//   step A  every thread takes columns n2, transforms their M1 points (strided, coalesced loads) in registers and
//           multiplies by W_M^(n2*k1) (powers of one table entry),
//   step B  warp k1 runs the 1024-point warp-level FFT (32 x 32, as above) on row k1, in place,
//   step C  untangling with k = t + T*i: coalesced stores, -i W_N^k = (per-thread table entry) * (immediate).
// Rows are kRowC2 apart so that the row FFT's exchange layout fits and step C reads M1 rows without bank conflicts.
// ---------------------------------------------------------------------------------------------------------------
template <int M1>
struct CtaFftCfg {
  static constexpr int M2 = 1024;
  static constexpr int M = M1 * M2;
  static constexpr int T = 32 * M1;  // one warp per row
  static constexpr int kRowC2 = 16 * kXStride + (M1 == 4 ? 4 : 2);
  static constexpr int kRows = 7 + 3;  // row-FFT tables: wb[1..7], wa[1..3]
};

template <int K, int M1>
__device__ __forceinline__ c2 cpow_p(const c2 (&w)[M1]) {  // w[K] = w[K/2] * w[K - K/2] (w[1] given): depth <= log2(M1)
  const float2 a = upk(w[K / 2]);
  return cmul_tw(w[K - K / 2], TwP{a.x, pk(-a.y, a.y)});
}
template <int M1, int... K>
__device__ __forceinline__ void cpow_all(c2 (&w)[M1], std::integer_sequence<int, K...>) {
  ((w[K + 2] = cpow_p<K + 2, M1>(w)), ...);
}

template <int I, int M1>
__device__ __forceinline__ void cta_untangle(const c2* __restrict__ Zs, float2* __restrict__ o, const TwP& wpb, int t) {
  using C = CtaFftCfg<M1>;
  constexpr int M = C::M;
  const int k = t + C::T * I;
  const int m = (M - k) & (M - 1);
  const c2 zk = Zs[(k % M1) * C::kRowC2 + k / M1], zm = Zs[(m % M1) * C::kRowC2 + m / M1];
  const c2 E = add2(zk, cconj(zm));
  c2 D = csub(zk, cconj(zm));
  if constexpr (I != 0) D = cmul_imm<Tw<2 * M, C::T * I>>(D);  // W_N^(T i)
  const c2 Tt = cmul_tw(D, wpb);                                // wpb = 0.5 * (-i) * W_N^t
  o[k] = upk(fma2(bc(0.5f), E, Tt));
  o[M - k] = upk(cconj(fma2(bc(0.5f), E, cneg(Tt))));
}
template <int M1, int... I>
__device__ __forceinline__ void cta_untangle_all(const c2* Zs, float2* o, const TwP& wpb, int t, std::integer_sequence<int, I...>) {
  (cta_untangle<I, M1>(Zs, o, wpb, t), ...);
}

template <int M1>
__global__ void __launch_bounds__(32 * M1) rfft_cta_kernel(const float* __restrict__ in, float2* __restrict__ out, const int64_t batch) {
  using C = CtaFftCfg<M1>;
  constexpr int M = C::M, M2 = C::M2, T = C::T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* tab4 = reinterpret_cast<float4*>(smem_raw);          // [kRows][32] row-FFT twiddles W_1024^(l*b), W_1024^(l*8a)
  c2* Ys = reinterpret_cast<c2*>(tab4 + C::kRows * 32);        // [M1][kRowC2]
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int i = t; i < C::kRows * 32; i += T) {
    const int r = i >> 5, l = i & 31;
    const float2 w = r < 7 ? cis_neg_w(2.f * static_cast<float>(l * (r + 1)) / M2) : cis_neg_w(2.f * static_cast<float>(l * 8 * (r - 6)) / M2);
    tab4[i] = make_float4(w.x, w.y, -w.y, w.y);
  }
  TwP wpb;
  {
    float s, c;
    sincospif(static_cast<float>(t) / M, &s, &c);  // W_N^t
    wpb = TwP{-0.5f * s, pk(0.5f * c, -0.5f * c)};  // 0.5 * (-i) * W_N^t = (-s, -c)/2 in the (wr, (-wi, wi)) form
  }
  // (Tried: requesting the CTA's next frame before the untangling / store phase, 32 packed values per thread held in
  // registers. It spills at 128 registers and measured 5-12 % slower at all three lengths, profiles/fft_prefetch_sweep_r01.txt.)
  for (int64_t frame = blockIdx.x; frame < batch; frame += gridDim.x) {
    const float* x = in + frame * (2 * M);
    __syncthreads();  // tables ready / previous frame's step C done
    // ---- step A: columns ----
#pragma unroll
    for (int j = 0; j < M2 / T; ++j) {
      const int c = t + T * j;
      c2 v[M1];
#pragma unroll
      for (int n1 = 0; n1 < M1; ++n1) v[bitrev(n1, ilog2(M1))] = *reinterpret_cast<const c2*>(x + 2 * (n1 * M2 + c));
      DitP<M1, M1, 0>::run(v);
      c2 w[M1];
      {
        float s, cs;
        sincospif(2.f * static_cast<float>(c) / M, &s, &cs);  // W_M^c
        w[1] = pk(cs, -s);
      }
      cpow_all<M1>(w, std::make_integer_sequence<int, M1 - 2>{});
#pragma unroll
      for (int k1 = 1; k1 < M1; ++k1) {
        const float2 a = upk(w[k1]);
        v[k1] = cmul_tw(v[k1], TwP{a.x, pk(-a.y, a.y)});
      }
#pragma unroll
      for (int k1 = 0; k1 < M1; ++k1) Ys[k1 * C::kRowC2 + c] = v[k1];
    }
    __syncthreads();
    // ---- step B: warp k1 transforms row k1 (1024 points = 32 x 32), in place ----
    {
      c2* row = Ys + warp * C::kRowC2;
      c2 v[32];
#pragma unroll
      for (int n1 = 0; n1 < 32; ++n1) v[bitrev(n1, 5)] = row[32 * n1 + lane];
      __syncwarp();
      DitP<32, 32, 0>::run(v);
      tww_chain<32, 0>(v, tab4, lane, make_twp(tab4[lane]));
#pragma unroll
      for (int m = 0; m < 16; ++m) *reinterpret_cast<ulonglong2*>(row + m * kXStride + 2 * lane) = make_ulonglong2(v[2 * m], v[2 * m + 1]);
      __syncwarp();
      c2 A[32];
#pragma unroll
      for (int n2 = 0; n2 < 32; ++n2) A[bitrev(n2, 5)] = row[(lane >> 1) * kXStride + 2 * n2 + (lane & 1)];
      __syncwarp();
      DitP<32, 32, 0>::run(A);
#pragma unroll
      for (int k2 = 0; k2 < 32; ++k2) row[lane + 32 * k2] = A[k2];  // Z[k1 + M1*(lane + 32*k2')] at position k2 = lane + 32*k2'
    }
    __syncthreads();
    {
      // ---- step C: untangle, k = t + T*i ----
      float2* o = out + frame * (M + 1);
      cta_untangle_all<M1>(Ys, o, wpb, t, std::make_integer_sequence<int, M / (2 * T)>{});
      if (t == 0) {
        const float2 z = upk(Ys[((M / 2) % M1) * C::kRowC2 + (M / 2) / M1]);
        o[M / 2] = make_float2(z.x, -z.y);
      }
    }
  }
}

template <int M1>
int launch_cta(ca_ctx* ctx, const float* d_in, int64_t batch, float* d_out, cudaStream_t stream) {
  using C = CtaFftCfg<M1>;
  auto kern = rfft_cta_kernel<M1>;
  const size_t smem = sizeof(float4) * C::kRows * 32 + sizeof(c2) * C::kRowC2 * M1;
  int per_sm = 0;
  CA_KERNEL_SETUP(ctx, kern, C::T, smem, per_sm);
  if (per_sm < 1) per_sm = 1;
  const int64_t grid = std::min<int64_t>(static_cast<int64_t>(ctx->sm_count) * per_sm, batch);
  kern<<<static_cast<unsigned>(grid), C::T, smem, stream>>>(d_in, reinterpret_cast<float2*>(d_out), batch);
  CA_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches += 1;
  return CA_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// N >= 65536, two passes through HBM: M = N/2 = MA x 1024 complex points, n = 1024*r + c, k = kA + MA*k2.
//   pass 1  big2_cols_kernel: a CTA takes 16 adjacent columns c (128-byte runs in HBM) and all MA rows: MA = A x B point FFT
//           down each column in two register stages with one shared-memory exchange, twiddle W_M^(c*kA), work[kA][c]
//   pass 2  big2_rows_kernel: a CTA takes 8 adjacent rows kA, the 8 mirrored rows MA - kA (and row MA/2 in the last CTA of a
//           frame): one warp per row runs the 1024-point warp-level FFT in shared memory, then the whole CTA untangles
//           X[k], X[M-k] from the two row sets and writes them in runs of 8 consecutive bins (64 bytes)
// 32M bytes moved for 16M algorithmic: the ceiling is half of the HBM roofline. (The three-pass path this replaced — thread per
// column MA-point FFTs, 8192-point rows in place, untangling transposition — reached 24-27 %, profiles/README.md.)
// (Tried and removed: ONE persistent launch that runs both passes with in-order tickets and keeps the work frames in an L2-sized
// ring. Correct, but 25-70 % slower, profiles/fft_fused_sweep_r01.txt: both passes must share one CTA shape, which leaves the
// column pass at 9-17 warps per SM instead of 30, and the row pass is issue-bound rather than HBM-bound, profiles/ncu_big2_r01.json.)
// ---------------------------------------------------------------------------------------------------------------
constexpr int kBig2Row = 1024;

// CW: columns per CTA (CW*8-byte runs in HBM)
template <int A, int B, int CW>
__global__ void __launch_bounds__(256) big2_cols_kernel(const float2* __restrict__ zin, float2* __restrict__ work) {
  constexpr int MA = A * B, R = kBig2Row, M = MA * R, QT = 256 / CW;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  c2* u = reinterpret_cast<c2*>(smem_raw);  // [A][B][CW]: element (k1, r2, column)
  const int t = threadIdx.x, cl = t % CW, q = t / CW;
  const int c = blockIdx.x * CW + cl;
  const c2* src = reinterpret_cast<const c2*>(zin) + static_cast<int64_t>(blockIdx.y) * M + c;
  c2* dst = reinterpret_cast<c2*>(work) + static_cast<int64_t>(blockIdx.y) * M + c;
  // ---- stage 1: r = B*r1 + r2; A-point FFT over r1 for every (column, r2), twiddle W_MA^(r2*k1) ----
  for (int r2 = q; r2 < B; r2 += QT) {
    c2 v[A];
#pragma unroll
    for (int r1 = 0; r1 < A; ++r1) v[bitrev(r1, ilog2(A))] = __ldg(src + static_cast<int64_t>(B * r1 + r2) * R);
    DitP<A, A, 0>::run(v);
    if (r2 != 0) {
      c2 w[A];
      float sn, cs;
      sincospif(2.f * static_cast<float>(r2) / MA, &sn, &cs);
      w[1] = pk(cs, -sn);
      if constexpr (A > 2) cpow_all<A>(w, std::make_integer_sequence<int, A - 2>{});
#pragma unroll
      for (int k1 = 1; k1 < A; ++k1) {
        const float2 a = upk(w[k1]);
        v[k1] = cmul_tw(v[k1], TwP{a.x, pk(-a.y, a.y)});
      }
    }
#pragma unroll
    for (int k1 = 0; k1 < A; ++k1) u[(k1 * B + r2) * CW + cl] = v[k1];
  }
  __syncthreads();
  // ---- stage 2: B-point FFT over r2 for every (column, k1): Y[k1 + A*k2], twiddle W_M^(c*(k1 + A*k2)) ----
  for (int k1 = q; k1 < A; k1 += QT) {
    c2 y[B];
#pragma unroll
    for (int r2 = 0; r2 < B; ++r2) y[bitrev(r2, ilog2(B))] = u[(k1 * B + r2) * CW + cl];
    DitP<B, B, 0>::run(y);
    c2 pw[B];  // pw[j] = W_M^(c*A*j) by binary powers (depth <= log2 B)
    {
      float sn, cs;
      sincospif(2.f * static_cast<float>(c * A) / M, &sn, &cs);  // exact argument: c*A < 2^15, M a power of two
      pw[1] = pk(cs, -sn);
    }
    if constexpr (B > 2) cpow_all<B>(pw, std::make_integer_sequence<int, B - 2>{});
    TwP base;
    {
      float sn, cs;
      sincospif(2.f * static_cast<float>(c * k1) / M, &sn, &cs);
      base = TwP{cs, pk(sn, -sn)};  // W_M^(c*k1) = (cs, -sn) in the (wr, (-wi, wi)) form
    }
#pragma unroll
    for (int k2 = 0; k2 < B; ++k2) {
      c2 val = cmul_tw(y[k2], base);
      if (k2) {
        const float2 a = upk(pw[k2]);
        val = cmul_tw(val, TwP{a.x, pk(-a.y, a.y)});
      }
      dst[static_cast<int64_t>(k1 + A * k2) * R] = val;
    }
  }
}

// G: adjacent rows per CTA (G*8-byte runs in HBM)
template <int MA, int G>
__global__ void __launch_bounds__(32 * (2 * G + 1)) big2_rows_kernel(const float2* __restrict__ work, float2* __restrict__ out) {
  constexpr int R = kBig2Row, M = MA * R, CPF = MA / (2 * G), T = 32 * (2 * G + 1);
  constexpr int kRowC2 = 16 * kXStride + 2;  // = 2 (mod 16): 8 rows x 2 consecutive k2 cover the 32 banks once (G = 8)
  constexpr int kTabRows = 7 + 3;            // row-FFT tables: wb[1..7], wa[1..3] (WarpFftCfg<32>)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* tab4 = reinterpret_cast<float4*>(smem_raw);
  c2* rows = reinterpret_cast<c2*>(tab4 + kTabRows * 32);  // [2G+1][kRowC2]
  float2* tw2r = reinterpret_cast<float2*>(rows + (2 * G + 1) * kRowC2);  // [R] W_2R^k2 for the untangling step
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  static_assert(T % G == 0, "a thread keeps its kA through the untangling loop");
  for (int i = t; i < R; i += T) tw2r[i] = cis_neg_w(static_cast<float>(i) / R);  // exp(-i pi k2 / R)
  const int64_t frame = blockIdx.x / CPF;
  const int cc = blockIdx.x % CPF;
  for (int i = t; i < kTabRows * 32; i += T) {
    const int r = i >> 5, l = i & 31;
    const float2 w = r < 7 ? cis_neg_w(2.f * static_cast<float>(l * (r + 1)) / R) : cis_neg_w(2.f * static_cast<float>(l * 8 * (r - 6)) / R);
    tab4[i] = make_float4(w.x, w.y, -w.y, w.y);
  }
  // slot -> row of the frame: [0, G) the CTA's own rows, [G, 2G) their mirrors MA - kA, 2G the self-paired row MA/2
  int kA_row;
  bool valid = true;
  if (warp < G) kA_row = cc * G + warp;
  else if (warp < 2 * G) {
    kA_row = (MA - cc * G - (warp - G)) & (MA - 1);
    valid = !(cc == 0 && warp == G);  // row 0 pairs with itself (slot 0)
  } else {
    kA_row = MA / 2;
    valid = cc == CPF - 1;
  }
  c2 v[32];
  if (valid) {
    const c2* src = reinterpret_cast<const c2*>(work) + frame * M + static_cast<int64_t>(kA_row) * R;
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) v[bitrev(n1, 5)] = __ldg(src + 32 * n1 + lane);
  }
  __syncthreads();  // tables ready
  if (valid) {      // 1024 = 32 x 32 point FFT of the row, warp-level (as step B of rfft_cta_kernel)
    c2* row = rows + warp * kRowC2;
    DitP<32, 32, 0>::run(v);
    tww_chain<32, 0>(v, tab4, lane, make_twp(tab4[lane]));
#pragma unroll
    for (int m = 0; m < 16; ++m) *reinterpret_cast<ulonglong2*>(row + m * kXStride + 2 * lane) = make_ulonglong2(v[2 * m], v[2 * m + 1]);
    __syncwarp();
    c2 Acol[32];
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) Acol[bitrev(n2, 5)] = row[(lane >> 1) * kXStride + 2 * n2 + (lane & 1)];
    __syncwarp();
    DitP<32, 32, 0>::run(Acol);
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) row[lane + 32 * k2] = Acol[k2];  // Z[kA + MA*k2] at position k2
  }
  __syncthreads();
  // ---- untangle: X[k] and X[M-k], k = kA + MA*k2; consecutive threads take consecutive kA ----
  // 0.5 * (-i) * W_N^k = [0.5 * (-i) * W_N^kA] * W_2R^k2: kA is fixed per thread (T is a multiple of G), W_2R^k2 comes from a
  // table in shared memory, so an output pair costs one complex multiply instead of a sincospif
  float2* o = out + frame * (M + 1);
  auto emit = [&](int64_t k, c2 zk, c2 zm, const TwP& wp) {
    const c2 E = add2(zk, cconj(zm));
    const c2 Tt = cmul_tw(csub(zk, cconj(zm)), wp);
    o[k] = upk(fma2(bc(0.5f), E, Tt));
    o[M - k] = upk(cconj(fma2(bc(0.5f), E, cneg(Tt))));
  };
  auto twiddle = [&](int kA, int k2) {  // 0.5 * (-i) * W_N^(kA + MA*k2) in the (wr, (-wi, wi)) form
    float sn, cs;
    sincospif(static_cast<float>(kA) / M, &sn, &cs);  // W_N^kA = (cs, -sn)
    const float2 w2 = tw2r[k2];                        // W_2R^k2
    const float wr = cs * w2.x + sn * w2.y, wi = cs * w2.y - sn * w2.x;  // W_N^k
    return TwP{0.5f * wi, pk(0.5f * wr, -0.5f * wr)};  // (-i) * w = (wi, -wr)
  };
  {
    const int i = t & (G - 1);
    const int kA = cc * G + i;
    float sn, cs;
    sincospif(static_cast<float>(kA) / M, &sn, &cs);
    for (int item = t; item < G * R; item += T) {  // T % G == 0: item & (G-1) == i
      const int k2 = item / G;
      const float2 w2 = tw2r[k2];
      const float wr = cs * w2.x + sn * w2.y, wi = cs * w2.y - sn * w2.x;
      const TwP wp{0.5f * wi, pk(0.5f * wr, -0.5f * wr)};
      if (kA == 0) {  // M - k = MA*(R - k2): the same row
        if (k2 > R / 2) continue;
        emit(static_cast<int64_t>(MA) * k2, rows[k2], rows[(R - k2) & (R - 1)], wp);
      } else {        // M - k = (MA - kA) + MA*(R - 1 - k2)
        emit(kA + static_cast<int64_t>(MA) * k2, rows[i * kRowC2 + k2], rows[(G + i) * kRowC2 + R - 1 - k2], wp);
      }
    }
  }
  if (cc == CPF - 1) {  // kA = MA/2: M - k = MA/2 + MA*(R - 1 - k2)
    const c2* rs = rows + 2 * G * kRowC2;
    for (int k2 = t; k2 < R / 2; k2 += T) emit(MA / 2 + static_cast<int64_t>(MA) * k2, rs[k2], rs[R - 1 - k2], twiddle(MA / 2, k2));
  }
}

template <int A, int B, int CW, int G>
int launch_big2(ca_ctx* ctx, const float* d_in, int64_t batch, float* d_out, cudaStream_t stream) {
  constexpr int MA = A * B, M = MA * kBig2Row;
  static_assert(MA % (2 * G) == 0, "rows per frame must split into CTAs of 2G rows");
  const size_t want = static_cast<size_t>(512) << 20;  // bounded work buffer: frames are processed in groups
  const int64_t group = std::max<int64_t>(1, std::min<int64_t>(batch, static_cast<int64_t>(want / (sizeof(float2) * M))));
  const size_t need = sizeof(float2) * M * static_cast<size_t>(group);
  if (ctx->fft_work_bytes < need) {
    if (ctx->d_fft_work) cudaFree(ctx->d_fft_work);
    ctx->d_fft_work = nullptr;
    ctx->fft_work_bytes = 0;
    CA_CUDA_OK(ctx, cudaMalloc(&ctx->d_fft_work, need));
    ctx->fft_work_bytes = need;
  }
  auto cols = big2_cols_kernel<A, B, CW>;
  auto rowsk = big2_rows_kernel<MA, G>;
  const size_t smem_cols = sizeof(c2) * MA * CW;
  const size_t smem_rows = sizeof(float4) * 10 * 32 + sizeof(c2) * (16 * kXStride + 2) * (2 * G + 1) + sizeof(float2) * kBig2Row;
  CA_CUDA_OK(ctx, cudaFuncSetAttribute(cols, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_cols)));
  CA_CUDA_OK(ctx, cudaFuncSetAttribute(rowsk, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_rows)));
  for (int64_t b0 = 0; b0 < batch; b0 += group) {
    const int64_t nb = std::min(group, batch - b0);
    const float2* zin = reinterpret_cast<const float2*>(d_in) + b0 * M;
    float2* out = reinterpret_cast<float2*>(d_out) + b0 * (M + 1);
    cols<<<dim3(kBig2Row / CW, static_cast<unsigned>(nb)), 256, smem_cols, stream>>>(zin, ctx->d_fft_work);
    CA_CUDA_OK(ctx, cudaGetLastError());
    rowsk<<<static_cast<unsigned>(nb * (MA / (2 * G))), 32 * (2 * G + 1), smem_rows, stream>>>(ctx->d_fft_work, out);
    CA_CUDA_OK(ctx, cudaGetLastError());
    ctx->launches += 2;
  }
  return CA_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// N = 2^15 ... 2^17 in ONE pass through HBM, the frame held on chip across a thread-block cluster (CA_FFT_CLUSTER=0 restores the
// CTA-per-frame / two-pass kernels above). M = N/2 = P x Q complex points, n = Q*r + c, k = kA + P*k2, with P = A*B rows and
// Q = 32*R1 points per row chosen "square" so that both sides of HBM see long runs: a CTA of the C-CTA cluster reads a slab of P
// rows x 16-32 columns (128-256-byte runs) and owns 2G = P/C rows of the transform (G adjacent rows and their G mirrors P - kA;
// CTA 0 keeps row P/2 in the slot of row 0's mirror), so the untangled bins leave in runs of G >= 8 (64-128 bytes):
//   columns  as big2_cols_kernel: A-point FFTs, exchange through the CTA's own shared memory, B-point FFTs, twiddle W_M^(c*kA);
//            the 32 results per thread stay in registers across a cluster barrier (every CTA has started and is done with its
//            exchange buffer, which is the row storage) and then go to the owners of their rows through distributed shared
//            memory: st.shared::cluster + a second cluster barrier, or st.async counted on the owner's mbarrier (ASYNC),
//   rows     one warp per 32/R1 rows (the warp-level transform of rfft_warp_kernel: R1 points per lane, 32-point columns),
//   untangle X[k], X[M-k] from a row and its mirror, both local; consecutive threads take consecutive kA of the CTA's row set.
// 16M bytes per frame instead of the 32M of the two-pass form; 35-40 KB of shared memory, 128 threads and 96 registers per CTA,
// so five CTAs of different frames share an SM and one's loads overlap another's row transforms.
// Measured and removed (profiles/fft_cluster_sweep_r02.txt): rows of 1024 points with P = 16/32/64-point columns in registers
// (no column exchange, but G = 2-4: 16-32-byte output runs cost 2.6x the bytes on the L1->L2 path): 39-48 % of the HBM peak.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  return ra;
}
// 8-byte store into a peer CTA's shared memory that also counts 8 bytes on the peer's mbarrier when it lands
__device__ __forceinline__ void st_async_c2(uint32_t remote_addr, uint32_t remote_mbar, c2 v) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr), "l"(v), "r"(remote_mbar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint32_t mbar, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(mbar),
      "r"(parity)
      : "memory");
}
template <int A_, int B_, int R1_, int C_, bool ASYNC_ = false>
struct Cluster2Cfg {
  static constexpr bool ASYNC = ASYNC_;  // rows delivered by st.async counted on the owner's mbarrier, or by plain stores + a cluster barrier
  static constexpr int A = A_, B = B_, R1 = R1_, C = C_;
  static constexpr int P = A * B, Q = 32 * R1, M = P * Q;
  static constexpr int G = P / (2 * C);        // adjacent rows per CTA (and as many mirrored ones)
  static constexpr int FPW = 32 / R1;          // rows per warp
  static constexpr int CWC = Q / C;            // columns per CTA (16 or 32: CWC*8-byte runs in HBM)
  static constexpr int SUB = 32 / CWC;         // row groups per warp in the column phase
  static constexpr int WARPS = 2 * G / FPW, T = 32 * WARPS, QT = WARPS * SUB;
  // Row storage [2G][RS], RS a multiple of 16 (the warp's exchange layout fits in its rows). The transformed rows stay in place with
  // the 16-point groups of row slot s rotated by rho(s), so that the last row pass (16/R1 slots x R1 columns per half-warp) and the
  // untangling loads (G slots x 16/G values of k2) both touch each of the sixteen 8-byte banks once: R1 = 8, G = 8:
  // rho = 2*((s >> 1) + 4*(s & 1)); R1 = 4, G = 16: rho = 4*(s & 3) + ((s >> 2) & 3). Slots 0 and G (the self-paired rows of
  // CTA 0) keep rho = 0. (Measured against it, profiles/fft_cluster_sweep_r02.txt: the rows stored transposed behind one more CTA
  // barrier, and in place without the rotation, i.e. with 2-4-way conflicts on the last store: 1-3 % slower.)
  static constexpr int RS = Q + 16;
  static __device__ __forceinline__ int rho(int s) {
    if constexpr (R1 == 8) return 2 * (((s & 7) >> 1) + 4 * (s & 1));
    else return 4 * (s & 3) + ((s >> 2) & 3);
  }
  using W = WarpFftCfg<R1>;
  static constexpr size_t smem = 16 + sizeof(float4) * W::kRows * 32 + sizeof(c2) * RS * 2 * G;  // mbarrier, row tables, rows
  static constexpr int kMinCtas = 5;  // 96 registers; six CTAs per SM (80 registers, 80-200 bytes of spills) measured 15 % slower
  static_assert(CWC == 32 || CWC == 16, "a CTA takes 16 or 32 columns");
  static_assert((G == 8 && R1 == 8) || (G == 16 && R1 == 4), "bank rotation of the row storage");
  static_assert(FPW * RS >= FPW * (R1 / 2) * kXStride && RS * 2 * G >= P * CWC, "exchange layouts fit in the row storage");
  static_assert(A % QT == 0 && B % QT == 0 && (A / QT) * B == 32 && (FPW * RS) % 2 == 0, "32 points per thread");
};

// row kA of the frame -> (CTA of the cluster, row slot)
template <typename Cf>
__device__ __forceinline__ void cluster2_row_dest(int kA, uint32_t& rank, int& slot) {
  constexpr int P = Cf::P, G = Cf::G;
  if (kA < P / 2) {
    rank = kA / G;
    slot = kA % G;
  } else if (kA == P / 2) {
    rank = 0;
    slot = G;
  } else {
    const int m = P - kA;
    rank = m / G;
    slot = G + m % G;
  }
}

template <typename Cf, int I>
__device__ __forceinline__ void cluster2_untangle(const c2* __restrict__ rows, float2* __restrict__ o, const TwP& wpb, int kA, int i, int k2_0,
                                                  int zk_sw, int zm_sw) {
  constexpr int P = Cf::P, Q = Cf::Q, M = Cf::M, G = Cf::G, RS = Cf::RS, KS = Cf::T / G;
  static_assert(KS == 16 || KS == 8, "k2_0 + (KS*I & 15) stays inside a 16-point group");
  const int k2 = k2_0 + KS * I;
  c2 zk, zm;
  int64_t k;
  if (kA == 0) {  // M - k = P*(Q - k2): the same row (slot 0, not rotated)
    if (k2 > Q / 2) return;
    zk = rows[k2];
    zm = rows[(Q - k2) & (Q - 1)];
    k = static_cast<int64_t>(P) * k2;
  } else {        // M - k = (P - kA) + P*(Q - 1 - k2); k2 and Q-1-k2 inside their rotated groups: zk_sw = k2_0 + rho, zm_sw = 15 - k2_0 + rho
    constexpr int g0 = (KS * I) & ~15, w0 = (KS * I) & 15;
    zk = rows[i * RS + g0 + ((zk_sw + w0) & 15)];
    zm = rows[(G + i) * RS + Q - 16 - g0 + ((zm_sw - w0) & 15)];
    k = kA + static_cast<int64_t>(P) * k2;
  }
  const c2 E = add2(zk, cconj(zm));
  c2 D = csub(zk, cconj(zm));
  if constexpr (I != 0) D = cmul_imm<Tw<2 * Q, KS * I>>(D);  // W_N^(P*KS*I) = W_2Q^(KS*I)
  const c2 Tt = cmul_tw(D, wpb);
  o[k] = upk(fma2(bc(0.5f), E, Tt));
  o[M - k] = upk(cconj(fma2(bc(0.5f), E, cneg(Tt))));
}
template <typename Cf, int... I>
__device__ __forceinline__ void cluster2_untangle_all(const c2* rows, float2* o, const TwP& wpb, int kA, int i, int k2_0, int zk_sw, int zm_sw,
                                                      std::integer_sequence<int, I...>) {
  (cluster2_untangle<Cf, I>(rows, o, wpb, kA, i, k2_0, zk_sw, zm_sw), ...);
}

template <typename Cf>
__global__ void __launch_bounds__(Cf::T, Cf::kMinCtas) rfft_cluster2_kernel(const float* __restrict__ in, float2* __restrict__ out) {
  using W = typename Cf::W;
  constexpr int A = Cf::A, B = Cf::B, R1 = Cf::R1, C = Cf::C, P = Cf::P, Q = Cf::Q, M = Cf::M, G = Cf::G, FPW = Cf::FPW, T = Cf::T,
                QT = Cf::QT, RS = Cf::RS, NI1 = B / QT, NI2 = A / QT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* tab4 = reinterpret_cast<float4*>(smem_raw + 16);    // [W::kRows][32] row-transform twiddles
  c2* rows = reinterpret_cast<c2*>(tab4 + W::kRows * 32);     // the column exchange u[A][B][CWC], then the rows [2G][RS]
  constexpr int CWC = Cf::CWC;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int q = warp * Cf::SUB + lane / CWC, cl = lane % CWC;  // column phase: row group and column of the slab
  const uint32_t rank = cluster_ctarank();
  // the rows arrive from the whole cluster (this CTA included) as st.async stores that count their bytes on this mbarrier
  const uint32_t mbar = static_cast<uint32_t>(__cvta_generic_to_shared(smem_raw));
  if (Cf::ASYNC && t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(static_cast<uint32_t>(sizeof(c2) * 2 * G * Q)) : "memory");
  }
  const int64_t frame = blockIdx.x / C;
  const int c = static_cast<int>(rank) * CWC + cl;  // this thread's column
  const c2* src = reinterpret_cast<const c2*>(in) + frame * M + c;
  // ---- columns, stage 1: r = B*r1 + r2; A-point FFT over r1 for r2 = q, q + QT, ...; twiddle W_P^(r2*k1) ----
  c2 v[NI1][A];
#pragma unroll
  for (int it = 0; it < NI1; ++it)
#pragma unroll
    for (int r1 = 0; r1 < A; ++r1) v[it][bitrev(r1, ilog2(A))] = __ldg(src + (B * r1 + q + QT * it) * Q);
  for (int i = t; i < (W::kRows - 1) * 32; i += T) {
    const int r = i >> 5, l = i & 31;
    const float2 w = r < W::B - 1 ? cis_neg_w(2.f * static_cast<float>(l * (r + 1)) / Q) : cis_neg_w(2.f * static_cast<float>(l * W::B * (r - (W::B - 2))) / Q);
    tab4[i] = make_float4(w.x, w.y, -w.y, w.y);
  }
  c2* u = rows;
#pragma unroll
  for (int it = 0; it < NI1; ++it) {
    const int r2 = q + QT * it;
    DitP<A, A, 0>::run(v[it]);
    if (r2 != 0) {
      c2 w[A];
      float sn, cs;
      sincospif(2.f * static_cast<float>(r2) / P, &sn, &cs);
      w[1] = pk(cs, -sn);
      if constexpr (A > 2) cpow_all<A>(w, std::make_integer_sequence<int, A - 2>{});
#pragma unroll
      for (int k1 = 1; k1 < A; ++k1) {
        const float2 a = upk(w[k1]);
        v[it][k1] = cmul_tw(v[it][k1], TwP{a.x, pk(-a.y, a.y)});
      }
    }
#pragma unroll
    for (int k1 = 0; k1 < A; ++k1) u[(k1 * B + r2) * CWC + cl] = v[it][k1];
  }
  __syncthreads();
  // ---- stage 2: B-point FFT over r2 for k1 = q, q + QT, ...: Y[k1 + A*k2], twiddle W_M^(c*(k1 + A*k2)); results stay in registers ----
  c2 y[NI2][B];
#pragma unroll
  for (int it = 0; it < NI2; ++it)
#pragma unroll
    for (int r2 = 0; r2 < B; ++r2) y[it][bitrev(r2, ilog2(B))] = u[((q + QT * it) * B + r2) * CWC + cl];
#pragma unroll
  for (int it = 0; it < NI2; ++it) DitP<B, B, 0>::run(y[it]);  // (consumes the loads from u before the barrier is signalled)
  cluster_arrive_relaxed();  // this CTA runs, its mbarrier is initialised (fence above) and it is done with its exchange buffer
  {
    c2 pw[B];  // pw[j] = W_M^(c*A*j)
    {
      float sn, cs;
      sincospif(2.f * static_cast<float>(c * A) / M, &sn, &cs);
      pw[1] = pk(cs, -sn);
    }
    if constexpr (B > 2) cpow_all<B>(pw, std::make_integer_sequence<int, B - 2>{});
#pragma unroll
    for (int it = 0; it < NI2; ++it) {
      float sn, cs;
      sincospif(2.f * static_cast<float>(c * (q + QT * it)) / M, &sn, &cs);
      const TwP base{cs, pk(sn, -sn)};
#pragma unroll
      for (int k2 = 0; k2 < B; ++k2) {
        c2 val = cmul_tw(y[it][k2], base);
        if (k2) {
          const float2 a = upk(pw[k2]);
          val = cmul_tw(val, TwP{a.x, pk(-a.y, a.y)});
        }
        y[it][k2] = val;
      }
    }
  }
  cluster_wait();
  {
    const uint32_t rows_addr = static_cast<uint32_t>(__cvta_generic_to_shared(rows));
#pragma unroll
    for (int it = 0; it < NI2; ++it)
#pragma unroll
      for (int k2 = 0; k2 < B; ++k2) {
        uint32_t dr;
        int slot;
        cluster2_row_dest<Cf>(q + QT * it + A * k2, dr, slot);
        const uint32_t rbase = mapa_u32(rows_addr, dr);  // the peer's window is a translation of this one: same offsets
        if constexpr (Cf::ASYNC) st_async_c2(rbase + static_cast<uint32_t>(sizeof(c2)) * (slot * RS + c), rbase - (rows_addr - mbar), y[it][k2]);
        else asm volatile("st.shared::cluster.b64 [%0], %1;" ::"r"(rbase + static_cast<uint32_t>(sizeof(c2)) * (slot * RS + c)), "l"(y[it][k2]) : "memory");
      }
  }
  if constexpr (Cf::ASYNC) {
    mbar_wait_parity(mbar, 0);  // all 2G rows of this CTA have landed
  } else {
    cluster_arrive();  // the rows are complete once every CTA has stored its columns
    cluster_wait();
  }
  // ---- rows: warp q transforms the FPW rows in slots [q*FPW, (q+1)*FPW) (Q = R1 x 32 points each), in place ----
  {
    c2* buf = rows + warp * FPW * RS;
    c2 w1[FPW][R1];
#pragma unroll
    for (int fw = 0; fw < FPW; ++fw)
#pragma unroll
      for (int n1 = 0; n1 < R1; ++n1) w1[fw][bitrev(n1, ilog2(R1))] = buf[fw * RS + 32 * n1 + lane];
    __syncwarp();
#pragma unroll
    for (int fw = 0; fw < FPW; ++fw) {
      DitP<R1, R1, 0>::run(w1[fw]);
      tww_chain<R1, 0>(w1[fw], tab4, lane, make_twp(tab4[lane]));
      c2* xb = buf + fw * (R1 / 2) * kXStride;
#pragma unroll
      for (int m = 0; m < R1 / 2; ++m)
        *reinterpret_cast<ulonglong2*>(xb + m * kXStride + 2 * lane) = make_ulonglong2(w1[fw][2 * m], w1[fw][2 * m + 1]);
    }
    __syncwarp();
    c2 Acol[32];
    const int col = lane & (R1 - 1), fr = lane / R1;
    {
      const c2* xb = buf + fr * (R1 / 2) * kXStride;
#pragma unroll
      for (int n2 = 0; n2 < 32; ++n2) Acol[bitrev(n2, 5)] = xb[(col >> 1) * kXStride + 2 * n2 + (col & 1)];
    }
    DitP<32, 32, 0>::run(Acol);
    __syncwarp();
    c2* Z = buf + fr * RS;
    const int cr = col + Cf::rho(warp * FPW + fr);  // position col + R1*k2 inside its rotated 16-point group
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) Z[((R1 * k2) & ~15) + ((cr + ((R1 * k2) & 15)) & 15)] = Acol[k2];  // Z[kA + P*k2]
  }
  __syncthreads();
  // ---- untangle: consecutive threads take consecutive kA; k2 = k2_0 + (T/G)*I ----
  float2* o = out + frame * (M + 1);
  {
    const int i = t & (G - 1), k2_0 = t / G;
    const int kA = static_cast<int>(rank) * G + i;
    float sn, cs;
    sincospif(static_cast<float>(kA + P * k2_0) / M, &sn, &cs);  // W_N^(kA + P*k2_0) = (cs, -sn)
    const TwP wpb{-0.5f * sn, pk(0.5f * cs, -0.5f * cs)};
    const int rho = Cf::rho(i);
    cluster2_untangle_all<Cf>(rows, o, wpb, kA, i, k2_0, k2_0 + rho, 15 - k2_0 + rho, std::make_integer_sequence<int, Q * G / T>{});
  }
  if (rank == 0) {  // kA = P/2 (slot G, not rotated): M - k = P/2 + P*(Q - 1 - k2)
    const c2* rs = rows + G * RS;
    for (int k2 = t; k2 < Q / 2; k2 += T) {
      float sn, cs;
      sincospif(static_cast<float>(P / 2 + P * k2) / M, &sn, &cs);
      const TwP wp{-0.5f * sn, pk(0.5f * cs, -0.5f * cs)};
      const c2 zk = rs[k2], zm = rs[Q - 1 - k2];
      const c2 E = add2(zk, cconj(zm));
      const c2 Tt = cmul_tw(csub(zk, cconj(zm)), wp);
      const int64_t k = P / 2 + static_cast<int64_t>(P) * k2;
      o[k] = upk(fma2(bc(0.5f), E, Tt));
      o[M - k] = upk(cconj(fma2(bc(0.5f), E, cneg(Tt))));
    }
  }
}

// Can this device co-schedule one cluster of this shape (the 16-CTA clusters of 2^17 are a non-portable size)? Asked once per
// instantiation and device; a "no" sends the length to the CTA-per-frame / two-pass kernels instead of failing the launch.
template <typename Cf>
bool cluster2_available(ca_ctx* ctx) {
  static int s_dev = -1;
  static bool s_ok = false;
  if (s_dev != ctx->device) {
    auto kern = rfft_cluster2_kernel<Cf>;
    s_ok = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(Cf::smem)) == cudaSuccess;
    if (s_ok && Cf::C > 8) s_ok = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    if (s_ok) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(Cf::C);
      cfg.blockDim = dim3(Cf::T);
      cfg.dynamicSmemBytes = Cf::smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = Cf::C;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int n = 0;
      s_ok = cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n >= 1;
    }
    if (!s_ok) (void)cudaGetLastError();
    s_dev = ctx->device;
  }
  return s_ok;
}

template <typename Cf>
int launch_cluster2(ca_ctx* ctx, const float* d_in, int64_t batch, float* d_out, cudaStream_t stream) {
  auto kern = rfft_cluster2_kernel<Cf>;
  static int s_dev = -1;
  if (s_dev != ctx->device) {
    CA_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(Cf::smem)));
    if (Cf::C > 8) CA_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    s_dev = ctx->device;
  }
  const int64_t max_frames = (static_cast<int64_t>(1) << 31) / Cf::C - 1;
  for (int64_t b0 = 0; b0 < batch; b0 += max_frames) {
    const int64_t nb = std::min(max_frames, batch - b0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(nb * Cf::C));
    cfg.blockDim = dim3(Cf::T);
    cfg.dynamicSmemBytes = Cf::smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = Cf::C;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    // CTAs of a cluster packed on as few SMs as the scheduler likes (measured: +3 % on 4- and 8-CTA clusters, -7 % on 16)
    at[1].id = cudaLaunchAttributeClusterSchedulingPolicyPreference;
    at[1].val.clusterSchedulingPolicyPreference = Cf::C <= 8 ? cudaClusterSchedulingPolicyLoadBalancing : cudaClusterSchedulingPolicySpread;
    cfg.numAttrs = 2;
    CA_CUDA_OK(ctx, cudaLaunchKernelEx(&cfg, kern, d_in + b0 * 2 * Cf::M, reinterpret_cast<float2*>(d_out) + b0 * (Cf::M + 1)));
    ctx->launches += 1;
  }
  return CA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Fused analysis (window -> real FFT -> |X|^2 -> peaks) for the FFT lengths other than 4096: the transforms above with
// the window applied on load and the untangling step producing |X|^2 (HBM row + a shared-memory copy) instead of complex
// bins; the peak list is extracted from the shared-memory copy exactly as in stft.cu (strict local maxima above the
// threshold, ascending, parabolic interpolation on log2 ratios). One pass over HBM instead of the three of
// window_frames_kernel -> K1 -> sqmag_peaks_kernel.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ const float* frame_ptr(const StftArgs& p, int64_t f) {
  if (p.total_frames < (1ll << 31)) {  // the usual case: a 32-bit divide (the 64-bit one is a ~75-instruction subroutine)
    const unsigned uf = static_cast<unsigned>(f), fps = static_cast<unsigned>(p.frames_per_stream);
    const unsigned stream = uf / fps, fr = uf - stream * fps;
    return p.x + stream * p.stream_stride + static_cast<int64_t>(fr) * p.hop;
  }
  const int64_t stream = f / p.frames_per_stream;
  const int64_t fr = f - stream * p.frames_per_stream;
  return p.x + stream * p.stream_stride + fr * p.hop;
}
// windowed pair (x[idx], x[idx+1]) * (w[idx], w[idx+1]); samples at or beyond the window are zero padding. kSkipPad: neither the
// sample nor the window entry of the padding is fetched (+7 % at the reference's default window, 8708 samples in 16384 points; the
// 32768-point kernel compiles 7 % slower with the branch and keeps the predicated sample load with an unconditional window load)
template <bool kVec2, bool kSkipPad = false>
__device__ __forceinline__ c2 load_windowed(const StftArgs& p, const float* __restrict__ xs, int idx) {
  if constexpr (kSkipPad) {
    if (idx >= p.window) return pk(0.f, 0.f);
    c2 s;
    if (kVec2) s = ldg_c2(xs + idx);
    else s = pk(__ldg(xs + idx), idx + 1 < p.window ? __ldg(xs + idx + 1) : 0.f);
    return mul2(s, ldg_c2(p.win + idx));
  } else {
    c2 s;
    if (kVec2) s = idx < p.window ? ldg_c2(xs + idx) : pk(0.f, 0.f);
    else s = pk(idx < p.window ? __ldg(xs + idx) : 0.f, idx + 1 < p.window ? __ldg(xs + idx + 1) : 0.f);
    return mul2(s, ldg_c2(p.win + idx));
  }
}
// (freq, dB) of the peak at bin k from its neighbours (SPEC S6), as in stft.cu
__device__ __forceinline__ void peak_interp_store(const StftArgs& p, int64_t out_base, int idx, int k, float sa, float sb, float sc,
                                                  bool store_bin) {
  const float fl = 1e-37f;
  sa = fmaxf(sa, fl);
  sb = fmaxf(sb, fl);
  sc = fmaxf(sc, fl);
  const float inv_b = __frcp_rn(sb);
  const float la = __log2f(sa * inv_b), lc = __log2f(sc * inv_b);
  const float num = la - lc, den = la + lc;
  const float pp = den != 0.f ? 0.5f * __fdividef(num, den) : 0.f;
  constexpr float kDbPerLog2 = 3.01029995663981195f;  // 10*log10(2)
  if (store_bin) p.peak_bin[out_base + idx] = k;
  p.peak_freq[out_base + idx] = p.freq_as_offset ? pp : (static_cast<float>(k) + pp) * p.bin_to_hz;
  p.peak_db[out_base + idx] = kDbPerLog2 * __log2f(sb) - 0.25f * (kDbPerLog2 * num) * pp;
}

// |X[k]|^2 and |X[M-k]|^2 from Z[k], Z[M-k] of a frame whose window carries the 1/2 of the untangling (stft_plan);
// wpb = (-i) * W_N^(k mod stride), imm = W_N^(stride * I)
template <typename Imm, bool kHasImm>
__device__ __forceinline__ void untangle_sq(c2 zk, c2 zm, const TwP& wpb, float& sk, float& sm) {
  const c2 E = add2(zk, cconj(zm));
  c2 D = csub(zk, cconj(zm));
  if constexpr (kHasImm) D = cmul_imm<Imm>(D);
  const c2 T = cmul_tw(D, wpb);
  // (P, Q) = E +- T formed component-wise so that |P|^2 and |Q|^2 come out of two packed instructions (as post_pair_p in stft.cu)
  const float2 e = upk(E), t = upk(T);
  const c2 pm = pk(1.f, -1.f);
  const c2 X = fma2(bc(t.x), pm, bc(e.x));
  const c2 Y = fma2(bc(t.y), pm, bc(e.y));
  const float2 sq2 = upk(fma2(X, X, mul2(Y, Y)));
  sk = sq2.x;
  sm = sq2.y;
}

// ---- one warp per frame: n_fft = 1024 (R1 = 16, two frames per warp), 2048 (R1 = 32) ----
template <int R1>
struct WarpStftCfg {
  using F = WarpFftCfg<R1>;
  static constexpr int SB = R1 + 4;             // floats per lane block of the |X|^2 buffer (R1 bins + pad: LDS.128, no conflicts)
  static constexpr int kSqFloats = 32 * SB + 4;  // bins 0..M (bin M at 32*SB)
  static constexpr int kList = 256;              // peak-bin list capacity in shared memory (longer lists: through the output)
  static constexpr int kWarpBytes = 8 * F::kBufC2 + 4 * (F::FPW * kSqFloats + kList);
};
template <int I, int R1>
__device__ __forceinline__ void warp_untangle_sq(const c2* __restrict__ Z, float* __restrict__ row, float* __restrict__ sq,
                                                 const TwP& wpb, int lane) {
  using C = WarpFftCfg<R1>;
  using S = WarpStftCfg<R1>;
  constexpr int M = C::M;
  const int k = lane + 32 * I, m = M - k;
  float sk, sm;
  untangle_sq<Tw<2 * M, 32 * I>, (I != 0)>(Z[k], Z[m & (M - 1)], wpb, sk, sm);
  if (row) {
    row[k] = sk;
    row[m] = sm;
  }
  sq[(k / R1) * S::SB + (k % R1)] = sk;
  sq[(m / R1) * S::SB + (m % R1)] = sm;
}
template <int R1, int... I>
__device__ __forceinline__ void warp_untangle_sq_all(const c2* Z, float* row, float* sq, const TwP& wpb, int lane,
                                                     std::integer_sequence<int, I...>) {
  (warp_untangle_sq<I, R1>(Z, row, sq, wpb, lane), ...);
}

template <int R1, int WARPS, bool kVec2>
__global__ void __launch_bounds__(WARPS * 32, 3) stft_warp_kernel(const __grid_constant__ StftArgs p) {
  using C = WarpFftCfg<R1>;
  using S = WarpStftCfg<R1>;
  constexpr int M = C::M, FPW = C::FPW, CPL = C::CPL;
  static_assert(CPL == 1, "the fused warp kernel covers R1 <= 32");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* tab4 = reinterpret_cast<float4*>(smem_raw);
  unsigned char* warp_base = reinterpret_cast<unsigned char*>(tab4 + C::kRows * 32);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < C::kRows * 32; i += WARPS * 32) {
    const int r = i >> 5, l = i & 31;
    float2 w;
    if (r < C::B - 1) w = cis_neg_w(2.f * static_cast<float>(l * (r + 1)) / M);
    else if (r < C::B - 1 + C::A - 1) w = cis_neg_w(2.f * static_cast<float>(l * C::B * (r - (C::B - 2))) / M);
    else {
      float s, c;
      sincospif(static_cast<float>(l) / M, &s, &c);
      w = make_float2(-s, -c);  // (-i) * W_N^l (the window carries the 1/2)
    }
    tab4[i] = make_float4(w.x, w.y, -w.y, w.y);
  }
  __syncthreads();
  c2* buf = reinterpret_cast<c2*>(warp_base + static_cast<size_t>(warp) * S::kWarpBytes);
  float* sq_all = reinterpret_cast<float*>(buf + C::kBufC2);
  int* klist = reinterpret_cast<int*>(sq_all + FPW * S::kSqFloats);
  const TwP wpb = make_twp(tab4[(C::kRows - 1) * 32 + lane]);

  // raw samples of the warp's next group are requested before the |X|^2 / peak phase of the current one (see rfft_warp_kernel)
  const int64_t groups = (p.total_frames + FPW - 1) / FPW;
  const int64_t gstep = static_cast<int64_t>(gridDim.x) * WARPS;
  c2 v[FPW][R1];
  auto load_group = [&](int64_t g) {
#pragma unroll
    for (int fw = 0; fw < FPW; ++fw) {
      const int64_t frame = g * FPW + fw;
      if (frame < p.total_frames) {
        const float* xs = frame_ptr(p, frame);
#pragma unroll
        for (int n1 = 0; n1 < R1; ++n1) {
          const int idx = 64 * n1 + 2 * lane;
          c2 s;
          if (kVec2) s = idx < p.window ? ldg_c2(xs + idx) : pk(0.f, 0.f);
          else s = pk(idx < p.window ? __ldg(xs + idx) : 0.f, idx + 1 < p.window ? __ldg(xs + idx + 1) : 0.f);
          v[fw][bitrev(n1, ilog2(R1))] = s;
        }
      } else {
#pragma unroll
        for (int n1 = 0; n1 < R1; ++n1) v[fw][n1] = pk(0.f, 0.f);
      }
    }
  };
  int64_t g = static_cast<int64_t>(blockIdx.x) * WARPS + warp;
  if (g < groups) load_group(g);
  for (; g < groups; g += gstep) {
#pragma unroll
    for (int fw = 0; fw < FPW; ++fw) {
#pragma unroll
      for (int n1 = 0; n1 < R1; ++n1) v[fw][bitrev(n1, ilog2(R1))] = mul2(v[fw][bitrev(n1, ilog2(R1))], ldg_c2(p.win + 64 * n1 + 2 * lane));
      DitP<R1, R1, 0>::run(v[fw]);
      tww_chain<R1, 0>(v[fw], tab4, lane, make_twp(tab4[lane]));
      c2* xb = buf + fw * (R1 / 2) * kXStride;
#pragma unroll
      for (int m = 0; m < R1 / 2; ++m)
        *reinterpret_cast<ulonglong2*>(xb + m * kXStride + 2 * lane) = make_ulonglong2(v[fw][2 * m], v[fw][2 * m + 1]);
    }
    __syncwarp();
    c2 A[32];
    {
      const int col = FPW > 1 ? (lane & (R1 - 1)) : lane;
      const c2* xb = buf + (FPW > 1 ? (lane / R1) * (R1 / 2) * kXStride : 0);
#pragma unroll
      for (int n2 = 0; n2 < 32; ++n2) A[bitrev(n2, 5)] = xb[(col >> 1) * kXStride + 2 * n2 + (col & 1)];
      __syncwarp();
      DitP<32, 32, 0>::run(A);
      c2* Z = buf + (FPW > 1 ? (lane / R1) * M : 0);
#pragma unroll
      for (int k2 = 0; k2 < 32; ++k2) Z[col + R1 * k2] = A[k2];
    }
    __syncwarp();
    if (g + gstep < groups) load_group(g + gstep);  // in flight during the |X|^2 / peak phase
#pragma unroll
    for (int fw = 0; fw < FPW; ++fw) {
      const int64_t frame = g * FPW + fw;
      if (frame >= p.total_frames) break;
      const c2* Z = buf + fw * M;
      float* sq = sq_all + fw * S::kSqFloats;
      float* row = p.sqmag ? p.sqmag + frame * p.sqmag_row_stride : nullptr;
      warp_untangle_sq_all<R1>(Z, row, sq, wpb, lane, std::make_integer_sequence<int, M / 64>{});
      if (lane == 0) {  // X[M/2] = conj(Z[M/2])
        const float2 z = upk(Z[M / 2]);
        const float s = 4.f * fmaf(z.x, z.x, z.y * z.y);  // X[M/2] = conj(Z[M/2]) of the unhalved frame (the window carries a 1/2)
        if (row) row[M / 2] = s;
        sq[((M / 2) / R1) * S::SB + ((M / 2) % R1)] = s;
      }
      if (!p.peak_count) continue;
      __syncwarp();
      // ---- peaks: lane owns bins [R1*lane, R1*lane + R1) ----
      const int base = S::SB * lane;
      float w[R1 + 2];
      w[0] = lane ? sq[base - 5] : __int_as_float(0x7f800000);  // bin 0 is never a peak
#pragma unroll
      for (int q = 0; q < R1 / 4; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(sq + base + 4 * q);
        w[1 + 4 * q] = t.x;
        w[2 + 4 * q] = t.y;
        w[3 + 4 * q] = t.z;
        w[4 + 4 * q] = t.w;
      }
      w[R1 + 1] = sq[base + S::SB];  // first bin of the next block (bin M for the last lane)
      uint32_t mask = 0;
#pragma unroll
      for (int j = R1 - 1; j >= 0; --j)  // sign bit of max(l, r, thr) - c, shifted in (see frame_peaks_p in stft.cu)
        mask = __funnelshift_l(__float_as_uint(fmaxf(fmaxf(w[j], w[j + 2]), p.thr_below) - w[j + 1]), mask, 1);
      const int cnt = __popc(mask);
      int incl = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      int slot = incl - cnt;
      if (lane == 0) p.peak_count[frame] = total;
      const int64_t out_base = frame * static_cast<int64_t>(p.max_peaks);
      const bool list_in_smem = p.max_peaks <= S::kList;
      const int n_store = min(total, p.max_peaks);
      while (mask && slot < n_store) {
        const int j = __ffs(static_cast<int>(mask)) - 1;
        mask &= mask - 1;
        if (list_in_smem) klist[slot] = R1 * lane + j;
        else p.peak_bin[out_base + slot] = R1 * lane + j;
        ++slot;
      }
      __syncwarp();
      for (int idx = lane; idx < n_store; idx += 32) {
        const int k = list_in_smem ? klist[idx] : *reinterpret_cast<volatile int*>(p.peak_bin + out_base + idx);
        auto at = [&](int b) { return sq[(b / R1) * S::SB + (b % R1)]; };
        peak_interp_store(p, out_base, idx, k, at(k - 1), at(k), at(k + 1), list_in_smem);
      }
      __syncwarp();
    }
    __syncwarp();
  }
}

// ---- one CTA per frame: n_fft = 8192, 16384, 32768 (M = M1 x 1024, one warp per row, 32 bins per thread) ----
template <int M1>
struct CtaStftCfg {
  using F = CtaFftCfg<M1>;
  static constexpr int SB = 36;                          // 32 bins per thread + pad
  static constexpr int kSqFloats = F::T * SB + 4;         // bins 0..M (bin M at T*SB)
  static constexpr int kList = 256;
  static constexpr size_t kSmem = sizeof(float4) * F::kRows * 32 + sizeof(c2) * F::kRowC2 * M1 + 4 * (kSqFloats + kList + 32);
};
template <int I, int M1>
__device__ __forceinline__ void cta_untangle_sq(const c2* __restrict__ Zs, float* __restrict__ row, float* __restrict__ sq,
                                                const TwP& wpb, int t) {
  using C = CtaFftCfg<M1>;
  constexpr int M = C::M, SB = CtaStftCfg<M1>::SB;
  const int k = t + C::T * I, m = M - k, mm = m & (M - 1);
  float sk, sm;
  untangle_sq<Tw<2 * M, C::T * I>, (I != 0)>(Zs[(k % M1) * C::kRowC2 + k / M1], Zs[(mm % M1) * C::kRowC2 + mm / M1], wpb, sk, sm);
  if (row) {
    row[k] = sk;
    row[m] = sm;
  }
  sq[(k >> 5) * SB + (k & 31)] = sk;
  sq[(m >> 5) * SB + (m & 31)] = sm;
}
template <int M1, int... I>
__device__ __forceinline__ void cta_untangle_sq_all(const c2* Zs, float* row, float* sq, const TwP& wpb, int t,
                                                    std::integer_sequence<int, I...>) {
  (cta_untangle_sq<I, M1>(Zs, row, sq, wpb, t), ...);
}

template <int M1, bool kVec2>
__global__ void __launch_bounds__(32 * M1, 16 / M1 > 4 ? 4 : 16 / M1) stft_cta_kernel(const __grid_constant__ StftArgs p) {
  using C = CtaFftCfg<M1>;
  using S = CtaStftCfg<M1>;
  constexpr int M = C::M, M2 = C::M2, T = C::T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* tab4 = reinterpret_cast<float4*>(smem_raw);
  c2* Ys = reinterpret_cast<c2*>(tab4 + C::kRows * 32);
  float* sq = reinterpret_cast<float*>(Ys + C::kRowC2 * M1);
  int* klist = reinterpret_cast<int*>(sq + S::kSqFloats);
  int* warp_tot = klist + S::kList;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int i = t; i < C::kRows * 32; i += T) {
    const int r = i >> 5, l = i & 31;
    const float2 w = r < 7 ? cis_neg_w(2.f * static_cast<float>(l * (r + 1)) / M2) : cis_neg_w(2.f * static_cast<float>(l * 8 * (r - 6)) / M2);
    tab4[i] = make_float4(w.x, w.y, -w.y, w.y);
  }
  TwP wpb;
  {
    float s, c;
    sincospif(static_cast<float>(t) / M, &s, &c);
    wpb = TwP{-s, pk(c, -c)};  // (-i) * W_N^t (the window carries the 1/2)
  }
  for (int64_t frame = blockIdx.x; frame < p.total_frames; frame += gridDim.x) {
    const float* xs = frame_ptr(p, frame);
    __syncthreads();
    // ---- step A: windowed columns ----
#pragma unroll
    for (int j = 0; j < M2 / T; ++j) {
      const int c = t + T * j;
      c2 v[M1];
#pragma unroll
      for (int n1 = 0; n1 < M1; ++n1) v[bitrev(n1, ilog2(M1))] = load_windowed<kVec2, (M1 <= 8)>(p, xs, 2 * (n1 * M2 + c));
      DitP<M1, M1, 0>::run(v);
      c2 w[M1];
      {
        float s, cs;
        sincospif(2.f * static_cast<float>(c) / M, &s, &cs);
        w[1] = pk(cs, -s);
      }
      cpow_all<M1>(w, std::make_integer_sequence<int, M1 - 2>{});
#pragma unroll
      for (int k1 = 1; k1 < M1; ++k1) {
        const float2 a = upk(w[k1]);
        v[k1] = cmul_tw(v[k1], TwP{a.x, pk(-a.y, a.y)});
      }
#pragma unroll
      for (int k1 = 0; k1 < M1; ++k1) Ys[k1 * C::kRowC2 + c] = v[k1];
    }
    __syncthreads();
    // ---- step B: warp k1 transforms row k1, in place (as rfft_cta_kernel) ----
    {
      c2* row = Ys + warp * C::kRowC2;
      c2 v[32];
#pragma unroll
      for (int n1 = 0; n1 < 32; ++n1) v[bitrev(n1, 5)] = row[32 * n1 + lane];
      __syncwarp();
      DitP<32, 32, 0>::run(v);
      tww_chain<32, 0>(v, tab4, lane, make_twp(tab4[lane]));
#pragma unroll
      for (int m = 0; m < 16; ++m) *reinterpret_cast<ulonglong2*>(row + m * kXStride + 2 * lane) = make_ulonglong2(v[2 * m], v[2 * m + 1]);
      __syncwarp();
      c2 A[32];
#pragma unroll
      for (int n2 = 0; n2 < 32; ++n2) A[bitrev(n2, 5)] = row[(lane >> 1) * kXStride + 2 * n2 + (lane & 1)];
      __syncwarp();
      DitP<32, 32, 0>::run(A);
#pragma unroll
      for (int k2 = 0; k2 < 32; ++k2) row[lane + 32 * k2] = A[k2];
    }
    __syncthreads();
    // ---- step C: |X|^2 ----
    float* rowo = p.sqmag ? p.sqmag + frame * p.sqmag_row_stride : nullptr;
    cta_untangle_sq_all<M1>(Ys, rowo, sq, wpb, t, std::make_integer_sequence<int, M / (2 * T)>{});
    if (t == 0) {
      const float2 z = upk(Ys[((M / 2) % M1) * C::kRowC2 + (M / 2) / M1]);
      const float s = 4.f * fmaf(z.x, z.x, z.y * z.y);  // X[M/2] = conj(Z[M/2]) of the unhalved frame (the window carries a 1/2)
      if (rowo) rowo[M / 2] = s;
      sq[((M / 2) >> 5) * S::SB + ((M / 2) & 31)] = s;
    }
    if (p.band_amps) {  // vocoder: amplitude of every band of this frame from the shared-memory copy of the spectrum
      __syncthreads();
      for (int b = warp; b < p.n_bands; b += M1) {
        const int lo = p.band_begin[b], hi = p.band_begin[b + 1];
        float acc = 0.f;
        for (int i = lo + lane; i < hi; i += 32) acc += sq[(i >> 5) * S::SB + (i & 31)];
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) p.band_amps[frame * p.n_bands + b] = sqrtf(acc);
      }
    }
    if (!p.peak_count) continue;
    __syncthreads();
    // ---- peaks: thread t owns bins [32t, 32t+32) ----
    const int base = S::SB * t;
    uint32_t mask = 0;
    {
      float w[34];
      w[0] = t ? sq[base - 5] : __int_as_float(0x7f800000);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 u = *reinterpret_cast<const float4*>(sq + base + 4 * q);
        w[1 + 4 * q] = u.x;
        w[2 + 4 * q] = u.y;
        w[3 + 4 * q] = u.z;
        w[4 + 4 * q] = u.w;
      }
      w[33] = sq[base + S::SB];
#pragma unroll
      for (int j = 31; j >= 0; --j)  // sign bit of max(l, r, thr) - c, shifted in (see frame_peaks_p in stft.cu)
        mask = __funnelshift_l(__float_as_uint(fmaxf(fmaxf(w[j], w[j + 2]), p.thr_below) - w[j + 1]), mask, 1);
    }
    const int cnt = __popc(mask);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int wv = 0; wv < M1; ++wv) {
      const int tw = warp_tot[wv];
      if (wv < warp) before += tw;
      total += tw;
    }
    int slot = before + incl - cnt;
    if (t == 0) p.peak_count[frame] = total;
    const int64_t out_base = frame * static_cast<int64_t>(p.max_peaks);
    const bool list_in_smem = p.max_peaks <= S::kList;
    const int n_store = min(total, p.max_peaks);
    while (mask && slot < n_store) {
      const int j = __ffs(static_cast<int>(mask)) - 1;
      mask &= mask - 1;
      if (list_in_smem) klist[slot] = 32 * t + j;
      else p.peak_bin[out_base + slot] = 32 * t + j;
      ++slot;
    }
    if (!list_in_smem) __threadfence_block();
    __syncthreads();
    for (int idx = t; idx < n_store; idx += T) {
      const int k = list_in_smem ? klist[idx] : *reinterpret_cast<volatile int*>(p.peak_bin + out_base + idx);
      auto at = [&](int b) { return sq[(b >> 5) * S::SB + (b & 31)]; };
      peak_interp_store(p, out_base, idx, k, at(k - 1), at(k), at(k + 1), list_in_smem);
    }
  }
}

template <int R1, int WARPS>
int launch_stft_warp(ca_ctx* ctx, const StftArgs& a, bool vec2, cudaStream_t stream) {
  using C = WarpFftCfg<R1>;
  using S = WarpStftCfg<R1>;
  void (*kern)(StftArgs) = vec2 ? stft_warp_kernel<R1, WARPS, true> : stft_warp_kernel<R1, WARPS, false>;
  const size_t smem = sizeof(float4) * C::kRows * 32 + static_cast<size_t>(S::kWarpBytes) * WARPS;
  // (two kernels share this function: set up per call)
  CA_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int per_sm = 0;
  CA_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
  if (per_sm < 1) per_sm = 1;
  const int64_t groups = (a.total_frames + C::FPW - 1) / C::FPW;
  const int64_t grid = std::min<int64_t>(static_cast<int64_t>(ctx->sm_count) * per_sm, (groups + WARPS - 1) / WARPS);
  kern<<<static_cast<unsigned>(grid), WARPS * 32, smem, stream>>>(a);
  CA_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches += 1;
  return CA_OK;
}
template <int M1>
int launch_stft_cta(ca_ctx* ctx, const StftArgs& a, bool vec2, cudaStream_t stream) {
  using C = CtaFftCfg<M1>;
  using S = CtaStftCfg<M1>;
  void (*kern)(StftArgs) = vec2 ? stft_cta_kernel<M1, true> : stft_cta_kernel<M1, false>;
  CA_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(S::kSmem)));
  int per_sm = 0;
  CA_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C::T, S::kSmem));
  if (per_sm < 1) per_sm = 1;
  const int64_t grid = std::min<int64_t>(static_cast<int64_t>(ctx->sm_count) * per_sm, a.total_frames);
  kern<<<static_cast<unsigned>(grid), C::T, S::kSmem, stream>>>(a);
  CA_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches += 1;
  return CA_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// Inverse real FFT (N <= 32768), unscaled like the forward one: x[n] = scale * sum_k X[k] exp(+2 pi i k n / N) over the
// Hermitian extension of the N/2+1 input bins (irfft(rfft(x)) = N*x at scale 1). Used by the vocoder's carrier path
// (fft.inverse at source/rt.resynth.lib.vocoder.cpp:438-470). The same transforms as above run on
//     Zc[k] = conj( (X[k] + conj X[M-k]) + i W_N^-k (X[k] - conj X[M-k]) ),   k < M = N/2,
// formed while loading, and the result is conjugated while storing (IFFT = conj . FFT . conj); its real and imaginary
// parts are the even and odd output samples.
// ---------------------------------------------------------------------------------------------------------------
// wmb = conj-rotation table entry i * W_N^-(k mod stride) in the (wr, (-wi, wi)) form; Imm = W_N^-(stride * I)
template <typename Imm, bool kHasImm>
__device__ __forceinline__ c2 tangle_conj(c2 xk, c2 xm, const TwP& wmb) {
  const c2 E = add2(xk, cconj(xm));
  c2 D = csub(xk, cconj(xm));
  if constexpr (kHasImm) D = cmul_imm<Imm>(D);
  return cconj(add2(E, cmul_tw(D, wmb)));
}
template <int N, int I>
struct TwInv {  // W_N^-I = conj(W_N^I)
  static constexpr float re = Tw<N, I>::re;
  static constexpr float im = -Tw<N, I>::im;
};

template <int R1, int N1>
__device__ __forceinline__ void irfft_load(c2 (&v)[R1], const c2* __restrict__ X, const TwP& wmb, int lane) {
  using C = WarpFftCfg<R1>;
  constexpr int M = C::M;
  const int k = 32 * N1 + lane;
  v[bitrev(N1, ilog2(R1))] = tangle_conj<TwInv<2 * M, 32 * N1>, (N1 != 0)>(__ldg(X + k), __ldg(X + (M - k)), wmb);
}
template <int R1, int... N1>
__device__ __forceinline__ void irfft_load_all(c2 (&v)[R1], const c2* X, const TwP& wmb, int lane, std::integer_sequence<int, N1...>) {
  (irfft_load<R1, N1>(v, X, wmb, lane), ...);
}

template <int R1, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) irfft_warp_kernel(const float2* __restrict__ in, float* __restrict__ out,
                                                                   const int64_t batch, const float scale) {
  using C = WarpFftCfg<R1>;
  constexpr int M = C::M, FPW = C::FPW, CPL = C::CPL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* tab4 = reinterpret_cast<float4*>(smem_raw);
  c2* buf_all = reinterpret_cast<c2*>(tab4 + C::kRows * 32);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < C::kRows * 32; i += WARPS * 32) {
    const int r = i >> 5, l = i & 31;
    float2 w;
    if (r < C::B - 1) w = cis_neg_w(2.f * static_cast<float>(l * (r + 1)) / M);
    else if (r < C::B - 1 + C::A - 1) w = cis_neg_w(2.f * static_cast<float>(l * C::B * (r - (C::B - 2))) / M);
    else {
      float s, c;
      sincospif(static_cast<float>(l) / M, &s, &c);  // W_N^-l = c + i s;  i * W_N^-l = -s + i c
      w = make_float2(-s, c);
    }
    tab4[i] = make_float4(w.x, w.y, -w.y, w.y);
  }
  __syncthreads();
  c2* buf = buf_all + warp * C::kBufC2;
  const TwP wmb = make_twp(tab4[(C::kRows - 1) * 32 + lane]);
  const c2 sc = pk(scale, -scale);  // conj and scale at once

  const int64_t groups = (batch + FPW - 1) / FPW;
  for (int64_t g = static_cast<int64_t>(blockIdx.x) * WARPS + warp; g < groups; g += static_cast<int64_t>(gridDim.x) * WARPS) {
#pragma unroll
    for (int fw = 0; fw < FPW; ++fw) {
      const int64_t frame = g * FPW + fw;
      c2 v[R1];
      if (frame < batch) {
        irfft_load_all<R1>(v, reinterpret_cast<const c2*>(in) + frame * (M + 1), wmb, lane, std::make_integer_sequence<int, R1>{});
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
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int col = FPW > 1 ? (lane & (R1 - 1)) : lane + 32 * c;
      const c2* xb = buf + (FPW > 1 ? (lane / R1) * (R1 / 2) * kXStride : 0);
      c2 A[32];
#pragma unroll
      for (int n2 = 0; n2 < 32; ++n2) A[bitrev(n2, 5)] = xb[(col >> 1) * kXStride + 2 * n2 + (col & 1)];
      DitP<32, 32, 0>::run(A);
      // y[col + R1*k2] = (x[2n], -x[2n+1]) at n = col + R1*k2: consecutive lanes, consecutive pairs
      const int64_t frame = g * FPW + (FPW > 1 ? lane / R1 : 0);
      if (frame < batch) {
        c2* o = reinterpret_cast<c2*>(out + frame * (2 * M));
#pragma unroll
        for (int k2 = 0; k2 < 32; ++k2) o[col + R1 * k2] = mul2(A[k2], sc);
      }
    }
    __syncwarp();
  }
}

template <int M1, int J, int N1>
__device__ __forceinline__ c2 irfft_cta_elem(const c2* __restrict__ X, const TwP& wmb, int t) {
  using C = CtaFftCfg<M1>;
  constexpr int M = C::M, OFF = N1 * C::M2 + C::T * J;
  const int k = OFF + t;
  return tangle_conj<TwInv<2 * M, OFF>, (OFF != 0)>(__ldg(X + k), __ldg(X + (M - k)), wmb);
}
template <int M1, int J, int... N1>
__device__ __forceinline__ void irfft_cta_col(c2 (&v)[M1], const c2* X, const TwP& wmb, int t, std::integer_sequence<int, N1...>) {
  ((v[bitrev(N1, ilog2(M1))] = irfft_cta_elem<M1, J, N1>(X, wmb, t)), ...);
}

template <int M1, int J>
__device__ __forceinline__ void irfft_cta_step_a(const c2* X, c2* Ys, const TwP& wmb, int t) {
  using C = CtaFftCfg<M1>;
  constexpr int M = C::M;
  const int c = t + C::T * J;
  c2 v[M1];
  irfft_cta_col<M1, J>(v, X, wmb, t, std::make_integer_sequence<int, M1>{});
  DitP<M1, M1, 0>::run(v);
  c2 w[M1];
  {
    float s, cs;
    sincospif(2.f * static_cast<float>(c) / M, &s, &cs);
    w[1] = pk(cs, -s);
  }
  cpow_all<M1>(w, std::make_integer_sequence<int, M1 - 2>{});
#pragma unroll
  for (int k1 = 1; k1 < M1; ++k1) {
    const float2 a = upk(w[k1]);
    v[k1] = cmul_tw(v[k1], TwP{a.x, pk(-a.y, a.y)});
  }
#pragma unroll
  for (int k1 = 0; k1 < M1; ++k1) Ys[k1 * C::kRowC2 + c] = v[k1];
}
template <int M1, int... J>
__device__ __forceinline__ void irfft_cta_step_a_all(const c2* X, c2* Ys, const TwP& wmb, int t, std::integer_sequence<int, J...>) {
  (irfft_cta_step_a<M1, J>(X, Ys, wmb, t), ...);
}

template <int M1>
__global__ void __launch_bounds__(32 * M1) irfft_cta_kernel(const float2* __restrict__ in, float* __restrict__ out, const int64_t batch,
                                                            const float scale) {
  using C = CtaFftCfg<M1>;
  constexpr int M = C::M, M2 = C::M2, T = C::T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* tab4 = reinterpret_cast<float4*>(smem_raw);
  c2* Ys = reinterpret_cast<c2*>(tab4 + C::kRows * 32);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int i = t; i < C::kRows * 32; i += T) {
    const int r = i >> 5, l = i & 31;
    const float2 w = r < 7 ? cis_neg_w(2.f * static_cast<float>(l * (r + 1)) / M2) : cis_neg_w(2.f * static_cast<float>(l * 8 * (r - 6)) / M2);
    tab4[i] = make_float4(w.x, w.y, -w.y, w.y);
  }
  TwP wmb;
  {
    float s, c;
    sincospif(static_cast<float>(t) / M, &s, &c);  // i * W_N^-t = -s + i c
    wmb = TwP{-s, pk(-c, c)};
  }
  const c2 sc = pk(scale, -scale);
  for (int64_t frame = blockIdx.x; frame < batch; frame += gridDim.x) {
    const c2* X = reinterpret_cast<const c2*>(in) + frame * (M + 1);
    __syncthreads();
    irfft_cta_step_a_all<M1>(X, Ys, wmb, t, std::make_integer_sequence<int, M2 / T>{});
    __syncthreads();
    {
      c2* row = Ys + warp * C::kRowC2;
      c2 v[32];
#pragma unroll
      for (int n1 = 0; n1 < 32; ++n1) v[bitrev(n1, 5)] = row[32 * n1 + lane];
      __syncwarp();
      DitP<32, 32, 0>::run(v);
      tww_chain<32, 0>(v, tab4, lane, make_twp(tab4[lane]));
#pragma unroll
      for (int m = 0; m < 16; ++m) *reinterpret_cast<ulonglong2*>(row + m * kXStride + 2 * lane) = make_ulonglong2(v[2 * m], v[2 * m + 1]);
      __syncwarp();
      c2 A[32];
#pragma unroll
      for (int n2 = 0; n2 < 32; ++n2) A[bitrev(n2, 5)] = row[(lane >> 1) * kXStride + 2 * n2 + (lane & 1)];
      __syncwarp();
      DitP<32, 32, 0>::run(A);
#pragma unroll
      for (int k2 = 0; k2 < 32; ++k2) row[lane + 32 * k2] = A[k2];
    }
    __syncthreads();
    c2* o = reinterpret_cast<c2*>(out + frame * (2 * M));
#pragma unroll 8
    for (int n = t; n < M; n += T) o[n] = mul2(Ys[(n % M1) * C::kRowC2 + n / M1], sc);
  }
}

template <int R1, int WARPS>
int launch_iwarp(ca_ctx* ctx, const float* d_in, int64_t batch, float* d_out, float scale, cudaStream_t stream) {
  using C = WarpFftCfg<R1>;
  auto kern = irfft_warp_kernel<R1, WARPS>;
  const size_t smem = sizeof(float4) * C::kRows * 32 + sizeof(c2) * C::kBufC2 * WARPS;
  int per_sm = 0;
  CA_KERNEL_SETUP(ctx, kern, WARPS * 32, smem, per_sm);
  if (per_sm < 1) per_sm = 1;
  const int64_t groups = (batch + C::FPW - 1) / C::FPW;
  const int64_t grid = std::min<int64_t>(static_cast<int64_t>(ctx->sm_count) * per_sm, (groups + WARPS - 1) / WARPS);
  kern<<<static_cast<unsigned>(grid), WARPS * 32, smem, stream>>>(reinterpret_cast<const float2*>(d_in), d_out, batch, scale);
  CA_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches += 1;
  return CA_OK;
}
template <int M1>
int launch_icta(ca_ctx* ctx, const float* d_in, int64_t batch, float* d_out, float scale, cudaStream_t stream) {
  using C = CtaFftCfg<M1>;
  auto kern = irfft_cta_kernel<M1>;
  const size_t smem = sizeof(float4) * C::kRows * 32 + sizeof(c2) * C::kRowC2 * M1;
  int per_sm = 0;
  CA_KERNEL_SETUP(ctx, kern, C::T, smem, per_sm);
  if (per_sm < 1) per_sm = 1;
  const int64_t grid = std::min<int64_t>(static_cast<int64_t>(ctx->sm_count) * per_sm, batch);
  kern<<<static_cast<unsigned>(grid), C::T, smem, stream>>>(reinterpret_cast<const float2*>(d_in), d_out, batch, scale);
  CA_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches += 1;
  return CA_OK;
}

// tile shapes of the two-pass path (CA_FFT_CW = 16|32 columns per CTA in pass 1, CA_FFT_G = 4|8 rows per CTA in pass 2;
// defaults from profiles/fft_tiles_sweep_r01.txt: the shapes differ by a few percent only, every one moves ~5.3 TB/s)
template <int A, int B>
int big2_variant(ca_ctx* ctx, const float* d_in, int64_t batch, float* d_out, cudaStream_t stream) {
  static const int cw = [] { const char* e = getenv("CA_FFT_CW"); return e ? atoi(e) : (A * B <= 64 ? 32 : 16); }();
  static const int g = [] { const char* e = getenv("CA_FFT_G"); return e ? atoi(e) : (A * B <= 64 ? 4 : 8); }();
  if (cw == 32 && A * B < 512) return g == 4 ? launch_big2<A, B, 32, 4>(ctx, d_in, batch, d_out, stream) : launch_big2<A, B, 32, 8>(ctx, d_in, batch, d_out, stream);
  return g == 4 ? launch_big2<A, B, 16, 4>(ctx, d_in, batch, d_out, stream) : launch_big2<A, B, 16, 8>(ctx, d_in, batch, d_out, stream);
}

}  // namespace

int rfft_warp(ca_ctx* ctx, const float* d_in, int n_fft, int64_t batch, float* d_out, cudaStream_t stream) {
  // A/B switch: CA_FFT_CLUSTER = 0 (CTA-per-frame / two-pass kernels) / 1 (cluster kernels, the measured default) / 2 (cluster
  // kernels with the other delivery: st.async + mbarrier <-> plain stores + cluster barrier)
  static const int cl = [] { const char* e = getenv("CA_FFT_CLUSTER"); return e ? atoi(e) : 1; }();
  switch (n_fft) {
    case 512: return launch_warp<8, 8>(ctx, d_in, batch, d_out, stream);
    case 1024: return launch_warp<16, 8>(ctx, d_in, batch, d_out, stream);
    case 2048: return launch_warp<32, 8>(ctx, d_in, batch, d_out, stream);
    case 4096: return launch_warp<64, 12>(ctx, d_in, batch, d_out, stream);
    case 8192: return launch_cta<4>(ctx, d_in, batch, d_out, stream);
    case 16384: return launch_cta<8>(ctx, d_in, batch, d_out, stream);
    case 32768:
      if (cl == 1 && cluster2_available<Cluster2Cfg<16, 8, 4, 4, false>>(ctx)) return launch_cluster2<Cluster2Cfg<16, 8, 4, 4, false>>(ctx, d_in, batch, d_out, stream);
      if (cl == 2 && cluster2_available<Cluster2Cfg<16, 8, 4, 4, true>>(ctx)) return launch_cluster2<Cluster2Cfg<16, 8, 4, 4, true>>(ctx, d_in, batch, d_out, stream);
      return launch_cta<16>(ctx, d_in, batch, d_out, stream);
    case 1 << 16:
      if (cl == 1 && cluster2_available<Cluster2Cfg<16, 8, 8, 8, false>>(ctx)) return launch_cluster2<Cluster2Cfg<16, 8, 8, 8, false>>(ctx, d_in, batch, d_out, stream);
      if (cl == 2 && cluster2_available<Cluster2Cfg<16, 8, 8, 8, true>>(ctx)) return launch_cluster2<Cluster2Cfg<16, 8, 8, 8, true>>(ctx, d_in, batch, d_out, stream);
      return big2_variant<4, 8>(ctx, d_in, batch, d_out, stream);
    case 1 << 17:
      if (cl == 1 && cluster2_available<Cluster2Cfg<16, 16, 8, 16, true>>(ctx)) return launch_cluster2<Cluster2Cfg<16, 16, 8, 16, true>>(ctx, d_in, batch, d_out, stream);
      if (cl == 2 && cluster2_available<Cluster2Cfg<16, 16, 8, 16, false>>(ctx)) return launch_cluster2<Cluster2Cfg<16, 16, 8, 16, false>>(ctx, d_in, batch, d_out, stream);
      return big2_variant<8, 8>(ctx, d_in, batch, d_out, stream);
    case 1 << 18: return big2_variant<8, 16>(ctx, d_in, batch, d_out, stream);
    case 1 << 19: return big2_variant<16, 16>(ctx, d_in, batch, d_out, stream);
    case 1 << 20: return big2_variant<32, 16>(ctx, d_in, batch, d_out, stream);
    default: break;
  }
  return fail(ctx, CA_ERR_UNSUPPORTED, "rfft_warp: unsupported n_fft");
}

}  // namespace ca

namespace ca {

int irfft_batched(ca_ctx* ctx, const float* d_in, int n_fft, int64_t batch, float* d_out, float scale, cudaStream_t stream) {
  if (!d_in || !d_out || batch < 0) return fail(ctx, CA_ERR_INVALID, "ca_irfft_batched: null pointer or negative batch");
  if (reinterpret_cast<uintptr_t>(d_in) % 8 || reinterpret_cast<uintptr_t>(d_out) % 8)
    return fail(ctx, CA_ERR_INVALID, "ca_irfft_batched: buffers must be 8-byte aligned");
  if (batch == 0) return CA_OK;
  switch (n_fft) {
    case 512: return launch_iwarp<8, 8>(ctx, d_in, batch, d_out, scale, stream);
    case 1024: return launch_iwarp<16, 8>(ctx, d_in, batch, d_out, scale, stream);
    case 2048: return launch_iwarp<32, 8>(ctx, d_in, batch, d_out, scale, stream);
    case 4096: return launch_iwarp<64, 12>(ctx, d_in, batch, d_out, scale, stream);
    case 8192: return launch_icta<4>(ctx, d_in, batch, d_out, scale, stream);
    case 16384: return launch_icta<8>(ctx, d_in, batch, d_out, scale, stream);
    case 32768: return launch_icta<16>(ctx, d_in, batch, d_out, scale, stream);
    default: break;
  }
  return fail(ctx, CA_ERR_UNSUPPORTED, "ca_irfft_batched: n_fft must be a power of two in [2^10, 2^15]");
}

int stft_fused_generic(ca_ctx* ctx, const StftArgs& a, int n_fft, bool vec2, cudaStream_t stream) {
  switch (n_fft) {
    case 1024: return launch_stft_warp<16, 4>(ctx, a, vec2, stream);  // 3 CTAs x 4 warps per SM (shared memory bound)
    case 2048: return launch_stft_warp<32, 4>(ctx, a, vec2, stream);
    case 8192: return launch_stft_cta<4>(ctx, a, vec2, stream);
    case 16384: return launch_stft_cta<8>(ctx, a, vec2, stream);
    case 32768: return launch_stft_cta<16>(ctx, a, vec2, stream);
    default: break;
  }
  return CA_ERR_UNSUPPORTED;
}

}  // namespace ca
