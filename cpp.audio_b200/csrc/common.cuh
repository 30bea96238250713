// common.cuh — shared device helpers: compile-time twiddles, in-register radix-2 FFTs, small PTX wrappers.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <utility>

namespace ca {

// ---------------------------------------------------------------------------------------------------------------
// compile-time exp(-2 pi i k / n) (double precision Taylor series after reduction to |angle| <= pi/4)
// ---------------------------------------------------------------------------------------------------------------
struct cpxd {
  double re, im;
};

constexpr double kPiD = 3.141592653589793238462643383279502884;

constexpr double taylor_sin(double x) {
  double x2 = x * x, term = x, sum = x;
  for (int i = 1; i <= 12; ++i) {
    term *= -x2 / ((2 * i) * (2 * i + 1));
    sum += term;
  }
  return sum;
}
constexpr double taylor_cos(double x) {
  double x2 = x * x, term = 1, sum = 1;
  for (int i = 1; i <= 12; ++i) {
    term *= -x2 / ((2 * i - 1) * (2 * i));
    sum += term;
  }
  return sum;
}
// cos/sin of 2*pi*k/n for integers, exact symmetries
constexpr cpxd cis2pi(long long k, long long n) {
  k %= n;
  if (k < 0) k += n;
  // nearest multiple of a quarter turn: m = round(4k/n)
  long long m = (8 * k + n) / (2 * n);
  // delta = k/n - m/4 turns, |delta| <= 1/8
  double delta = static_cast<double>(4 * k - m * n) / static_cast<double>(4 * n);
  double a = 2 * kPiD * delta;
  double c = taylor_cos(a), s = taylor_sin(a);
  switch (m & 3) {
    case 0: return {c, s};
    case 1: return {-s, c};
    case 2: return {-c, -s};
    default: return {s, -c};
  }
}

// W_N^I = exp(-2 pi i I / N) as float constants
template <int N, int I>
struct Tw {
  static constexpr float re = static_cast<float>(cis2pi(I, N).re);
  static constexpr float im = static_cast<float>(-cis2pi(I, N).im);
};

constexpr int bitrev(int v, int bits) {
  int r = 0;
  for (int b = 0; b < bits; ++b)
    if (v & (1 << b)) r |= 1 << (bits - 1 - b);
  return r;
}
constexpr int ilog2(int n) {
  int b = 0;
  while ((1 << b) < n) ++b;
  return b;
}

__device__ __forceinline__ uint32_t smem_u32(const void* ptr) { return static_cast<uint32_t>(__cvta_generic_to_shared(ptr)); }
// 1-D bulk asynchronous copy global -> shared (TMA engine, SASS UBLKCP), completion counted in bytes on an mbarrier.
// Unlike cp.async (LDGSTS) it is not tracked by the warp's scoreboard, so the shared-memory reads of the peak scan that
// follow do not wait for it: the copy really overlaps them.
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(-a.y, b.y, a.x * b.x), fmaf(a.x, b.y, a.y * b.x));
}

// ---------------------------------------------------------------------------------------------------------------
// In-register radix-2 decimation-in-frequency FFT of N points held in v[OFF .. OFF+N) (natural order in,
// bit-reversed order out: on return v[OFF + j] holds the output of index bitrev(j)). Every index and twiddle is a
// compile-time constant so the array lives in registers and trivial twiddles cost nothing.
// ---------------------------------------------------------------------------------------------------------------
template <int SIZE, int N, int OFF>
struct Dif {
  template <int I>
  static __device__ __forceinline__ void butterfly(float2 (&v)[SIZE]) {
    const float2 a = v[OFF + I];
    const float2 b = v[OFF + I + N / 2];
    v[OFF + I] = make_float2(a.x + b.x, a.y + b.y);
    const float dx = a.x - b.x, dy = a.y - b.y;
    constexpr float kHalfSqrt2 = 0.70710678118654752440f;
    if constexpr (I == 0) {
      v[OFF + I + N / 2] = make_float2(dx, dy);
    } else if constexpr (4 * I == N) {  // * (-i)
      v[OFF + I + N / 2] = make_float2(dy, -dx);
    } else if constexpr (8 * I == N) {  // * (1-i)/sqrt2
      v[OFF + I + N / 2] = make_float2(kHalfSqrt2 * (dx + dy), kHalfSqrt2 * (dy - dx));
    } else if constexpr (8 * I == 3 * N) {  // * (-1-i)/sqrt2
      v[OFF + I + N / 2] = make_float2(kHalfSqrt2 * (dy - dx), -kHalfSqrt2 * (dx + dy));
    } else {
      constexpr float wr = Tw<N, I>::re, wi = Tw<N, I>::im;
      v[OFF + I + N / 2] = make_float2(fmaf(-dy, wi, dx * wr), fmaf(dx, wi, dy * wr));
    }
  }
  template <int... I>
  static __device__ __forceinline__ void stage(float2 (&v)[SIZE], std::integer_sequence<int, I...>) {
    (butterfly<I>(v), ...);
  }
  static __device__ __forceinline__ void run(float2 (&v)[SIZE]) {
    if constexpr (N > 1) {
      stage(v, std::make_integer_sequence<int, N / 2>{});
      Dif<SIZE, N / 2, OFF>::run(v);
      Dif<SIZE, N / 2, OFF + N / 2>::run(v);
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// In-register radix-2 decimation-in-time FFT: on entry v[OFF + i] holds the input of index bitrev(i), on return
// v[OFF + k] holds output k (natural order). A general butterfly is 6 FFMA whose twiddles are immediates
// (a + w*b in 4, a - w*b = 2a - (a + w*b) in 2): fewer instructions than the 8 of the DIF form and only two register
// sources per instruction, which matters on B200 where FFMA with three distinct registers issues at half rate
// (profiles/microbench/fp32_ports_b200.txt).
// ---------------------------------------------------------------------------------------------------------------
template <int SIZE, int N, int OFF>
struct Dit {
  template <int J>
  static __device__ __forceinline__ void butterfly(float2 (&v)[SIZE]) {
    const float2 a = v[OFF + J];
    const float2 b = v[OFF + J + N / 2];
    if constexpr (J == 0) {
      v[OFF + J] = make_float2(a.x + b.x, a.y + b.y);
      v[OFF + J + N / 2] = make_float2(a.x - b.x, a.y - b.y);
    } else if constexpr (4 * J == N) {  // w = -i: w*b = (b.y, -b.x)
      v[OFF + J] = make_float2(a.x + b.y, a.y - b.x);
      v[OFF + J + N / 2] = make_float2(a.x - b.y, a.y + b.x);
    } else {
      constexpr float wr = Tw<N, J>::re, wi = Tw<N, J>::im;
      const float ox = fmaf(wr, b.x, fmaf(-wi, b.y, a.x));
      const float oy = fmaf(wr, b.y, fmaf(wi, b.x, a.y));
      v[OFF + J] = make_float2(ox, oy);
      v[OFF + J + N / 2] = make_float2(fmaf(2.f, a.x, -ox), fmaf(2.f, a.y, -oy));
    }
  }
  template <int... J>
  static __device__ __forceinline__ void stage(float2 (&v)[SIZE], std::integer_sequence<int, J...>) {
    (butterfly<J>(v), ...);
  }
  static __device__ __forceinline__ void run(float2 (&v)[SIZE]) {
    if constexpr (N > 1) {
      Dit<SIZE, N / 2, OFF>::run(v);
      Dit<SIZE, N / 2, OFF + N / 2>::run(v);
      stage(v, std::make_integer_sequence<int, N / 2>{});
    }
  }
};

__device__ __forceinline__ float2 ldg2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }

}  // namespace ca
