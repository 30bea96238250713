// packed.cuh — complex arithmetic on the packed FP32 instructions of sm_100 (fma/add/mul.rn.f32x2 -> SASS FFMA2 / FADD2 /
// FMUL2). A complex number is one 64-bit register pair (re = low word, im = high word), so a complex add is ONE
// instruction and a radix-2 butterfly with a constant twiddle is THREE:
//     o  = a + w*b = fma2(wi, i*b, fma2(wr, b, a))        o' = a - w*b = fma2(2, a, -o)
// ptxas folds the swap / negation of i*b, -o and conj() into the operand modifiers of the packed instruction
// (R.F32x2.LO_HI, .NP, leading '-') and takes wr, wi as 32-bit immediates, so nothing else is issued.
// Measured on B200 (profiles/microbench/f32x2.cu): packed instructions run at the same FLOP rate as the scalar forms
// (3.9 scalar-equivalent warp-instr/clk/SM with two register-pair operands) in HALF the issue slots and half the code
// bytes — the analysis kernel was bound by instruction fetch (70 KB body vs the 32 KB instruction cache) and issue.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <utility>

#include "common.cuh"

namespace ca {

typedef unsigned long long c2;  // packed complex: {re, im}

__device__ __forceinline__ c2 pk(float x, float y) {
  c2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ float2 upk(c2 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ c2 fma2(c2 a, c2 b, c2 c) {
  c2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ c2 add2(c2 a, c2 b) {
  c2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ c2 mul2(c2 a, c2 b) {
  c2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ c2 bc(float x) { return pk(x, x); }  // broadcast (R.F32 operand or an immediate)
__device__ __forceinline__ c2 cneg(c2 a) {
  const float2 v = upk(a);
  return pk(-v.x, -v.y);
}
__device__ __forceinline__ c2 cmuli(c2 a) {  // i*a
  const float2 v = upk(a);
  return pk(-v.y, v.x);
}
__device__ __forceinline__ c2 cmulni(c2 a) {  // -i*a
  const float2 v = upk(a);
  return pk(v.y, -v.x);
}
__device__ __forceinline__ c2 cconj(c2 a) {
  const float2 v = upk(a);
  return pk(v.x, -v.y);
}
__device__ __forceinline__ c2 cswap(c2 a) {
  const float2 v = upk(a);
  return pk(v.y, v.x);
}
__device__ __forceinline__ c2 csub(c2 a, c2 b) { return add2(a, cneg(b)); }

// w*b for a compile-time w (two instructions, immediates)
template <typename W>
__device__ __forceinline__ c2 cmul_imm(c2 b) {
  return fma2(bc(W::im), cmuli(b), mul2(bc(W::re), b));
}
// w*b for a run-time twiddle kept as wr (scalar register, broadcast operand) and P = (-wi, wi) (register pair):
// wr*(bx,by) + (-wi,wi)*(by,bx). Two instructions; the pair form avoids a modifier combination ptxas does not fold.
struct TwP {
  float wr;
  c2 P;
};
__device__ __forceinline__ c2 cmul_tw(c2 b, const TwP& w) { return fma2(w.P, cswap(b), mul2(bc(w.wr), b)); }
__device__ __forceinline__ TwP make_twp(const float4 t) { return TwP{t.x, pk(t.z, t.w)}; }  // table entry (wr, wi, -wi, wi)

// ---------------------------------------------------------------------------------------------------------------
// In-register radix-2 decimation-in-time FFT on packed complex values: on entry v[OFF + i] holds the input of index
// bitrev(i), on return v[OFF + k] holds output k. Every index and twiddle is a compile-time constant.
// ---------------------------------------------------------------------------------------------------------------
template <int SIZE, int N, int OFF>
struct DitP {
  template <int J>
  static __device__ __forceinline__ void butterfly(c2 (&v)[SIZE]) {
    const c2 a = v[OFF + J];
    const c2 b = v[OFF + J + N / 2];
    if constexpr (J == 0) {
      v[OFF + J] = add2(a, b);
      v[OFF + J + N / 2] = csub(a, b);
    } else if constexpr (4 * J == N) {  // w = -i
      v[OFF + J] = add2(a, cmulni(b));
      v[OFF + J + N / 2] = add2(a, cmuli(b));
    } else {
      constexpr float wr = Tw<N, J>::re, wi = Tw<N, J>::im;
      const c2 o = fma2(bc(wi), cmuli(b), fma2(bc(wr), b, a));
      v[OFF + J] = o;
      v[OFF + J + N / 2] = fma2(bc(2.f), a, cneg(o));
    }
  }
  template <int... J>
  static __device__ __forceinline__ void stage(c2 (&v)[SIZE], std::integer_sequence<int, J...>) {
    (butterfly<J>(v), ...);
  }
  static __device__ __forceinline__ void run(c2 (&v)[SIZE]) {
    if constexpr (N > 1) {
      DitP<SIZE, N / 2, OFF>::run(v);
      DitP<SIZE, N / 2, OFF + N / 2>::run(v);
      stage(v, std::make_integer_sequence<int, N / 2>{});
    }
  }
};

__device__ __forceinline__ c2 ldg_c2(const float* p) { return __ldg(reinterpret_cast<const c2*>(p)); }

}  // namespace ca
