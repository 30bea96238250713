// f32x2.cu — microbenchmark: issue rate of the packed FP32 instructions of sm_100 (fma/add/mul.rn.f32x2 -> SASS
// FFMA2 / FADD2 / FMUL2) against the scalar forms, in flop-equivalent warp-instructions / clk / SM.
// Question: a complex butterfly is two-wide by nature; does the packed form lift the register-port limit measured in
// fp32_ports.cu (FFMA with three distinct registers: 2.2 of 4 warp-instr/clk/SM)?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int NACC = 16;
constexpr int ITERS = 2048;

__device__ __forceinline__ uint64_t pk(float x, float y) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void upk(uint64_t v, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(const float* __restrict__ in, float* out, int iters) {
  uint64_t a[NACC], b[NACC], c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    a[i] = pk(in[threadIdx.x + 32 * i], in[threadIdx.x + 32 * i + 512]);
    b[i] = pk(in[threadIdx.x + 32 * i + 1024], in[threadIdx.x + 32 * i + 1536]);
    c[i] = pk(in[threadIdx.x + 32 * i + 2048], in[threadIdx.x + 32 * i + 2560]);
  }
  const uint64_t m = pk(in[4000], in[4002]), cc = pk(in[4001], in[4003]);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (MODE == 0) a[i] = add2(a[i], b[i]);             // FADD2, 2 distinct register pairs
      if (MODE == 1) a[i] = fma2(a[i], b[i], c[i]);       // FFMA2, 3 distinct register pairs
      if (MODE == 2) a[i] = fma2(a[i], m, cc);            // FFMA2, 2 operands shared (reuse cache)
      if (MODE == 3) a[i] = mul2(a[i], b[i]);             // FMUL2
      if (MODE == 5) a[i] = fma2(a[i], b[i], a[i]);       // FFMA2, 2 distinct register pairs
      if (MODE == 6) {                                    // mix: one packed + one scalar-side ALU op (issue slots)
        a[i] = fma2(a[i], b[i], c[i]);
        c[i] = c[i] ^ 0x0000000100000000ull;
      }
      if (MODE == 7) {  // complex DIT butterfly, packed: t = wr*b + wi*swap(b)*(-1,1); o = a + t; o2 = 2a - o
        float bx, by;
        upk(b[i], bx, by);
        const uint64_t bs = pk(-by, bx);
        uint64_t t = fma2(m, b[i], a[i]);
        t = fma2(cc, bs, t);
        b[i] = fma2(pk(2.f, 2.f), a[i], t ^ 0x8000000080000000ull);
        a[i] = t;
      }
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s ^= a[i] ^ c[i] ^ b[i];
  if (s == 0x123456789ull) out[0] = 1.f;
}

template <int MODE>
void run(const char* name, const float* d_in, float* d_out, int sms, float clk_ghz, double flop_instr_per_iter, int warps_per_sm = 64) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int threads = warps_per_sm >= 8 ? 256 : 32 * warps_per_sm;
  const int blocks = sms * (warps_per_sm * 32 / threads);
  k<MODE><<<blocks, threads>>>(d_in, d_out, 16);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(d_in, d_out, ITERS);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double winst = double(blocks) * (threads / 32) * ITERS * NACC * flop_instr_per_iter;
  printf("%-58s %2d warps/SM %8.3f ms  %6.3f scalar-equivalent warp-instr/clk/SM\n", name, warps_per_sm, ms,
         winst / (ms * 1e-3) / (clk_ghz * 1e9) / sms);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  float* d_in;
  float* d_out;
  cudaMalloc(&d_in, 8192 * 4);
  cudaMalloc(&d_out, 4);
  float h[8192];
  for (int i = 0; i < 8192; ++i) h[i] = 1.0f + 1e-6f * (i % 97);
  h[4000] = 0.9999f;
  h[4001] = 1e-5f;
  h[4002] = 0.9998f;
  h[4003] = 2e-5f;
  cudaMemcpy(d_in, h, sizeof(h), cudaMemcpyHostToDevice);
  const float ghz = p.clockRate * 1e-6f;
  printf("%s, %d SMs, clockRate %.3f GHz (a packed instruction counts as 2 scalar-equivalents)\n", p.name, p.multiProcessorCount, ghz);
  for (int w : {8, 16, 64}) {
    run<0>("FADD2 a=a+b           (2 pairs)", d_in, d_out, p.multiProcessorCount, ghz, 2, w);
    run<3>("FMUL2 a=a*b           (2 pairs)", d_in, d_out, p.multiProcessorCount, ghz, 2, w);
    run<1>("FFMA2 a=fma(a,b,c)    (3 distinct pairs)", d_in, d_out, p.multiProcessorCount, ghz, 2, w);
    run<5>("FFMA2 a=fma(a,b,a)    (2 distinct pairs)", d_in, d_out, p.multiProcessorCount, ghz, 2, w);
    run<2>("FFMA2 a=fma(a,m,c)    (m,c shared)", d_in, d_out, p.multiProcessorCount, ghz, 2, w);
    run<6>("FFMA2 3 pairs + one LOP3 (counted: 2)", d_in, d_out, p.multiProcessorCount, ghz, 2, w);
    run<7>("packed DIT butterfly (3 FFMA2 + swap; counted: 6)", d_in, d_out, p.multiProcessorCount, ghz, 6, w);
  }
  return 0;
}
