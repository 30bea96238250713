// fp32_ports.cu — microbenchmark: issue rate of register-operand FP32 instructions on B200 (warp-instr / clk / SM).
// Question: are FADD/FMUL/FFMA with 2 or 3 DISTINCT register sources limited below 4 warp-instr/clk/SM (register-file
// read ports), as profiles of the STFT / oscillator kernels suggest (IPC ~2.3 with ~75 % FP32 instructions)?
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NACC = 16;
constexpr int ITERS = 2048;

template <int MODE>
__global__ void __launch_bounds__(256) k(const float* __restrict__ in, float* out, int iters) {
  float a[NACC], b[NACC], c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    a[i] = in[threadIdx.x + 32 * i];
    b[i] = in[threadIdx.x + 32 * i + 1024];
    c[i] = in[threadIdx.x + 32 * i + 2048];
  }
  const float m = in[4000], cc = in[4001];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (MODE == 0) a[i] = a[i] + b[i];                 // FADD, 2 distinct regs
      if (MODE == 1) a[i] = fmaf(a[i], b[i], c[i]);      // FFMA, 3 distinct regs, none shared between instructions
      if (MODE == 2) a[i] = fmaf(a[i], m, cc);           // FFMA, 2 operands shared by every instruction (reuse cache)
      if (MODE == 3) a[i] = a[i] * b[i];                 // FMUL, 2 distinct regs
      if (MODE == 4) a[i] = fmaf(a[i], 0.999f, 1e-6f);   // FFMA with immediates
      if (MODE == 5) a[i] = fmaf(a[i], b[i], a[i]);      // FFMA, 2 distinct regs
      if (MODE == 6) { a[i] = fmaf(-b[i], a[i], c[i]); c[i] = c[i] + a[i]; }  // the oscillator pair: d -= k*y ; y += d
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += a[i] + c[i];
  if (s == 123.456f) out[0] = s;
}

template <int MODE>
void run(const char* name, const float* d_in, float* d_out, int sms, float clk_ghz, int per_iter, int warps_per_sm = 64) {
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
  const double winst = double(blocks) * (threads / 32) * ITERS * NACC * per_iter;
  printf("%-52s %2d warps/SM %8.3f ms  %6.3f warp-instr/clk/SM (at %.3f GHz)\n", name, warps_per_sm, ms, winst / (ms * 1e-3) / (clk_ghz * 1e9) / sms, clk_ghz);
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
  cudaMemcpy(d_in, h, sizeof(h), cudaMemcpyHostToDevice);
  const float ghz = p.clockRate * 1e-6f;
  printf("%s, %d SMs, clockRate %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
  run<0>("FADD  a=a+b           (2 regs)", d_in, d_out, p.multiProcessorCount, ghz, 1);
  run<3>("FMUL  a=a*b           (2 regs)", d_in, d_out, p.multiProcessorCount, ghz, 1);
  run<1>("FFMA  a=fma(a,b,c)    (3 distinct regs)", d_in, d_out, p.multiProcessorCount, ghz, 1);
  run<5>("FFMA  a=fma(a,b,a)    (2 distinct regs)", d_in, d_out, p.multiProcessorCount, ghz, 1);
  run<2>("FFMA  a=fma(a,m,c)    (m,c shared: reuse cache)", d_in, d_out, p.multiProcessorCount, ghz, 1);
  run<4>("FFMA  a=fma(a,imm,imm)", d_in, d_out, p.multiProcessorCount, ghz, 1);
  run<6>("pair  d=fma(-k,y,d); y=y+d (oscillator step)", d_in, d_out, p.multiProcessorCount, ghz, 2);
  // occupancy sensitivity: the production kernels run 8-16 warps/SM with long independent chains
  for (int w : {4, 8, 12, 16, 32}) {
    run<0>("FADD  a=a+b  (16 independent chains per thread)", d_in, d_out, p.multiProcessorCount, ghz, 1, w);
    run<4>("FFMA  imm    (16 independent chains per thread)", d_in, d_out, p.multiProcessorCount, ghz, 1, w);
    run<6>("pair  oscillator step (16 chains)", d_in, d_out, p.multiProcessorCount, ghz, 2, w);
  }
  return 0;
}
