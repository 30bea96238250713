// icache.cu — does straight-line code that is larger than the instruction caches cap the issue rate on B200?
// Same FFMA-imm work (16 independent chains per thread), loop body unrolled to different sizes.
#include <cstdio>
#include <cuda_runtime.h>

template <int UNROLL>
__global__ void __launch_bounds__(128) k(float* out, int iters) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], 0.999f + 1e-6f * (u % 7), 1e-6f * (i + 1));
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456f) out[0] = s;
}

template <int UNROLL>
void run(float* d_out, int sms, float ghz, int warps_per_sm) {
  const int threads = 128, blocks = sms * (warps_per_sm / 4);
  const int iters = (1 << 18) / UNROLL;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<UNROLL><<<blocks, threads>>>(d_out, 4);
  cudaEventRecord(e0);
  k<UNROLL><<<blocks, threads>>>(d_out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double winst = double(blocks) * 4 * double(iters) * UNROLL * 16;
  printf("body %6d instr (%5.0f KB)  %2d warps/SM  %8.3f ms  %5.2f warp-instr/clk/SM\n", UNROLL * 16, UNROLL * 16 * 16 / 1024.0,
         warps_per_sm, ms, winst / (ms * 1e-3) / (ghz * 1e9) / sms);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  float* d;
  cudaMalloc(&d, 4);
  const float ghz = p.clockRate * 1e-6f;
  for (int w : {4, 12, 16}) {
    run<4>(d, p.multiProcessorCount, ghz, w);
    run<32>(d, p.multiProcessorCount, ghz, w);
    run<64>(d, p.multiProcessorCount, ghz, w);
    run<128>(d, p.multiProcessorCount, ghz, w);
    run<256>(d, p.multiProcessorCount, ghz, w);
    run<512>(d, p.multiProcessorCount, ghz, w);
  }
  return 0;
}
