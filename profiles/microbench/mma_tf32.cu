// Throughput of the legacy tensor path (mma.sync m16n8k8 tf32, SASS HMMA.1688.F32.TF32) and of shared-memory operand loads on
// B200: how fast can a 3xTF32 (hi/lo split) product run next to the FFMA2 product of oscbank.cu?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_tf32 mma_tf32.cu && ./mma_tf32
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int TILES, int KIND>
__global__ void mma16_kernel(float* out, int iters) {
  float d[TILES][4];
  uint32_t a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = 0x3c003c00u + threadIdx.x + i;
  for (int i = 0; i < 2; ++i) b[i] = 0x38003800u + threadIdx.x + i;
  for (int t = 0; t < TILES; ++t)
    for (int i = 0; i < 4; ++i) d[t][i] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < TILES; ++t) {
      if (KIND == 0) mma_f16(d[t], a, b);
      else mma_bf16(d[t], a, b);
    }
  }
  float s = 0;
  for (int t = 0; t < TILES; ++t)
    for (int i = 0; i < 4; ++i) s += d[t][i];
  if (s == 1234.5f) out[0] = s;
}

template <int TILES>
__global__ void mma_kernel(float* out, int iters) {
  float d[TILES][4];
  uint32_t a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i);
  for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.5f + threadIdx.x * 1e-3f + i);
  for (int t = 0; t < TILES; ++t)
    for (int i = 0; i < 4; ++i) d[t][i] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < TILES; ++t) mma_tf32(d[t], a, b);
  }
  float s = 0;
  for (int t = 0; t < TILES; ++t)
    for (int i = 0; i < 4; ++i) s += d[t][i];
  if (s == 1234.5f) out[0] = s;
}

// shared-memory operand delivery: every lane issues LDS.128 / LDS.64 / LDS.32; "same" = one address per 8 lanes
template <int BYTES, bool BROADCAST>
__global__ void lds_kernel(float* out, int iters) {
  __shared__ __align__(16) float s[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) s[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int idx = BROADCAST ? (lane >> 3) * (BYTES / 4) : lane * (BYTES / 4);
  float acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int off = ((it + u) & 7) * 256 + idx;
      if (BYTES == 16) { float4 v; asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((uint32_t)__cvta_generic_to_shared(s + off))); acc += (v.x + v.y) + (v.z + v.w); }
      else if (BYTES == 8) { float2 v; asm volatile("ld.volatile.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((uint32_t)__cvta_generic_to_shared(s + off))); acc += v.x + v.y; }
      else { float v; asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"((uint32_t)__cvta_generic_to_shared(s + off))); acc += v; }
    }
  }
  if (acc == 1234.5f) out[0] = acc;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const double ghz = 1.965;
  printf("%s, %d SMs (rates at %.3f GHz)\n", p.name, p.multiProcessorCount, ghz);
  float* d; cudaMalloc(&d, 4);
  const int sms = p.multiProcessorCount, iters = 4096;
  for (int warps : {4, 8, 16}) {
    auto report = [&](const char* name, int tiles, float ms) {
      const double n = double(tiles) * iters * warps * sms;  // warp-level mma instructions
      printf("%-34s %2d warps/SM  %.3f ms  %.3f mma/clk/SM  %.1f TFLOP/s dense tf32 (1x)\n", name, warps, ms, n / (ms * 1e-3) / (ghz * 1e9) / sms,
             n * 2048 / (ms * 1e-3) / 1e12);
    };
    report("mma.m16n8k8.tf32, 8 acc tiles", 8, time_ms([&] { mma_kernel<8><<<sms, warps * 32>>>(d, iters); }));
    report("mma.m16n8k8.tf32, 16 acc tiles", 16, time_ms([&] { mma_kernel<16><<<sms, warps * 32>>>(d, iters); }));
    report("mma.m16n8k16.f16 (x2 flop), 16 tiles", 16, time_ms([&] { mma16_kernel<16, 0><<<sms, warps * 32>>>(d, iters); }));
    report("mma.m16n8k16.bf16 (x2 flop), 16 tiles", 16, time_ms([&] { mma16_kernel<16, 1><<<sms, warps * 32>>>(d, iters); }));
  }
  for (int warps : {8, 16}) {
    auto report = [&](const char* name, float ms) {
      const double n = 16.0 * iters * warps * sms;
      printf("%-34s %2d warps/SM  %.3f ms  %.3f warp-loads/clk/SM\n", name, warps, ms, n / (ms * 1e-3) / (ghz * 1e9) / sms);
    };
    report("LDS.32  distinct", time_ms([&] { lds_kernel<4, false><<<sms, warps * 32>>>(d, iters); }));
    report("LDS.32  4 addresses per warp", time_ms([&] { lds_kernel<4, true><<<sms, warps * 32>>>(d, iters); }));
    report("LDS.64  distinct", time_ms([&] { lds_kernel<8, false><<<sms, warps * 32>>>(d, iters); }));
    report("LDS.64  4 addresses per warp", time_ms([&] { lds_kernel<8, true><<<sms, warps * 32>>>(d, iters); }));
    report("LDS.128 distinct", time_ms([&] { lds_kernel<16, false><<<sms, warps * 32>>>(d, iters); }));
    report("LDS.128 4 addresses per warp", time_ms([&] { lds_kernel<16, true><<<sms, warps * 32>>>(d, iters); }));
  }
  return 0;
}
