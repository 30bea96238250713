// rfft.cu — K1 entry point: batched forward real FFT, N = 2^10 .. 2^20, float32, unscaled.
//
// Stands in for fft::Algo_<Tag,T>::forward as FftOperation<T> uses it (source/rt.resynth.lib.periodicfft.cpp:210-235) and
// forward_fft (source/main.test_fft.cpp:78). The kernels live in rfftw.cu:
//   N <= 4096           one warp per frame (two frames per warp at N = 1024), frame in registers + one shared-memory exchange
//   N = 8192 .. 32768   one CTA per frame, frame resident in shared memory between its single read and single write of HBM
//   N >= 65536          two passes through HBM: column FFTs (big2_cols_kernel), then 1024-point row FFTs with the untangling
//                       (big2_rows_kernel)
#include "common.cuh"
#include "ctx.hpp"

namespace ca {

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
  if (n_fft < 512 || n_fft > (1 << 20) || (n_fft & (n_fft - 1)))
    return fail(ctx, CA_ERR_UNSUPPORTED, "ca_rfft_batched: n_fft must be a power of two in [2^9, 2^20]");
  if (batch == 0) return CA_OK;
  if (reinterpret_cast<uintptr_t>(d_in) % 8 || reinterpret_cast<uintptr_t>(d_out) % 8)
    return fail(ctx, CA_ERR_INVALID, "ca_rfft_batched: buffers must be 8-byte aligned");
  return rfft_warp(ctx, d_in, n_fft, batch, d_out, stream);
}

}  // namespace ca
