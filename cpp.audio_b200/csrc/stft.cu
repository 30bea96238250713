// stft.cu — K2: fused periodic-FFT analysis for 4096-point frames.
//
// One warp owns one frame from load to peak list; the only synchronisation is __syncwarp().
//   load x[f*hop + n]*w[n]  (float2 per lane, coalesced; window table pre-scaled by 1/(2W))
//   -> pack as 2048 complex z[n] = (x[2n], x[2n+1])
//   -> lane n2 runs a 64-point in-register FFT over n1 (z[32*n1+n2])                      [stage 1]
//   -> twiddle W_2048^(n2*k1), transpose through shared memory (64 x 33 float2 per warp)
//   -> lane t runs two 32-point in-register FFTs, columns k1 = t and 64-t (lane 0: 32)   [stage 2]
//      so that Z[k] and Z[2048-k] sit in the same lane and the real-FFT untangling needs no exchange;
//      column k1 = 0 (bins 64*j) is a 32-point FFT across the lanes done with shuffles
//   -> |X[k]|^2 written straight from registers to HBM (coalesced 128 B rows) and to the warp's shared buffer
//   -> each lane scans 64 contiguous bins for strict local maxima above the threshold, a warp prefix sum orders
//      them, parabolic interpolation on log-magnitudes gives (freq, dB).
//
// Restates PeriodicFFT<SqMagFftOperation<Window,T>> + extractLocalMaxFreqsMags for whole streams
// (source/rt.resynth.lib.periodicfft.cpp:13-181,251-325; call site source/rt.resynth.lib.cpp:1592-1595);
// the unpinned arithmetic follows oracle/SPEC.md S1-S6.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "ctx.hpp"

namespace ca {

namespace {

constexpr int kN = 4096;         // real FFT length
constexpr int kM = 2048;         // complex FFT length
constexpr int kBins = kM + 1;
constexpr int kExchStride = 33;  // float2 per row of the 64 x 32 transpose buffer (conflict-free both ways)
constexpr int kExchFloat2 = 64 * kExchStride;
constexpr int kTabRows = 21;

struct StftArgs {
  const float* x;
  int64_t n_samples;
  int64_t stream_stride;
  int64_t frames_per_stream;
  int64_t total_frames;
  int32_t n_streams;
  int32_t hop;
  int32_t window;       // samples actually read per frame (<= 4096); the rest is zero padding
  const float* win;     // [4096] window * 1/(2W), zero beyond `window`
  float* sqmag;         // nullable
  int64_t sqmag_row_stride;
  int32_t* peak_bin;    // nullable
  float* peak_freq;
  float* peak_db;
  int32_t* peak_count;
  int32_t max_peaks;
  float thr_below;  // largest float strictly below the sqmag threshold (-1 when there is no threshold)
  float bin_to_hz;
  int32_t freq_as_offset;  // internal: peak_freq receives the sub-bin offset p (round trip computes k+p in double)
};

__device__ __forceinline__ float2 cis_neg(float turns_times_two) {  // exp(-i*pi*arg)
  float s, c;
  sincospif(turns_times_two, &s, &c);
  return make_float2(c, -s);
}

// sqmag of the pair (k, M-k) from Z[k] and Z[M-k]; wp = -i*W_N^k. (The shorter form |X|^2 = S +- C with
// S = |E|^2+|D|^2, C = 2 Re(E conj(wp D)) cancels catastrophically on bins far below their mirror bin: not used.)
__device__ __forceinline__ void untangle(float2 zk, float2 zm, float2 wp, float& sq_k, float& sq_m) {
  const float ex = zk.x + zm.x, ey = zk.y - zm.y;  // Z[k] + conj(Z[M-k])
  const float dx = zk.x - zm.x, dy = zk.y + zm.y;  // Z[k] - conj(Z[M-k])
  const float tx = fmaf(-dy, wp.y, dx * wp.x), ty = fmaf(dx, wp.y, dy * wp.x);
  const float px = ex + tx, py = ey + ty, qx = ex - tx, qy = ey - ty;
  sq_k = fmaf(px, px, py * py);
  sq_m = fmaf(qx, qx, qy * qy);
}

// untangle bins k = colA + 64*K2 and 2048-k = colB + 64*(31-K2) and store them at once (HBM row and/or the warp's padded
// shared buffer), so that the stores are spread over the arithmetic instead of queueing up behind it
template <int K2>
__device__ __forceinline__ void post_pair(const float2 (&A)[32], const float2 (&B)[32], float2 wpbase, float* __restrict__ rowA,
                                          float* __restrict__ rowB, float* __restrict__ sqA, float* __restrict__ sqB) {
  constexpr float2 w64 = {Tw<64, K2>::re, Tw<64, K2>::im};
  float2 wp;
  if constexpr (K2 == 0) wp = wpbase;
  else wp = cmul(wpbase, w64);
  float sk, sm;
  untangle(A[K2], B[31 - K2], wp, sk, sm);
  if (rowA) {
    rowA[64 * K2] = sk;
    rowB[64 * (31 - K2)] = sm;
  }
  if (sqA) {
    sqA[65 * K2] = sk;
    sqB[65 * (31 - K2)] = sm;
  }
}
template <int... K2>
__device__ __forceinline__ void post_all(const float2 (&A)[32], const float2 (&B)[32], float2 wpbase, float* rowA, float* rowB,
                                         float* sqA, float* sqB, std::integer_sequence<int, K2...>) {
  (post_pair<K2>(A, B, wpbase, rowA, rowB, sqA, sqB), ...);
}

// stage-1 twiddle + transposed store for output k1 = J
template <int J>
__device__ __forceinline__ void tw_store(const float2 (&v)[64], const float2 (&wa)[8], const float2 (&wb)[8], float2* exch,
                                         int lane) {
  constexpr int k1 = J;
  constexpr int a = k1 >> 3, b = k1 & 7;
  float2 y = v[J];
  if constexpr (a != 0 && b != 0) y = cmul(y, cmul(wa[a], wb[b]));
  else if constexpr (a != 0) y = cmul(y, wa[a]);
  else if constexpr (b != 0) y = cmul(y, wb[b]);
  exch[k1 * kExchStride + lane] = y;
}
// split variant: twiddle in place, then the real and imaginary planes cross the transpose buffer one after the other
// (half the shared memory per warp: 64 x 33 floats)
template <int J>
__device__ __forceinline__ void tw_inplace(float2 (&v)[64], const float2 (&wa)[8], const float2 (&wb)[8]) {
  constexpr int k1 = J;
  constexpr int a = k1 >> 3, b = k1 & 7;
  if constexpr (a != 0 && b != 0) v[J] = cmul(v[J], cmul(wa[a], wb[b]));
  else if constexpr (a != 0) v[J] = cmul(v[J], wa[a]);
  else if constexpr (b != 0) v[J] = cmul(v[J], wb[b]);
}
template <int... J>
__device__ __forceinline__ void tw_inplace_all(float2 (&v)[64], const float2 (&wa)[8], const float2 (&wb)[8],
                                               std::integer_sequence<int, J...>) {
  (tw_inplace<J>(v, wa, wb), ...);
}
template <bool kImag, int... J>
__device__ __forceinline__ void plane_store_all(const float2 (&v)[64], float* xb, int lane, std::integer_sequence<int, J...>) {
  ((xb[J * kExchStride + lane] = kImag ? v[J].y : v[J].x), ...);
}

template <int... J>
__device__ __forceinline__ void tw_store_all(const float2 (&v)[64], const float2 (&wa)[8], const float2 (&wb)[8],
                                             float2* exch, int lane, std::integer_sequence<int, J...>) {
  (tw_store<J>(v, wa, wb, exch, lane), ...);
}

__device__ __forceinline__ int phys(int k) { return k + (k >> 6); }  // padded index of bin k in the warp's sqmag buffer

// Everything after the load: stage-1 FFT, transpose, stage-2 FFTs, untangling, stores, peak extraction (see file header).
template <int kSync, bool kSplit>
__device__ __forceinline__ void frame_body(float2 (&v)[64], const StftArgs& p, const float2* __restrict__ tab,
                                           float2* __restrict__ exch, const int lane, const int64_t f, const bool active) {
  float* sq = reinterpret_cast<float*>(exch);
  const int colA = lane ? lane : 32;
  const int colB = lane ? 64 - lane : 32;
  const int k2c = static_cast<int>(__brev(static_cast<unsigned>(lane)) >> 27);  // brev5(lane)
  // ---- stage 1: 64-point FFT over n1 ----
  if (kSync == 1 || kSync == 3) __syncthreads();
  if constexpr (kSync == 3) {  // finer lock step: the 64-point FFT in three pieces
    Dit<64, 32, 0>::run(v);
    __syncthreads();
    Dit<64, 32, 32>::run(v);
    __syncthreads();
    Dit<64, 64, 0>::stage(v, std::make_integer_sequence<int, 32>{});
  } else {
    Dit<64, 64, 0>::run(v);  // natural order out: v[k1]
  }
  const float2 c0 = v[0];  // column k1 = 0 (no twiddle), one value per lane

  if (kSync) __syncthreads();
  float2 A[32], B[32];
  {
    float2 wa[8], wb[8];
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      wb[i] = tab[(i - 1) * 32 + lane];
      wa[i] = tab[(6 + i) * 32 + lane];
    }
    wa[0] = wb[0] = make_float2(1.f, 0.f);
    if constexpr (!kSplit) {
      tw_store_all(v, wa, wb, exch, lane, std::make_integer_sequence<int, 64>{});
    } else {
      tw_inplace_all(v, wa, wb, std::make_integer_sequence<int, 64>{});
    }
  }
  // ---- stage 2 inputs: columns colA, colB of the transposed data ----
  if constexpr (!kSplit) {
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) {
      A[bitrev(n2, 5)] = exch[colA * kExchStride + n2];
      B[bitrev(n2, 5)] = exch[colB * kExchStride + n2];
    }
  } else {
    float* xb = reinterpret_cast<float*>(exch);
    plane_store_all<false>(v, xb, lane, std::make_integer_sequence<int, 64>{});
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) {
      A[bitrev(n2, 5)].x = xb[colA * kExchStride + n2];
      B[bitrev(n2, 5)].x = xb[colB * kExchStride + n2];
    }
    __syncwarp();
    plane_store_all<true>(v, xb, lane, std::make_integer_sequence<int, 64>{});
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) {
      A[bitrev(n2, 5)].y = xb[colA * kExchStride + n2];
      B[bitrev(n2, 5)].y = xb[colB * kExchStride + n2];
    }
  }
  __syncwarp();  // exch is reused as the sqmag buffer below
  if (kSync) __syncthreads();
  Dit<32, 32, 0>::run(A);
  if (kSync == 1 || kSync == 3) __syncthreads();
  Dit<32, 32, 0>::run(B);

  // ---- column 0: 32-point FFT across lanes (DIF, bit-reversed result: lane L holds k2 = brev5(L)) ----
  float2 u = c0;
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int d = 16 >> s;
    const float ox = __shfl_xor_sync(0xffffffffu, u.x, d);
    const float oy = __shfl_xor_sync(0xffffffffu, u.y, d);
    if (lane & d) {
      const float2 w = tab[(14 + s) * 32 + lane];
      u = cmul(make_float2(ox - u.x, oy - u.y), w);
    } else {
      u = make_float2(u.x + ox, u.y + oy);
    }
  }
  float sq0, sqM;
  {
    const int pk2 = (32 - k2c) & 31;
    const int plane = static_cast<int>(__brev(static_cast<unsigned>(pk2)) >> 27);
    const float mx = __shfl_sync(0xffffffffu, u.x, plane);
    const float my = __shfl_sync(0xffffffffu, u.y, plane);
    untangle(u, make_float2(mx, my), tab[19 * 32 + lane], sq0, sqM);
  }

  // ---- untangle the real FFT (bins k = colA + 64*k2 and 2048 - k) and write |X|^2: HBM row (coalesced across lanes)
  //      and the warp's shared buffer (padded, conflict-free) ----
  if (kSync) __syncthreads();
  float* row = (p.sqmag && active) ? p.sqmag + f * p.sqmag_row_stride : nullptr;
  float* sqw = p.peak_count ? sq : nullptr;
  post_all(A, B, tab[20 * 32 + lane], row ? row + colA : nullptr, row ? row + colB : nullptr, sqw ? sqw + colA : nullptr,
           sqw ? sqw + colB : nullptr, std::make_integer_sequence<int, 32>{});
  if (row) {
    row[64 * k2c] = sq0;
    if (lane == 0) row[kM] = sqM;
  }
  if (kSync == 1 || kSync == 3) __syncthreads();
  if (p.peak_count) {
    sq[65 * k2c] = sq0;
    if (lane == 0) sq[phys(kM)] = sqM;
    __syncwarp();

    // ---- peak scan: lane owns bins [64*lane, 64*lane+64) ----
    // strict local maximum at or above the threshold  <=>  c > max(l, r, largest float below the threshold)
    const int base = 65 * lane;
    uint32_t mlo = 0, mhi = 0;
    float l = lane ? sq[base - 2] : __int_as_float(0x7f800000);  // bin 0 is never a peak
    float c = sq[base];
#pragma unroll
    for (int j = 0; j < 64; ++j) {
      const float r = sq[j == 63 ? base + 65 : base + j + 1];
      const bool is_peak = c > fmaxf(fmaxf(l, r), p.thr_below);
      if (j < 32) mlo |= is_peak ? (1u << j) : 0u;
      else mhi |= is_peak ? (1u << (j - 32)) : 0u;
      l = c;
      c = r;
    }
    uint64_t mask = (static_cast<uint64_t>(mhi) << 32) | mlo;
    const int cnt = __popcll(mask);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int slot = incl - cnt;
    if (lane == 0 && active) p.peak_count[f] = total;
    // each lane lists the bins of its peaks (ascending), then the peaks are spread over the lanes: the interpolation
    // runs once per 32 peaks instead of once per peak of the busiest lane, and the stores are coalesced
    // (the list lives behind the sqmag buffer, or, in the split variant whose buffer is half as large, in the output itself)
    const int64_t out_base = f * static_cast<int64_t>(p.max_peaks);
    int* klist = kSplit ? p.peak_bin + out_base : reinterpret_cast<int*>(sq + 2112);
    const int n_store = active ? min(total, p.max_peaks) : 0;
    while (mask && slot < n_store) {
      const int j = __ffsll(static_cast<long long>(mask)) - 1;
      mask &= mask - 1;
      klist[slot++] = 64 * lane + j;
    }
    __syncwarp();
    for (int idx = lane; idx < n_store; idx += 32) {
      const int k = kSplit ? *reinterpret_cast<volatile int*>(klist + idx) : klist[idx];
      const float fl = 1e-37f;
      const float sa = fmaxf(sq[phys(k - 1)], fl), sb = fmaxf(sq[phys(k)], fl), sc = fmaxf(sq[phys(k + 1)], fl);
      // dB differences in log2 units, from ratios to the centre bin (well conditioned); MUFU.LG2 / MUFU.RCP precision
      const float inv_b = __frcp_rn(sb);
      const float la = __log2f(sa * inv_b), lc = __log2f(sc * inv_b);  // (alpha-beta), (gamma-beta) / (10 log10 2)
      const float num = la - lc;
      const float den = la + lc;
      const float pp = den != 0.f ? 0.5f * __fdividef(num, den) : 0.f;
      constexpr float kDbPerLog2 = 3.01029995663981195f;  // 10*log10(2)
      const float beta = kDbPerLog2 * __log2f(sb);
      if (!kSplit) p.peak_bin[out_base + idx] = k;
      p.peak_freq[out_base + idx] = p.freq_as_offset ? pp : (static_cast<float>(k) + pp) * p.bin_to_hz;
      p.peak_db[out_base + idx] = beta - 0.25f * (kDbPerLog2 * num) * pp;
    }
    __syncwarp();  // the next frame overwrites the buffer
  }
}

template <int WARPS>
__device__ __forceinline__ void build_tables(float2* tab) {
  for (int i = threadIdx.x; i < kTabRows * 32; i += WARPS * 32) {
    const int r = i >> 5, l = i & 31;
    float2 w;
    if (r < 7) {
      w = cis_neg(2.f * static_cast<float>(l * (r + 1)) / kM);  // W_2048^(l*b), b = r+1
    } else if (r < 14) {
      w = cis_neg(2.f * static_cast<float>(l * 8 * (r - 6)) / kM);  // W_2048^(l*8a), a = r-6
    } else if (r < 19) {
      const int d = 16 >> (r - 14);
      w = cis_neg(static_cast<float>(l % d) / d);  // exp(-2 pi i (l%d)/(2d))
    } else if (r == 19) {
      const int k = 64 * static_cast<int>(__brev(static_cast<unsigned>(l)) >> 27);  // k = 64*brev5(l)
      float s, c;
      sincospif(2.f * static_cast<float>(k) / kN, &s, &c);
      w = make_float2(-s, -c);  // -i * W_4096^k
    } else {
      const int col = l ? l : 32;
      float s, c;
      sincospif(2.f * static_cast<float>(col) / kN, &s, &c);
      w = make_float2(-s, -c);
    }
    tab[i] = w;
  }
}

// kFull: window == 4096 (no zero padding, no load predicates); kVec2: 8-byte aligned float2 loads are legal.
// kSync: the warps of the CTA walk the frame code in step (one __syncthreads per phase). The frame body is ~70 KB of
// straight-line SASS; unsynchronised warps each stream it from L2 on their own and instruction fetch becomes the limiter
// (ncu: stall_no_inst 26 % of samples, profiles/), in step they share the fetches.
template <int WARPS, int MINB, int kSync, bool kVec2, bool kFull>
__global__ void __launch_bounds__(WARPS * 32, MINB) stft4096_kernel(const __grid_constant__ StftArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tab = reinterpret_cast<float2*>(smem_raw);  // [kTabRows][32]
  float2* exch_all = tab + kTabRows * 32;             // [WARPS][kExchFloat2]

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;

  build_tables<WARPS>(tab);
  __syncthreads();

  float2* exch = exch_all + warp * kExchFloat2;

  // CTA-uniform rounds: round r covers frames base .. base+WARPS-1, one per warp (the tail warps of the last round idle)
  const int fps = static_cast<int>(p.frames_per_stream);
  for (int64_t fbase = static_cast<int64_t>(blockIdx.x) * WARPS; fbase < p.total_frames;
       fbase += static_cast<int64_t>(gridDim.x) * WARPS) {
    const int64_t f = fbase + warp;
    const bool active = f < p.total_frames;
    if (kSync) __syncthreads();
    if (!kSync && !active) break;
    int stream, fr;
    if (p.total_frames < (1ll << 31)) {
      const unsigned uf = static_cast<unsigned>(active ? f : 0);
      stream = static_cast<int>(uf / static_cast<unsigned>(fps));
      fr = static_cast<int>(uf - static_cast<unsigned>(stream) * static_cast<unsigned>(fps));
    } else {
      stream = static_cast<int>((active ? f : 0) / fps);
      fr = static_cast<int>((active ? f : 0) - static_cast<int64_t>(stream) * fps);
    }
    const float* xs = p.x + stream * p.stream_stride + static_cast<int64_t>(fr) * p.hop;

    // ---- load + window: lane n2 takes z[32*n1 + n2] = (x[64*n1+2*n2], x[64*n1+2*n2+1]) ----
    float2 v[64];
#pragma unroll
    for (int n1 = 0; n1 < 64; ++n1) {
      const int idx = 64 * n1 + 2 * lane;
      float2 s;
      if (kVec2 && kFull) {
        s = ldg2(xs + idx);
      } else if (kVec2) {
        s = (idx < p.window) ? ldg2(xs + idx) : make_float2(0.f, 0.f);
      } else {
        s.x = (kFull || idx < p.window) ? __ldg(xs + idx) : 0.f;
        s.y = (kFull || idx + 1 < p.window) ? __ldg(xs + idx + 1) : 0.f;
      }
      const float2 w = ldg2(p.win + idx);
      v[bitrev(n1, 6)] = make_float2(s.x * w.x, s.y * w.y);  // DIT: bit-reversed register order in
    }

    frame_body<kSync, false>(v, p, tab, exch, lane, f, active);
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Ring variant (window == 4096, hop % 4 == 0, hop <= 1024, 16-byte aligned streams): the CTA walks a CONTIGUOUS range of
// frames of one stream, WARPS frames per round, and keeps the samples it needs in an 8192-float shared-memory ring that is
// fed by 1-D bulk async copies (cp.async.bulk + mbarrier, SASS UBLKCP): the 4*hop new samples of the next round are in
// flight during the whole arithmetic of the current round, so the load phase reads shared memory instead of waiting on
// L2/HBM. Frames overlap by (window-hop)/window, so each sample crosses L2->SM once per CTA instead of 16 times.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kRing = 8192;  // floats; >= 4096 + 3*hop + 4*hop for hop <= 585, and window + 3*hop <= 8192 for hop <= 1024

__device__ __forceinline__ uint32_t smem_u32(const void* ptr) { return static_cast<uint32_t>(__cvta_generic_to_shared(ptr)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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

// copy samples [s0, s1) of a stream into the ring (position = sample index mod kRing); s0, s1 multiples of 4
__device__ __forceinline__ uint32_t ring_fill(float* ring, const float* stream_base, int64_t s0, int64_t s1, uint64_t* bar) {
  if (s1 <= s0) return 0;
  const uint32_t bytes = static_cast<uint32_t>(s1 - s0) * 4u;
  mbar_expect_tx(bar, bytes);
  int64_t s = s0;
  while (s < s1) {
    const int pos = static_cast<int>(s & (kRing - 1));
    const int64_t n = min(s1 - s, static_cast<int64_t>(kRing - pos));
    bulk_g2s(ring + pos, stream_base + s, static_cast<uint32_t>(n) * 4u, bar);
    s += n;
  }
  return bytes;
}

struct RingPlan {
  int32_t units_per_stream;
  int32_t frames_per_unit;  // multiple of WARPS
  int32_t total_units;
};

template <int WARPS, int MINB, bool kSplit>
__global__ void __launch_bounds__(WARPS * 32, MINB) stft4096_ring_kernel(const __grid_constant__ StftArgs p,
                                                                        const __grid_constant__ RingPlan rp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* ring = reinterpret_cast<float*>(smem_raw);                    // [kRing]
  float2* tab = reinterpret_cast<float2*>(ring + kRing);               // [kTabRows][32]
  float2* exch_all = tab + kTabRows * 32;                              // [WARPS][kExchFloat2 (half of it when kSplit)]
  __shared__ __align__(8) uint64_t bar;

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  build_tables<WARPS>(tab);
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  float2* exch = exch_all + warp * (kSplit ? kExchFloat2 / 2 : kExchFloat2);
  const int fps = static_cast<int>(p.frames_per_stream);
  uint32_t parity = 0;

  for (int u = blockIdx.x; u < rp.total_units; u += gridDim.x) {
    const int stream = u / rp.units_per_stream;
    const int fa = (u - stream * rp.units_per_stream) * rp.frames_per_unit;
    const int fb = min(fps, fa + rp.frames_per_unit);
    const float* xs = p.x + static_cast<int64_t>(stream) * p.stream_stride;
    // all warps are past the load phase of the previous unit (barrier after its last load): the ring can be refilled
    bool pending = false;
    if (threadIdx.x == 0) {
      const int last = min(fa + WARPS - 1, fb - 1);
      ring_fill(ring, xs, static_cast<int64_t>(fa) * p.hop, static_cast<int64_t>(last) * p.hop + kN, &bar);
    }
    pending = true;
    for (int f0 = fa; f0 < fb; f0 += WARPS) {
      const int fr = f0 + warp;
      const bool active = fr < fb;
      const int64_t f = static_cast<int64_t>(stream) * fps + (active ? fr : fa);
      if (pending) {
        mbar_wait(&bar, parity);
        parity ^= 1u;
      }
      // ---- load + window from the ring ----
      float2 v[64];
      const int pos0 = static_cast<int>((static_cast<int64_t>(active ? fr : fa) * p.hop) & (kRing - 1));
#pragma unroll
      for (int n1 = 0; n1 < 64; ++n1) {
        const int idx = 64 * n1 + 2 * lane;
        const float2 sm = *reinterpret_cast<const float2*>(ring + ((pos0 + idx) & (kRing - 1)));
        const float2 w = ldg2(p.win + idx);
        v[bitrev(n1, 6)] = make_float2(sm.x * w.x, sm.y * w.y);
      }
      __syncthreads();  // every warp has its samples in registers: the ring may take the next round's data
      {
        const int next_last = min(f0 + 2 * WARPS - 1, fb - 1);
        const int this_last = min(f0 + WARPS - 1, fb - 1);
        pending = next_last > this_last;  // CTA-uniform
        if (pending && threadIdx.x == 0)
          ring_fill(ring, xs, static_cast<int64_t>(this_last) * p.hop + kN, static_cast<int64_t>(next_last) * p.hop + kN, &bar);
      }
      frame_body<1, kSplit>(v, p, tab, exch, lane, f, active);
    }
    __syncthreads();
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Generic path for FFT lengths other than 4096 (1024 .. 32768; the reference default window of 0.1814 s is 8192 points
// at 44.1 kHz): frames are windowed into a work buffer, transformed by K1 (rfft.cu) and a third kernel forms |X|^2 and
// extracts the peaks. Three passes over HBM instead of one — correct for every size, tuned for none.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) window_frames_kernel(const float* __restrict__ x, const float* __restrict__ win,
                                                            float* __restrict__ out, int64_t stream_stride, int frames_per_stream,
                                                            int hop, int window, int n_fft, int64_t f_begin, int64_t n_frames) {
  // grid.x: frames of this group, threads stride over the frame
  const int64_t f = f_begin + blockIdx.x;
  if (blockIdx.x >= n_frames) return;
  const int64_t stream = f / frames_per_stream;
  const int64_t fr = f - stream * frames_per_stream;
  const float* xs = x + stream * stream_stride + fr * hop;
  float* o = out + static_cast<int64_t>(blockIdx.x) * n_fft;
  for (int n = threadIdx.x; n < n_fft; n += 256) o[n] = n < window ? __ldg(xs + n) * __ldg(win + n) : 0.f;
}

// one warp per frame: |X|^2 (scale folded into the window), HBM row, peaks as in the fused kernel
__global__ void __launch_bounds__(128) sqmag_peaks_kernel(const float2* __restrict__ bins, int n_fft, int64_t f_begin,
                                                          int64_t n_frames, const __grid_constant__ StftArgs p) {
  const int lane = threadIdx.x & 31;
  const int64_t local = static_cast<int64_t>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (local >= n_frames) return;
  const int64_t f = f_begin + local;
  const int half = n_fft / 2;
  const float2* X = bins + local * (half + 1);
  float* row = p.sqmag ? p.sqmag + f * p.sqmag_row_stride : nullptr;
  if (row)
    for (int k = lane; k <= half; k += 32) {
      const float2 z = X[k];
      row[k] = fmaf(z.x, z.x, z.y * z.y);
    }
  if (!p.peak_count) return;
  auto sq = [&](int k) {
    const float2 z = X[k];
    return fmaf(z.x, z.x, z.y * z.y);
  };
  // lane owns bins [lane*B, lane*B + B), B = half/32; bins 1 .. half-1 are candidates
  const int B = half / 32;
  const int k0 = lane * B;
  int cnt = 0;
  {
    float l = k0 > 0 ? sq(k0 - 1) : __int_as_float(0x7f800000);
    float c = sq(k0);
    for (int j = 0; j < B; ++j) {
      const float r = sq(k0 + j + 1);
      cnt += (c > fmaxf(fmaxf(l, r), p.thr_below)) ? 1 : 0;
      l = c;
      c = r;
    }
  }
  int incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  int slot = incl - cnt;
  if (lane == 0) p.peak_count[f] = total;
  const int64_t out_base = f * static_cast<int64_t>(p.max_peaks);
  float l = k0 > 0 ? sq(k0 - 1) : __int_as_float(0x7f800000);
  float c = sq(k0);
  for (int j = 0; j < B && slot < p.max_peaks; ++j) {
    const float r = sq(k0 + j + 1);
    if (c > fmaxf(fmaxf(l, r), p.thr_below)) {
      const int k = k0 + j;
      const float fl = 1e-37f;
      const float sa = fmaxf(l, fl), sb = fmaxf(c, fl), sc = fmaxf(r, fl);
      const float inv_b = __frcp_rn(sb);
      const float la = __log2f(sa * inv_b), lc = __log2f(sc * inv_b);
      const float num = la - lc, den = la + lc;
      const float pp = den != 0.f ? 0.5f * __fdividef(num, den) : 0.f;
      constexpr float kDbPerLog2 = 3.01029995663981195f;
      p.peak_bin[out_base + slot] = k;
      p.peak_freq[out_base + slot] = p.freq_as_offset ? pp : (static_cast<float>(k) + pp) * p.bin_to_hz;
      p.peak_db[out_base + slot] = kDbPerLog2 * __log2f(sb) - 0.25f * (kDbPerLog2 * num) * pp;
      ++slot;
    }
    l = c;
    c = r;
  }
}

}  // namespace

int stft_plan(ca_ctx* ctx, const ca_stft_params* p) {
  if (!p || p->window < 2 || (p->window & 1) || p->hop < 1 || p->max_peaks < 0 || p->sample_rate <= 0)
    return fail(ctx, CA_ERR_INVALID, "ca_stft_plan: window must be even and >= 2, hop >= 1");
  if (p->window_type != CA_WINDOW_RECTANGULAR && p->window_type != CA_WINDOW_GAUSSIAN)
    return fail(ctx, CA_ERR_INVALID, "ca_stft_plan: unknown window type");
  const int n_fft = ca_fft_length_for(p->window, 1);
  if (n_fft < 1024 || n_fft > 32768)
    return fail(ctx, CA_ERR_UNSUPPORTED, "ca_stft_plan: fft length must be within 1024 .. 32768 (512 < window <= 32768)");
  // half window, normalised to mean 1 (oracle/SPEC.md S2,S3), then scaled by 1/W (S4) and by the 1/2 of the real-FFT
  // untangling so that the kernel's |E+T|^2 is |X|^2 directly.
  const int W = p->window, H = W / 2;
  std::vector<double> hw(H);
  long double sum = 0;
  for (int j = 0; j < H; ++j) {
    long double v = 1;
    if (p->window_type == CA_WINDOW_GAUSSIAN) {
      long double u = static_cast<long double>(p->sigma_factor) * (j + 0.5L) / H;
      v = expl(-0.5L * u * u);
    }
    hw[j] = static_cast<double>(v);
    sum += v;
  }
  const double norm = static_cast<double>(H / sum) * (n_fft == kN ? 0.5 : 1.0) / W;
  std::vector<float> full(n_fft, 0.f);
  for (int n = 0; n < W; ++n) full[n] = static_cast<float>((n < H ? hw[H - 1 - n] : hw[n - H]) * norm);
  if (!ctx->d_window || ctx->n_fft != n_fft) {
    if (ctx->d_window) cudaFree(ctx->d_window);
    ctx->d_window = nullptr;
    CA_CUDA_OK(ctx, cudaMalloc(&ctx->d_window, sizeof(float) * n_fft));
  }
  CA_CUDA_OK(ctx, cudaMemcpy(ctx->d_window, full.data(), sizeof(float) * n_fft, cudaMemcpyHostToDevice));
  ctx->n_fft = n_fft;
  ctx->stft = *p;
  ctx->thr_sq = isinf(p->min_db) && p->min_db < 0 ? 0.f : static_cast<float>(pow(10.0, p->min_db / 10.0));
  ctx->stft_planned = true;
  return CA_OK;
}

void stft_release(ca_ctx* ctx) {
  if (ctx->d_stft_work) cudaFree(ctx->d_stft_work);
  ctx->d_stft_work = nullptr;
  ctx->stft_work_bytes = 0;
  if (ctx->d_window) cudaFree(ctx->d_window);
  ctx->d_window = nullptr;
  ctx->stft_planned = false;
}

int stft_analyze(ca_ctx* ctx, const float* d_x, int64_t n_samples, int64_t n_streams, int64_t stream_stride,
                 float* d_sqmag, int64_t sqmag_row_stride, int32_t* d_peak_bin, float* d_peak_freq, float* d_peak_db,
                 int32_t* d_peak_count, cudaStream_t stream) {
  if (!ctx->stft_planned) return fail(ctx, CA_ERR_NOT_PLANNED, "ca_stft_analyze: call ca_stft_plan first");
  if (!d_x || n_streams < 0 || n_samples < 0 || stream_stride < n_samples)
    return fail(ctx, CA_ERR_INVALID, "ca_stft_analyze: bad input shape");
  const bool any_peak = d_peak_bin || d_peak_freq || d_peak_db || d_peak_count;
  if (any_peak && !(d_peak_bin && d_peak_freq && d_peak_db && d_peak_count))
    return fail(ctx, CA_ERR_INVALID, "ca_stft_analyze: peak outputs must be all null or all non-null");
  if (d_sqmag && sqmag_row_stride < ctx->n_fft / 2 + 1)
    return fail(ctx, CA_ERR_INVALID, "ca_stft_analyze: sqmag_row_stride < n_fft/2+1");
  const int64_t frames = ca_stft_num_frames(n_samples, ctx->stft.window, ctx->stft.hop);
  const int64_t total = frames * n_streams;
  if (total == 0) return CA_OK;
  if (frames >= (1ll << 31) - (1 << 20) || n_streams >= (1ll << 31))
    return fail(ctx, CA_ERR_UNSUPPORTED, "ca_stft_analyze: more than 2^31 frames per stream or streams per call");

  StftArgs a{};
  a.x = d_x;
  a.n_samples = n_samples;
  a.stream_stride = stream_stride;
  a.frames_per_stream = frames;
  a.total_frames = total;
  a.n_streams = static_cast<int32_t>(n_streams);
  a.hop = ctx->stft.hop;
  a.window = ctx->stft.window;
  a.win = ctx->d_window;
  a.sqmag = d_sqmag;
  a.sqmag_row_stride = sqmag_row_stride;
  a.peak_bin = d_peak_bin;
  a.peak_freq = d_peak_freq;
  a.peak_db = d_peak_db;
  a.peak_count = d_peak_count;
  a.max_peaks = ctx->stft.max_peaks;
  a.thr_below = ctx->thr_sq > 0.f ? nextafterf(ctx->thr_sq, 0.f) : -1.f;
  a.bin_to_hz = static_cast<float>(static_cast<double>(ctx->stft.sample_rate) / ctx->n_fft);
  if (ctx->n_fft != kN) {
    // generic three-pass path, frames processed in groups that bound the work buffers to ~256 MiB each
    const int n_fft = ctx->n_fft, half = n_fft / 2;
    if (frames >= (1ll << 31)) return fail(ctx, CA_ERR_UNSUPPORTED, "ca_stft_analyze: too many frames per stream");
    const int64_t group = std::max<int64_t>(1, std::min<int64_t>(total, (256ll << 20) / (sizeof(float) * n_fft)));
    const size_t need = sizeof(float) * n_fft * group + sizeof(float2) * (half + 1) * group;
    if (ctx->stft_work_bytes < need) {
      if (ctx->d_stft_work) cudaFree(ctx->d_stft_work);
      ctx->d_stft_work = nullptr;
      ctx->stft_work_bytes = 0;
      CA_CUDA_OK(ctx, cudaMalloc(&ctx->d_stft_work, need));
      ctx->stft_work_bytes = need;
    }
    float* wbuf = static_cast<float*>(ctx->d_stft_work);
    float* bbuf = wbuf + static_cast<size_t>(n_fft) * group;
    a.thr_below = ctx->thr_sq > 0.f ? nextafterf(ctx->thr_sq, 0.f) : -1.f;
    a.max_peaks = ctx->stft.max_peaks;
    a.freq_as_offset = ctx->stft_freq_as_offset ? 1 : 0;
    for (int64_t f0 = 0; f0 < total; f0 += group) {
      const int64_t nf = std::min(group, total - f0);
      window_frames_kernel<<<static_cast<unsigned>(nf), 256, 0, stream>>>(d_x, ctx->d_window, wbuf, stream_stride, static_cast<int>(frames),
                                                                         ctx->stft.hop, ctx->stft.window, n_fft, f0, nf);
      CA_CUDA_OK(ctx, cudaGetLastError());
      int rc = rfft_batched(ctx, wbuf, n_fft, nf, bbuf, stream);
      if (rc) return rc;
      sqmag_peaks_kernel<<<static_cast<unsigned>((nf + 3) / 4), 128, 0, stream>>>(reinterpret_cast<const float2*>(bbuf), n_fft, f0, nf, a);
      CA_CUDA_OK(ctx, cudaGetLastError());
      ctx->launches += 2;
    }
    return CA_OK;
  }
  a.freq_as_offset = ctx->stft_freq_as_offset ? 1 : 0;

  // tuning knobs (defaults are the measured best): CA_STFT_VARIANT = 0..4
  static const int variant = [] {
    const char* e = getenv("CA_STFT_VARIANT");
    return e ? atoi(e) : 4;
  }();
  const bool vec2 = (ctx->stft.hop % 2 == 0) && (stream_stride % 2 == 0) && (reinterpret_cast<uintptr_t>(d_x) % 8 == 0) &&
                    (ctx->stft.window % 2 == 0);
  const bool full = ctx->stft.window == kN;
  void (*kern)(StftArgs) = nullptr;
  int WARPS = 4;
#define CA_PICK(W, MB, SY)                                                                                       \
  do {                                                                                                           \
    WARPS = W;                                                                                                   \
    kern = vec2 ? (full ? stft4096_kernel<W, MB, SY, true, true> : stft4096_kernel<W, MB, SY, true, false>)     \
                : (full ? stft4096_kernel<W, MB, SY, false, true> : stft4096_kernel<W, MB, SY, false, false>);  \
  } while (0)
  // measured on B200, 32 x 60 s streams, Mframes/s (profiles/README.md): 0 (4 warps x 3 CTAs, free running) 144;
  // 1 (8x1, in step) 127; 2 (8x1, free) 127; 3 (12x1, in step) 140; 4 (4x3, in step) 153 <- default;
  // 5 (4x3, in step, bulk-copy ring + split transpose) 147; 6 (4x2, in step, bulk-copy ring) 150
  switch (variant) {
    case 0: CA_PICK(4, 3, 0); break;
    case 1: CA_PICK(8, 1, 1); break;
    case 3: CA_PICK(12, 1, 1); break;
    case 7: CA_PICK(4, 3, 2); break;
    case 8: CA_PICK(4, 3, 3); break;
    default: CA_PICK(4, 3, 1); break;
  }
#undef CA_PICK
  const bool ring_ok = full && ctx->stft.hop % 4 == 0 && ctx->stft.hop <= 1024 && stream_stride % 4 == 0 &&
                       reinterpret_cast<uintptr_t>(d_x) % 16 == 0 && frames < (1ll << 30) && variant >= 5;
  if (ring_ok) {
    constexpr int RW = 4;
    const bool split = variant != 6;
    auto rk = split ? stft4096_ring_kernel<RW, 3, true> : stft4096_ring_kernel<RW, 2, false>;
    const size_t rsmem = sizeof(float) * kRing + sizeof(float2) * (kTabRows * 32 + RW * (split ? kExchFloat2 / 2 : kExchFloat2));
    CA_CUDA_OK(ctx, cudaFuncSetAttribute(rk, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(rsmem)));
    int rper_sm = 0;
    CA_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&rper_sm, rk, RW * 32, rsmem));
    if (rper_sm < 1) rper_sm = 1;
    const int64_t rgrid_full = static_cast<int64_t>(ctx->sm_count) * rper_sm;
    // units: contiguous frame ranges inside a stream, about 4 per CTA for balance
    RingPlan rp{};
    int64_t ups = (4 * rgrid_full + n_streams - 1) / n_streams;
    if (ups < 1) ups = 1;
    int64_t fpu = (frames + ups - 1) / ups;
    fpu = (fpu + RW - 1) / RW * RW;
    if (fpu < 8 * RW) fpu = 8 * RW;  // amortise the unit prologue (a full window) over at least 8 rounds
    ups = (frames + fpu - 1) / fpu;
    rp.units_per_stream = static_cast<int32_t>(ups);
    rp.frames_per_unit = static_cast<int32_t>(fpu);
    const int64_t units = ups * n_streams;
    if (units < (1ll << 31)) {
      rp.total_units = static_cast<int32_t>(units);
      const int64_t rgrid = std::min<int64_t>(rgrid_full, units);
      rk<<<static_cast<unsigned>(rgrid), RW * 32, rsmem, stream>>>(a, rp);
      CA_CUDA_OK(ctx, cudaGetLastError());
      ctx->launches += 1;
      return CA_OK;
    }
  }
  const size_t smem = sizeof(float2) * (kTabRows * 32 + WARPS * kExchFloat2);
  CA_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int per_sm = 0;
  CA_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
  if (per_sm < 1) per_sm = 1;
  if (const char* e = getenv("CA_STFT_CTAS_PER_SM")) per_sm = std::max(1, std::min(per_sm, atoi(e)));  // tuning experiments
  int64_t grid = static_cast<int64_t>(ctx->sm_count) * per_sm;
  const int64_t need = (total + WARPS - 1) / WARPS;
  if (grid > need) grid = need;
  kern<<<static_cast<unsigned>(grid), WARPS * 32, smem, stream>>>(a);
  CA_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches += 1;
  return CA_OK;
}

}  // namespace ca
