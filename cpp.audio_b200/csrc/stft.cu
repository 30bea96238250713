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
#include <cuda.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "ctx.hpp"
#include "packed.cuh"

namespace ca {

namespace {

constexpr int kN = 4096;         // real FFT length
constexpr int kM = 2048;         // complex FFT length
constexpr int kBins = kM + 1;
// transpose buffer of a warp, 64 x 32 complex values in 32 * kExchStride slots.
// Default: element (k1, n2) at (k1 >> 1)*kExchStride + 2*n2 + (k1 & 1); a lane stores two consecutive k1 with one 128-bit
// store, the readers (one column k1 per lane, consecutive columns in consecutive lanes) are conflict-free because
// kExchStride = 2 (mod 16). ptxas gathers each register quad of a 128-bit store with MOVs (about 100 per frame).
// CA_STFT_EXCH64 (A/B): element (k1, n2) at k1*33 + n2 with 64-bit stores: no MOVs, 32 more store instructions per frame,
// the same number of shared-memory wavefronts; measured 1 % slower (211.6 against 213.6 M frames/s): the load/store unit's
// instruction queue weighs more than the ALU's.
constexpr int kExchStride = 66;
constexpr int kExchFloat2 = 32 * kExchStride;
constexpr int kWinStride = 66;  // window table: 64 pairs per lane (+2 pad), the 32 even-n1 pairs first, then the odd ones
#ifdef CA_STFT_EXCH64
constexpr bool kExch128 = false;
__device__ __forceinline__ int exch_idx(int k1, int n2) { return k1 * (kExchStride / 2) + n2; }
#else
constexpr bool kExch128 = true;
__device__ __forceinline__ int exch_idx(int k1, int n2) { return (k1 >> 1) * kExchStride + 2 * n2 + (k1 & 1); }
#endif

#ifdef CA_STFT_TIMELINE
#define CA_STAMP(i)                                                                                        \
  do {                                                                                                     \
    if (p.dbg && lane == 0 && blockIdx.x < 2 && dbg_frame < 24) p.dbg[((blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 24 + dbg_frame) * 12 + (i)] = clock64(); \
  } while (0)
#else
#define CA_STAMP(i) \
  do {              \
  } while (0)
#endif

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

// ---------------------------------------------------------------------------------------------------------------
// K2: stft4096p_kernel. One warp owns one frame; 2048 complex points = 64 (registers of a lane) x 32 (lanes), all complex
// arithmetic on FFMA2 / FADD2 / FMUL2 (packed.cuh): a complex value is one register pair, a constant-twiddle butterfly is
// 3 instructions. The inter-stage twiddle W_2048^(lane*k1), k1 = 8a+b, is applied as two table twiddles (wb[b], then
// wa[a]: 112 two-instruction multiplies, no products to form); the untangling twiddle -i W_4096^k = wpbase(lane)*W_64^k2
// is an immediate multiply followed by a per-lane one. The frame loop is software-pipelined by hand: the samples of the
// warp's NEXT frame are requested right after the spectrum of the current one is stored (the 128 data registers are
// free then) and land while the peak list is extracted, so the FFT never waits on L2/HBM. The window lives in shared
// memory. Frame body ~2.3 K instructions (the scalar predecessor: 4.5 K, 70 KB — it thrashed the 32 KB instruction cache).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTab4Rows = 20;  // float4 (wr, s, -wi, wi): rows 0..6 wb[1..7], 7..13 wa[1..7], 14 wpbase, 15..19 column-0 FFT
constexpr int kSqBlock = 68;   // floats per 64-bin block of the warp's |X|^2 buffer (16-byte aligned, conflict-free)
constexpr int kSqFloats = 32 * kSqBlock;  // 2176: bins 0..2047 in 32 padded blocks; the 4 padding floats of each block hold
                                          // the peak-bin list (127 entries) and, in the very last one, bin 2048
constexpr int kKlistCap = 127;
constexpr int kStageOff = kSqFloats / 2;  // in c2 units: the upper 8 KB of the warp's buffer stage the next frame's even half
static_assert(kSqFloats * 4 + 32 * 32 * 8 == kExchFloat2 * 8, "spectrum + staged half frame must fill the exchange buffer");

__device__ __forceinline__ int physp(int k) { return k == kM ? kSqFloats - 1 : k + 4 * (k >> 6); }
__device__ __forceinline__ int klist_slot(int s) { return kSqBlock * (s >> 2) + 64 + (s & 3); }

// 3-D tiled tensor copy global -> shared (SASS UTMALDG): ONE instruction moves a strided box, here the 32 rows of 64
// samples that make the even (or odd) half of a frame
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

template <int WARPS>
__device__ __forceinline__ void build_tables_p(float4* tab4, float2* tab2, float2* win_s, const float* __restrict__ win) {
  for (int i = threadIdx.x; i < kTab4Rows * 32; i += WARPS * 32) {
    const int r = i >> 5, l = i & 31;
    float2 w;
    float sgn = 0.f;
    if (r < 7) {
      w = cis_neg(2.f * static_cast<float>(l * (r + 1)) / kM);  // W_2048^(l*b), b = r+1
    } else if (r < 14) {
      w = cis_neg(2.f * static_cast<float>(l * 8 * (r - 6)) / kM);  // W_2048^(l*8a), a = r-6
    } else if (r == 14) {
      const int col = l ? l : 32;
      float s, c;
      sincospif(2.f * static_cast<float>(col) / kN, &s, &c);
      w = make_float2(-s, -c);  // -i * W_4096^col
    } else {
      // column-0 FFT across lanes, DIF stage s = r-15, distance d = 16 >> s: lanes with bit d take (other - own) * w,
      // the others (own + other) * 1; branch-free as (other + sgn*own) * w'
      const int d = 16 >> (r - 15);
      if (l & d) {
        w = cis_neg(static_cast<float>(l % d) / d);  // exp(-2 pi i (l%d)/(2d))
        sgn = -1.f;
      } else {
        w = make_float2(1.f, 0.f);
        sgn = 1.f;
      }
    }
    tab4[i] = make_float4(w.x, sgn, -w.y, w.y);
  }
  for (int l = threadIdx.x; l < 32; l += WARPS * 32) {
    const int k = 64 * static_cast<int>(__brev(static_cast<unsigned>(l)) >> 27);  // k = 64*brev5(l)
    float s, c;
    sincospif(2.f * static_cast<float>(k) / kN, &s, &c);
    tab2[l] = make_float2(-s, -c);  // -i * W_4096^k
  }
  // win_s[lane][par][r] = (w[64*n1 + 2*lane], w[64*n1 + 2*lane + 1]), n1 = 2r + par: a lane reads two of its pairs with
  // one conflict-free 128-bit load
  for (int i = threadIdx.x; i < kM; i += WARPS * 32) {
    const int n1 = i >> 5, l = i & 31;
    win_s[l * kWinStride + (n1 & 1) * 32 + (n1 >> 1)] = __ldg(reinterpret_cast<const float2*>(win) + i);
  }
}

// inter-stage twiddle W_2048^(lane*k1), k1 = 8a + b, as wb[b] then wa[a]. Two table twiddles are live at a time (the one
// sweeping its 8 elements and the next one, already requested): with all 14 resident the kernel spilled its bookkeeping
// registers, with one the shared-memory latency of each load was exposed (14 x ~35 clocks per frame).
// Table rows: 0..6 wb[1..7], 7..13 wa[1..7].
template <int R, int... E>
__device__ __forceinline__ void tw_row_p(c2 (&v)[64], const TwP& w, std::integer_sequence<int, E...>) {
  if constexpr (R < 7) ((v[8 * E + R + 1] = cmul_tw(v[8 * E + R + 1], w)), ...);  // wb[R+1]: elements 8a + b, all a
  else ((v[8 * (R - 6) + E] = cmul_tw(v[8 * (R - 6) + E], w)), ...);              // wa[R-6]: elements 8a + b, all b
}
template <int A>
__device__ __forceinline__ void tw_store8_p(const c2 (&v)[64], c2* exch, int lane) {
  // (column 0 is stored but never read back)
  if constexpr (kExch128) {
#pragma unroll
    for (int b = 0; b < 8; b += 2)
      *reinterpret_cast<ulonglong2*>(exch + ((8 * A + b) >> 1) * kExchStride + 2 * lane) = make_ulonglong2(v[8 * A + b], v[8 * A + b + 1]);
  } else {
#pragma unroll
    for (int b = 0; b < 8; ++b) exch[exch_idx(8 * A + b, lane)] = v[8 * A + b];
  }
}
template <int R>
__device__ __forceinline__ void tw_chain_p(c2 (&v)[64], const float4* __restrict__ tab4, c2* exch, int lane, const TwP& w) {
  TwP wn = w;
  if constexpr (R + 1 < 14) wn = make_twp(tab4[(R + 1) * 32 + lane]);  // requested before the current one is applied
  tw_row_p<R>(v, w, std::make_integer_sequence<int, 8>{});
  if constexpr (R == 6) tw_store8_p<0>(v, exch, lane);  // row a = 0 is complete once every wb has been applied
  if constexpr (R >= 7) tw_store8_p<R - 6>(v, exch, lane);
  if constexpr (R + 1 < 14) tw_chain_p<R + 1>(v, tab4, exch, lane, wn);
}
__device__ __forceinline__ void tw_store_all_p(c2 (&v)[64], const float4* __restrict__ tab4, c2* exch, int lane) {
  tw_chain_p<0>(v, tab4, exch, lane, make_twp(tab4[lane]));
}

// untangle bins k = colA + 64*K2 and 2048-k = colB + 64*(31-K2) and store them at once (HBM row and/or the warp's padded
// shared buffer), so that the stores are spread over the arithmetic instead of queueing up behind it.
// (The shorter form |X|^2 = S +- C, S = |E|^2+|D|^2, cancels catastrophically on bins far below their mirror bin.)
template <int K2>
__device__ __forceinline__ void post_pair_p(const c2 (&A)[32], const c2 (&B)[32], const TwP& wpb, float* __restrict__ rowA,
                                            float* __restrict__ rowB, float* __restrict__ sqA, float* __restrict__ sqB) {
  const c2 zk = A[K2], zm = B[31 - K2];
  const c2 E = add2(zk, cconj(zm));  // Z[k] + conj(Z[M-k])
  c2 D = csub(zk, cconj(zm));        // Z[k] - conj(Z[M-k])
  if constexpr (K2 != 0) D = cmul_imm<Tw<64, K2>>(D);
  const c2 T = cmul_tw(D, wpb);
#ifdef CA_STFT_SQ_SCALAR
  const float2 P = upk(add2(E, T)), Q = upk(csub(E, T));
  const float sk = fmaf(P.x, P.x, P.y * P.y);
  const float sm = fmaf(Q.x, Q.x, Q.y * Q.y);
#else
  // (P, Q) = E +- T formed component-wise, (P.x, Q.x) and (P.y, Q.y), so that |P|^2 and |Q|^2 come out of two packed
  // instructions: four packed instructions per bin pair instead of two packed and four scalar ones
  const float2 e = upk(E), t = upk(T);
  const c2 pm = pk(1.f, -1.f);
  const c2 X = fma2(bc(t.x), pm, bc(e.x));
  const c2 Y = fma2(bc(t.y), pm, bc(e.y));
  const float2 sq2 = upk(fma2(X, X, mul2(Y, Y)));
  const float sk = sq2.x, sm = sq2.y;
#endif
  if (rowA) {
    rowA[64 * K2] = sk;
    rowB[64 * (31 - K2)] = sm;
  }
  if (sqA) {
    sqA[kSqBlock * K2] = sk;
    sqB[kSqBlock * (31 - K2)] = sm;
  }
}
template <int... K2>
__device__ __forceinline__ void post_all_p(const c2 (&A)[32], const c2 (&B)[32], const TwP& wpb, float* rowA, float* rowB,
                                           float* sqA, float* sqB, std::integer_sequence<int, K2...>) {
  (post_pair_p<K2>(A, B, wpb, rowA, rowB, sqA, sqB), ...);
}

// Even half of frame (stream, fr) -> the staging area of the warp's buffer. The input is described to the copy engine as
// a tensor [stream][row][128 samples]; with hop % 128 == 0 a frame starts on a row, its even half (n1 = 2r: samples
// 128r .. 128r+63) is the box {64 samples, 32 rows} at column 0 and ONE tensor copy moves it. (32 separate 256-byte bulk
// copies cost ~85 clocks of issue each, measured with the timeline build; cp.async / LDGSTS is tracked by the warp's
// scoreboard, so the shared-memory reads of the peak scan would wait for it.)
// No fence.proxy.async: the last generic-proxy accesses to the staging area are the transpose loads, which returned the
// values of the stores before them, and every lane is past them (__syncwarp in the caller).
__device__ __forceinline__ void frame_stage_even_p(c2* __restrict__ stage, const CUtensorMap* map, const StftArgs& p,
                                                   const int stream, const int fr, const int lane, uint64_t* bar) {
  if (lane == 0) {
    mbar_expect_tx(bar, 32 * 256);
    tma_load_3d(stage, map, 0, fr * (p.hop >> 7), stream, bar);
  }
}

// (stream, frame in stream) of a warp's current frame, advanced by the grid stride without a division per frame (the 64-bit
// divide is a ~75-instruction subroutine, and it used to run twice per frame: once for the sample pointer, once for the copy)
struct FramePos {
  int stream, fr;  // position of the warp's frame
  __device__ __forceinline__ FramePos(const int64_t f, const int fps) {
    stream = static_cast<int>(f / fps);
    fr = static_cast<int>(f - static_cast<int64_t>(stream) * fps);
  }
  // grid stride = p.step_streams streams + p.step_frames frames (split on the host)
  __device__ __forceinline__ void advance(const StftArgs& p, const int fps) {
    stream += p.step_streams;
    fr += p.step_frames;
    if (fr >= fps) {
      fr -= fps;
      ++stream;
    }
  }
  __device__ __forceinline__ const float* samples(const StftArgs& p) const {
    return p.x + stream * p.stream_stride + static_cast<int64_t>(fr) * p.hop;
  }
};

// windowed spectrum of the frame held (bit-reversed over n1, already windowed) in v[]: stage-1 FFT, transpose, stage-2
// FFTs, untangling, |X|^2 to HBM and to the warp's shared buffer
__device__ __forceinline__ void frame_spectrum_p(c2 (&v)[64], const StftArgs& p, const float4* __restrict__ tab4,
                                                 const float2* __restrict__ tab2, c2* __restrict__ exch, const int lane,
                                                 const int64_t f, const CUtensorMap* map, const bool stage_next,
                                                 const FramePos& next, uint64_t* bar, const int dbg_frame = 0) {
  float* sq = reinterpret_cast<float*>(exch);
  const int colA = lane ? lane : 32;
  const int colB = lane ? 64 - lane : 32;
  const int k2c = static_cast<int>(__brev(static_cast<unsigned>(lane)) >> 27);  // brev5(lane)
  // ---- stage 1: last stage of the 64-point FFT over n1 (the two 32-point halves are done by the caller) ----
  DitP<64, 64, 0>::stage(v, std::make_integer_sequence<int, 32>{});
  c2 u = v[0];  // column k1 = 0 (no twiddle), one value per lane
  tw_store_all_p(v, tab4, exch, lane);
  CA_STAMP(3);
  // ---- stage 2: columns colA, colB of the transposed data ----
  c2 A[32], B[32];
  __syncwarp();
#pragma unroll
  for (int n2 = 0; n2 < 32; ++n2) {
    A[bitrev(n2, 5)] = exch[exch_idx(colA, n2)];
    B[bitrev(n2, 5)] = exch[exch_idx(colB, n2)];
  }
  __syncwarp();  // exch is reused as the sqmag buffer below
  // the upper half of the buffer is dead from here to the next transpose: stage the even half of the warp's next frame
  // in it now, so that the copy has the rest of this frame (FFTs, untangling, peak extraction) to land
  if (stage_next) frame_stage_even_p(exch + kStageOff, map, p, next.stream, next.fr, lane, bar);
  CA_STAMP(4);
  DitP<32, 32, 0>::run(A);
  DitP<32, 32, 0>::run(B);
  CA_STAMP(5);

  // ---- column 0: 32-point FFT across lanes (DIF, bit-reversed result: lane L holds k2 = brev5(L)), branch-free ----
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int d = 16 >> s;
    const float2 uu = upk(u);
    const c2 o = pk(__shfl_xor_sync(0xffffffffu, uu.x, d), __shfl_xor_sync(0xffffffffu, uu.y, d));
    const float4 t = tab4[(15 + s) * 32 + lane];
    u = cmul_tw(fma2(bc(t.y), u, o), make_twp(t));
  }
  float sq0, sqM;
  {
    const int pk2 = (32 - k2c) & 31;
    const int plane = static_cast<int>(__brev(static_cast<unsigned>(pk2)) >> 27);
    const float2 uu = upk(u);
    const float mx = __shfl_sync(0xffffffffu, uu.x, plane);
    const float my = __shfl_sync(0xffffffffu, uu.y, plane);
    untangle(uu, make_float2(mx, my), tab2[lane], sq0, sqM);
  }

  CA_STAMP(6);
  // ---- untangle (bins k = colA + 64*k2 and 2048 - k), |X|^2 to the HBM row and to the warp's shared buffer ----
  float* row = p.sqmag ? p.sqmag + f * p.sqmag_row_stride : nullptr;
  float* sqw = p.peak_count ? sq : nullptr;
  const TwP wpb = make_twp(tab4[14 * 32 + lane]);
  post_all_p(A, B, wpb, row ? row + colA : nullptr, row ? row + colB : nullptr, sqw ? sqw + colA : nullptr,
             sqw ? sqw + colB : nullptr, std::make_integer_sequence<int, 32>{});
  if (row) {
    row[64 * k2c] = sq0;
    if (lane == 0) row[kM] = sqM;
  }
  if (sqw) {
    sq[kSqBlock * k2c] = sq0;
    if (lane == 0) sq[kSqFloats - 1] = sqM;
  }
}

// peak list of the frame whose |X|^2 sits in the warp's shared buffer: lane owns bins [64*lane, 64*lane+64)
__device__ __forceinline__ void frame_peaks_p(const StftArgs& p, float* __restrict__ sq, const int lane, const int64_t f) {
  __syncwarp();
  // strict local maximum at or above the threshold  <=>  c > max(l, r, largest float below the threshold).
  // All 64 bins of the lane are fetched first (16 x LDS.128): one memory latency instead of one per chunk.
  const int base = kSqBlock * lane;
  float w[66];
  w[0] = lane ? sq[base - 5] : __int_as_float(0x7f800000);  // bin 0 is never a peak
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const float4 t = *reinterpret_cast<const float4*>(sq + base + 4 * q);
    w[1 + 4 * q] = t.x;
    w[2 + 4 * q] = t.y;
    w[3 + 4 * q] = t.z;
    w[4 + 4 * q] = t.w;
  }
  w[65] = sq[base + (lane == 31 ? 67 : 68)];  // first bin of the next block (bin 2048 sits in the very last pad)
  uint32_t mlo = 0, mhi = 0;
#ifndef CA_STFT_PEAK_COMPARE
  // the sign bit of max(l, r, thr) - c shifted into the mask: FMNMX3 + FADD + SHF per bin (FSETP + SEL + half a LOP3 for the
  // compare-and-select form, CA_STFT_PEAK_COMPARE: 2 % slower); c == max gives +0, so the test stays strict. kChains independent
  // shift chains per mask word keep the dependent SHFs apart.
#ifndef CA_STFT_PEAK_CHAINS
#define CA_STFT_PEAK_CHAINS 1
#endif
  constexpr int kChains = CA_STFT_PEAK_CHAINS, kLen = 32 / kChains;
  uint32_t part[2][kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) part[0][c] = part[1][c] = 0;
#pragma unroll
  for (int j = kLen - 1; j >= 0; --j)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int c = 0; c < kChains; ++c) {
        const int b = 32 * h + kLen * c + j;
        part[h][c] = __funnelshift_l(__float_as_uint(fmaxf(fmaxf(w[b], w[b + 2]), p.thr_below) - w[b + 1]), part[h][c], 1);
      }
#pragma unroll
  for (int c = 0; c < kChains; ++c) {
    mlo |= part[0][c] << (kLen * c);
    mhi |= part[1][c] << (kLen * c);
  }
#else
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    const bool is_peak = w[j + 1] > fmaxf(fmaxf(w[j], w[j + 2]), p.thr_below);
    if (j < 32) mlo |= is_peak ? (1u << j) : 0u;
    else mhi |= is_peak ? (1u << (j - 32)) : 0u;
  }
#endif
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
  if (lane == 0) p.peak_count[f] = total;
  // each lane lists the bins of its peaks (ascending), then the peaks are spread over the lanes: the interpolation
  // runs once per 32 peaks instead of once per peak of the busiest lane, and the stores are coalesced
  // (the list lives in the padding floats of the spectrum buffer; longer lists go through the output array itself)
  const int64_t out_base = f * static_cast<int64_t>(p.max_peaks);
  int* sqi = reinterpret_cast<int*>(sq);
  const bool list_in_smem = p.max_peaks <= kKlistCap;
  const int n_store = min(total, p.max_peaks);
  while (mask && slot < n_store) {
    const int j = __ffsll(static_cast<long long>(mask)) - 1;
    mask &= mask - 1;
    if (list_in_smem) sqi[klist_slot(slot)] = 64 * lane + j;
    else p.peak_bin[out_base + slot] = 64 * lane + j;
    ++slot;
  }
  __syncwarp();
  for (int idx = lane; idx < n_store; idx += 32) {
    const int k = list_in_smem ? sqi[klist_slot(idx)] : *reinterpret_cast<volatile int*>(p.peak_bin + out_base + idx);
    const float fl = 1e-37f;
    const float sa = fmaxf(sq[physp(k - 1)], fl), sb = fmaxf(sq[physp(k)], fl), sc = fmaxf(sq[physp(k + 1)], fl);
    // dB differences in log2 units, from ratios to the centre bin (well conditioned); MUFU.LG2 / MUFU.RCP precision
    const float inv_b = __frcp_rn(sb);
    const float la = __log2f(sa * inv_b), lc = __log2f(sc * inv_b);  // (alpha-beta), (gamma-beta) / (10 log10 2)
    const float num = la - lc;
    const float den = la + lc;
    const float pp = den != 0.f ? 0.5f * __fdividef(num, den) : 0.f;
    constexpr float kDbPerLog2 = 3.01029995663981195f;  // 10*log10(2)
    const float beta = kDbPerLog2 * __log2f(sb);
    if (list_in_smem) p.peak_bin[out_base + idx] = k;
    p.peak_freq[out_base + idx] = p.freq_as_offset ? pp : (static_cast<float>(k) + pp) * p.bin_to_hz;
    p.peak_db[out_base + idx] = beta - 0.25f * (kDbPerLog2 * num) * pp;
  }
  __syncwarp();  // the next frame overwrites the buffer
}

// raw samples of frame f into v[] (bit-reversed over n1): lane n2 takes z[32*n1 + n2] = (x[64*n1+2*n2], x[64*n1+2*n2+1]).
// PAR selects the half with n1 = PAR (mod 2), i.e. registers v[32*PAR .. 32*PAR+31]: the first five DIT stages stay inside
// a half, so the even half can be prefetched early (64 registers) and the odd half fetched while the even half is transformed.
// kFull: window == 4096 (no zero padding, no load predicates); kVec2: 8-byte aligned 64-bit loads are legal.
template <int PAR, bool kVec2, bool kFull>
__device__ __forceinline__ void frame_load_p(c2 (&v)[64], const StftArgs& p, const float* __restrict__ xs, const int lane) {
#pragma unroll
  for (int n1 = PAR; n1 < 64; n1 += 2) {
    const int idx = 64 * n1 + 2 * lane;
    c2 s;
    if (kVec2 && kFull) {
      s = ldg_c2(xs + idx);
    } else if (kVec2) {
      s = (idx < p.window) ? ldg_c2(xs + idx) : pk(0.f, 0.f);
    } else {
      s = pk((kFull || idx < p.window) ? __ldg(xs + idx) : 0.f, (kFull || idx + 1 < p.window) ? __ldg(xs + idx + 1) : 0.f);
    }
    v[bitrev(n1, 6)] = s;  // DIT: bit-reversed register order in
  }
}
// staged even half -> registers, windowed
__device__ __forceinline__ void frame_take_even_p(c2 (&v)[64], const c2* __restrict__ stage, const c2* __restrict__ win_c,
                                                  const int lane) {
#pragma unroll
  for (int r = 0; r < 32; r += 2) {
    const ulonglong2 w = *reinterpret_cast<const ulonglong2*>(win_c + lane * kWinStride + r);
    v[bitrev(2 * r, 6)] = mul2(stage[32 * r + lane], w.x);
    v[bitrev(2 * r + 2, 6)] = mul2(stage[32 * r + 32 + lane], w.y);
  }
}
template <int PAR>
__device__ __forceinline__ void frame_window_p(c2 (&v)[64], const c2* __restrict__ win_c, const int lane) {
#pragma unroll
  for (int r = 0; r < 32; r += 2) {
    const ulonglong2 w = *reinterpret_cast<const ulonglong2*>(win_c + lane * kWinStride + 32 * PAR + r);
    v[bitrev(2 * r + PAR, 6)] = mul2(v[bitrev(2 * r + PAR, 6)], w.x);
    v[bitrev(2 * r + 2 + PAR, 6)] = mul2(v[bitrev(2 * r + 2 + PAR, 6)], w.y);
  }
}

// kBulk: the even half of the warp's next frame is staged by bulk copies during the peak extraction (needs window ==
// 4096, hop % 4 == 0 and 16-byte aligned streams); otherwise both halves are loaded at the top of the frame.
template <int WARPS, int MINB, bool kVec2, bool kFull, bool kBulk>
__global__ void __launch_bounds__(WARPS * 32, MINB) stft4096p_kernel(const __grid_constant__ StftArgs p,
                                                                     const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* tab4 = reinterpret_cast<float4*>(smem_raw);               // [kTab4Rows][32]
  float2* tab2 = reinterpret_cast<float2*>(tab4 + kTab4Rows * 32);  // [32]
  float2* win_s = tab2 + 32;                                        // [32][kWinStride] window (pairs), pre-scaled
  c2* exch_all = reinterpret_cast<c2*>(win_s + 32 * kWinStride);    // [WARPS][kExchFloat2]
  __shared__ __align__(8) uint64_t bars[WARPS];

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  build_tables_p<WARPS>(tab4, tab2, win_s, p.win);
  if (kBulk && lane == 0) {
    mbar_init(&bars[warp], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  c2* exch = exch_all + warp * kExchFloat2;
  const c2* win_c = reinterpret_cast<const c2*>(win_s);
  uint64_t* bar = &bars[warp];

  const int64_t fstep = static_cast<int64_t>(gridDim.x) * WARPS;
  int64_t f = static_cast<int64_t>(blockIdx.x) * WARPS + warp;
  if (f >= p.total_frames) return;
  c2* stage = exch + kStageOff;
  c2 v[64];
  const int fps = static_cast<int>(p.frames_per_stream);
  FramePos pos(f, fps);
  uint32_t parity = 0;
  int dbg_frame = 0;
  if (kBulk) frame_stage_even_p(stage, &tmap, p, pos.stream, pos.fr, lane, bar);
  for (;;) {
    CA_STAMP(0);
    const float* xs = pos.samples(p);
    frame_load_p<1, kVec2, kFull>(v, p, xs, lane);  // odd half: lands while the even half is transformed
    if (kBulk) {
      mbar_wait(bar, parity);
      parity ^= 1u;
      CA_STAMP(1);
      frame_take_even_p(v, stage, win_c, lane);
    } else {
      frame_load_p<0, kVec2, kFull>(v, p, xs, lane);
      frame_window_p<0>(v, win_c, lane);
    }
    DitP<64, 32, 0>::run(v);  // (ptxas contracts the window multiply into the first butterfly by itself: a hand-fused form
    frame_window_p<1>(v, win_c, lane);  // compiles to the same instruction mix)
    DitP<64, 32, 32>::run(v);
    __syncwarp();  // every lane has taken its staged samples: the buffer may be overwritten by the transpose
    CA_STAMP(2);
    const bool more = f + fstep < p.total_frames;
    pos.advance(p, fps);  // from here on: the position of the warp's next frame
    frame_spectrum_p(v, p, tab4, tab2, exch, lane, f, &tmap, kBulk && more, pos, bar, dbg_frame);
    CA_STAMP(7);
    CA_STAMP(8);
    if (p.peak_count) frame_peaks_p(p, reinterpret_cast<float*>(exch), lane, f);
    CA_STAMP(9);
    ++dbg_frame;
    if (!more) break;
    f += fstep;
  }
}

}  // namespace

int stft_plan(ca_ctx* ctx, StftPlan& plan, const ca_stft_params* p) {
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
  const double norm = static_cast<double>(H / sum) * 0.5 / W;
  std::vector<float> full(n_fft, 0.f);
  for (int n = 0; n < W; ++n) full[n] = static_cast<float>((n < H ? hw[H - 1 - n] : hw[n - H]) * norm);
  plan.planned = false;
  if (!plan.d_window || plan.n_fft != n_fft) {
    if (plan.d_window) cudaFree(plan.d_window);
    plan.d_window = nullptr;
    CA_CUDA_OK(ctx, cudaMalloc(&plan.d_window, sizeof(float) * n_fft));
  }
  CA_CUDA_OK(ctx, cudaMemcpy(plan.d_window, full.data(), sizeof(float) * n_fft, cudaMemcpyHostToDevice));
  plan.n_fft = n_fft;
  plan.prm = *p;
  plan.thr_sq = isinf(p->min_db) && p->min_db < 0 ? 0.f : static_cast<float>(pow(10.0, p->min_db / 10.0));
  plan.planned = true;
  return CA_OK;
}

void stft_release(ca_ctx* ctx) {
  for (StftPlan* pl : {&ctx->plan_user, &ctx->plan_rt, &ctx->plan_voc}) {
    if (pl->d_window) cudaFree(pl->d_window);
    if (pl->d_window64) cudaFree(pl->d_window64);
    if (pl->d_tw64) cudaFree(pl->d_tw64);
    *pl = StftPlan{};
  }
}

int stft_analyze(ca_ctx* ctx, const StftPlan& plan, const float* d_x, int64_t n_samples, int64_t n_streams, int64_t stream_stride,
                 float* d_sqmag, int64_t sqmag_row_stride, int32_t* d_peak_bin, float* d_peak_freq, float* d_peak_db,
                 int32_t* d_peak_count, cudaStream_t stream) {
  if (!plan.planned) return fail(ctx, CA_ERR_NOT_PLANNED, "ca_stft_analyze: call ca_stft_plan first");
  if (!d_x || n_streams < 0 || n_samples < 0 || stream_stride < n_samples)
    return fail(ctx, CA_ERR_INVALID, "ca_stft_analyze: bad input shape");
  const bool any_peak = d_peak_bin || d_peak_freq || d_peak_db || d_peak_count;
  if (any_peak && !(d_peak_bin && d_peak_freq && d_peak_db && d_peak_count))
    return fail(ctx, CA_ERR_INVALID, "ca_stft_analyze: peak outputs must be all null or all non-null");
  if (d_sqmag && sqmag_row_stride < plan.n_fft / 2 + 1)
    return fail(ctx, CA_ERR_INVALID, "ca_stft_analyze: sqmag_row_stride < n_fft/2+1");
  const int64_t frames = ca_stft_num_frames(n_samples, plan.prm.window, plan.prm.hop);
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
  a.hop = plan.prm.hop;
  a.window = plan.prm.window;
  a.win = plan.d_window;
  a.sqmag = d_sqmag;
  a.sqmag_row_stride = sqmag_row_stride;
  a.peak_bin = d_peak_bin;
  a.peak_freq = d_peak_freq;
  a.peak_db = d_peak_db;
  a.peak_count = d_peak_count;
  a.max_peaks = plan.prm.max_peaks;
  a.thr_below = plan.thr_sq > 0.f ? nextafterf(plan.thr_sq, 0.f) : -1.f;
  a.bin_to_hz = static_cast<float>(static_cast<double>(plan.prm.sample_rate) / plan.n_fft);
  if (plan.band_amps) {
    if (plan.n_fft < 8192) return fail(ctx, CA_ERR_UNSUPPORTED, "stft_analyze: fused band amplitudes need an FFT length >= 8192");
    a.band_begin = plan.band_begin;
    a.n_bands = plan.n_bands;
    a.band_amps = plan.band_amps;
  }
  if (plan.n_fft != kN) {
    a.max_peaks = plan.prm.max_peaks;
    a.freq_as_offset = plan.freq_as_offset ? 1 : 0;
    {  // fused single-pass kernels for 1024, 2048, 8192, 16384, 32768 (rfftw.cu)
      const bool v2 = (plan.prm.hop % 2 == 0) && (stream_stride % 2 == 0) && (reinterpret_cast<uintptr_t>(d_x) % 8 == 0);
      const int rcf = stft_fused_generic(ctx, a, plan.n_fft, v2, stream);
      if (rcf != CA_ERR_UNSUPPORTED) return rcf;
    }
    return fail(ctx, CA_ERR_UNSUPPORTED, "ca_stft_analyze: no kernel for this fft length");
  }
  a.freq_as_offset = plan.freq_as_offset ? 1 : 0;

  const bool vec2 = (plan.prm.hop % 2 == 0) && (stream_stride % 2 == 0) && (reinterpret_cast<uintptr_t>(d_x) % 8 == 0) &&
                    (plan.prm.window % 2 == 0);
  const bool full = plan.prm.window == kN;
  // 12 warps x 1 CTA per SM: 12 x 16.5 KB exchange buffers + 16 KB window + tables = 224 KB of shared memory, 168 registers
  static const int warps_env = [] {
    const char* e = getenv("CA_STFT_WARPS");  // 4 = one warp per scheduler: the uncontended timeline (profiles/build_timeline.sh)
    return e ? atoi(e) : 12;
  }();
  // staged path: window 4096, hop a multiple of 128 samples, 16-byte aligned streams (see frame_stage_even_p)
  bool bulk = full && plan.prm.hop % 128 == 0 && stream_stride % 4 == 0 && reinterpret_cast<uintptr_t>(d_x) % 16 == 0 &&
              n_samples / 128 < (1ll << 31);
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (bulk) {
    using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = [] {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
          qres != cudaDriverEntryPointSuccess)
        fn = nullptr;
      return reinterpret_cast<EncodeFn>(fn);
    }();
    const cuuint64_t gdim[3] = {128, static_cast<cuuint64_t>(n_samples / 128), static_cast<cuuint64_t>(n_streams)};
    const cuuint64_t gstride[2] = {512, static_cast<cuuint64_t>(stream_stride) * sizeof(float)};
    const cuuint32_t box[3] = {64, 32, 1};
    const cuuint32_t estride[3] = {1, 1, 1};
    if (!encode || encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(d_x), gdim, gstride, box, estride,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      bulk = false;  // no tensor map: every lane loads its samples itself
  }
  int PW = 12;
  void (*kern)(StftArgs, CUtensorMap) =
      bulk   ? stft4096p_kernel<12, 1, true, true, true>
      : vec2 ? (full ? stft4096p_kernel<12, 1, true, true, false> : stft4096p_kernel<12, 1, true, false, false>)
             : (full ? stft4096p_kernel<12, 1, false, true, false> : stft4096p_kernel<12, 1, false, false, false>);
  if (bulk && warps_env == 4) {
    PW = 4;
    kern = stft4096p_kernel<4, 1, true, true, true>;
  }
  const size_t smem = sizeof(float4) * kTab4Rows * 32 + sizeof(float2) * (32 + 32 * kWinStride + PW * kExchFloat2);
  CA_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int64_t grid = ctx->sm_count;
  const int64_t need = (total + PW - 1) / PW;
  if (grid > need) grid = need;
#ifdef CA_STFT_TIMELINE
  static long long* d_dbg = nullptr;
  const size_t dbg_n = 2 * 12 * 24 * 12;
  if (!d_dbg) cudaMalloc(&d_dbg, dbg_n * sizeof(long long));
  cudaMemset(d_dbg, 0, dbg_n * sizeof(long long));
  a.dbg = d_dbg;
#endif
  a.step_streams = static_cast<int32_t>((grid * PW) / frames);
  a.step_frames = static_cast<int32_t>((grid * PW) % frames);
  kern<<<static_cast<unsigned>(grid), PW * 32, smem, stream>>>(a, tmap);
  CA_CUDA_OK(ctx, cudaGetLastError());
#ifdef CA_STFT_TIMELINE
  {
    std::vector<long long> h(dbg_n);
    cudaMemcpy(h.data(), d_dbg, dbg_n * sizeof(long long), cudaMemcpyDeviceToHost);
    const char* names[9] = {"wait_stage", "fft32x2", "last+tw+sts", "lds", "fft32x2", "col0", "untangle+st", "stage_issue", "peaks"};
    for (int w = 0; w < 2 * PW; w += 1) {
      double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      int n = 0;
      double period = 0;
      for (int fr = 4; fr < 20; ++fr) {
        const long long* t = &h[(static_cast<size_t>(w) * 24 + fr) * 12];
        if (!t[9]) continue;
        for (int i = 0; i < 9; ++i) acc[i] += static_cast<double>(t[i + 1] - t[i]);
        period += static_cast<double>(h[(static_cast<size_t>(w) * 24 + fr + 1) * 12] - t[0]);
        ++n;
      }
      if (!n) continue;
      fprintf(stderr, "timeline cta %d warp %2d:", w / PW, w % PW);
      for (int i = 0; i < 9; ++i) fprintf(stderr, " %s %.0f", names[i], acc[i] / n);
      fprintf(stderr, " | frame period %.0f\n", period / n);
    }
  }
#endif
  ctx->launches += 1;
  return CA_OK;
}

}  // namespace ca
