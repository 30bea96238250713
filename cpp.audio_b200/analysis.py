"""Host side of the analysis path: torch tensors in, torch tensors out, all compute in libcpp_audio_b200.so.

Mirrors the reference's operator vocabulary: `SqMagAnalyzer` plays the role of
PeriodicFFT<SqMagFftOperation<Window,T>> + extractLocalMaxFreqsMags (source/rt.resynth.lib.periodicfft.cpp:13-181,251-325;
source/rt.resynth.lib.cpp:1592-1595) for whole streams; `rfft_batched` of FftOperation<T> (periodicfft.cpp:183-244).
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from ._capi import Context, StftParams, lib

WINDOW_RECTANGULAR = 0
WINDOW_GAUSSIAN = 1


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def rfft_batched(ctx: Context, x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """Forward real FFT of every row of x [batch, n_fft] (float32, CUDA). Returns complex64 [batch, n_fft/2+1], unscaled.
    `out`: a complex64 tensor of that shape to write into (saves the allocation for small, latency-bound batches)."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous()):
        raise ValueError("x must be a contiguous float32 CUDA tensor [batch, n_fft]")
    batch, n = x.shape
    if out is None:
        out = torch.view_as_complex(torch.empty((batch, n // 2 + 1, 2), dtype=torch.float32, device=x.device))
    elif not (out.is_cuda and out.dtype == torch.complex64 and tuple(out.shape) == (batch, n // 2 + 1) and out.is_contiguous()):
        raise ValueError("out must be a contiguous complex64 CUDA tensor [batch, n_fft/2+1]")
    ctx.check(lib().ca_rfft_batched(ctx.handle, x.data_ptr(), n, batch, out.data_ptr(), _stream_ptr(x.device)),
              "ca_rfft_batched")
    return out


def irfft_batched(ctx: Context, X: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """Inverse real FFT of every row of X [batch, n_fft/2+1] (complex64, CUDA). Returns float32 [batch, n_fft] =
    scale * (unnormalised inverse): irfft_batched(rfft_batched(x)) = n_fft * x at scale 1."""
    if not (X.is_cuda and X.dtype == torch.complex64 and X.dim() == 2 and X.is_contiguous()):
        raise ValueError("X must be a contiguous complex64 CUDA tensor [batch, n_fft/2+1]")
    batch, bins = X.shape
    n = 2 * (bins - 1)
    out = torch.empty((batch, n), dtype=torch.float32, device=X.device)
    ctx.check(lib().ca_irfft_batched(ctx.handle, torch.view_as_real(X).data_ptr(), n, batch, out.data_ptr(), scale, _stream_ptr(X.device)),
              "ca_irfft_batched")
    return out


class SqMagAnalyzer:
    """Windowed squared-magnitude spectra + spectral peaks of overlapping frames of many streams.

    init_data(window) has the meaning of SqMagFftOperation::init_data (periodicfft.cpp:273-299): it (re)builds the
    normalised half window; here it also uploads it (ca_stft_plan), the only allocating call.
    """

    def __init__(self, ctx: Context, window: int, hop: int, window_type: int = WINDOW_GAUSSIAN, sigma_factor: float = 4.0,
                 sample_rate: int = 48000, min_db: float = -80.0, max_peaks: int = 200):
        self.ctx = ctx
        self.hop = hop
        self.window_type = window_type
        self.sigma_factor = sigma_factor
        self.sample_rate = sample_rate
        self.min_db = min_db
        self.max_peaks = max_peaks
        self.window = 0
        self.n_fft = 0
        self.init_data(window)

    def init_data(self, window: int) -> bool:
        if window == self.window:
            return False
        self._plan(window)
        return True

    def _plan(self, window: int) -> None:
        p = StftParams(window, self.hop, self.window_type, self.sigma_factor, self.sample_rate,
                       self.min_db if math.isfinite(self.min_db) else -math.inf, self.max_peaks)
        self.ctx.check(lib().ca_stft_plan(self.ctx.handle, C.byref(p)), "ca_stft_plan")
        self._generation = int(lib().ca_stft_plan_generation(self.ctx.handle))
        self.window = window
        self.n_fft = int(lib().ca_fft_length_for(window, 1))

    @property
    def bins(self) -> int:
        return self.n_fft // 2 + 1

    def num_frames(self, n_samples: int) -> int:
        return int(lib().ca_stft_num_frames(n_samples, self.window, self.hop))

    def allocate(self, n_streams: int, n_samples: int, device, want_sqmag=True, want_peaks=True):
        rows = n_streams * self.num_frames(n_samples)
        out = {}
        if want_sqmag:
            out["sqmag"] = torch.empty((rows, self.bins), dtype=torch.float32, device=device)
        if want_peaks:
            out["peak_bin"] = torch.empty((rows, self.max_peaks), dtype=torch.int32, device=device)
            out["peak_freq"] = torch.empty((rows, self.max_peaks), dtype=torch.float32, device=device)
            out["peak_db"] = torch.empty((rows, self.max_peaks), dtype=torch.float32, device=device)
            out["peak_count"] = torch.empty((rows,), dtype=torch.int32, device=device)
        return out

    def __call__(self, x: torch.Tensor, out: dict | None = None, want_sqmag=True, want_peaks=True) -> dict:
        """x [streams, samples] float32 CUDA. Outputs are flattened over (stream, frame). Asynchronous on the current stream."""
        if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1):
            raise ValueError("x must be a float32 CUDA tensor [streams, samples] with unit inner stride")
        n_streams, n_samples = x.shape
        if int(lib().ca_stft_plan_generation(self.ctx.handle)) != self._generation:
            self._plan(self.window)  # another analyzer planned on this context since: the context holds one plan
        if out is None:
            out = self.allocate(n_streams, n_samples, x.device, want_sqmag, want_peaks)
        sq = out.get("sqmag")
        pk = out.get("peak_count")
        stream_stride = x.stride(0) if n_streams > 1 else n_samples  # a size-1 dimension may carry any stride
        self.ctx.check(lib().ca_stft_analyze(
            self.ctx.handle, x.data_ptr(), n_samples, n_streams, stream_stride,
            sq.data_ptr() if sq is not None else None, sq.stride(0) if sq is not None else 0,
            out["peak_bin"].data_ptr() if pk is not None else None,
            out["peak_freq"].data_ptr() if pk is not None else None,
            out["peak_db"].data_ptr() if pk is not None else None,
            pk.data_ptr() if pk is not None else None, _stream_ptr(x.device)), "ca_stft_analyze")
        return out


def analyze_streams(ctx: Context, x: torch.Tensor, window: int, hop: int, **kw) -> dict:
    want_sqmag = kw.pop("want_sqmag", True)
    want_peaks = kw.pop("want_peaks", True)
    an = SqMagAnalyzer(ctx, window, hop, **kw)
    return an(x, want_sqmag=want_sqmag, want_peaks=want_peaks)
