"""FFT vocoder (source/rt.resynth.lib.vocoder.cpp) through ca_vocoder_host: band amplitudes of the modulator shape the
carrier's spectrum frame by frame; host audio in, host audio out."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._capi import Context, VocoderParams, lib


def vocoder(ctx: Context, modulator: np.ndarray, carrier: np.ndarray, sample_rate: int = 48000,
            modulator_window_seconds: float = 0.10, stride_seconds: float = 0.005, count_bands: int = 5, min_freq: float = 100.0,
            max_freq: float = 20000.0, vol_modulator: float = 0.0, vol_carrier: float = 0.0, vol_vocoder: float = 1.0,
            out: np.ndarray | None = None) -> np.ndarray:
    """modulator, carrier float32 [streams, samples] (host). Returns float32 [streams, samples]. Defaults are the reference's
    (source/rt.resynth.lib.cpp:986-990)."""
    m = np.ascontiguousarray(np.atleast_2d(modulator), dtype=np.float32)
    c = np.ascontiguousarray(np.atleast_2d(carrier), dtype=np.float32)
    if m.shape != c.shape:
        raise ValueError("modulator and carrier must have the same shape")
    streams, n = m.shape
    if out is None:
        out = np.empty((streams, n), dtype=np.float32)
    assert out.shape == (streams, n) and out.dtype == np.float32 and out.flags.c_contiguous
    p = VocoderParams(sample_rate, modulator_window_seconds, stride_seconds, count_bands, min_freq, max_freq, vol_modulator,
                      vol_carrier, vol_vocoder)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    ctx.check(lib().ca_vocoder_host(ctx.handle, C.byref(p), vp(m), vp(c), n, streams, vp(out)), "ca_vocoder_host")
    return out
