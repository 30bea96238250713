"""Analysis -> pitch pipeline -> resynthesis round trip (BASELINE config 4) through ca_resynth_roundtrip_host.

Plays the role of the offline RtResynth job (source/rt.resynth.lib.cpp:1185-1235): host audio in, host stereo audio out."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from ._capi import Ahdsr, Context, RoundtripParams, StftParams, lib


def resynth_roundtrip(ctx: Context, x: np.ndarray, window: int = 4096, hop: int = 256, sample_rate: int = 48000,
                      min_db: float = -80.0, max_peaks: int = 200, gain: float = 1.0, stereo_spread: float = 1.0,
                      env: Ahdsr | None = None, nearby_distance_tones: float = 0.4, min_volume: float = 1e-4,
                      max_track_pitches: float = 1.0, max_channels: int = 254, host_threads: int = 0, out: np.ndarray | None = None):
    """x float32 [streams, samples] (host). Returns (audio float32 [streams, samples, 2], stats int64 [streams, 2]).

    Any host memory works; page-locked buffers (`pinned_empty`) let the copies run at PCIe speed: pageable memory is
    staged by the driver at a fraction of it."""
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float32)
    streams, n = x.shape
    if out is None:
        out = np.empty((streams, n, 2), dtype=np.float32)
    assert out.shape == (streams, n, 2) and out.dtype == np.float32 and out.flags.c_contiguous
    stats = np.zeros((streams, 2), dtype=np.int64)
    e = env if env is not None else Ahdsr(0, 0, 0, 0, 0, 0, 0, 1.0)
    p = RoundtripParams(StftParams(window, hop, 1, 4.0, sample_rate, min_db if math.isfinite(min_db) else -math.inf, max_peaks),
                        nearby_distance_tones, min_volume, max_track_pitches, gain, stereo_spread, e, max_channels, host_threads)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    ctx.check(lib().ca_resynth_roundtrip_host(ctx.handle, C.byref(p), vp(x), n, streams, vp(out), vp(stats)),
              "ca_resynth_roundtrip_host")
    return out, stats


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """Page-locked host array (torch's caching pinned allocator); keep a reference to it while it is in use."""
    import torch

    t = torch.empty(tuple(shape), dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
    a = t.numpy()
    _PINNED_KEEPALIVE[a.ctypes.data] = t
    return a


_PINNED_KEEPALIVE: dict = {}
