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
                      max_track_pitches: float = 1.0, max_channels: int = 254, host_threads: int = 0):
    """x float32 [streams, samples] (host). Returns (audio float32 [streams, samples, 2], stats int64 [streams, 2])."""
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float32)
    streams, n = x.shape
    out = np.empty((streams, n, 2), dtype=np.float32)
    stats = np.zeros((streams, 2), dtype=np.int64)
    e = env if env is not None else Ahdsr(0, 0, 0, 0, 0, 0, 0, 1.0)
    p = RoundtripParams(StftParams(window, hop, 1, 4.0, sample_rate, min_db if math.isfinite(min_db) else -math.inf, max_peaks),
                        nearby_distance_tones, min_volume, max_track_pitches, gain, stereo_spread, e, max_channels, host_threads)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    ctx.check(lib().ca_resynth_roundtrip_host(ctx.handle, C.byref(p), vp(x), n, streams, vp(out), vp(stats)),
              "ca_resynth_roundtrip_host")
    return out, stats
