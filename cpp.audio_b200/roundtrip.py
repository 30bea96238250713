"""Analysis -> pitch pipeline -> resynthesis round trip (BASELINE config 4) through ca_resynth_roundtrip_host.

Plays the role of the offline RtResynth job (source/rt.resynth.lib.cpp:1185-1235): host audio in, host stereo audio out."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from ._capi import Ahdsr, Context, PitchOptions, RoundtripParams, StftParams, lib


def resynth_roundtrip(ctx: Context, x: np.ndarray, window: int = 4096, hop: int = 256, sample_rate: int = 48000,
                      min_db: float = -80.0, max_peaks: int = 200, gain: float = 1.0, stereo_spread: float = 1.0,
                      env: Ahdsr | None = None, nearby_distance_tones: float = 0.4, min_volume: float = 1e-4,
                      max_track_pitches: float = 1.0, max_channels: int = 254, host_threads: int = 0, out: np.ndarray | None = None,
                      opts: PitchOptions | None = None):
    """x float32 [streams, samples] (host). Returns (audio float32 [streams, samples, 2], stats int64 [streams, 2]).
    opts: shift / harmonize / autotune stages (`pitch_options`); None = off, as rt.resynth out of the box.

    Any host memory works; page-locked buffers (`pinned_empty`) let the copies run at PCIe speed: pageable memory is
    staged by the driver at a fraction of it."""
    pcm16 = np.asarray(x).dtype == np.int16   # 16-bit PCM: ca_resynth_roundtrip_pcm16_host (x / 32768 on the device)
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.int16 if pcm16 else np.float32)
    streams, n = x.shape
    if out is None:
        out = np.empty((streams, n, 2), dtype=np.float32)
    assert out.shape == (streams, n, 2) and out.dtype == np.float32 and out.flags.c_contiguous
    stats = np.zeros((streams, 2), dtype=np.int64)
    e = env if env is not None else Ahdsr(0, 0, 0, 0, 0, 0, 0, 1.0)
    p = RoundtripParams(StftParams(window, hop, 1, 4.0, sample_rate, min_db if math.isfinite(min_db) else -math.inf, max_peaks),
                        nearby_distance_tones, min_volume, max_track_pitches, gain, stereo_spread, e, max_channels, host_threads,
                        opts if opts is not None else PitchOptions())
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    fn = lib().ca_resynth_roundtrip_pcm16_host if pcm16 else lib().ca_resynth_roundtrip_host
    ctx.check(fn(ctx.handle, C.byref(p), vp(x), n, streams, vp(out), vp(stats)), "ca_resynth_roundtrip_host")
    return out, stats


def truncated_frames(ctx: Context) -> int:
    """Frames of the last round trip whose peak list exceeded max_peaks (ca_resynth_truncated_frames); 0 = nothing was cut."""
    return int(lib().ca_resynth_truncated_frames(ctx.handle))


def pitch_options(shift_pre=0.0, harmonize_pre=0.0, shift_post=0.0, harmonize_post=0.0, allowed_pitches=None, scale=None,
                  scale_root_pitch=0.0, max_pitch=151.0, tolerance=0.0) -> PitchOptions:
    """ca_pitch_options. allowed_pitches (ascending float64): the vector mkAutotuneFunction builds for AutotuneType::Chord /
    ::FixedSizeIntervals (rt.resynth.lib.cpp:1767-1858); scale: pitches 0 ... 12 of a MusicalScalePitches
    (rt.resynth.lib.autotune.cpp:142-175). The returned struct keeps the table alive."""
    assert allowed_pitches is None or scale is None
    o = PitchOptions(shift_pre, harmonize_pre, shift_post, harmonize_post, 0, max_pitch, tolerance, 0, None, scale_root_pitch)
    tab = allowed_pitches if allowed_pitches is not None else scale
    if tab is not None:
        tab = np.ascontiguousarray(tab, dtype=np.float64)
        o.autotune = 1 if allowed_pitches is not None else 2
        o.n_table = int(tab.shape[0])
        o.table = tab.ctypes.data
        o._keep = tab
    return o


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """Page-locked host array (torch's caching pinned allocator); keep a reference to it while it is in use."""
    import torch

    t = torch.empty(tuple(shape), dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
    a = t.numpy()
    _PINNED_KEEPALIVE[a.ctypes.data] = t
    return a


_PINNED_KEEPALIVE: dict = {}
