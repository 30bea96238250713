"""ctypes front-end of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs. The product package (cpp.audio_b200/) never imports this module.
Parity status: analysis arithmetic "parity unpinned" (oracle/SPEC.md), synthesis pinned by in-tree reference code.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    src_newer = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("oracle.hpp", "oracle_capi.cpp", "Makefile")
    )
    if force or src_newer:
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class Ahdsr(C.Structure):
    """Mirror of audioelement::AHDSR (include/audioelement.h:314-335); itp values per oracle/SPEC.md S7."""

    _fields_ = [
        ("attack", C.c_int32), ("attack_itp", C.c_int32),
        ("hold", C.c_int32),
        ("decay", C.c_int32), ("decay_itp", C.c_int32),
        ("release", C.c_int32), ("release_itp", C.c_int32),
        ("sustain", C.c_float),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        p = C.c_void_p
        L.orc_hardware_threads.restype = C.c_int
        L.orc_rfft_f64.argtypes = [p, C.c_int, C.c_int64, p, C.c_int]
        L.orc_rfft_f32.argtypes = [p, C.c_int, C.c_int64, p, C.c_int]
        L.orc_num_frames.argtypes = [C.c_int64, C.c_int, C.c_int]
        L.orc_num_frames.restype = C.c_int64
        L.orc_fft_length_for.argtypes = [C.c_int, C.c_int]
        L.orc_fft_length_for.restype = C.c_int
        L.orc_window_f64.argtypes = [C.c_int, C.c_double, C.c_int, p]
        L.orc_stft_analyze.argtypes = [p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double,
                                       C.c_int, C.c_double, C.c_int, p, p, p, p, p, C.c_int]
        L.orc_stft_analyze.restype = C.c_int64
        L.orc_oscbank_render.argtypes = [C.c_int64, C.c_int, C.c_int64, C.c_int, p, p, p, p, p, p, p, p, C.c_int64,
                                         p, C.c_int, p, C.c_int]
        L.orc_ramp_render.argtypes = [C.c_int64, C.c_int64, C.c_int, p, p, p, p, p, p, p, p, p, C.c_int, p, C.c_int]
        L.orc_envelope_trace.argtypes = [p, C.c_int, C.c_double, C.c_int32, C.c_int64, C.c_int64, p, p]
        L.orc_freq_aliasing_multiplicator.argtypes = [C.c_double]
        L.orc_freq_aliasing_multiplicator.restype = C.c_double
        L.orc_sine_trace.argtypes = [C.c_double, C.c_double, C.c_int64, p]
        L.orc_roundtrip.argtypes = [p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double, C.c_float, C.c_float, p, p, p,
                                    C.c_int]
        L.orc_pitch_stages.argtypes = [p, p, p, C.c_int, C.c_int]
        L.orc_pitch_stages.restype = C.c_int
        L.orc_roundtrip_opts.argtypes = [p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double, C.c_float, C.c_float, p, p,
                                         p, p, C.c_int]
        L.orc_roundtrip_full.argtypes = [p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double, C.c_float, C.c_float, p, p,
                                         p, C.c_int, p, p, C.c_int]
        L.orc_render_tracks.argtypes = [p, C.c_int64, C.c_int64, C.c_int, C.c_int, p, C.c_int, C.c_float, C.c_int, p, p]
        L.orc_sound_engine_render.argtypes = [p, C.c_int, p, C.c_int64, C.c_int64, C.c_float, C.c_int, C.c_float, C.c_int64, p]
        L.orc_equal_loudness_volume_from_freq.argtypes = [C.c_float, C.c_int, C.c_float, C.c_float]
        L.orc_equal_loudness_volume_from_freq.restype = C.c_float
        L.orc_equal_loudness_db.argtypes = [C.c_double, C.c_float]
        L.orc_equal_loudness_db.restype = C.c_float
        L.orc_set_sync_phase.argtypes = [C.c_int]
        L.orc_sync_count.argtypes = [C.c_int]
        L.orc_sync_count.restype = C.c_int64
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def hardware_threads() -> int:
    return int(lib().orc_hardware_threads())


def rfft(x: np.ndarray, dtype=np.float64, threads: int = 1) -> np.ndarray:
    """Forward real FFT of each row, unscaled. Returns complex array [batch, n/2+1]."""
    x = np.ascontiguousarray(x, dtype=dtype)
    batch, n = x.shape
    out = np.empty((batch, n // 2 + 1, 2), dtype=dtype)
    fn = lib().orc_rfft_f64 if dtype == np.float64 else lib().orc_rfft_f32
    fn(_ptr(x), n, batch, _ptr(out), threads)
    return out[..., 0] + 1j * out[..., 1]


def num_frames(n_samples: int, window: int, hop: int) -> int:
    return int(lib().orc_num_frames(n_samples, window, hop))


def window(window_type: int, sigma_factor: float, size: int) -> np.ndarray:
    out = np.empty(size, dtype=np.float64)
    lib().orc_window_f64(window_type, float(sigma_factor), size, _ptr(out))
    return out


def stft_analyze(x: np.ndarray, window: int, hop: int, window_type: int = 1, sigma_factor: float = 4.0,
                 sample_rate: int = 48000, min_db: float = -80.0, max_peaks: int = 200, want_sqmag: bool = True,
                 want_peaks: bool = True, threads: int = 1, repeat: int = 0, count_only: bool = False):
    """x: float32 [streams, samples] (or 1-D). Returns dict(frames, sqmag, peak_bin, peak_freq, peak_db, peak_count).
    repeat > 0: analyse the single stream x `repeat` times as independent streams (timing runs; stride 0, no copies);
    count_only: run the peak extraction but keep only the per-frame counts."""
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float32)
    streams, n = x.shape
    stride = n
    if repeat > 0:
        assert streams == 1
        streams, stride = repeat, 0
    frames = num_frames(n, window, hop)
    n_fft = int(lib().orc_fft_length_for(window, 1))
    bins = n_fft // 2 + 1
    sqmag = np.zeros((streams, frames, bins), dtype=np.float64) if want_sqmag else None
    full = want_peaks and not count_only
    pb = np.full((streams, frames, max_peaks), -1, dtype=np.int32) if full else None
    pf = np.zeros((streams, frames, max_peaks), dtype=np.float64) if full else None
    pd = np.zeros((streams, frames, max_peaks), dtype=np.float64) if full else None
    pc = np.zeros((streams, frames), dtype=np.int32) if want_peaks else None
    got = lib().orc_stft_analyze(_ptr(x), n, streams, stride, window, hop, window_type, float(sigma_factor), sample_rate,
                                 float(min_db), max_peaks, _ptr(sqmag), _ptr(pb), _ptr(pf), _ptr(pd), _ptr(pc), threads)
    assert got == frames
    return dict(frames=frames, n_fft=n_fft, sqmag=sqmag, peak_bin=pb, peak_freq=pf, peak_db=pd, peak_count=pc)


class PitchOptions(C.Structure):
    """Mirror of oracle::PitchOptions / ca_pitch_options: optional stages of RtResynth::step (rt.resynth.lib.cpp:1687-1707)."""
    _fields_ = [("shift_pre_autotune", C.c_float), ("harmonize_pre_autotune", C.c_float), ("shift_post_autotune", C.c_float),
                ("harmonize_post_autotune", C.c_float), ("autotune", C.c_int32), ("autotune_max_pitch", C.c_float),
                ("autotune_tolerance_pitches", C.c_float), ("n_table", C.c_int32), ("table", C.c_void_p),
                ("scale_root_pitch", C.c_double)]


def make_ahdsr(attack=0, hold=0, decay=0, release=0, sustain=1.0, attack_itp=0, decay_itp=0, release_itp=0) -> Ahdsr:
    return Ahdsr(attack, attack_itp, hold, decay, decay_itp, release, release_itp, sustain)


def oscbank_render(f0_hz, volume, partial_volume, partial_phase, n_samples: int, sample_rate: int = 48000, env=None,
                   start_angle=None, key_delay=None, release_at=None, pan=None, mix: int = 0, threads: int = 1):
    """Render harmonic-bank voices. partial_volume/phase: [voices, partials]. env: Ahdsr or list of Ahdsr per voice."""
    f0 = np.ascontiguousarray(f0_hz, dtype=np.float32)
    nv = f0.shape[0]
    vol = np.ascontiguousarray(volume, dtype=np.float32)
    pv = np.ascontiguousarray(partial_volume, dtype=np.float32)
    pp = np.ascontiguousarray(partial_phase, dtype=np.float32)
    npart = pv.shape[1]
    if env is None:
        env = make_ahdsr()
    if isinstance(env, Ahdsr):
        envs = (Ahdsr * 1)(env)
        n_env = 1
    else:
        envs = (Ahdsr * len(env))(*env)
        n_env = len(env)
        assert n_env == nv
    sa = None if start_angle is None else np.ascontiguousarray(start_angle, dtype=np.float32)
    kd = None if key_delay is None else np.ascontiguousarray(key_delay, dtype=np.int32)
    ra = None if release_at is None else np.ascontiguousarray(release_at, dtype=np.int64)
    pn = None if pan is None else np.ascontiguousarray(pan, dtype=np.float32)
    if mix == 0:
        out = np.zeros((nv, n_samples), dtype=np.float64)
    elif mix == 1:
        out = np.zeros((n_samples, 2), dtype=np.float64)
    else:
        out = np.zeros((n_samples,), dtype=np.float64)
    lib().orc_oscbank_render(nv, npart, n_samples, sample_rate, _ptr(f0), _ptr(vol), _ptr(sa), _ptr(kd), _ptr(ra),
                             _ptr(pv), _ptr(pp), C.cast(envs, C.c_void_p), n_env, _ptr(pn), mix, _ptr(out), threads)
    return out


def ramp_render(from_hz, to_hz, duration, n_samples: int, sample_rate: int = 48000, volume=None, start_sample=None,
                start_angle=None, key_delay=None, release_at=None, env=None, mix: int = 0, threads: int = 0) -> np.ndarray:
    """gen.voice ramp voices (LogRamp + sine + AHDSR, constant volume), float64 [voices, samples] or [samples] (mix=2)."""
    f = lambda a, t: None if a is None else np.ascontiguousarray(a, dtype=t)
    from_hz, to_hz, duration = f(from_hz, np.float32), f(to_hz, np.float32), f(duration, np.float32)
    V = from_hz.shape[0]
    volume = np.ones(V, np.float32) if volume is None else f(volume, np.float32)
    start_sample, start_angle = f(start_sample, np.float32), f(start_angle, np.float32)
    key_delay, release_at = f(key_delay, np.int32), f(release_at, np.int64)
    e = env if env is not None else make_ahdsr()
    out = np.zeros((V, n_samples) if mix == 0 else (n_samples,), dtype=np.float64)
    q = lambda a: None if a is None else _ptr(a)
    lib().orc_ramp_render(V, n_samples, sample_rate, _ptr(from_hz), _ptr(to_hz), _ptr(duration), q(start_sample), _ptr(volume),
                          q(start_angle), q(key_delay), q(release_at), C.byref(e), mix, _ptr(out),
                          threads if threads > 0 else hardware_threads())
    return out


def envelope_trace(env: Ahdsr, n_steps: int, angle_increment: float, sample_rate: int = 48000, key_delay: int = 0,
                   release_at: int = -1):
    vals = np.zeros(n_steps, dtype=np.float64)
    states = np.zeros(n_steps, dtype=np.int32)
    lib().orc_envelope_trace(C.byref(env), sample_rate, float(angle_increment), key_delay, release_at, n_steps,
                             _ptr(vals), _ptr(states))
    return vals, states


def freq_aliasing_multiplicator(ai: float) -> float:
    return float(lib().orc_freq_aliasing_multiplicator(float(ai)))


def sine_trace(start_angle: float, angle_increment: float, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.float64)
    lib().orc_sine_trace(float(start_angle), float(angle_increment), n, _ptr(out))
    return out


def roundtrip(x: np.ndarray, window: int = 4096, hop: int = 256, sample_rate: int = 48000, min_db: float = -80.0,
              gain: float = 1.0, stereo_spread: float = 1.0, env: Ahdsr | None = None, threads: int = 1, opts=None,
              nearby_distance_tones: float = 0.4, min_volume: float = 1e-4, max_track_pitches: float = 1.0, max_channels: int = 254):
    """Analysis -> pitch pipeline -> resynthesis of each row of x (offline loop of the reference). Returns
    (audio float64 [streams, samples, 2], stats int64 [streams, 2] = notes started, notes dropped)."""
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float32)
    streams, n = x.shape
    out = np.zeros((streams, n, 2), dtype=np.float64)
    stats = np.zeros((streams, 2), dtype=np.int64)
    e = env if env is not None else make_ahdsr()
    # the tracking parameters travel as the floats the product's parameter block holds (ca_roundtrip_params)
    trk = np.array([np.float32(nearby_distance_tones), np.float32(min_volume), np.float32(max_track_pitches)], dtype=np.float64)
    lib().orc_roundtrip_full(_ptr(x), n, streams, sample_rate, window, hop, float(min_db), float(gain), float(stereo_spread),
                             C.byref(e), None if opts is None else C.byref(opts), _ptr(trk), int(max_channels), _ptr(out), _ptr(stats),
                             threads)
    return out, stats


class TrackEvent(C.Structure):
    """orc::TrackEvent = ca_track_event: type 0 NoteOn, 1 NoteOff, 2 NoteChange (EventType, include/events.h:12-17)."""
    _fields_ = [("at_sample", C.c_int64), ("noteid", C.c_int64), ("type", C.c_int32), ("frequency", C.c_float),
                ("velocity", C.c_float), ("reserved", C.c_int32)]


def render_tracks(events, n_samples: int, sample_rate: int = 48000, glide_samples: int = 256, env: Ahdsr | None = None,
                  max_channels: int = 254, stereo_spread: float = 1.0, mono: bool = False):
    """events: [(at_sample, noteid, type, frequency_hz, velocity)] played on the rt.resynth voice pool one sample at a time.
    Returns (audio float64 [n_samples, 2] or [n_samples], (notes started, notes dropped))."""
    arr = (TrackEvent * max(len(events), 1))()
    for i, (at, nid, typ, freq, vel) in enumerate(events):
        arr[i] = TrackEvent(int(at), int(nid), int(typ), float(freq), float(vel), 0)
    out = np.zeros((n_samples,) if mono else (n_samples, 2), dtype=np.float64)
    stats = np.zeros(2, dtype=np.int64)
    e = env if env is not None else make_ahdsr()
    lib().orc_render_tracks(C.cast(arr, C.c_void_p), len(events), n_samples, sample_rate, glide_samples, C.byref(e), max_channels,
                            float(stereo_spread), 1 if mono else 0, _ptr(out), _ptr(stats))
    return out, (int(stats[0]), int(stats[1]))


class RampSpec(C.Structure):
    """orc::RampSpec = ca_ramp_spec: one entry of SoundEngine::RampSpecs (include/soundengine.h:1030-1122)."""
    _fields_ = [("from_hz", C.c_float), ("to_hz", C.c_float), ("duration", C.c_float), ("start_sample", C.c_float),
                ("itp", C.c_int32), ("volume", C.c_float), ("silence_follows", C.c_int32)]


class SoundEngineParams(C.Structure):
    _fields_ = [("sample_rate", C.c_int), ("xfade_len", C.c_int), ("articulative_pause", C.c_int), ("loudness_on", C.c_int),
                ("loudness_ref_index", C.c_int), ("loudness_compensation", C.c_float), ("loudness_level", C.c_float),
                ("sine_gain", C.c_float)]


def sound_engine_render(voices, n_samples: int, sample_rate: int = 48000, xfade_len: int = 480, articulative_pause: int = 0,
                        loudness_on: bool = True, loudness_ref_index: int = 0, loudness_compensation: float = 1.0,
                        loudness_level: float = 40.0, sine_gain: float = 1.0, stereo: bool = False) -> np.ndarray:
    """voices: [dict(specs=[(from_hz, to_hz, duration, start_sample, itp, volume, silence_follows), ...], key_on=int,
    key_off=int or -1, velocity=float, pan=float)]. Returns the sum of the voices, float64 [n_samples] or [n_samples, 2]."""
    out = np.zeros((n_samples, 2) if stereo else (n_samples,), dtype=np.float64)
    prm = SoundEngineParams(sample_rate, xfade_len, articulative_pause, 1 if loudness_on else 0, loudness_ref_index,
                            loudness_compensation, loudness_level, sine_gain)
    for v in voices:
        sp = v["specs"]
        arr = (RampSpec * max(len(sp), 1))()
        for i, t in enumerate(sp):
            arr[i] = RampSpec(float(t[0]), float(t[1]), float(t[2]), float(t[3]), int(t[4]), float(t[5]), int(t[6]))
        lib().orc_sound_engine_render(C.cast(arr, C.c_void_p), len(sp), C.byref(prm), int(v.get("key_on", 0)), int(v.get("key_off", -1)),
                                      float(v.get("velocity", 1.0)), 1 if stereo else 0, float(v.get("pan", 0.0)), n_samples, _ptr(out))
    return out


def equal_loudness_volume_from_freq(freq: float, ref_index: int = 0, log_ratio: float = 1.0, phons: float = 40.0) -> float:
    return float(lib().orc_equal_loudness_volume_from_freq(float(freq), int(ref_index), float(log_ratio), float(phons)))


def equal_loudness_db(pitch: float, phons: float = 60.0) -> float:
    """The equal-loudness contour of `phons` at a midi pitch, in dB (include/loudness.h:232-244)."""
    return float(lib().orc_equal_loudness_db(float(pitch), float(phons)))


def set_sync_phase(on: bool) -> None:
    """Test hook: SynchronizePhase::Yes (the reference's setting, default) or off."""
    lib().orc_set_sync_phase(1 if on else 0)


def sync_count(reset: bool = True) -> int:
    """Test hook: how often a NoteOn took over the phase of a channel with the same angle increment."""
    return int(lib().orc_sync_count(1 if reset else 0))


def pitch_stages(opts: PitchOptions, pitch, volume, cap: int = 1024):
    """shift / harmonize / autotune / shift / harmonize on one (pitch, volume) list; returns the resulting arrays."""
    n = len(pitch)
    pp = np.zeros(max(cap, n), dtype=np.float64)
    vv = np.zeros(max(cap, n), dtype=np.float64)
    pp[:n] = pitch
    vv[:n] = volume
    m = lib().orc_pitch_stages(C.byref(opts), _ptr(pp), _ptr(vv), n, len(pp))
    assert m <= len(pp)
    return pp[:m].copy(), vv[:m].copy()


def make_pitch_options(shift_pre=0.0, harmonize_pre=0.0, shift_post=0.0, harmonize_post=0.0, allowed_pitches=None, scale=None,
                       scale_root_pitch=0.0, max_pitch=151.0, tolerance=0.0) -> PitchOptions:
    o = PitchOptions(shift_pre, harmonize_pre, shift_post, harmonize_post, 0, max_pitch, tolerance, 0, None, scale_root_pitch)
    tab = allowed_pitches if allowed_pitches is not None else scale
    if tab is not None:
        tab = np.ascontiguousarray(tab, dtype=np.float64)
        o.autotune = 1 if allowed_pitches is not None else 2
        o.n_table = int(tab.shape[0])
        o.table = tab.ctypes.data
        o._keep = tab
    return o


def vocoder(modulator: np.ndarray, carrier: np.ndarray, sample_rate: int = 48000, modulator_window_seconds: float = 0.04,
            stride_seconds: float = 0.005, count_bands: int = 20, min_freq: float = 100.0, max_freq: float = 5000.0,
            vol_modulator: float = 0.0, vol_carrier: float = 0.0, vol_vocoder: float = 1.0, threads: int = 1):
    """Vocoder of the reference (source/rt.resynth.lib.vocoder.cpp), one sample at a time in double. modulator / carrier
    float32 [streams, samples]. Returns (out float64 [streams, samples], stride in samples)."""
    m = np.ascontiguousarray(np.atleast_2d(modulator), dtype=np.float32)
    c = np.ascontiguousarray(np.atleast_2d(carrier), dtype=np.float32)
    assert m.shape == c.shape
    streams, n = m.shape
    out = np.zeros((streams, n), dtype=np.float64)
    L = lib()
    L.orc_vocoder.restype = C.c_int
    L.orc_vocoder.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_float, C.c_float, C.c_int, C.c_float,
                              C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int]
    stride = L.orc_vocoder(_ptr(m), _ptr(c), n, streams, sample_rate, modulator_window_seconds, stride_seconds, count_bands,
                           min_freq, max_freq, vol_modulator, vol_carrier, vol_vocoder, _ptr(out), threads)
    return out, int(stride)
