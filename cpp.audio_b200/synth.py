"""Host side of the synthesis path (sine oscillator bank), torch tensors in/out, compute in libcpp_audio_b200.so.

`SineBank` plays the role of a pool of VolumeAdjusted<MultiEnveloped<SineOscillatorAlgo,AHDSREnvelope>> elements driven by
ImplCRTP (include/gen.crtp.h:350-378, :382-531): notes are described up front (frequency, velocity, start phase, key delay,
release time, harmonics) and rendered for the whole duration instead of 16-frame blocks.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._capi import Ahdsr, Context, OscBankDesc, lib

MIX_NONE, MIX_STEREO, MIX_MONO = 0, 1, 2
# enum ca_itp: 0 LINEAR, then 1 + 3*family + mode (family QUAD, CUBIC, QUART, QUINT, SINE, EXPO, CIRC; mode IN, OUT, INOUT)
ITP_LINEAR, ITP_EASE_OUT_CUBIC, ITP_COUNT = 0, 5, 22
BASE_VOLUME_SINE = 0.1  # reduceUnadjustedVolumes * soundBaseVolume(SINE), include/audioelement.h:2370,1270; include/sound.h:84-85


def _opt_ptr(t):
    return None if t is None else t.data_ptr()


class SineBank:
    def __init__(self, ctx: Context, max_voices: int, max_partials: int, sample_rate: int = 48000):
        self.ctx = ctx
        self.sample_rate = sample_rate
        self.max_voices = max_voices
        self.max_partials = max_partials
        ctx.check(lib().ca_oscbank_plan(ctx.handle, max_voices, max_partials), "ca_oscbank_plan")

    def render(self, n_samples: int, f0_hz: torch.Tensor, volume: torch.Tensor, partial_volume: torch.Tensor,
               partial_phase: torch.Tensor, env, start_angle=None, key_delay=None, release_at=None, pan=None,
               mix: int = MIX_NONE, out: torch.Tensor | None = None, first_sample: int = 0,
               env_kind: int | None = None) -> torch.Tensor:
        """All tensors on the same CUDA device. env: one Ahdsr or a list of Ahdsr (one per voice). env_kind: ca_env_kind
        override (0 = unknown: both the LINEAR and the eased pass are launched); by default it is worked out from `env`."""
        dev = f0_hz.device
        nv, npart = partial_volume.shape
        for t, dt in ((f0_hz, torch.float32), (volume, torch.float32), (partial_volume, torch.float32),
                      (partial_phase, torch.float32)):
            if not (t.is_cuda and t.dtype == dt and t.is_contiguous()):
                raise ValueError("note parameters must be contiguous CUDA tensors of the documented dtype")
        envs = [env] if isinstance(env, Ahdsr) else list(env)
        for e in envs:  # the descriptor's env block lives on the device: the library cannot check it there
            if not all(0 <= t < ITP_COUNT for t in (e.attack_itp, e.decay_itp, e.release_itp)):
                raise ValueError("envelope interpolation is not a ca_itp value")
        n_eased = sum((e.attack_itp | e.decay_itp | e.release_itp) != 0 for e in envs)
        if env_kind is None:
            env_kind = 1 if n_eased == 0 else 2 if n_eased == len(envs) else 3  # ca_env_kind: ALL_LINEAR / ALL_EASED / MIXED
        arr = (Ahdsr * len(envs))(*envs)
        env_host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        env_dev = env_host.to(dev)
        if mix == MIX_NONE:
            shape = (nv, n_samples)
        elif mix == MIX_STEREO:
            shape = (n_samples, 2)
        else:
            shape = (n_samples,)
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=dev)
        d = OscBankDesc(nv, npart, n_samples, self.sample_rate, f0_hz.data_ptr(), volume.data_ptr(), _opt_ptr(start_angle),
                        _opt_ptr(key_delay), _opt_ptr(release_at), partial_volume.data_ptr(), partial_phase.data_ptr(),
                        env_dev.data_ptr(), len(envs), _opt_ptr(pan), mix, first_sample, env_kind)
        self.ctx.check(lib().ca_oscbank_render(self.ctx.handle, C.byref(d), out.data_ptr(),
                                               torch.cuda.current_stream(dev).cuda_stream), "ca_oscbank_render")
        self._keep = (env_dev,)  # the render is asynchronous: keep the parameter block alive
        return out


def render_bank(ctx: Context, n_samples: int, f0_hz, volume, partial_volume, partial_phase, env, sample_rate=48000, **kw):
    bank = SineBank(ctx, partial_volume.shape[0], partial_volume.shape[1], sample_rate)
    return bank.render(n_samples, f0_hz, volume, partial_volume, partial_phase, env, **kw)


def render_ramps(ctx: Context, from_hz, to_hz, duration, n_samples: int, sample_rate: int = 48000, volume=None, start_sample=None,
                 start_angle=None, key_delay=None, release_at=None, env: Ahdsr | None = None, mix: int = 0) -> np.ndarray:
    """gen.voice ramp voices (ca_ramp_render_host): host arrays in, float32 [voices, samples] (mix=0) or [samples] (mix=2) out."""
    import ctypes as C

    from ._capi import RampDesc, lib

    f = lambda a, t: None if a is None else np.ascontiguousarray(a, dtype=t)
    from_hz, to_hz, duration = f(from_hz, np.float32), f(to_hz, np.float32), f(duration, np.float32)
    V = int(from_hz.shape[0])
    arrs = dict(volume=f(volume, np.float32), start_sample=f(start_sample, np.float32), start_angle=f(start_angle, np.float32),
                key_delay=f(key_delay, np.int32), release_at=f(release_at, np.int64))
    q = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    e = env if env is not None else Ahdsr(0, 0, 0, 0, 0, 0, 0, 1.0)
    d = RampDesc(V, n_samples, sample_rate, mix, q(from_hz), q(to_hz), q(duration), q(arrs["start_sample"]), q(arrs["volume"]),
                 q(arrs["start_angle"]), q(arrs["key_delay"]), q(arrs["release_at"]), e)
    out = np.zeros((V, n_samples) if mix == 0 else (n_samples,), dtype=np.float32)
    ctx.check(lib().ca_ramp_render_host(ctx.handle, C.byref(d), out.ctypes.data_as(C.c_void_p)), "ca_ramp_render_host")
    return out


NOTE_ON, NOTE_OFF, NOTE_CHANGE = 0, 1, 2  # enum ca_event_type = EventType of include/events.h:12-17


def track_events(events):
    """[(at_sample, noteid, type, frequency_hz, velocity), ...] -> ctypes array of ca_track_event."""
    from ._capi import TrackEvent

    arr = (TrackEvent * max(len(events), 1))()
    for i, (at, nid, typ, freq, vel) in enumerate(events):
        arr[i] = TrackEvent(int(at), int(nid), int(typ), float(freq), float(vel), 0)
    return arr


def render_tracks(ctx: Context, events, n_samples: int, sample_rate: int = 48000, glide_samples: int = 256, env: Ahdsr | None = None,
                  max_channels: int = 254, stereo_spread: float = 1.0, mono: bool = False, device_out: torch.Tensor | None = None):
    """A note schedule on the rt.resynth voice pool (ca_oscbank_render_tracks[_host]). Returns (audio, stats): audio float32
    [n_samples, 2] (or [n_samples] when mono) as a numpy array, or the CUDA tensor `device_out` when given; stats = (notes
    started, notes dropped)."""
    from ._capi import TracksDesc

    arr = track_events(events)
    e = env if env is not None else Ahdsr(0, 0, 0, 0, 0, 0, 0, 1.0)
    d = TracksDesc(len(events), C.cast(arr, C.c_void_p), n_samples, sample_rate, glide_samples, e, max_channels, stereo_spread,
                   MIX_MONO if mono else MIX_STEREO)
    stats = (C.c_int64 * 2)()
    shape = (n_samples,) if mono else (n_samples, 2)
    if device_out is not None:
        assert device_out.is_cuda and device_out.dtype == torch.float32 and device_out.is_contiguous() and tuple(device_out.shape) == shape
        ctx.check(lib().ca_oscbank_render_tracks(ctx.handle, C.byref(d), device_out.data_ptr(), stats,
                                                 torch.cuda.current_stream(device_out.device).cuda_stream), "ca_oscbank_render_tracks")
        return device_out, (int(stats[0]), int(stats[1]))
    out = np.empty(shape, dtype=np.float32)
    ctx.check(lib().ca_oscbank_render_tracks_host(ctx.handle, C.byref(d), out.ctypes.data_as(C.c_void_p), stats),
              "ca_oscbank_render_tracks_host")
    return out, (int(stats[0]), int(stats[1]))


MODE_BIRDS, MODE_ROBOTS, MODE_SWEEP = 0, 1, 2  # SoundEngineMode, include/soundengine.h:30-39
WAVE_SAW, WAVE_SQUARE, WAVE_TRIANGLE = 0, 1, 2


def render_sound_engine(ctx: Context, voices, n_samples: int, sample_rate: int = 48000, xfade_len: int = 480, articulative_pause: int = 0,
                        loudness_on: bool = True, loudness_ref_index: int = 0, loudness_compensation: float = 1.0,
                        loudness_level: float = 40.0, sine_gain: float = 1.0, mix: int = MIX_MONO) -> np.ndarray:
    """gen.voice voices (ca_sound_engine_render_host). voices: [dict(specs=[(from_hz, to_hz, duration, start_sample, itp, volume,
    silence_follows), ...], key_on=int, key_off=int or -1, velocity=float, pan=float)]. Returns float32 [voices, samples] (MIX_NONE),
    [samples] (MIX_MONO) or [samples, 2] (MIX_STEREO)."""
    from ._capi import RampSpec, SoundEngineDesc

    V = len(voices)
    begin = np.zeros(V + 1, dtype=np.int32)
    flat = []
    for i, v in enumerate(voices):
        flat += list(v["specs"])
        begin[i + 1] = len(flat)
    arr = (RampSpec * max(len(flat), 1))()
    for i, t in enumerate(flat):
        arr[i] = RampSpec(float(t[0]), float(t[1]), float(t[2]), float(t[3]), int(t[4]), float(t[5]), int(t[6]))
    key_on = np.array([v.get("key_on", 0) for v in voices], dtype=np.int64)
    key_off = np.array([v.get("key_off", -1) for v in voices], dtype=np.int64)
    vel = np.array([v.get("velocity", 1.0) for v in voices], dtype=np.float32)
    pan = np.array([v.get("pan", 0.0) for v in voices], dtype=np.float32)
    q = lambda a: a.ctypes.data_as(C.c_void_p)
    d = SoundEngineDesc(V, q(begin), C.cast(arr, C.c_void_p), q(key_on), q(key_off), q(vel), q(pan), n_samples, sample_rate, mix, xfade_len,
                        articulative_pause, 1 if loudness_on else 0, loudness_ref_index, loudness_compensation, loudness_level, sine_gain)
    shape = (V, n_samples) if mix == MIX_NONE else (n_samples, 2) if mix == MIX_STEREO else (n_samples,)
    out = np.zeros(shape, dtype=np.float32)
    ctx.check(lib().ca_sound_engine_render_host(ctx.handle, C.byref(d), q(out)), "ca_sound_engine_render_host")
    return out


def sound_engine_specs(mode: int, seed: int, base_freq: float = 440.0, length_ms: float = 100.0, length_exp=(0.0, 0.0),
                       freq_scatter: float = 0.0, phase_ratio1: float = 0.0, itp: int = 0, xfade_freq: int = 0, freq_xfade_samples: int = 480,
                       freq_itp: int = 16, d1: float = 0.0, d2: float = 0.0, har_att: float = 0.9, low_freq: float = 100.0,
                       high_freq: float = 1000.0, start_node: int = 0, pre_tries: int = 0, min_path_length: int = 4,
                       additional_tries: int = 0, sample_rate: int = 48000):
    """The ramp specs of one note (ca_sound_engine_specs): list of (from_hz, to_hz, duration, start_sample, itp, volume, silence_follows)."""
    from ._capi import RampSpec, SoundEngineGen

    g = SoundEngineGen(mode, sample_rate, seed, base_freq, length_ms, length_exp[0], length_exp[1], freq_scatter, phase_ratio1, itp, xfade_freq,
                       freq_xfade_samples, freq_itp, d1, d2, har_att, low_freq, high_freq, start_node, pre_tries, min_path_length, additional_tries)
    arr = (RampSpec * 30)()
    n = lib().ca_sound_engine_specs(C.byref(g), C.cast(arr, C.c_void_p), 30)
    if n < 0:
        raise ValueError(f"ca_sound_engine_specs failed ({n})")
    return [(a.from_hz, a.to_hz, a.duration, a.start_sample, a.itp, a.volume, a.silence_follows) for a in arr[:n]]


def waveform_harmonics(waveform: int, n_harmonics: int):
    """(volume, phase) float32 arrays of the first n harmonics of the reference's saw / square / triangle (ca_waveform_harmonics)."""
    vol = np.zeros(n_harmonics, dtype=np.float32)
    ph = np.zeros(n_harmonics, dtype=np.float32)
    rc = lib().ca_waveform_harmonics(waveform, n_harmonics, vol.ctypes.data_as(C.c_void_p), ph.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise ValueError(f"ca_waveform_harmonics failed ({rc})")
    return vol, ph
