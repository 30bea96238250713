"""ctypes binding of include/cpp_audio_b200.h (one Python name per C entry point, same argument order)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.environ.get("CA_B200_LIB") or os.path.join(_HERE, "libcpp_audio_b200.so")  # override: instrumented builds (profiles/)

DECLARED_SYMBOLS = [
    "ca_version", "ca_init", "ca_destroy", "ca_last_error", "ca_sm_count", "ca_kernel_launches", "ca_host_alloc",
    "ca_host_free", "ca_rfft_batched", "ca_irfft_batched", "ca_rfft_batched_host", "ca_fft_length_for", "ca_stft_num_frames", "ca_stft_plan",
    "ca_stft_analyze", "ca_stft_analyze_host", "ca_oscbank_plan", "ca_oscbank_render", "ca_oscbank_render_host",
    "ca_measure_fp32_peak", "ca_resynth_roundtrip_host", "ca_ramp_render_host", "ca_vocoder_host",
    "ca_oscbank_render_tracks", "ca_oscbank_render_tracks_host", "ca_tracks_open", "ca_tracks_post", "ca_tracks_pull",
    "ca_tracks_pull_host", "ca_tracks_state", "ca_tracks_close", "ca_resynth_truncated_frames", "ca_stft_plan_generation", "ca_sound_engine_render_host", "ca_sound_engine_specs",
    "ca_waveform_harmonics", "ca_resynth_roundtrip_pcm16_host",
]


class CaError(RuntimeError):
    pass


class StftParams(C.Structure):
    _fields_ = [("window", C.c_int32), ("hop", C.c_int32), ("window_type", C.c_int32), ("sigma_factor", C.c_float),
                ("sample_rate", C.c_int32), ("min_db", C.c_float), ("max_peaks", C.c_int32)]


class Ahdsr(C.Structure):
    _fields_ = [("attack", C.c_int32), ("attack_itp", C.c_int32), ("hold", C.c_int32), ("decay", C.c_int32),
                ("decay_itp", C.c_int32), ("release", C.c_int32), ("release_itp", C.c_int32), ("sustain", C.c_float)]


class PitchOptions(C.Structure):
    """ca_pitch_options: shift / harmonize / autotune stages of RtResynth::step (rt.resynth.lib.cpp:1687-1707); zero = off."""
    _fields_ = [("shift_pre_autotune", C.c_float), ("harmonize_pre_autotune", C.c_float), ("shift_post_autotune", C.c_float),
                ("harmonize_post_autotune", C.c_float), ("autotune", C.c_int32), ("autotune_max_pitch", C.c_float),
                ("autotune_tolerance_pitches", C.c_float), ("n_table", C.c_int32), ("table", C.c_void_p),
                ("scale_root_pitch", C.c_double)]


class RoundtripParams(C.Structure):
    _fields_ = [("stft", StftParams), ("nearby_distance_tones", C.c_float), ("min_volume", C.c_float),
                ("max_track_pitches", C.c_float), ("gain", C.c_float), ("stereo_spread", C.c_float), ("env", Ahdsr),
                ("max_channels", C.c_int32), ("host_threads", C.c_int32), ("opts", PitchOptions)]


class VocoderParams(C.Structure):
    _fields_ = [("sample_rate", C.c_int32), ("modulator_window_size_seconds", C.c_float), ("stride_seconds", C.c_float),
                ("count_bands", C.c_int32), ("min_freq", C.c_float), ("max_freq", C.c_float), ("vol_modulator", C.c_float),
                ("vol_carrier", C.c_float), ("vol_vocoder", C.c_float)]


class RampDesc(C.Structure):
    _fields_ = [("n_voices", C.c_int64), ("n_samples", C.c_int64), ("sample_rate", C.c_int32), ("mix", C.c_int32),
                ("from_hz", C.c_void_p), ("to_hz", C.c_void_p), ("duration", C.c_void_p), ("start_sample", C.c_void_p),
                ("volume", C.c_void_p), ("start_angle", C.c_void_p), ("key_delay", C.c_void_p), ("release_at", C.c_void_p),
                ("env", Ahdsr)]


class TrackEvent(C.Structure):
    """ca_track_event: Event of include/events.h:12-107 with the sample time at which it takes effect."""
    _fields_ = [("at_sample", C.c_int64), ("noteid", C.c_int64), ("type", C.c_int32), ("frequency", C.c_float),
                ("velocity", C.c_float), ("reserved", C.c_int32)]


class TracksDesc(C.Structure):
    _fields_ = [("n_events", C.c_int64), ("events", C.c_void_p), ("n_samples", C.c_int64), ("sample_rate", C.c_int32),
                ("glide_samples", C.c_int32), ("env", Ahdsr), ("max_channels", C.c_int32), ("stereo_spread", C.c_float),
                ("mix", C.c_int32)]


class RampSpec(C.Structure):
    """ca_ramp_spec: one entry of SoundEngine::RampSpecs (include/soundengine.h:1030-1122)."""
    _fields_ = [("from_hz", C.c_float), ("to_hz", C.c_float), ("duration", C.c_float), ("start_sample", C.c_float),
                ("itp", C.c_int32), ("volume", C.c_float), ("silence_follows", C.c_int32)]


class SoundEngineDesc(C.Structure):
    _fields_ = [("n_voices", C.c_int64), ("spec_begin", C.c_void_p), ("specs", C.c_void_p), ("key_on", C.c_void_p),
                ("key_off", C.c_void_p), ("velocity", C.c_void_p), ("pan", C.c_void_p), ("n_samples", C.c_int64),
                ("sample_rate", C.c_int32), ("mix", C.c_int32), ("xfade_len", C.c_int32), ("articulative_pause", C.c_int32),
                ("loudness_on", C.c_int32), ("loudness_ref_index", C.c_int32), ("loudness_compensation", C.c_float),
                ("loudness_level", C.c_float), ("sine_gain", C.c_float)]


class SoundEngineGen(C.Structure):
    _fields_ = [("mode", C.c_int32), ("sample_rate", C.c_int32), ("seed", C.c_uint32), ("base_freq", C.c_float),
                ("length_ms", C.c_float), ("length_exp_min", C.c_float), ("length_exp_max", C.c_float), ("freq_scatter", C.c_float),
                ("phase_ratio1", C.c_float), ("itp", C.c_int32), ("xfade_freq", C.c_int32), ("freq_xfade_samples", C.c_int32),
                ("freq_itp", C.c_int32), ("d1", C.c_float), ("d2", C.c_float), ("har_att", C.c_float), ("low_freq", C.c_float),
                ("high_freq", C.c_float), ("start_node", C.c_int32), ("pre_tries", C.c_int32), ("min_path_length", C.c_int32),
                ("additional_tries", C.c_int32)]


class OscBankDesc(C.Structure):
    _fields_ = [("n_voices", C.c_int64), ("n_partials", C.c_int32), ("n_samples", C.c_int64), ("sample_rate", C.c_int32),
                ("f0_hz", C.c_void_p), ("volume", C.c_void_p), ("start_angle", C.c_void_p), ("key_delay", C.c_void_p),
                ("release_at", C.c_void_p), ("partial_volume", C.c_void_p), ("partial_phase", C.c_void_p),
                ("env", C.c_void_p), ("n_env", C.c_int64), ("pan", C.c_void_p), ("mix", C.c_int32), ("first_sample", C.c_int64),
                ("env_kind", C.c_int32)]


_lib = None


def lib_path() -> str:
    return _LIB


def lib():
    """Load the shared library. Raises (loudly) if it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        raise CaError(f"{_LIB} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
    L = C.CDLL(_LIB)
    p, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    L.ca_version.restype = C.c_char_p
    L.ca_init.argtypes = [i32, C.POINTER(p)]
    L.ca_destroy.argtypes = [p]
    L.ca_destroy.restype = None
    L.ca_last_error.argtypes = [p]
    L.ca_last_error.restype = C.c_char_p
    L.ca_sm_count.argtypes = [p]
    L.ca_kernel_launches.argtypes = [p]
    L.ca_kernel_launches.restype = i64
    L.ca_host_alloc.argtypes = [C.c_size_t]
    L.ca_host_alloc.restype = p
    L.ca_host_free.argtypes = [p]
    L.ca_host_free.restype = None
    L.ca_rfft_batched.argtypes = [p, p, i32, i64, p, p]
    L.ca_rfft_batched_host.argtypes = [p, p, i32, i64, p]
    L.ca_irfft_batched.argtypes = [p, p, i32, i64, p, C.c_float, p]
    L.ca_fft_length_for.argtypes = [i32, i32]
    L.ca_stft_num_frames.argtypes = [i64, i32, i32]
    L.ca_stft_num_frames.restype = i64
    L.ca_stft_plan.argtypes = [p, C.POINTER(StftParams)]
    L.ca_stft_analyze.argtypes = [p, p, i64, i64, i64, p, i64, p, p, p, p, p]
    L.ca_stft_analyze_host.argtypes = [p, p, i64, i64, i64, p, i64, p, p, p, p]
    L.ca_oscbank_plan.argtypes = [p, i64, C.c_int32]
    L.ca_oscbank_render.argtypes = [p, C.POINTER(OscBankDesc), p, p]
    L.ca_oscbank_render_host.argtypes = [p, C.POINTER(OscBankDesc), p]
    L.ca_measure_fp32_peak.argtypes = [p, C.POINTER(C.c_double)]
    L.ca_resynth_roundtrip_host.argtypes = [p, C.POINTER(RoundtripParams), p, i64, i64, p, p]
    L.ca_resynth_roundtrip_pcm16_host.argtypes = [p, C.POINTER(RoundtripParams), p, i64, i64, p, p]
    L.ca_ramp_render_host.argtypes = [p, C.POINTER(RampDesc), p]
    L.ca_vocoder_host.argtypes = [p, C.POINTER(VocoderParams), p, p, i64, i64, p]
    L.ca_oscbank_render_tracks.argtypes = [p, C.POINTER(TracksDesc), p, p, p]
    L.ca_oscbank_render_tracks_host.argtypes = [p, C.POINTER(TracksDesc), p, p]
    L.ca_sound_engine_render_host.argtypes = [p, C.POINTER(SoundEngineDesc), p]
    L.ca_sound_engine_specs.argtypes = [C.POINTER(SoundEngineGen), p, C.c_int32]
    L.ca_waveform_harmonics.argtypes = [C.c_int32, C.c_int32, p, p]
    L.ca_stft_plan_generation.argtypes = [p]
    L.ca_stft_plan_generation.restype = C.c_uint64
    L.ca_resynth_truncated_frames.argtypes = [p]
    L.ca_resynth_truncated_frames.restype = i64
    L.ca_tracks_open.argtypes = [p, C.POINTER(TracksDesc), C.POINTER(p)]
    L.ca_tracks_post.argtypes = [p, p, i64]
    L.ca_tracks_pull.argtypes = [p, i64, p, p]
    L.ca_tracks_pull_host.argtypes = [p, i64, p]
    L.ca_tracks_state.argtypes = [p, p]
    L.ca_tracks_close.argtypes = [p]
    L.ca_tracks_close.restype = None
    _lib = L
    return L


def version() -> str:
    return lib().ca_version().decode()


def fft_length_for(n: int, zpf: int = 1) -> int:
    return int(lib().ca_fft_length_for(n, zpf))


def stft_num_frames(n_samples: int, window: int, hop: int) -> int:
    return int(lib().ca_stft_num_frames(n_samples, window, hop))


class Context:
    """One per GPU (ca_ctx). Not thread-safe, like the reference operators."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        rc = lib().ca_init(device, C.byref(self._h))
        if rc != 0:
            raise CaError(f"ca_init(device={device}) failed with {rc}: a B200 (sm_100) is required, there is no CPU fallback")
        self.device = device

    def close(self):
        if self._h:
            lib().ca_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def check(self, rc: int, what: str):
        if rc != 0:
            raise CaError(f"{what} failed ({rc}): {lib().ca_last_error(self._h).decode()}")

    @property
    def sm_count(self) -> int:
        return int(lib().ca_sm_count(self._h))

    @property
    def kernel_launches(self) -> int:
        return int(lib().ca_kernel_launches(self._h))

    def measure_fp32_peak(self) -> float:
        v = C.c_double()
        self.check(lib().ca_measure_fp32_peak(self._h, C.byref(v)), "ca_measure_fp32_peak")
        return v.value
