"""cpp.audio_b200 — B200-native rt.resynth hot path (periodic-FFT analysis + sine-bank resynthesis).

The directory name contains a dot, so it is loaded by path: `import __graft_entry__; pkg = __graft_entry__.load_package()`
(module name `cpp_audio_b200`). The product is the C-ABI shared library `libcpp_audio_b200.so` built from `csrc/`; this
Python layer only binds it with ctypes for tests and benchmarks, and uses torch for device memory, streams and
torch.distributed. There is no CPU fallback: importing works without a GPU (so symbols can be checked), every compute
call raises when the library or a B200 is missing.
"""
from ._capi import (  # noqa: F401
    Ahdsr,
    CaError,
    Context,
    OscBankDesc,
    PitchOptions,
    RampDesc,
    RoundtripParams,
    StftParams,
    TrackEvent,
    TracksDesc,
    VocoderParams,
    fft_length_for,
    lib,
    lib_path,
    stft_num_frames,
    version,
    DECLARED_SYMBOLS,
)
from .analysis import SqMagAnalyzer, analyze_streams, irfft_batched, rfft_batched  # noqa: F401
from .synth import (  # noqa: F401
    NOTE_CHANGE, NOTE_OFF, NOTE_ON, SineBank, render_bank, render_ramps, render_sound_engine, render_tracks, sound_engine_specs,
    waveform_harmonics,
)
from .roundtrip import pinned_empty, pitch_options, resynth_roundtrip, truncated_frames  # noqa: F401
from .vocoder import vocoder  # noqa: F401
from .sharding import frame_shard, gather_rows, reduce_bus, shard_range  # noqa: F401
