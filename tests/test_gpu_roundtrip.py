"""GPU parity of the analysis -> pitch pipeline -> resynthesis round trip (BASELINE config 4) against the CPU oracle,
which runs the reference's offline loop one sample at a time in double (oracle.hpp RoundTrip)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_two_streams_match_oracle(pkg, orc, ctx, wl):
    x = wl.c4_streams(2, 1.0)
    out, stats = pkg.resynth_roundtrip(ctx, x)
    ref, rstats = orc.roundtrip(x)
    assert np.array_equal(stats, rstats)
    assert stats[:, 0].min() >= 16 and stats[:, 1].max() == 0
    assert np.abs(ref).max() > 0.1
    err = np.abs(out - ref).max()
    assert err < 1e-5, err
    assert np.abs(out[:, :4095]).max() == 0.0   # nothing sounds before the first window is full


def test_changing_signal_glides_and_releases(pkg, orc, ctx):
    # a vibrato tone + a tone that stops half way (note off, release) + a late entry (note on): exercises every event
    sr, n = 48000, 60000
    t = np.arange(n) / sr
    a = 0.4 * np.sin(2 * np.pi * (440 * t + 3.0 * np.sin(2 * np.pi * 3 * t) / (2 * np.pi * 3) * 10))
    b = 0.3 * np.sin(2 * np.pi * 1300 * t) * (t < 0.55)
    c = 0.2 * np.sin(2 * np.pi * 150 * t) * (t > 0.7)      # below sr/hop: volume filter in its frequency-limited regime
    x = (a + b + c).astype(np.float32)[None, :]
    out, stats = pkg.resynth_roundtrip(ctx, x)
    ref, rstats = orc.roundtrip(x)
    assert np.array_equal(stats, rstats) and stats[0, 0] >= 3
    err = np.abs(out - ref).max()
    assert err < 2e-5, err


def test_envelope_with_decay_and_pool_limit(pkg, orc, ctx, wl):
    x = wl.c4_streams(1, 0.6)
    env = pkg.Ahdsr(300, 0, 100, 600, 0, 2000, 0, 0.5)
    out, stats = pkg.resynth_roundtrip(ctx, x, env=env, max_channels=8, gain=0.7, stereo_spread=0.3)
    ref, rstats = orc.roundtrip(x, env=orc.make_ahdsr(300, 100, 600, 2000, 0.5), gain=0.7, stereo_spread=0.3) \
        if False else (None, None)
    # the oracle entry point has no max_channels knob: check the pool limit on its own
    assert stats[0, 1] == stats[0, 0] - 8 and stats[0, 1] > 0
    out2, stats2 = pkg.resynth_roundtrip(ctx, x, env=env, gain=0.7, stereo_spread=0.3)
    ref2, rstats2 = orc.roundtrip(x, env=orc.make_ahdsr(300, 100, 600, 2000, 0.5), gain=0.7, stereo_spread=0.3)
    assert np.array_equal(stats2, rstats2)
    assert np.abs(out2 - ref2).max() < 1e-5


def test_edge_cases(pkg, ctx):
    out, stats = pkg.resynth_roundtrip(ctx, np.zeros((2, 3000), np.float32))      # shorter than a window: silence
    assert out.shape == (2, 3000, 2) and np.abs(out).max() == 0 and stats.sum() == 0
    out, stats = pkg.resynth_roundtrip(ctx, np.zeros((1, 20000), np.float32))     # silence in, silence out
    assert np.abs(out).max() == 0 and stats.sum() == 0


def test_reference_default_window_and_stride(pkg, orc, ctx, wl):
    """window 0.1814 s / stride 0.09 s at 48 kHz (source/rt.resynth.lib.cpp:945-946 -> 8706 samples, FFT length 16384, hop
    4320): the frame interval is longer than a render tile (1024 samples) and is cut into several tiles."""
    x = wl.c4_streams(2, 2.0)
    out, stats = pkg.resynth_roundtrip(ctx, x, window=8706, hop=4320)
    ref, rstats = orc.roundtrip(x, window=8706, hop=4320)
    assert np.array_equal(stats, rstats) and stats[:, 0].min() >= 8
    assert np.abs(ref).max() > 0.1
    assert np.abs(out - ref).max() < 2e-5
