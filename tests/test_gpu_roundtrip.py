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
    ref, rstats = orc.roundtrip(x, env=orc.make_ahdsr(300, 100, 600, 2000, 0.5), gain=0.7, stereo_spread=0.3, max_channels=8)
    assert np.array_equal(stats, rstats)
    assert stats[0, 1] == stats[0, 0] - 8 and stats[0, 1] > 0
    assert np.abs(out - ref).max() < 1e-5
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


# ---- the pitch pipeline on the device (csrc/pitch.cu) against its host form (host_threads != 0 selects the host form) ----

def _both(pkg, ctx, x, **kw):
    dev, dstats = pkg.resynth_roundtrip(ctx, x, host_threads=0, **kw)
    host, hstats = pkg.resynth_roundtrip(ctx, x, host_threads=-1, **kw)
    return dev, dstats, host, hstats


def test_device_pitch_pipeline_matches_host_form(pkg, ctx, wl):
    x = wl.c4_streams(3, 1.5)
    dev, dstats, host, hstats = _both(pkg, ctx, x)
    assert np.array_equal(dstats, hstats) and dstats[:, 0].min() >= 16
    assert np.abs(host).max() > 0.1
    assert np.abs(dev - host).max() < 2e-6     # same formulas; libm vs CUDA math differ in the last bits only


def test_device_pitch_pipeline_dense_spectrum(pkg, orc, ctx):
    """120 partials, a quarter of them switching on and off: more than 32 tuned entries and notes per frame (the loops over
    lane groups), competing matches (the serial fallback), more than 48 sounding notes per tile (segment buffer retry),
    notes that die (pool compaction) and a pool limit below the demand (drops)."""
    sr, n = 48000, 40000
    rng = np.random.default_rng(77)
    t = np.arange(n) / sr
    x = np.zeros(n)
    for h in range(120):
        f = 130.0 * (h + 1) ** 0.93 + rng.uniform(-3, 3)
        amp = 0.3 / (1 + 0.05 * h)
        gate = 1.0
        if h % 4 == 1:
            gate = ((t * (3 + h % 5)) % 1.0) < 0.6
        vib = 1.0 + 0.004 * np.sin(2 * np.pi * (2 + h % 3) * t)
        x += amp * gate * np.sin(2 * np.pi * f * t * vib + rng.uniform(0, 6))
    x = (x / np.abs(x).max() * 0.9).astype(np.float32)
    x = np.stack([x, 0.8 * np.roll(x, 777)])   # two streams: the stream-group pipeline (and its segment-buffer fallback) runs too
    dev, dstats, host, hstats = _both(pkg, ctx, x, nearby_distance_tones=0.05, max_track_pitches=0.3)
    assert np.array_equal(dstats, hstats)
    assert dstats[:, 0].min() > 120 and dstats[:, 1].max() == 0
    assert np.abs(dev - host).max() < 5e-6
    dev1, dstats1 = pkg.resynth_roundtrip(ctx, x)             # reference defaults: checked against the oracle
    ref, rstats = orc.roundtrip(x)
    assert np.array_equal(dstats1, rstats)
    assert np.abs(dev1 - ref).max() < 5e-5
    dev2, dstats2, host2, hstats2 = _both(pkg, ctx, x, nearby_distance_tones=0.05, max_track_pitches=0.3, max_channels=40)
    assert np.array_equal(dstats2, hstats2) and dstats2[:, 1].min() > 0
    assert np.abs(dev2 - host2).max() < 5e-6


def test_device_pitch_pipeline_wide_tracking_window(pkg, ctx, wl):
    # a tracking window wider than the spacing of the partials makes neighbouring entries compete for the same played note
    x = wl.c4_streams(2, 1.0)
    dev, dstats, host, hstats = _both(pkg, ctx, x, max_track_pitches=9.0)
    assert np.array_equal(dstats, hstats)
    assert np.abs(dev - host).max() < 2e-6


def test_device_pitch_pipeline_limits(pkg, ctx):
    x = np.zeros((1, 8192), np.float32)
    pkg.resynth_roundtrip(ctx, x, max_peaks=1024)        # n_fft/4: every local maximum of a frame fits
    with pytest.raises(Exception):
        pkg.resynth_roundtrip(ctx, x, max_peaks=20000)
    with pytest.raises(Exception):
        pkg.resynth_roundtrip(ctx, x, max_channels=300)


def test_peak_capacity_beyond_the_tuned_list(pkg, orc, ctx, wl):
    """With stft.max_peaks = n_fft/4 the weak ripples of the window's side lobes all reach the pitch pipeline (hundreds per frame,
    as in the reference, whose peak list is unbounded); min_volume removes them and the result equals the oracle's, which
    thresholds nothing either (min_db = -inf)."""
    x = wl.c4_streams(2, 1.0)
    out, stats = pkg.resynth_roundtrip(ctx, x, min_db=-np.inf, max_peaks=1024)
    assert pkg.truncated_frames(ctx) == 0
    ref, rstats = orc.roundtrip(x, min_db=-np.inf)
    assert np.array_equal(stats, rstats)
    assert np.abs(out - ref).max() < 1e-5
    pkg.resynth_roundtrip(ctx, x, min_db=-np.inf, max_peaks=64)
    assert pkg.truncated_frames(ctx) > 0                 # 64 slots do not hold them: reported


def test_full_length_streams_against_oracle(pkg, orc, ctx, wl):
    """BASELINE config 4 length (60 s per stream), five streams (uneven stream groups in the pipelined path) against the
    oracle's offline loop in double, at the north star's bar of 1e-5 max-abs. The peaks come from the FP64 form of the
    analysis (csrc/stft64.cu): with float32 spectra the resynthesised phase random-walks to 1e-4 over 60 s
    (profiles/roundtrip_error_vs_length_r02.txt). Also: device and host forms of the pitch pipeline agree, nothing sounds
    before the first window is full, and a stream's result does not depend on what it is batched with."""
    x = wl.c4_streams(5, 60.0)
    dev, dstats = pkg.resynth_roundtrip(ctx, x)
    ref, rstats = orc.roundtrip(x, threads=orc.hardware_threads())
    assert np.array_equal(dstats, rstats) and dstats[:, 0].min() >= 16 and dstats[:, 1].max() == 0
    assert np.abs(ref).max() > 0.1
    err = np.abs(dev - ref).max()
    assert err < 1e-5, err
    host, hstats = pkg.resynth_roundtrip(ctx, x, host_threads=-1)
    assert np.array_equal(dstats, hstats)
    assert np.abs(dev[:, :4095]).max() == 0.0
    assert np.abs(dev - host).max() < 1e-5
    alone, astats = pkg.resynth_roundtrip(ctx, x[3:4])
    assert np.array_equal(astats[0], dstats[3]) and np.array_equal(alone[0], dev[3])


def test_full_length_reference_default_window_against_oracle(pkg, orc, ctx, wl):
    """60 s at the window / stride the reference ships (0.1814 s / 0.09 s -> 8706 samples, 16384-point frames, hop 4320)."""
    x = wl.c4_streams(2, 60.0, first=100)
    dev, dstats = pkg.resynth_roundtrip(ctx, x, window=8706, hop=4320)
    ref, rstats = orc.roundtrip(x, window=8706, hop=4320, threads=orc.hardware_threads())
    assert np.array_equal(dstats, rstats) and dstats[:, 0].min() >= 8
    assert np.abs(ref).max() > 0.1
    err = np.abs(dev - ref).max()
    assert err < 1e-5, err


def test_float32_analysis_switch_still_tracks(pkg, orc, ctx, wl):
    """The FP32 form of the analysis feeds the same pipeline (shapes the FP64 kernel does not take: odd hops)."""
    x = wl.c4_streams(2, 1.0)
    out, stats = pkg.resynth_roundtrip(ctx, x, hop=255)
    ref, rstats = orc.roundtrip(x, hop=255)
    assert np.array_equal(stats, rstats)
    assert np.abs(out - ref).max() < 2e-5


def test_start_phase_synchronisation(pkg, orc, ctx):
    """setPhase<SynchronizePhase::Yes> (include/gen.crtp.h:127-156): note-ons at the increment of a sounding channel take over
    its phase, on fresh channels and -- the second click -- on reused ones, whose oscillators still hold the increment of
    their previous note (tests/test_oracle_sync_phase.py explains the signal). Device and host forms against the oracle."""
    from test_oracle_sync_phase import sync_signal
    x = sync_signal()
    orc.sync_count()
    ref, rstats = orc.roundtrip(x, env=orc.make_ahdsr(0, 0, 0, 20000, 1.0), max_track_pitches=0.0)
    assert orc.sync_count() >= 50                 # 34 on stream 0, most of stream 1
    env = pkg.Ahdsr(0, 0, 0, 0, 0, 20000, 0, 1.0)
    dev, dstats = pkg.resynth_roundtrip(ctx, x, env=env, max_track_pitches=0.0)
    host, hstats = pkg.resynth_roundtrip(ctx, x, env=env, max_track_pitches=0.0, host_threads=-1)
    assert np.array_equal(dstats, rstats) and np.array_equal(hstats, rstats)
    assert np.abs(ref).max() > 0.5
    assert np.abs(host - ref).max() < 1e-5
    assert np.abs(dev - ref).max() < 1e-5
    # a pool smaller than the demand: drops and channel reuse while the rule is at work
    ref8, rstats8 = orc.roundtrip(x, env=orc.make_ahdsr(0, 0, 0, 20000, 1.0), max_track_pitches=0.0, max_channels=8)
    dev8, dstats8 = pkg.resynth_roundtrip(ctx, x, env=env, max_track_pitches=0.0, max_channels=8)
    assert np.array_equal(dstats8, rstats8) and rstats8[:, 1].min() > 0
    assert np.abs(dev8 - ref8).max() < 1e-5


def test_pcm16_input(pkg, orc, ctx, wl):
    """16-bit PCM input (the job's WAV files): converted on the device exactly as a WAV reader would, x / 32768."""
    x = wl.c4_streams(3, 1.0)
    pcm = np.round(x * 32767.0).astype(np.int16)
    as_float = pcm.astype(np.float32) / np.float32(32768.0)
    out, stats = pkg.resynth_roundtrip(ctx, pcm)
    same, sstats = pkg.resynth_roundtrip(ctx, as_float)
    assert np.array_equal(stats, sstats) and np.array_equal(out, same)
    ref, rstats = orc.roundtrip(as_float)
    assert np.array_equal(stats, rstats) and np.abs(out - ref).max() < 1e-5
    odd = pcm[:1, :47999]                                          # odd length, one stream: the sequential path and its scalar tail
    a, _ = pkg.resynth_roundtrip(ctx, odd)
    b, _ = pkg.resynth_roundtrip(ctx, odd.astype(np.float32) / np.float32(32768.0))
    assert np.array_equal(a, b)
