"""GPU parity of the round trip with the optional pitch stages (shift / harmonize / autotune, source/rt.resynth.lib.cpp:1687-1707)
against the CPU oracle, and of the device pitch pipeline against its host form with the same options."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MAJOR = [0., 2., 4., 5., 7., 9., 11., 12.]  # rt.resynth.lib.autotune.cpp:142-151


def tones(pitches, amps, seconds=1.0, sr=48000, stop_at=None):
    """sum of sines at midi pitches chosen away from every decision threshold of the stages under test"""
    t = np.arange(int(seconds * sr)) / sr
    x = np.zeros_like(t)
    for i, (p, a) in enumerate(zip(pitches, amps)):
        f = 440.0 * 2.0 ** ((p - 69.0) / 12.0)
        s = a * np.sin(2 * np.pi * f * t + 0.3 * i)
        if stop_at is not None and i == 0:  # the first tone fades out over 50 ms (note off, release)
            s = s * np.clip((stop_at + 0.05 - t) / 0.05, 0.0, 1.0)
        x += s
    return x.astype(np.float32)


def both_opts(pkg, orc, **kw):
    return pkg.pitch_options(**kw), orc.make_pitch_options(**{{"shift_pre": "shift_pre", "harmonize_pre": "harmonize_pre",
                                                              "shift_post": "shift_post", "harmonize_post": "harmonize_post",
                                                              "allowed_pitches": "allowed_pitches", "scale": "scale",
                                                              "scale_root_pitch": "scale_root_pitch", "max_pitch": "max_pitch",
                                                              "tolerance": "tolerance"}[k]: v for k, v in kw.items()})


def check(pkg, orc, ctx, x, tol=2e-5, **kw):
    po, oo = both_opts(pkg, orc, **kw)
    out, stats = pkg.resynth_roundtrip(ctx, x, opts=po)
    ref, rstats = orc.roundtrip(x, opts=oo)
    assert np.array_equal(stats, rstats), (stats, rstats)
    assert np.abs(ref).max() > 0.05
    err = np.abs(out - ref).max()
    assert err < tol, err
    host, hstats = pkg.resynth_roundtrip(ctx, x, opts=po, host_threads=-1)  # host form of the same pipeline
    assert np.array_equal(hstats, stats)
    assert np.abs(host - out).max() < 1e-5
    return out, stats


def test_shift_and_harmonize(pkg, orc, ctx):
    x = np.stack([tones([57.3, 64.6, 71.2], [0.3, 0.2, 0.1]), tones([50.1, 66.4], [0.3, 0.25], stop_at=0.5)])
    base, bstats = pkg.resynth_roundtrip(ctx, x)
    out, stats = check(pkg, orc, ctx, x, shift_pre=3.5, harmonize_post=7.0)
    assert (stats[:, 0] == 2 * bstats[:, 0]).all() and stats[0, 0] == 6  # every note is doubled by its fifth
    assert np.abs(out - base).max() > 0.05
    check(pkg, orc, ctx, x, harmonize_pre=-12.0, shift_post=-2.25)


def test_autotune_chord_table(pkg, orc, ctx):
    # AutotuneType::Chord, OctavePeriodic, a major triad on C (rt.resynth.lib.cpp:1767-1800): root_pitch + i + 12*octave
    allowed = np.sort(np.array([60.0 + i + 12 * o for o in range(-5, 6) for i in (0, 4, 7)]))
    x = np.stack([tones([59.6, 64.3, 66.8, 73.5], [0.3, 0.2, 0.15, 0.1]), tones([48.4, 55.2, 90.3], [0.3, 0.2, 0.1])])
    out, stats = check(pkg, orc, ctx, x, allowed_pitches=allowed, max_pitch=84.0, tolerance=1.0)
    # 73.5 is 1.5 away from 72 and 2.5 from 76: not played; 90.3 is above max_pitch: played untuned
    assert stats[0, 0] == 3 and stats[1, 0] == 3


def test_autotune_scale_with_everything_else(pkg, orc, ctx):
    x = np.stack([tones([58.7, 61.3, 66.2], [0.3, 0.2, 0.2]), tones([70.4, 75.3], [0.3, 0.2])])
    check(pkg, orc, ctx, x, scale=MAJOR, scale_root_pitch=60.0, max_pitch=151.0, tolerance=2.0, shift_pre=0.1,
          harmonize_pre=12.0, shift_post=-12.0, harmonize_post=4.0)


def test_many_streams_pipelined_groups(pkg, orc, ctx, wl):
    # enough streams for the four-group pipeline of the device path; harmonic tones (dense lists)
    x = wl.c4_streams(6, 0.7)
    check(pkg, orc, ctx, x, shift_pre=-5.0, harmonize_post=7.02, tol=3e-5)


def test_bad_options_are_rejected(pkg, ctx):
    x = tones([60.0], [0.3], 0.2)[None, :]
    with pytest.raises(pkg.CaError):
        pkg.resynth_roundtrip(ctx, x, opts=pkg.pitch_options(allowed_pitches=[64.0, 60.0], tolerance=1.0))  # not ascending
    o = pkg.pitch_options()
    o.autotune = 7
    with pytest.raises(pkg.CaError):
        pkg.resynth_roundtrip(ctx, x, opts=o)
    with pytest.raises(pkg.CaError):
        pkg.resynth_roundtrip(ctx, x, opts=pkg.pitch_options(scale=[0.0], scale_root_pitch=60.0))
    with pytest.raises(pkg.CaError):  # the round trip renders LINEAR envelopes only
        pkg.resynth_roundtrip(ctx, x, env=pkg.Ahdsr(100, 5, 0, 0, 0, 100, 0, 1.0))


def test_full_length_streams_with_options(pkg, ctx, wl):
    """BASELINE config 4 length (60 s per stream) with every optional stage on: the device and host forms of the pipeline
    agree, a shift by an octave after harmonising by an octave doubles the notes, and a stream's result does not depend on
    what it is batched with."""
    x = wl.c4_streams(3, 60.0)
    opts = pkg.pitch_options(shift_pre=-12.0, harmonize_post=12.0, scale=MAJOR, scale_root_pitch=57.0, max_pitch=151.0, tolerance=1.0)
    dev, dstats = pkg.resynth_roundtrip(ctx, x, opts=opts)
    host, hstats = pkg.resynth_roundtrip(ctx, x, opts=opts, host_threads=-1)
    plain, pstats = pkg.resynth_roundtrip(ctx, x)
    assert np.array_equal(dstats, hstats) and dstats[:, 1].max() == 0
    assert np.abs(dev[:, :4095]).max() == 0.0 and np.abs(dev).max() > 0.1
    assert np.abs(dev - host).max() < 1e-5
    assert np.abs(dev - plain).max() > 0.05
    alone, astats = pkg.resynth_roundtrip(ctx, x[1:2], opts=opts)
    assert np.array_equal(astats[0], dstats[1]) and np.array_equal(alone[0], dev[1])
