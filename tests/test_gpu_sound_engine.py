"""GPU parity of the gen.voice voices (ca_sound_engine_render_host) against the oracle's SoundEngineVoice, which steps the three
ramp elements and the orchestration of include/soundengine.h:738-803 one sample at a time in double."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LINEAR, OUT_CUBIC, INOUT_QUAD, IN_SINE, IN_EXPO, INOUT_CIRC = 0, 5, 3, 13, 16, 21


def both(pkg, orc, ctx, voices, n, stereo=False, **kw):
    mix = 1 if stereo else 2
    out = pkg.render_sound_engine(ctx, voices, n, mix=mix, **kw)
    ref = orc.sound_engine_render(voices, n, stereo=stereo, **kw)
    return out, ref


def test_three_crossfaded_ramps_over_two_seconds(pkg, orc, ctx):
    # chirps of 0.1 .. 0.3 s, cross-faded over 10 ms, LINEAR frequency interpolation, plain sine (no loudness volume)
    specs = [(2000, 1500, 9600, 0, LINEAR, 1.0, 1), (1000, 2000, 14400, 0, LINEAR, 0.8, 1), (1500, 1500, 4800, 0, LINEAR, 1.0, 1),
             (3000, 800, 19200, 0, LINEAR, 0.6, 1), (800, 3200, 24000, 6000, LINEAR, 0.9, 1), (440, 880, 12000, 0, LINEAR, 1.0, 1)]
    out, ref = both(pkg, orc, ctx, [dict(specs=specs, key_on=100)], 96000, loudness_on=False)
    assert np.abs(ref).max() > 0.5 and np.abs(ref[90000:]).max() == 0.0
    assert np.abs(out - ref).max() < 1e-5
    assert np.abs(out[:100]).max() == 0.0


@pytest.mark.parametrize("itp", list(range(22)))
def test_every_frequency_interpolation(pkg, orc, ctx, itp):
    # LogRamp with each of the 22 interpolations (include/audioelement.h:2464-2572), bouncing several times per note
    specs = [(600, 2400, 3000, 0, itp, 1.0, 1), (2400, 500, 1700, 300, itp, 0.7, 1)]
    out, ref = both(pkg, orc, ctx, [dict(specs=specs, key_on=0)], 24000, loudness_on=False, xfade_len=2000)
    assert np.abs(ref).max() > 0.3
    assert np.abs(out - ref).max() < 1e-5


def test_loudness_adjusted_sine(pkg, orc, ctx):
    # LoudnessVolumeAdjusted<Sine> (include/audioelement.h:1275-1316): the volume follows the equal-loudness contour of the gliding frequency
    specs = [(150, 6000, 20000, 0, OUT_CUBIC, 1.0, 1), (6000, 90, 20000, 0, LINEAR, 1.0, 1), (1000, 1000, 5000, 0, LINEAR, 0.5, 1)]
    for kw in (dict(loudness_ref_index=17, loudness_level=40.0, loudness_compensation=1.0),
               dict(loudness_ref_index=5, loudness_level=70.0, loudness_compensation=0.6)):
        out, ref = both(pkg, orc, ctx, [dict(specs=specs, key_on=10)], 50000, sine_gain=0.8, **kw)
        assert np.abs(ref).max() > 0.05
        assert np.abs(out - ref).max() < 1e-5
    flat = pkg.render_sound_engine(ctx, [dict(specs=specs, key_on=10)], 50000, loudness_on=False, sine_gain=0.8)
    assert np.abs(flat - out).max() > 0.05      # the loudness volume is at work


def test_orchestration_pauses_short_specs_and_note_off(pkg, orc, ctx):
    # articulative pauses, a spec shorter than the cross-fade (skipped, audioelement.h:774-777), specs short enough to make the
    # engine wait for a free ramp, a spec without a pause after it (robots), and a note-off in the middle of a ramp
    specs = [(900, 900, 2000, 0, LINEAR, 1.0, 0), (1200, 1200, 2000, 0, LINEAR, 0.7, 1), (700, 1400, 300, 0, LINEAR, 1.0, 1),
             (1000, 1000, 700, 0, LINEAR, 1.0, 1), (1300, 1300, 700, 0, LINEAR, 1.0, 1), (1600, 1600, 700, 0, LINEAR, 1.0, 1),
             (500, 2500, 9000, 0, INOUT_QUAD, 1.0, 1), (400, 400, 5000, 0, LINEAR, 1.0, 1)]
    for pause, key_off in ((0, -1), (333, -1), (100, 9000), (0, 1500)):
        out, ref = both(pkg, orc, ctx, [dict(specs=specs, key_on=50, key_off=key_off)], 30000, loudness_on=False, xfade_len=500,
                        articulative_pause=pause)
        assert np.abs(ref).max() > 0.3
        assert np.abs(out - ref).max() < 1e-5, (pause, key_off)


def test_generated_birds_and_robots_in_stereo(pkg, orc, ctx):
    voices = []
    for i in range(6):
        mode = i % 2
        specs = pkg.sound_engine_specs(mode, seed=100 + i, base_freq=300.0 * (1 + i), length_ms=60.0 + 10 * i, length_exp=(-0.5, 0.5),
                                       freq_scatter=0.1, itp=(LINEAR, OUT_CUBIC, IN_SINE)[i % 3], min_path_length=8, xfade_freq=i % 3,
                                       freq_xfade_samples=400, freq_itp=IN_EXPO, d1=7.0, d2=12.0, phase_ratio1=0.3)
        voices.append(dict(specs=specs, key_on=1000 * i, velocity=0.5 + 0.1 * i, pan=-1.0 + 0.4 * i))
    out, ref = both(pkg, orc, ctx, voices, 60000, stereo=True, loudness_ref_index=17)
    assert np.abs(ref).max() > 0.05
    assert np.abs(out - ref).max() < 1e-5
    rows = pkg.render_sound_engine(ctx, voices, 60000, mix=0, loudness_ref_index=17)       # one row per voice
    mono = pkg.render_sound_engine(ctx, voices, 60000, mix=2, loudness_ref_index=17)
    assert rows.shape == (6, 60000) and np.abs(rows.sum(0) - mono).max() < 1e-6


def test_edge_cases(pkg, ctx):
    out = pkg.render_sound_engine(ctx, [], 1000)
    assert out.shape == (1000,) and np.abs(out).max() == 0
    out = pkg.render_sound_engine(ctx, [dict(specs=[], key_on=0)], 1000)      # a note whose chain built nothing
    assert np.abs(out).max() == 0
    out = pkg.render_sound_engine(ctx, [dict(specs=[(500, 600, 4000, 0, 0, 1.0, 1)], key_on=5000)], 1000)   # starts after the end
    assert np.abs(out).max() == 0
    with pytest.raises(Exception):
        pkg.render_sound_engine(ctx, [dict(specs=[(0, 600, 4000, 0, 0, 1.0, 1)])], 1000)     # LogRamp asserts positive frequencies
    with pytest.raises(Exception):
        pkg.render_sound_engine(ctx, [dict(specs=[(500, 600, 4000, 0, 33, 1.0, 1)])], 1000)   # not a ca_itp
    with pytest.raises(Exception):
        pkg.render_sound_engine(ctx, [dict(specs=[(500, 600, 400, 0, 0, 1.0, 1)] * 31)], 1000)  # RampSpecs::n_specs = 30


def test_additive_carriers_on_the_bank(pkg, ctx):
    """Saw / square / triangle by additive synthesis on the oscillator bank (source/rt.resynth.txt:17-23): with every harmonic
    below sr/8 (no aliasing fade) the render is the wave's Fourier partial sum."""
    import torch
    sr, n, f0, H = 48000, 9600, 110.0, 48
    for wave in (0, 1, 2):
        vol, ph = pkg.waveform_harmonics(wave, H)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        out = pkg.render_bank(ctx, n, t(np.array([f0], np.float32)), t(np.array([1.0], np.float32)), t(vol[None, :]), t(ph[None, :]),
                              pkg.Ahdsr(0, 0, 0, 0, 0, 0, 0, 1.0), sample_rate=sr).cpu().numpy()[0]
        k = np.arange(3000, n)                       # past the attack (2.5 periods of the fundamental)
        a = np.float64(np.float32(2 * f0 / sr)) * (k + 1)
        h = np.arange(1, H + 1)[:, None]
        series = (vol[:, None].astype(np.float64) * np.sin(np.pi * (h * a[None, :] + ph[:, None]))).sum(0)
        assert np.abs(out[k] - series).max() < 2e-5, wave
