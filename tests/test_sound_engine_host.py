"""Host-side pieces of the gen.voice voices that need no GPU: the ramp specs of a note (ca_sound_engine_specs: the lambdas of
SoundEngine's Markov chains, include/soundengine.h:496-736) and the harmonic tables of the non-sine oscillators."""
import numpy as np
import pytest


def test_sweep_is_one_ramp(pkg):
    # create_sweep (:706-724): leaving node 0 builds the one ramp low -> high, phase_ratio1 of the way in
    s = pkg.sound_engine_specs(pkg.synth.MODE_SWEEP, seed=1, length_ms=250.0, low_freq=200.0, high_freq=3000.0, phase_ratio1=0.25, itp=5)
    assert len(s) == 1
    f, t, dur, start, itp, vol, sil = s[0]
    assert abs(f - 200.0) < 1e-3 and abs(t - 3000.0) < 1e-2 and dur == 12000.0 and start == 3000.0 and itp == 5 and vol == 1.0 and sil == 1


def test_birds_specs_follow_the_chain(pkg):
    # birds (:565-596): every spec goes from 4 f to 3 f or from 2 f to 4 f, times one scatter factor per run of equal base values
    for seed in range(20):
        s = pkg.sound_engine_specs(pkg.synth.MODE_BIRDS, seed=seed, base_freq=500.0, length_ms=80.0, length_exp=(-0.5, 0.5), freq_scatter=0.1,
                                   min_path_length=12, start_node=seed % 3)
        assert 1 <= len(s) <= 30
        for f, t, dur, start, itp, vol, sil in s:
            r = t / f
            assert abs(r - 0.75) < 1e-5 or abs(r - 2.0) < 1e-5
            base = f / 4 if abs(r - 0.75) < 1e-5 else f / 2
            assert 500.0 / 1.1 * 0.999 <= base <= 500.0 * 1.1 * 1.001           # the scatter factor lies in [1/1.1, 1.1]
            assert 80 * 48 * 2 ** -0.5 - 1 <= dur <= 80 * 48 * 2 ** 0.5 + 1     # length * 2^U(min_exp, max_exp)
            assert start == 0 and vol == 1.0 and sil == 1
    assert pkg.sound_engine_specs(0, seed=3, min_path_length=9) == pkg.sound_engine_specs(0, seed=3, min_path_length=9)   # a seed, a note
    assert pkg.sound_engine_specs(0, seed=3, min_path_length=9, freq_scatter=.2, length_exp=(-1, 1)) != \
        pkg.sound_engine_specs(0, seed=4, min_path_length=9, freq_scatter=.2, length_exp=(-1, 1))


def test_birds_frequency_transitions(pkg):
    # FreqXfade::All (:533-558): between two ramps a transition spec glides from the end of the first to the start of the second
    s = pkg.sound_engine_specs(pkg.synth.MODE_BIRDS, seed=5, base_freq=400.0, min_path_length=10, xfade_freq=2, freq_xfade_samples=300, freq_itp=16)
    glides = [x for x in s if x[2] == 300.0]
    assert glides and all(x[4] == 16 for x in glides)
    for a, b, c in zip(s, s[1:], s[2:]):
        if b[2] == 300.0 and a[2] != 300.0:
            assert abs(b[0] - a[1]) < 1e-2 * a[1] and abs(b[1] - c[0]) < 1e-3 * c[0]


def test_robot_specs(pkg):
    s = pkg.sound_engine_specs(pkg.synth.MODE_ROBOTS, seed=11, base_freq=300.0, length_ms=60.0, freq_scatter=0.05, d1=7.0, d2=12.0, har_att=0.9,
                               phase_ratio1=0.5, min_path_length=8)
    assert len(s) >= 2
    # node 0 is left first: two steady tones, the first without a pause after it (:604-622)
    assert s[0][0] == s[0][1] and s[1][0] == s[1][1] and s[0][6] == 0 and s[1][6] == 1
    assert all(x[3] == 0.5 * x[2] for x in s)                                   # start position = phase_ratio1 * duration
    v = sorted({round(x[5], 5) for x in s})
    assert v[-1] == 1.0 and (len(v) == 1 or abs(v[0] - 0.9 ** 7) < 1e-4 or abs(v[0] - 0.9 ** 12) < 1e-4)


def test_bad_generator_arguments(pkg):
    with pytest.raises(ValueError):
        pkg.sound_engine_specs(7, seed=1)
    with pytest.raises(ValueError):
        pkg.sound_engine_specs(0, seed=1, length_ms=0.0)
    with pytest.raises(ValueError):
        pkg.sound_engine_specs(0, seed=1, start_node=5)


@pytest.mark.parametrize("wave", [0, 1, 2])
def test_waveform_harmonics_are_the_fourier_series(pkg, wave):
    """sum_h volume sin(h pi a + pi phase) converges to the reference's wave shapes (include/sound.functions.h:86-139)."""
    n = 2000
    vol, ph = pkg.waveform_harmonics(wave, n)
    a = np.linspace(0.0, 2.0, 4001)[1:-1]
    h = np.arange(1, n + 1)[:, None]
    series = (vol[:, None] * np.sin(np.pi * h * a[None, :] + np.pi * ph[:, None])).sum(0)
    if wave == 0:
        ref = np.where(a <= 1.0, a, a - 2.0)
    elif wave == 1:
        ref = np.where((a > 0.5) & (a < 1.5), -1.0, 1.0)
    else:
        ref = np.where(a < 0.5, 2 * a, np.where(a < 1.5, 2 - 2 * a, -4 + 2 * a))
    away = np.ones_like(a, dtype=bool)
    for jump in ((1.0,) if wave == 0 else (0.5, 1.5) if wave == 1 else ()):
        away &= np.abs(a - jump) > 0.02          # Gibbs overshoot stays next to the jumps
    assert np.abs(series - ref)[away].max() < (2e-2 if wave < 2 else 1e-3)
