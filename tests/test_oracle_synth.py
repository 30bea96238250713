"""CPU tests pinning the oracle's synthesis half to the in-tree reference semantics."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oscillator_matches_closed_form(orc):
    # KAT-4: y[n] = sin(pi (a0 + (n+1) ai)) — step precedes read (gen.crtp.h:359-368)
    y = orc.sine_trace(0.25, 0.0123, 480000)
    n = np.arange(480000)
    assert np.abs(y - np.sin(np.pi * (0.25 + (n + 1) * 0.0123))).max() < 1e-9


def test_aliasing_gain_table(orc):
    # KAT-5 (audioelement.h:466-483): ai = 2 f / sr
    g = orc.freq_aliasing_multiplicator
    assert g(0.0) == 1.0
    assert g(2 * (1 / 16)) == 1.0                 # sr/16
    assert g(2 * (1 / 8)) == 1.0                  # sr/8 -> h = 4 -> 1
    assert abs(g(2 * (3 / 16)) - (1 / 0.375 - 1) / 3) < 1e-15
    assert abs(g(2 * (1 / 4)) - 1 / 3) < 1e-15    # sr/4 -> h = 2
    assert g(1.0) == 0.0                          # sr/2 -> h = 1 -> (1-1)/3
    assert g(1.5) == 0.0 and g(-0.1) == 0.0       # negative -> h < 1


def test_envelope_known_answer(orc):
    # KAT-6: A=48,H=0,D=0,S=1,R=48 @48 kHz, f=100 Hz -> min change 1200 samples dominates attack and release
    e = orc.make_ahdsr(attack=48, release=48)
    v, s = orc.envelope_trace(e, 4000, 2 * 100 / 48000, release_at=2000)
    k = np.arange(1, 1201)
    assert np.allclose(v[:1200], k / 1200.0, atol=1e-15)
    assert (v[1200:2000] == 1.0).all()
    assert np.allclose(v[2000:3200], 1.0 - k / 1200.0, atol=1e-15)
    assert (v[3200:] == 0).all()
    assert s[1999] == 1 and s[2000] == 2 and s[3199] == 3 and s[3199 + 17] == 4   # KeyReleased, Done1, then Done2 after 16 more steps


def test_envelope_hold_decay_and_delay(orc):
    e = orc.make_ahdsr(attack=100, hold=10, decay=200, release=300, sustain=0.5)
    v, _ = orc.envelope_trace(e, 2000, 2 * 4000 / 48000, key_delay=7, release_at=1000)
    assert (v[:7] == 0).all()
    assert np.allclose(v[7:107], np.arange(1, 101) / 100.0)
    assert (v[107:117] == 1.0).all()
    assert np.allclose(v[117:317], 1.0 - 0.5 * np.arange(1, 201) / 200.0)
    assert (v[317:1000] == 0.5).all()
    assert np.allclose(v[1000:1300], 0.5 * (1 - np.arange(1, 301) / 300.0))
    # released before the delay elapsed -> note skipped
    v, s = orc.envelope_trace(e, 50, 0.1, key_delay=20, release_at=5)
    assert (v == 0).all() and s[5] == 3


def _penner(itp, u):
    """Independent numpy statement of Penner's easing equations e(u), u = t/d in (0, 1]; enum values of SPEC S7 / ca_itp."""
    u = np.asarray(u, dtype=np.float64)
    if itp == 0:
        return u
    fam, mode = (itp - 1) // 3, (itp - 1) % 3
    if fam <= 3:
        ein = lambda x: x ** (fam + 2)
    elif fam == 4:
        ein = lambda x: 1 - np.cos(x * np.pi / 2)
    elif fam == 5:
        ein = lambda x: np.where(x == 0, 0.0, 2.0 ** (10 * (x - 1)))
    else:
        ein = lambda x: 1 - np.sqrt(np.maximum(0.0, 1 - x * x))
    if fam == 4 and mode == 2:
        return 0.5 * (1 - np.cos(np.pi * u))
    if mode == 0:
        return ein(u)
    if mode == 1:
        return 1 - ein(1 - u)
    return np.where(u < 0.5, 0.5 * ein(np.minimum(2 * u, 1.0)), 1 - 0.5 * ein(np.maximum(2 - 2 * u, 0.0)))


def test_envelope_easings_follow_penner(orc):
    # audioelement.h:733-737,994-998: value = itp::interpolate(type, counter, from, delta, duration); SPEC S7 for the set.
    # tune's envelopes are EASE_OUT_CUBIC attack / release with a LINEAR decay (main.tune.cpp:1146-1161).
    A, H, D, R, S = 100, 10, 200, 300, 0.5
    for itp in range(0, 22):
        e = orc.make_ahdsr(attack=A, hold=H, decay=D, release=R, sustain=S, attack_itp=itp, decay_itp=itp, release_itp=itp)
        v, _ = orc.envelope_trace(e, 2000, 2 * 4000 / 48000, key_delay=7, release_at=1000)
        assert (v[:7] == 0).all()
        assert np.allclose(v[7:107], _penner(itp, np.arange(1, A + 1) / A), atol=1e-12), itp
        assert (v[107:117] == 1.0).all()
        assert np.allclose(v[117:317], 1.0 + (S - 1.0) * _penner(itp, np.arange(1, D + 1) / D), atol=1e-12), itp
        assert (v[317:1000] == S).all()
        assert np.allclose(v[1000:1300], S - S * _penner(itp, np.arange(1, R + 1) / R), atol=1e-12), itp
        assert (v[1300:] == 0).all()
    # release in the middle of an eased attack: the release starts from the value reached (top), :836-841
    e = orc.make_ahdsr(attack=400, release=100, attack_itp=5, release_itp=5)
    v, _ = orc.envelope_trace(e, 600, 2 * 4000 / 48000, release_at=200)
    top = _penner(5, 200 / 400)
    assert abs(v[199] - top) < 1e-12
    assert np.allclose(v[200:300], top - top * _penner(5, np.arange(1, 101) / 100), atol=1e-12)


def test_bank_equals_sum_of_partials(orc):
    # KAT-7: 2 voices x 4 partials against the explicit MultiEnveloped formula (audioelement.h:529-545,590-594,614-619)
    f0 = np.array([440.0, 3000.0], np.float32)
    pv = np.array([[0.5, 0.25, 0.125, 0.0625]] * 2, np.float32)
    pp = np.array([[0.0, 0.5, -0.25, 1.0], [0.1, 0.2, 0.3, 0.4]], np.float32)
    sa = np.array([0.25, -0.5], np.float32)
    n = 6000
    out = orc.oscbank_render(f0, np.array([0.1, 0.05], np.float32), pv, pp, n, start_angle=sa)
    k = np.arange(n)
    for v in range(2):
        ai = float(np.float32(2) * f0[v] / np.float32(48000))
        ref = np.zeros(n)
        for i in range(1, 5):
            a = i * ai
            g = orc.freq_aliasing_multiplicator(a)
            dur = max(48, int(0.5 + 2.5 * 2 / a))
            env = np.minimum(1.0, (k + 1) / dur)
            ref += float(pv[v, i - 1]) * env * g * np.sin(np.pi * (float(pp[v, i - 1]) + i * float(sa[v]) + (k + 1) * a))
        ref *= float(np.float32([0.1, 0.05][v]))
        assert np.abs(out[v] - ref).max() < 1e-9


def test_mix_is_sum_of_voices(orc, wl):
    b = wl.c3_bank(5, 8)
    n = 3000
    pan = np.array([-1, -0.5, 0, 0.5, 1], np.float32)
    per = orc.oscbank_render(b["f0_hz"], b["volume"], b["partial_volume"], b["partial_phase"], n)
    st = orc.oscbank_render(b["f0_hz"], b["volume"], b["partial_volume"], b["partial_phase"], n, pan=pan, mix=1, threads=2)
    mono = orc.oscbank_render(b["f0_hz"], b["volume"], b["partial_volume"], b["partial_phase"], n, mix=2)
    ang = 0.25 * np.pi * (pan.astype(np.float64) + 1)
    assert np.abs(st[:, 0] - (np.cos(ang)[:, None] * per).sum(0)).max() < 1e-12
    assert np.abs(st[:, 1] - (np.sin(ang)[:, None] * per).sum(0)).max() < 1e-12
    assert np.abs(mono - per.sum(0)).max() < 1e-12


def test_golden_synthesis(orc):
    g = np.load(os.path.join(GOLD, "synth_bank.npz"))
    out = orc.oscbank_render(g["f0_hz"], g["volume"], g["partial_volume"], g["partial_phase"], int(g["n_samples"]),
                             env=orc.make_ahdsr(*[int(v) for v in g["env_adhr"]], sustain=float(g["sustain"])),
                             release_at=g["release_at"], key_delay=g["key_delay"], start_angle=g["start_angle"])
    assert np.abs(out[:, ::7] - g["out_decimated"]).max() < 1e-13


def test_log_ramp_known_answers(orc):
    """LogRamp (include/audioelement.h:2464-2572): a leg lasts about `duration` samples, the emitted increment runs from
    `from` to `to` and back, a constant ramp is a plain sine, the phase is the running sum of the increments."""
    sr, n = 48000, 20000
    # constant ramp == oscillator at that frequency: y[i] = sin(pi * (i+1) * ai)
    y = orc.ramp_render([1000.0], [1000.0], [777.0], n)[0]
    ai = float(np.float32(2 * np.float32(1000.0) / sr))
    i = np.arange(n)
    env_floor = max(sr // 1000, int(0.5 + 2.5 * 2 / ai))
    ref = np.sin(np.pi * (i + 1) * ai) * np.minimum(1.0, (i + 1) / env_floor)
    assert np.abs(y - ref).max() < 1e-9
    # instantaneous frequency of a slow up-ramp: phase difference between consecutive samples, away from the attack
    y = orc.ramp_render([500.0], [1000.0], [8000.0], n)[0]
    zc = np.where((y[:-1] < 0) & (y[1:] >= 0))[0]
    periods = np.diff(zc)
    f_first, f_turn = sr / periods[3], sr / periods[np.searchsorted(zc, 7900) - 1]
    assert abs(f_first - 500.0) < 25 and abs(f_turn - 1000.0) < 40
    # it bounces: around 2 * duration the frequency is back near `from`
    f_back = sr / periods[np.searchsorted(zc, 15900) - 1]
    assert abs(f_back - 500.0) < 40


# ---- vocoder restatement (source/rt.resynth.lib.vocoder.cpp): properties that follow from the in-tree lines ----

def test_vocoder_oracle_latency_linearity_and_band_membership(orc):
    sr, n = 48000, 12000
    t = np.arange(n) / sr
    rng = np.random.default_rng(2)
    mod = (0.5 * np.sin(2 * np.pi * 300 * t)).astype(np.float32)
    car = rng.uniform(-0.5, 0.5, n).astype(np.float32)
    kw = dict(modulator_window_seconds=0.04, stride_seconds=0.005, count_bands=20, min_freq=100.0, max_freq=5000.0)
    out, stride = orc.vocoder(mod, car, **kw)
    assert stride == 241 and stride % 2 == 1                         # good_stride, vocoder.cpp:84-92
    wm = 1920                                                        # 0.04 s at 48 kHz, even (:110-118)
    assert np.abs(out[0, :wm - 1]).max() == 0.0                      # band amplitudes are zero before the first modulator frame
    assert np.abs(out[0, wm + 2 * stride:]).max() > 1e-4
    out2, _ = orc.vocoder(mod, 2 * car, **kw)
    assert np.abs(out2 - 2 * out).max() < 1e-12                      # linear in the carrier
    dry, _ = orc.vocoder(mod, car, vol_modulator=1.0, vol_carrier=0.5, vol_vocoder=0.0, **kw)
    assert np.allclose(dry[0], mod.astype(np.float64) + 0.5 * car.astype(np.float64), atol=1e-12)
    # a modulator below every band (50 Hz) leaves all amplitudes near zero: the carrier is muted
    low = (0.5 * np.sin(2 * np.pi * 50 * t)).astype(np.float32)
    muted, _ = orc.vocoder(low, car, **kw)
    assert np.abs(muted).max() < 1e-3 * np.abs(out).max()
