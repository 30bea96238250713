"""GPU parity of the gen.voice ramp voices (ca_ramp_render_host) against the CPU oracle, which steps LogRamp, the
oscillator and the envelope one sample at a time in double (oracle.hpp LogRamp / RampVoice)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-5


def test_up_down_and_bouncing_ramps(pkg, orc, ctx):
    n = 48000
    f0 = np.array([440.0, 2000.0, 300.0, 5000.0, 700.0], np.float32)
    f1 = np.array([880.0, 500.0, 3000.0, 5200.0, 700.0], np.float32)      # the last one is a constant ramp
    dur = np.array([24000.0, 3000.0, 700.0, 120.0, 1000.0], np.float32)   # 2, 16, 68, 400 legs
    vol = np.array([1.0, 0.5, 0.25, 0.8, 0.3], np.float32)
    ang = np.array([0.0, 0.3, -0.7, 1.5, 0.1], np.float32)
    out = pkg.render_ramps(ctx, f0, f1, dur, n, volume=vol, start_angle=ang)
    ref = orc.ramp_render(f0, f1, dur, n, volume=vol, start_angle=ang)
    assert np.abs(ref).max() > 0.7
    err = np.abs(out - ref).max(axis=1)
    assert (err < TOL).all(), err


def test_start_position_envelope_and_release(pkg, orc, ctx):
    n = 30000
    f0 = np.array([600.0, 1500.0, 900.0, 900.0], np.float32)
    f1 = np.array([1200.0, 400.0, 1800.0, 1800.0], np.float32)
    dur = np.array([5000.0, 2500.0, 4000.0, 4000.0], np.float32)
    start = np.array([1250.0, 0.0, 3999.7, 100.0], np.float32)            # mid-ramp, start, about to turn around, early
    kd = np.array([0, 700, 0, 5000], np.int32)
    rel = np.array([20000, 15000, -1, 3000], np.int64)                     # the last voice is released before it starts
    env = pkg.Ahdsr(800, 0, 200, 1500, 0, 2500, 0, 0.6)
    out = pkg.render_ramps(ctx, f0, f1, dur, n, start_sample=start, key_delay=kd, release_at=rel, env=env)
    ref = orc.ramp_render(f0, f1, dur, n, start_sample=start, key_delay=kd, release_at=rel, env=orc.make_ahdsr(800, 200, 1500, 2500, 0.6))
    assert np.abs(ref[:3]).max() > 0.5 and np.abs(ref[3]).max() == 0.0
    err = np.abs(out - ref).max(axis=1)
    assert (err < TOL).all(), err
    assert np.abs(out[1, :700]).max() == 0.0 and np.abs(out[0, 20000 + 2500:]).max() == 0.0


def test_mixed_bus_is_the_sum_of_the_voices(pkg, orc, ctx):
    # a SoundEngine voice = up to three cross-faded ramps: staggered key presses and releases on one bus
    n = 24000
    f0 = np.array([500.0, 650.0, 800.0], np.float32)
    f1 = np.array([650.0, 800.0, 400.0], np.float32)
    dur = np.array([6000.0, 6000.0, 9000.0], np.float32)
    kd = np.array([0, 5000, 11000], np.int32)
    rel = np.array([6000, 12000, 20000], np.int64)
    env = pkg.Ahdsr(960, 0, 0, 0, 0, 960, 0, 1.0)
    out = pkg.render_ramps(ctx, f0, f1, dur, n, key_delay=kd, release_at=rel, env=env, mix=2)
    ref = orc.ramp_render(f0, f1, dur, n, key_delay=kd, release_at=rel, env=orc.make_ahdsr(960, 0, 0, 960, 1.0), mix=2)
    assert out.shape == (n,) and np.abs(ref).max() > 0.5
    assert np.abs(out - ref).max() < 2e-5


def test_rejects_non_positive_parameters(pkg, ctx):
    with pytest.raises(pkg.CaError):
        pkg.render_ramps(ctx, [0.0], [100.0], [1000.0], 100)
    with pytest.raises(pkg.CaError):
        pkg.render_ramps(ctx, [100.0], [200.0], [0.0], 100)
