"""GPU parity of the oscillator bank (K3) against the CPU oracle (line-by-line restatement of the reference elements)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-5


def T(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def gpu_render(pkg, ctx, b, n_samples, env, sr=48000, **kw):
    tk = {k: T(v) for k, v in kw.items() if k in ("start_angle", "key_delay", "release_at", "pan")}
    mix = kw.get("mix", 0)
    out = pkg.render_bank(ctx, n_samples, T(b["f0_hz"]), T(b["volume"]), T(b["partial_volume"]), T(b["partial_phase"]), env,
                          sample_rate=sr, mix=mix, first_sample=kw.get("first_sample", 0), env_kind=kw.get("env_kind"), **tk)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def default_env(pkg):
    return pkg.Ahdsr(0, 0, 0, 0, 0, 0, 0, 1.0)  # rt.resynth defaults, source/rt.resynth.lib.cpp:957-961


def orc_env(orc, e):
    return orc.make_ahdsr(e.attack, e.hold, e.decay, e.release, e.sustain, attack_itp=e.attack_itp, decay_itp=e.decay_itp,
                          release_itp=e.release_itp)


def test_small_bank_default_envelope(pkg, orc, ctx, wl):
    b = wl.c3_bank(8, 64)
    n = 20000
    out = gpu_render(pkg, ctx, b, n, default_env(pkg))
    ref = orc.oscbank_render(b["f0_hz"], b["volume"], b["partial_volume"], b["partial_phase"], n)
    err = np.abs(out - ref).max()
    assert err < TOL * 0.1, err  # volume is 0.1: hold the bar on the normalised signal


def test_single_partial_closed_form(pkg, ctx):
    # one voice, one partial, attack 48 samples at 1 kHz: y[n] = 0.1*env*sin(pi*(a0+(n+1)ai)), ai = float32(2f/sr)
    f0 = np.array([1000.0], np.float32)
    b = dict(f0_hz=f0, volume=np.array([0.1], np.float32), partial_volume=np.ones((1, 1), np.float32),
             partial_phase=np.array([[0.25]], np.float32))
    n = 100000
    out = gpu_render(pkg, ctx, b, n, default_env(pkg))[0]
    ai = float(np.float32(2) * f0[0] / np.float32(48000))
    k = np.arange(n)
    env = np.minimum(1.0, (k + 1) / 120.0)  # min change = int(0.5 + 2.5 * 48) = 120 samples > 48
    ref = 0.1 * env * np.sin(np.pi * (0.25 + (k + 1) * ai))
    assert np.abs(out - ref).max() < TOL * 0.1


def test_aliasing_fade_and_nyquist(pkg, orc, ctx):
    # partials between sr/8 and sr/2 are attenuated, above sr/2 silent (include/audioelement.h:466-483)
    b = dict(f0_hz=np.array([3000.0, 7000.0], np.float32), volume=np.array([0.1, 0.1], np.float32),
             partial_volume=np.ones((2, 8), np.float32) / 8, partial_phase=np.zeros((2, 8), np.float32))
    n = 6000
    out = gpu_render(pkg, ctx, b, n, default_env(pkg))
    ref = orc.oscbank_render(b["f0_hz"], b["volume"], b["partial_volume"], b["partial_phase"], n)
    assert np.abs(out - ref).max() < TOL * 0.1


def test_full_ahdsr_with_release(pkg, orc, ctx, wl):
    b = wl.c3_bank(6, 16)
    n = 30000
    e = pkg.Ahdsr(300, 0, 200, 500, 0, 2000, 0, 0.6)
    rel = np.array([9000, 100, 15000, -1, 29990, 400], np.int64)
    kd = np.array([0, 50, 1000, 0, 7, 500], np.int32)
    sa = np.array([0.0, 0.3, -0.7, 0.9, 0.1, 0.0], np.float32)
    out = gpu_render(pkg, ctx, b, n, e, release_at=rel, key_delay=kd, start_angle=sa)
    ref = orc.oscbank_render(b["f0_hz"], b["volume"], b["partial_volume"], b["partial_phase"], n, env=orc_env(orc, e),
                             release_at=rel, key_delay=kd, start_angle=sa)
    assert np.abs(out - ref).max() < TOL * 0.1
    # voice 5 is released before its key delay elapsed: the note is skipped (audioelement.h:774-777)
    assert np.abs(out[5]).max() == 0.0


def test_tune_envelope_ease_out_cubic(pkg, orc, ctx, wl):
    # tune's bank envelopes: EASE_OUT_CUBIC attack and release around a LINEAR decay (source/main.tune.cpp:1146-1161)
    b = wl.c3_bank(4, 16)
    n = 24000
    e = pkg.Ahdsr(2000, 5, 100, 3000, 0, 4000, 5, 0.7)
    rel = np.array([12000, 1000, -1, 6100], np.int64)   # voice 1 is released in the middle of its eased attack
    kd = np.array([0, 11, 300, 0], np.int32)
    out = gpu_render(pkg, ctx, b, n, e, release_at=rel, key_delay=kd)
    ref = orc.oscbank_render(b["f0_hz"], b["volume"], b["partial_volume"], b["partial_phase"], n, env=orc_env(orc, e),
                             release_at=rel, key_delay=kd)
    assert np.abs(out - ref).max() < TOL * 0.1
    lin = gpu_render(pkg, ctx, b, n, pkg.Ahdsr(2000, 0, 100, 3000, 0, 4000, 0, 0.7), release_at=rel, key_delay=kd)
    assert np.abs(out - lin).max() > 1e-3  # the easing is not a no-op


@pytest.mark.parametrize("family", range(7))
def test_every_easing_family(pkg, orc, ctx, wl, family):
    # one voice per (attack, decay, release) rotation of the family's IN / OUT / INOUT modes (ca_itp = 1 + 3*family + mode)
    b = wl.c3_bank(3, 8)
    n = 9000
    t = [1 + 3 * family + m for m in range(3)]
    envs = [pkg.Ahdsr(700 + 10 * i, t[i % 3], 30, 900, t[(i + 1) % 3], 1500, t[(i + 2) % 3], 0.4) for i in range(3)]
    rel = np.array([5000, 1200, 400], np.int64)  # after the decay, inside the decay, inside the attack
    out = gpu_render(pkg, ctx, b, n, envs, release_at=rel)
    ref = orc.oscbank_render(b["f0_hz"], b["volume"], b["partial_volume"], b["partial_phase"], n,
                             env=[orc_env(orc, e) for e in envs], release_at=rel)
    assert np.abs(out - ref).max() < TOL * 0.1


@pytest.mark.parametrize("mix", [0, 1, 2])
@pytest.mark.parametrize("env_kind", [None, 0])
def test_linear_and_eased_voices_in_one_render(pkg, orc, ctx, wl, mix, env_kind):
    # LINEAR and eased voices are rendered by two passes of the kernel (each skips the voices of the other kind); the mixed
    # buses are sums of both. env_kind 0 = the caller does not say what the envelopes are (ca_env_kind).
    b = wl.c3_bank(6, 12)
    n = 7000
    envs = [pkg.Ahdsr(500 + 50 * i, 5 * (i % 2), 0, 300, 2 * (i % 2), 900, 8 * (i % 2), 0.6) for i in range(6)]
    rel = np.array([4000, 4500, -1, 300, 6000, 2500], np.int64)
    pan = np.linspace(-1, 1, 6).astype(np.float32)
    out = gpu_render(pkg, ctx, b, n, envs, release_at=rel, pan=pan, mix=mix, env_kind=env_kind)
    ref = orc.oscbank_render(b["f0_hz"], b["volume"], b["partial_volume"], b["partial_phase"], n,
                             env=[orc_env(orc, e) for e in envs], release_at=rel, pan=pan, mix=mix)
    assert np.abs(out - ref).max() < TOL * 0.1 * (1 if mix == 0 else 6)


def test_unknown_easing_is_rejected(pkg, ctx, wl):
    b = wl.c3_bank(1, 2)
    with pytest.raises(ValueError):
        gpu_render(pkg, ctx, b, 100, pkg.Ahdsr(10, 22, 0, 0, 0, 10, 0, 1.0))
    import ctypes as C
    env = (pkg.Ahdsr * 1)(pkg.Ahdsr(10, 0, 0, 0, -1, 10, 0, 1.0))
    out = np.empty((1, 100), np.float32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p).value
    d = pkg.OscBankDesc(1, 2, 100, 48000, vp(b["f0_hz"]), vp(b["volume"]), None, None, None, vp(b["partial_volume"]),
                        vp(b["partial_phase"]), C.cast(env, C.c_void_p).value, 1, None, 0, 0)
    assert pkg.lib().ca_oscbank_render_host(ctx.handle, C.byref(d), out.ctypes.data_as(C.c_void_p)) == -1  # CA_ERR_INVALID
    # the ramp voices take LINEAR envelopes only and say so
    with pytest.raises(pkg.CaError):
        pkg.render_ramps(ctx, [200.0], [400.0], [1000.0], 500, env=pkg.Ahdsr(10, 5, 0, 0, 0, 10, 0, 1.0))


def test_per_voice_envelopes_and_odd_sizes(pkg, orc, ctx, wl):
    b = wl.c3_bank(5, 7)
    n = 5001
    envs = [pkg.Ahdsr(10 * (i + 1), 0, 5 * i, 100 * i, 0, 50, 0, 1.0 - 0.1 * i) for i in range(5)]
    out = gpu_render(pkg, ctx, b, n, envs)
    ref = orc.oscbank_render(b["f0_hz"], b["volume"], b["partial_volume"], b["partial_phase"], n,
                             env=[orc_env(orc, e) for e in envs])
    assert np.abs(out - ref).max() < TOL * 0.1


def test_inaudible_trailing_harmonics_dropped(pkg, orc, ctx):
    # trailing harmonics with |volume| <= 1e-6 are dropped (audioelement.h:401-415,512); inaudible ones in the middle are kept
    def bank(pv):
        pv = np.array([pv], np.float32)
        return dict(f0_hz=np.array([440.0], np.float32), volume=np.array([1.0], np.float32), partial_volume=pv,
                    partial_phase=np.zeros((1, pv.shape[1]), np.float32))
    a = gpu_render(pkg, ctx, bank([0.5, 9e-7, 0.25, 9e-7, 0.0, 5e-7]), 4000, default_env(pkg))
    b = gpu_render(pkg, ctx, bank([0.5, 9e-7, 0.25, 0.0, 0.0, 0.0]), 4000, default_env(pkg))
    c = gpu_render(pkg, ctx, bank([0.5, 0.0, 0.25, 0.0, 0.0, 0.0]), 4000, default_env(pkg))
    assert np.array_equal(a, b)          # the trailing ones contribute nothing at all
    assert np.abs(b - c).max() > 5e-7    # the one in the middle is played
    bb = bank([0.5, 9e-7, 0.25, 9e-7, 0.0, 5e-7])
    ref = orc.oscbank_render(bb["f0_hz"], bb["volume"], bb["partial_volume"], bb["partial_phase"], 4000)
    assert np.abs(a - ref).max() < 3e-6  # full-scale voice (volume 1)


@pytest.mark.parametrize("mix", [1, 2])
def test_mix_modes(pkg, orc, ctx, wl, mix):
    b = wl.c3_bank(40, 32)
    n = 9000
    pan = np.linspace(-1, 1, 40).astype(np.float32)
    rel = np.full(40, 6000, np.int64)
    out = gpu_render(pkg, ctx, b, n, default_env(pkg), pan=pan, release_at=rel, mix=mix)
    ref = orc.oscbank_render(b["f0_hz"], b["volume"], b["partial_volume"], b["partial_phase"], n, pan=pan, release_at=rel,
                             mix=mix)
    assert out.shape == ref.shape
    assert np.abs(out - ref).max() < TOL * 0.1 * 10  # 40 voices of volume 0.1 on one bus


def test_block_wise_rendering_equals_one_shot(pkg, orc, ctx, wl):
    # first_sample: rendering [a, b) of the timeline gives the same samples as the slice of a full render
    b = wl.c3_bank(4, 16)
    e = pkg.Ahdsr(300, 0, 100, 500, 0, 900, 0, 0.5)
    rel = np.array([5000, -1, 2500, 7000], np.int64)
    kd = np.array([0, 3000, 10, 0], np.int32)
    full = gpu_render(pkg, ctx, b, 9000, e, release_at=rel, key_delay=kd)
    ref = orc.oscbank_render(b["f0_hz"], b["volume"], b["partial_volume"], b["partial_phase"], 9000, env=orc_env(orc, e),
                             release_at=rel, key_delay=kd)
    assert np.abs(full - ref).max() < TOL * 0.1
    for a0, n in ((0, 1000), (1000, 2077), (3077, 5923)):
        part = gpu_render(pkg, ctx, b, n, e, release_at=rel, key_delay=kd, first_sample=a0)
        assert np.abs(part - ref[:, a0:a0 + n]).max() < TOL * 0.1


def test_long_note_phase_stays_exact(pkg, ctx):
    """10 s note (BASELINE config 3 length): fp32 must not drift — compare the last second with the closed form."""
    f0 = np.array([97.3, 3520.1], np.float32)
    b = dict(f0_hz=f0, volume=np.array([0.1, 0.1], np.float32), partial_volume=np.ones((2, 1), np.float32),
             partial_phase=np.array([[0.1], [-0.4]], np.float32))
    n = 480000
    out = gpu_render(pkg, ctx, b, n, default_env(pkg))
    k = np.arange(n - 48000, n)
    for v in range(2):
        ai = float(np.float32(2) * f0[v] / np.float32(48000))
        ref = 0.1 * np.sin(np.pi * (float(b["partial_phase"][v, 0]) + (k + 1) * ai))
        assert np.abs(out[v, -48000:] - ref).max() < TOL * 0.1


def test_c3_full_size_properties(pkg, orc, ctx, wl):
    """BASELINE config 3 at full size (1024 x 64 x 480000): spot-check voices against the oracle, energy for the rest."""
    b = wl.c3_bank()
    n = 480000
    out = pkg.render_bank(ctx, n, T(b["f0_hz"]), T(b["volume"]), T(b["partial_volume"]), T(b["partial_phase"]),
                          default_env(pkg))
    torch.cuda.synchronize()
    sel = [0, 511, 1023]
    ref = orc.oscbank_render(b["f0_hz"][sel], b["volume"][sel], b["partial_volume"][sel], b["partial_phase"][sel], n)
    got = out[sel].cpu().numpy()
    assert np.abs(got - ref).max() < TOL * 0.1
    # every voice: steady-state RMS = 0.1 * sqrt(sum (vol_p g_p)^2 / 2)
    rms = out[:, 48000:].square().mean(dim=1).sqrt().cpu().numpy()
    ai = (np.float32(2) * b["f0_hz"] / np.float32(48000)).astype(np.float64)[:, None] * np.arange(1, 65)[None, :]
    g = np.clip((1.0 / ai - 1.0) / 3.0, 0.0, 1.0)
    expect = 0.1 * np.sqrt(((b["partial_volume"] * g) ** 2).sum(axis=1) / 2)
    assert np.abs(rms - expect).max() < 2e-4


def test_host_entry_point(pkg, orc, ctx, wl):
    import ctypes as C
    b = wl.c3_bank(3, 16)
    n = 7000
    env = (pkg.Ahdsr * 1)(default_env(pkg))
    out = np.empty((3, n), np.float32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p).value
    d = pkg.OscBankDesc(3, 16, n, 48000, vp(b["f0_hz"]), vp(b["volume"]), None, None, None, vp(b["partial_volume"]),
                        vp(b["partial_phase"]), C.cast(env, C.c_void_p).value, 1, None, 0, 0)
    rc = pkg.lib().ca_oscbank_render_host(ctx.handle, C.byref(d), out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    ref = orc.oscbank_render(b["f0_hz"], b["volume"], b["partial_volume"], b["partial_phase"], n)
    assert np.abs(out - ref).max() < TOL * 0.1


@pytest.mark.parametrize("path", ["tensor", "fp32"])
def test_sustain_product_on_both_pipes(pkg, orc, ctx, wl, path, monkeypatch):
    # the sustain regime is a 32 x 64 x 2P matrix product per chunk: on the tensor cores (3 x FP16 split, the default) or on the
    # FP32 pipe (CA_OSC_PATH=fp32). Full-scale voices (volume 1) hold the bar of the normalised signal on either.
    monkeypatch.setenv("CA_OSC_PATH", path)
    b = wl.c3_bank(5, 64)
    b["volume"][:] = 1.0
    n = 3 * 2048 + 777
    out = gpu_render(pkg, ctx, b, n, default_env(pkg))
    ref = orc.oscbank_render(b["f0_hz"], b["volume"], b["partial_volume"], b["partial_phase"], n)
    assert np.abs(out - ref).max() < 2e-6


@pytest.mark.parametrize("mix", [0, 2])
def test_128_partials(pkg, orc, ctx, mix):
    # the largest partial count the bank takes (ca_oscbank_plan): four passes of 32 through the product
    rng = np.random.default_rng(9)
    V, P, n = 3, 128, 9000
    pv = rng.uniform(0.1, 1.0, (V, P)).astype(np.float32)
    b = dict(f0_hz=np.array([31.0, 60.5, 111.0], np.float32), volume=np.full(V, 0.1, np.float32),
             partial_volume=(pv / pv.sum(axis=1, keepdims=True)).astype(np.float32),
             partial_phase=rng.uniform(-1, 1, (V, P)).astype(np.float32))
    e = pkg.Ahdsr(400, 0, 50, 700, 0, 1500, 0, 0.5)
    rel = np.array([6000, -1, 2500], np.int64)
    out = gpu_render(pkg, ctx, b, n, e, release_at=rel, mix=mix)
    ref = orc.oscbank_render(b["f0_hz"], b["volume"], b["partial_volume"], b["partial_phase"], n, env=orc_env(orc, e),
                             release_at=rel, mix=mix)
    assert np.abs(out - ref).max() < TOL * 0.1 * (1 if mix == 0 else 3)


def test_loud_and_faint_voices_share_a_render(pkg, orc, ctx):
    # the FP16 operands of the tensor path are scaled per voice by a power of two chosen from the loudest partial: a voice far
    # above full scale and one at -100 dB keep their relative accuracy
    P, n = 16, 2 * 2048 + 100
    rng = np.random.default_rng(4)
    pv = np.tile((1.0 / np.arange(1, P + 1)).astype(np.float32), (3, 1))
    b = dict(f0_hz=np.array([220.0, 330.0, 95.0], np.float32), volume=np.array([40.0, 1e-5, 1.0], np.float32), partial_volume=pv,
             partial_phase=rng.uniform(-1, 1, (3, P)).astype(np.float32))
    out = gpu_render(pkg, ctx, b, n, default_env(pkg))
    ref = orc.oscbank_render(b["f0_hz"], b["volume"], b["partial_volume"], b["partial_phase"], n)
    for v in range(3):
        assert np.abs(out[v] - ref[v]).max() < 2e-6 * np.abs(ref[v]).max(), v


def test_long_release_and_slow_attack(pkg, orc, ctx, wl):
    # envelope transitions that span many chunks (1 s attack, 1.5 s release): every chunk on the way is product + corrections
    b = wl.c3_bank(3, 24)
    n = 4 * 48000
    e = pkg.Ahdsr(48000, 0, 1000, 20000, 0, 72000, 0, 0.7)
    rel = np.array([100000, 30000, -1], np.int64)   # after the decay; inside the attack; never
    kd = np.array([0, 5000, 47000], np.int32)
    out = gpu_render(pkg, ctx, b, n, e, release_at=rel, key_delay=kd)
    ref = orc.oscbank_render(b["f0_hz"], b["volume"], b["partial_volume"], b["partial_phase"], n, env=orc_env(orc, e),
                             release_at=rel, key_delay=kd)
    assert np.abs(out - ref).max() < TOL * 0.1


@pytest.mark.parametrize("path", ["umma", "mma"])
@pytest.mark.parametrize("n_partials", [64, 20])
def test_long_steady_stretches_on_tcgen05(pkg, orc, ctx, wl, path, n_partials, monkeypatch):
    # whole steady 16,384-sample chunks of a per-voice render go to the tcgen05 kernel (accumulator in tensor memory), the rest
    # (attack, release, tail) to the mma.sync kernel; CA_OSC_PATH=mma renders everything with the latter. Full-scale voices.
    monkeypatch.setenv("CA_OSC_PATH", path)
    b = wl.c3_bank(6, n_partials)
    b["volume"][:] = 1.0
    n = 5 * 16384 + 1000
    e = pkg.Ahdsr(300, 0, 200, 500, 0, 2000, 0, 0.8)
    rel = np.array([-1, 70000, 40000, -1, 100, 16384 * 3], np.int64)   # never; in chunk 4; in chunk 2; never; at once; on a boundary
    kd = np.array([0, 0, 20000, 16384, 0, 5], np.int32)
    out = gpu_render(pkg, ctx, b, n, e, release_at=rel, key_delay=kd)
    ref = orc.oscbank_render(b["f0_hz"], b["volume"], b["partial_volume"], b["partial_phase"], n, env=orc_env(orc, e),
                             release_at=rel, key_delay=kd)
    assert np.abs(out - ref).max() < 2e-6


def test_tcgen05_chunks_with_a_block_offset(pkg, orc, ctx, wl):
    # block-wise rendering (first_sample > 0): the chunk grid is relative to the block, the phases to the note
    b = wl.c3_bank(3, 40)
    n_total, first, n = 200000, 61234, 2 * 16384 + 77
    ref = orc.oscbank_render(b["f0_hz"], b["volume"], b["partial_volume"], b["partial_phase"], n_total)
    part = gpu_render(pkg, ctx, b, n, default_env(pkg), first_sample=first)
    assert np.abs(part - ref[:, first:first + n]).max() < TOL * 0.1
