"""GPU parity of the FFT vocoder (ca_vocoder_host) against the CPU oracle, which feeds the reference's FFTModulator /
FFTCarrier one sample at a time in double (oracle.hpp Vocoder; source/rt.resynth.lib.vocoder.cpp)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SR = 48000


def _signals(n, streams=1, seed=11):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / SR
    mod = np.stack([(0.6 * np.sin(2 * np.pi * (180 + 40 * s) * t) * (0.5 + 0.5 * np.sin(2 * np.pi * (2 + s) * t))
                     + 0.2 * np.sin(2 * np.pi * 2500 * t) * (t > 0.2)) for s in range(streams)])
    car = np.stack([0.3 * np.sign(np.sin(2 * np.pi * (110 + 7 * s) * t)) + rng.uniform(-0.2, 0.2, n) for s in range(streams)])
    return mod.astype(np.float32), car.astype(np.float32)


def test_reference_defaults_match_oracle(pkg, orc, ctx):
    # window 0.10 s (4800 samples, FFT 8192), stride 0.005 s (241 samples, carrier FFT 512), 5 bands 100 Hz .. 20 kHz
    mod, car = _signals(30000, streams=2)
    out = pkg.vocoder(ctx, mod, car)
    ref, stride = orc.vocoder(mod, car, modulator_window_seconds=0.10, stride_seconds=0.005, count_bands=5, max_freq=20000.0)
    assert stride == 241
    assert np.abs(ref).max() > 1e-3
    assert np.abs(out - ref).max() < 1e-5 * max(1.0, np.abs(ref).max())
    assert np.abs(out[:, :4799]).max() == 0.0          # nothing before the first modulator frame


@pytest.mark.parametrize("window_s,stride_s,bands,fmax", [(0.04, 0.01, 20, 5000.0), (0.03, 0.02, 50, 12000.0), (0.2, 0.05, 8, 8000.0)])
def test_other_shapes_match_oracle(pkg, orc, ctx, window_s, stride_s, bands, fmax):
    mod, car = _signals(40000, streams=1, seed=5)
    kw = dict(modulator_window_seconds=window_s, stride_seconds=stride_s, count_bands=bands, min_freq=80.0, max_freq=fmax)
    out = pkg.vocoder(ctx, mod, car, vol_modulator=0.25, vol_carrier=0.5, vol_vocoder=2.0, **kw)
    ref, _ = orc.vocoder(mod, car, vol_modulator=0.25, vol_carrier=0.5, vol_vocoder=2.0, **kw)
    assert np.abs(out - ref).max() < 2e-5 * max(1.0, np.abs(ref).max())


def test_properties_at_full_length(pkg, ctx):
    # 60 s streams: linear in the carrier, dry paths exact, silence in -> silence out
    n = 60 * SR
    mod, car = _signals(n, streams=2, seed=3)
    a = pkg.vocoder(ctx, mod, car)
    b = pkg.vocoder(ctx, mod, 2 * car)
    assert np.abs(b - 2 * a).max() < 1e-5 * np.abs(a).max()
    dry = pkg.vocoder(ctx, mod, car, vol_modulator=1.0, vol_carrier=1.0, vol_vocoder=0.0)
    assert np.array_equal(dry, mod + car)
    z = pkg.vocoder(ctx, np.zeros_like(mod), car)
    assert np.abs(z).max() == 0.0


def test_edge_cases(pkg, ctx):
    mod, car = _signals(300)                    # shorter than any frame: only the dry paths
    out = pkg.vocoder(ctx, mod, car, vol_carrier=1.0)
    assert np.array_equal(out, car)
    with pytest.raises(Exception):
        pkg.vocoder(ctx, mod, car, stride_seconds=0.001)      # carrier FFT below 512 points
    with pytest.raises(Exception):
        pkg.vocoder(ctx, mod, car, count_bands=51)


def test_band_sums_inside_the_analysis_kernel_equal_the_two_kernel_form():
    """CA_VOC_FUSED_BANDS=0 (|X|^2 rows through HBM + band_amps_kernel) and the default (band sums inside the analysis kernel at
    modulator FFT lengths >= 8192) give the same audio; the switch is read once per process, so each form runs in its own interpreter."""
    import os, subprocess, sys
    import numpy as np
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {root!r})\n"
        "import __graft_entry__ as g\n"
        "pkg = g.load_package(); ctx = pkg.Context(0)\n"
        "import cpp_audio_b200.workloads as wl\n"
        "mod = wl.tiled_streams(3, 2.0); car = np.random.default_rng(7).uniform(-0.5, 0.5, mod.shape).astype(np.float32)\n"
        "out = pkg.vocoder(ctx, mod, car)\n"
        "np.save(sys.argv[1], np.asarray(out))\n")
    outs = []
    for mode in ("1", "0"):
        path = f"/tmp/voc_fused_{mode}_{os.getpid()}.npy"
        r = subprocess.run([sys.executable, "-c", code, path], env=dict(os.environ, CA_VOC_FUSED_BANDS=mode), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(np.load(path))
        os.remove(path)
    assert outs[0].shape == outs[1].shape and np.abs(outs[0]).max() > 1e-3
    assert np.abs(outs[0] - outs[1]).max() < 1e-6
