"""GPU parity of the batched real FFT (K1) against the CPU oracle (and numpy as an independent anchor)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def gpu_rfft(pkg, ctx, x):
    out = pkg.rfft_batched(ctx, torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    return out.cpu().numpy()


def test_c1_1024_batch_4096(pkg, orc, ctx, wl):
    x = wl.c1_frames()
    X = gpu_rfft(pkg, ctx, x)
    ref = orc.rfft(x.astype(np.float64))
    # normalised: divide by N/2 so a full-scale sine has magnitude 1
    err = np.abs(X - ref).max() / (x.shape[1] / 2)
    assert err < 1e-5, err
    assert np.abs(ref - np.fft.rfft(x.astype(np.float64), axis=1)).max() < 1e-9


@pytest.mark.parametrize("p", [9, 10, 11, 12, 13, 14, 15])
def test_sizes_in_shared_memory(pkg, orc, ctx, p):
    n = 1 << p
    rng = np.random.default_rng(5000 + p)
    x = rng.uniform(-1, 1, (37, n)).astype(np.float32)
    X = gpu_rfft(pkg, ctx, x)
    ref = orc.rfft(x.astype(np.float64))
    err = np.abs(X - ref).max() / np.sqrt(n)  # white noise: bins scale like sqrt(N)
    assert err < 1e-5, (p, err)


@pytest.mark.parametrize("p,batch", [(16, 5), (17, 3), (18, 3), (19, 2), (20, 2)])
def test_sizes_four_step(pkg, orc, ctx, p, batch):
    n = 1 << p
    rng = np.random.default_rng(5000 + p)
    x = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
    X = gpu_rfft(pkg, ctx, x)
    ref = orc.rfft(x.astype(np.float64))
    err = np.abs(X - ref).max() / np.sqrt(n)
    assert err < 1e-5, (p, err)
    # a pure tone lands in one bin with magnitude N/2 (normalised error)
    t = np.arange(n)
    tone = np.sin(2 * np.pi * 12345 * t / n).astype(np.float32)[None, :]
    T = gpu_rfft(pkg, ctx, tone)[0]
    assert abs(abs(T[12345]) / (n / 2) - 1) < 1e-5
    T[12345] = 0
    assert np.abs(T).max() / (n / 2) < 1e-5


def test_many_frames_four_step_grouping(pkg, ctx):
    # more frames than one work-buffer group (256 MiB / (8*M) frames): exercises the group loop
    n = 1 << 16
    batch = 1100
    x = torch.randn((batch, n), device="cuda")
    X = pkg.rfft_batched(ctx, x)
    ref = torch.fft.rfft(x[[0, 1023, 1024, 1099]].double(), dim=1)   # torch/cuFFT only as an independent cross-check
    assert (X[[0, 1023, 1024, 1099]] - ref).abs().max().item() / np.sqrt(n) < 1e-5


def test_linearity_and_impulse(pkg, ctx):
    n = 4096
    x = np.zeros((3, n), np.float32)
    x[0, 0] = 1.0      # impulse -> flat spectrum of ones
    x[1, 1] = 1.0      # delayed impulse -> exp(-2 pi i k / n)
    x[2, :] = 1.0      # DC
    X = gpu_rfft(pkg, ctx, x)
    k = np.arange(n // 2 + 1)
    assert np.abs(X[0] - 1).max() < 1e-6
    assert np.abs(X[1] - np.exp(-2j * np.pi * k / n)).max() < 1e-6
    assert abs(X[2, 0] - n) < 1e-2 and np.abs(X[2, 1:]).max() < 1e-3


def test_batch_edge_cases(pkg, ctx):
    x = np.random.default_rng(0).uniform(-1, 1, (1, 1024)).astype(np.float32)
    X = gpu_rfft(pkg, ctx, x)
    assert np.abs(X - np.fft.rfft(x.astype(np.float64), axis=1)).max() < 1e-3
    with pytest.raises(pkg.CaError):
        pkg.rfft_batched(ctx, torch.zeros((2, 256), device="cuda"))
    with pytest.raises(pkg.CaError):
        pkg.rfft_batched(ctx, torch.zeros((2, 3000), device="cuda"))


@pytest.mark.parametrize("p", [9, 10, 11, 12, 13, 14, 15])
def test_inverse_real_fft(pkg, ctx, p):
    """ca_irfft_batched against numpy.fft.irfft (the anchor of the oracle's FFT, oracle/SPEC.md S5) and as the inverse of
    the forward transform: irfft(rfft(x)) = N x."""
    n = 1 << p
    rng = np.random.default_rng(7000 + p)
    x = rng.uniform(-1, 1, (19, n)).astype(np.float32)
    X = np.fft.rfft(x.astype(np.float64), axis=1)
    Xt = torch.from_numpy(X.astype(np.complex64)).cuda()
    y = pkg.irfft_batched(ctx, Xt, scale=1.0 / n)
    torch.cuda.synchronize()
    ref = np.fft.irfft(X.astype(np.complex64).astype(np.complex128), n=n, axis=1)
    assert np.abs(y.cpu().numpy() - ref).max() < 1e-5
    xt = torch.from_numpy(x).cuda()
    back = pkg.irfft_batched(ctx, pkg.rfft_batched(ctx, xt), scale=1.0 / n)
    torch.cuda.synchronize()
    assert (back - xt).abs().max().item() < 1e-5
    # a single bin k comes back as a cosine of amplitude 2 (1 for DC and Nyquist)
    S = torch.zeros((3, n // 2 + 1), dtype=torch.complex64, device="cuda")
    S[0, 0] = 1.0
    S[1, 37] = 1.0
    S[2, n // 2] = 1.0
    c = pkg.irfft_batched(ctx, S).cpu().numpy()
    t = np.arange(n)
    assert np.abs(c[0] - 1.0).max() < 1e-6
    assert np.abs(c[1] - 2 * np.cos(2 * np.pi * 37 * t / n)).max() < 1e-5
    assert np.abs(c[2] - np.cos(np.pi * t)).max() < 1e-6


def test_inverse_unsupported_length_fails_loudly(pkg, ctx):
    with pytest.raises(pkg.CaError):
        pkg.irfft_batched(ctx, torch.zeros((1, (1 << 16) // 2 + 1), dtype=torch.complex64, device="cuda"))


@pytest.mark.parametrize("p,batch", [(15, 301), (16, 77), (17, 41)])
def test_cluster_kernels_many_frames(pkg, ctx, p, batch):
    """2^15 ... 2^17 run as one pass with the frame held across a thread-block cluster (rfft_cluster2_kernel): more frames than
    clusters fit on the GPU at once, odd frame counts, every frame checked (torch.fft in double only as an independent anchor)."""
    n = 1 << p
    g = torch.Generator(device="cuda").manual_seed(9000 + p)
    x = torch.empty((batch, n), device="cuda").uniform_(-1, 1, generator=g)
    X = pkg.rfft_batched(ctx, x)
    ref = torch.fft.rfft(x.double(), dim=1)
    err = (X - ref).abs().max().item() / np.sqrt(n)
    assert err < 1e-5, (p, err)


@pytest.mark.parametrize("mode", ["0", "2"])
def test_fft_ab_switch_paths(mode):
    """CA_FFT_CLUSTER=0 (CTA-per-frame / two-pass kernels) and =2 (the other row delivery of the cluster kernels) stay correct:
    the switch is read once per process, so each mode runs in its own interpreter."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np, torch\n"
        f"sys.path.insert(0, {root!r})\n"
        "import __graft_entry__ as g\n"
        "pkg = g.load_package(); ctx = pkg.Context(0)\n"
        "for p, b in ((15, 9), (16, 5), (17, 3)):\n"
        "    x = torch.empty((b, 1 << p), device='cuda').uniform_(-1, 1)\n"
        "    X = pkg.rfft_batched(ctx, x)\n"
        "    err = (X - torch.fft.rfft(x.double(), dim=1)).abs().max().item() / np.sqrt(1 << p)\n"
        "    assert err < 1e-5, (p, err)\n"
        "print('ok')\n")
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, CA_FFT_CLUSTER=mode), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


@pytest.mark.parametrize("p", [15, 16, 17])
def test_cluster_kernels_special_rows(pkg, ctx, p):
    """Known answers that land on the rows the cluster kernels treat apart (row 0 and row P/2 pair with themselves; bins 0, N/4 and
    N/2): a constant, the Nyquist alternation, an impulse (|X[k]| = 1 for every k, phase -2 pi k n0 / N) and tones in bins that
    are multiples of the row count P, one off them, N/4 and N/2 - 1."""
    n = 1 << p
    t = np.arange(n)
    n0 = 12345
    frames = [np.ones(n), np.where(t % 2 == 0, 1.0, -1.0), (t == n0).astype(np.float64)]
    tone_bins = [128, 129, 256, n // 4, n // 4 + 1, n // 2 - 1, 3 * (n // 8) + 5]
    frames += [np.cos(2 * np.pi * k * t / n + 0.3) for k in tone_bins]
    x = np.stack(frames).astype(np.float32)
    X = gpu_rfft(pkg, ctx, x)
    assert X.shape == (len(frames), n // 2 + 1)
    # constant: X[0] = N, everything else 0
    assert abs(X[0, 0] - n) / n < 1e-6 and np.abs(X[0, 1:]).max() / n < 1e-6
    # alternation: X[N/2] = N
    assert abs(X[1, n // 2] - n) / n < 1e-6 and np.abs(X[1, : n // 2]).max() / n < 1e-6
    # impulse: unit magnitude in every bin, linear phase
    k = np.arange(n // 2 + 1)
    assert np.abs(X[2] - np.exp(-2j * np.pi * ((k * n0) % n) / n)).max() < 1e-5
    # tones: N/2 * exp(i 0.3) in their bin, nothing elsewhere
    for row, kb in zip(X[3:], tone_bins):
        assert abs(row[kb] / (n / 2) - np.exp(0.3j)) < 1e-5, (p, kb)
        rest = row.copy()
        rest[kb] = 0
        assert np.abs(rest).max() / (n / 2) < 1e-5, (p, kb)


def test_cluster_kernel_single_frame_and_tail(pkg, ctx):
    """One frame (one cluster) and a frame count that leaves the last wave of clusters partly empty."""
    for p, batch in ((15, 1), (16, 1), (17, 1), (15, 149)):
        n = 1 << p
        rng = np.random.default_rng(4200 + p + batch)
        x = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
        X = gpu_rfft(pkg, ctx, x)
        ref = np.fft.rfft(x.astype(np.float64), axis=1)
        assert np.abs(X - ref).max() / np.sqrt(n) < 1e-5, (p, batch)
