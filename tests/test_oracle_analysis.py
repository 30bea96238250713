"""CPU tests that pin the oracle's analysis half as far as it can be pinned without cpp.algorithms (oracle/SPEC.md):
numpy.fft as an independent FFT anchor, hand-derived known answers for window/scale/peaks, and the frozen golden vectors."""
import math
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_rfft_matches_numpy(orc):
    rng = np.random.default_rng(0)
    for n in (2, 4, 8, 1024, 4096, 32768):
        x = rng.uniform(-1, 1, (3, n))
        assert np.abs(orc.rfft(x) - np.fft.rfft(x, axis=1)).max() < 1e-10 * n
    x32 = rng.uniform(-1, 1, (4, 1024)).astype(np.float32)
    assert np.abs(orc.rfft(x32, np.float32) - np.fft.rfft(x32.astype(np.float64), axis=1)).max() < 2e-4


def test_window_spec(orc):
    w = orc.window(1, 4.0, 4096)
    assert abs(w.mean() - 1.0) < 1e-12                      # S3
    assert np.allclose(w, w[::-1])                          # symmetric
    assert abs(w[0] / w[2047] - math.exp(-0.5 * (4 * 2047.5 / 2048) ** 2) / math.exp(-0.5 * (4 * 0.5 / 2048) ** 2)) < 1e-12  # S2
    assert np.allclose(orc.window(0, 4.0, 64), 1.0)


def test_framing_matches_direct_formula(orc):
    """PeriodicFFT's feed/overlap-copy equals frame k = x[k*hop : k*hop+W] (periodicfft.cpp:55-67,167-172)."""
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, 3000).astype(np.float32)
    for W, hop in ((512, 128), (512, 512), (512, 700), (300, 7)):
        r = orc.stft_analyze(x, W, hop, want_peaks=False)
        w = orc.window(1, 4.0, W)
        n_fft = r["n_fft"]
        F = (len(x) - W) // hop + 1
        assert r["frames"] == F
        for k in (0, F // 2, F - 1):
            seg = np.zeros(n_fft)
            seg[:W] = x[k * hop:k * hop + W].astype(np.float64) * w
            ref = np.abs(np.fft.rfft(seg) / W) ** 2
            assert np.abs(r["sqmag"][0, k] - ref).max() < 1e-15


def test_kat_bin_centred_and_off_bin(orc):
    W = 4096
    n = np.arange(W)
    x = np.sin(2 * np.pi * 100 * n / W).astype(np.float32)
    r = orc.stft_analyze(x, W, 256)
    assert r["peak_count"][0, 0] == 1 and r["peak_bin"][0, 0, 0] == 100
    assert abs(math.sqrt(r["sqmag"][0, 0, 100]) - 0.5) < 1e-7          # S4: unit sine -> 1/2
    assert abs(r["peak_db"][0, 0, 0] - 20 * math.log10(0.5)) < 1e-5
    x = np.sin(2 * np.pi * 100.37 * n / W + 0.3).astype(np.float32)
    r = orc.stft_analyze(x, W, 256)
    assert abs(r["peak_freq"][0, 0, 0] * W / 48000 - 100.37) < 1e-4    # KAT-2
    assert abs(10 ** (r["peak_db"][0, 0, 0] / 20) - 0.5) < 1e-4        # parabola on dB is exact for a Gaussian lobe


def test_c2_expected_partials(orc, wl):
    x = wl.c2_signal(0.5)
    r = orc.stft_analyze(x, 4096, 256)
    assert (r["peak_count"] == 16).all()
    assert np.abs(r["peak_freq"][0, :, :16] - 220.0 * np.arange(1, 17)).max() < 0.02
    assert (r["peak_bin"][0, :, :16] == np.round(220.0 * np.arange(1, 17) * 4096 / 48000).astype(int)).all()


def test_golden_analysis(orc):
    g = np.load(os.path.join(GOLD, "analysis_c2_slice.npz"))
    r = orc.stft_analyze(g["x"], 4096, 256)
    assert np.array_equal(r["peak_bin"][0, :, :16], g["peak_bin"])
    assert np.abs(r["peak_freq"][0, :, :16] - g["peak_freq"]).max() < 1e-9
    assert np.abs(r["sqmag"][0][:, ::16] - g["sqmag_decimated"]).max() < 1e-15


def test_threads_do_not_change_results(orc, wl):
    x = wl.c4_streams(4, 0.2)
    a = orc.stft_analyze(x, 4096, 256, threads=1)
    b = orc.stft_analyze(x, 4096, 256, threads=4)
    assert np.array_equal(a["sqmag"], b["sqmag"]) and np.array_equal(a["peak_bin"], b["peak_bin"])
