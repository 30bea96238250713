"""The oracle's vocoder (sample-by-sample restatement of source/rt.resynth.lib.vocoder.cpp, oracle.hpp Vocoder) against an
independent frame-based statement of the same algorithm in numpy (numpy.fft is the FFT anchor, oracle/SPEC.md S5): the two were
written separately — streaming PeriodicFFT callbacks on one side, index arithmetic on whole arrays on the other."""
import numpy as np

SR = 48000


def numpy_vocoder(mod, car, window_s, stride_s, bands, fmin, fmax):
    S = max(1, int(np.float32(0.5) + np.float32(stride_s) * SR))
    S += (S % 2 == 0)                                             # good_stride, vocoder.cpp:84-92
    Wm = max(1, int(np.float32(0.5) + SR * np.float32(window_s)))
    Wm += (Wm % 2 == 1)                                           # :110-118
    Nm = 1 << int(np.ceil(np.log2(Wm)))
    Nc = 1 << int(np.ceil(np.log2(2 * S)))
    lo, hi = np.log(np.float32(fmin)), np.log(np.float32(fmax))   # float logs, double exp (:15-26)
    freqs = np.exp(np.float64(lo) + np.arange(bands + 1) / bands * (np.float64(hi) - np.float64(lo)))

    def band_of(hz):
        b = np.searchsorted(freqs, hz, side="left") - 1          # freqs[b] < hz <= freqs[b+1]
        b[(hz <= freqs[0]) | (hz > freqs[-1])] = -1
        return b

    # Gaussian window, sigma factor 4, mean 1 (SPEC S2/S3), spectrum scale 1/W (S4)
    j = np.arange(Wm // 2)
    hw = np.exp(-0.5 * (4.0 * (j + 0.5) / (Wm // 2)) ** 2)
    hw *= (Wm // 2) / hw.sum()
    w = np.concatenate([hw[::-1], hw])
    n = len(mod)
    xm = mod.astype(np.float32).astype(np.float64)
    xc = car.astype(np.float32).astype(np.float64)
    frames_m = 0 if n < Wm else 1 + (n - Wm) // S
    amps = np.zeros((frames_m, bands))
    bm = band_of(np.arange(Nm // 2 + 1) * SR / Nm)
    for m in range(frames_m):
        X = np.fft.rfft(xm[m * S:m * S + Wm] * w, Nm) / Wm
        sq = X.real ** 2 + X.imag ** 2
        for b in range(bands):
            amps[m, b] = np.sqrt(sq[bm == b].sum())
    frames_c = 0 if n < 2 * S else 1 + (n - 2 * S) // S
    bc = band_of(np.arange(Nc // 2 + 1) * SR / Nc)
    sig = np.zeros((frames_c, Nc))
    for jf in range(frames_c):
        m = (2 * S + jf * S - Wm) // S if 2 * S + jf * S - Wm >= 0 else -1
        X = np.fft.rfft(xc[jf * S:jf * S + 2 * S], Nc)
        g = np.where(bc >= 0, amps[m][np.maximum(bc, 0)], 0.0) if m >= 0 else np.zeros(len(bc))
        sig[jf] = np.fft.irfft(X * g, Nc)
    out = np.zeros(n)
    p = np.arange(1, S + 1)
    a = 0.5 * (1 - np.cos(np.pi * p / S))                        # SPEC S13
    for jf in range(frames_c):
        n0 = 2 * S - 1 + jf * S
        L = min(S, n - n0)
        old = sig[jf - 1][S:S + L] if jf > 0 else np.zeros(L)
        out[n0:n0 + L] = sig[jf][:L] * a[:L] + old * (1 - a[:L])
    return out, S


def test_oracle_vocoder_matches_numpy_statement(orc):
    rng = np.random.default_rng(9)
    n = 20000
    t = np.arange(n) / SR
    mod = (0.5 * np.sin(2 * np.pi * 220 * t) * (0.6 + 0.4 * np.sin(2 * np.pi * 4 * t)) + 0.1 * np.sin(2 * np.pi * 3100 * t)).astype(np.float32)
    car = rng.uniform(-0.5, 0.5, n).astype(np.float32)
    for window_s, stride_s, bands, fmin, fmax in [(0.10, 0.005, 5, 100.0, 20000.0), (0.04, 0.01, 20, 80.0, 5000.0)]:
        ref, S = numpy_vocoder(mod, car, window_s, stride_s, bands, fmin, fmax)
        out, stride = orc.vocoder(mod, car, modulator_window_seconds=window_s, stride_seconds=stride_s, count_bands=bands,
                                  min_freq=fmin, max_freq=fmax)
        assert stride == S
        assert np.abs(ref).max() > 1e-3
        assert np.abs(out[0] - ref).max() < 1e-10 * max(1.0, np.abs(ref).max())
