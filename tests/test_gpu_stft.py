"""GPU parity of the fused analysis kernel (K2) against the CPU oracle, through the C ABI.
Bars (BASELINE.json north_star): peak bins bit-exact on well separated partials; spectra and frequencies within 1e-5."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-5  # max-abs on normalised float32 (|X| of a full-scale sine is 0.5)


def run_gpu(pkg, ctx, x, window, hop, **kw):
    xt = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    res = pkg.analyze_streams(ctx, xt, window, hop, **kw)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in res.items()}


def compare(res, ref, freq_rtol=1e-5, check_db=True):
    S, F, B = ref["sqmag"].shape
    mag = np.sqrt(res["sqmag"].reshape(S, F, B).astype(np.float64))
    err = np.abs(mag - np.sqrt(ref["sqmag"])).max()
    assert err < TOL, f"|X| max-abs error {err}"
    cnt = res["peak_count"].reshape(S, F)
    assert (cnt == ref["peak_count"]).all(), "peak counts differ"
    P = ref["peak_bin"].shape[-1]
    m = np.arange(P)[None, None, :] < np.minimum(cnt, P)[..., None]
    assert (res["peak_bin"].reshape(S, F, P)[m] == ref["peak_bin"][m]).all(), "peak bins differ"
    f = res["peak_freq"].reshape(S, F, P)[m]
    rel = np.abs(f - ref["peak_freq"][m]) / ref["peak_freq"][m]
    assert rel.max() < freq_rtol, f"peak frequency rel error {rel.max()}"
    if check_db:
        d = np.abs(res["peak_db"].reshape(S, F, P)[m] - ref["peak_db"][m])
        assert d.max() < 1e-3, f"peak dB error {d.max()}"
    return err


def test_bin_centred_sine(pkg, orc, ctx):
    W, hop, k = 4096, 256, 100
    n = np.arange(W + 5 * hop)
    x = np.sin(2 * np.pi * k * n / W).astype(np.float32)[None, :]
    res = run_gpu(pkg, ctx, x, W, hop)
    ref = orc.stft_analyze(x, W, hop)
    compare(res, ref)
    assert (res["peak_count"] == 1).all() and (res["peak_bin"][:, 0] == k).all()
    assert abs(math.sqrt(res["sqmag"][0, k]) - 0.5) < 1e-5


def test_harmonic_tone_two_streams(pkg, orc, ctx, wl):
    x = wl.c4_streams(2, 1.0)
    res = run_gpu(pkg, ctx, x, 4096, 256)
    ref = orc.stft_analyze(x, 4096, 256)
    compare(res, ref)
    assert (ref["peak_count"] == 16).all()


def test_c2_slice_matches_expected_partials(pkg, orc, ctx, wl):
    x = wl.c2_signal(2.0)[None, :]
    res = run_gpu(pkg, ctx, x, 4096, 256)
    ref = orc.stft_analyze(x, 4096, 256)
    compare(res, ref)
    F = ref["frames"]
    f = res["peak_freq"].reshape(F, -1)[:, :16]
    assert np.abs(f - 220.0 * np.arange(1, 17)[None, :]).max() < 0.05


@pytest.mark.parametrize("window,hop", [(4096, 1024), (4096, 4096), (4096, 6000), (3000, 250), (2050, 64)])
def test_window_and_hop_variants(pkg, orc, ctx, wl, window, hop):
    x = wl.c4_streams(1, 0.5)
    res = run_gpu(pkg, ctx, x, window, hop)
    ref = orc.stft_analyze(x, window, hop)
    compare(res, ref)


def test_odd_hop_unaligned_path(pkg, orc, ctx, wl):
    x = wl.c4_streams(3, 0.3)[:, :14001]
    x = np.ascontiguousarray(x)
    res = run_gpu(pkg, ctx, x, 4096, 255)
    ref = orc.stft_analyze(x, 4096, 255)
    compare(res, ref)


def test_rectangular_window_bin_centred(pkg, orc, ctx):
    # bin-centred partials: the rectangular window has no leakage for them
    W = 4096
    n = np.arange(W + 512)
    x = (0.4 * np.sin(2 * np.pi * 64 * n / W) + 0.3 * np.sin(2 * np.pi * 300 * n / W + 1.0)).astype(np.float32)[None, :]
    res = run_gpu(pkg, ctx, x, W, 256, window_type=0, min_db=-60.0)
    ref = orc.stft_analyze(x, W, 256, window_type=0, min_db=-60.0)
    compare(res, ref, check_db=False)


def test_edge_cases(pkg, orc, ctx):
    # too short for a single frame: zero frames, nothing launched, no error
    x = np.zeros((2, 4000), dtype=np.float32)
    res = run_gpu(pkg, ctx, x, 4096, 256)
    assert res["sqmag"].shape[0] == 0 and res["peak_count"].shape[0] == 0
    # silence: no peaks, zero spectrum
    x = np.zeros((1, 8192), dtype=np.float32)
    res = run_gpu(pkg, ctx, x, 4096, 256)
    assert (res["peak_count"] == 0).all() and (res["sqmag"] == 0).all()
    # exactly one frame
    x = np.random.default_rng(3).uniform(-1, 1, (1, 4096)).astype(np.float32)
    res = run_gpu(pkg, ctx, x, 4096, 256, min_db=-30.0)
    ref = orc.stft_analyze(x, 4096, 256, min_db=-30.0)
    assert res["sqmag"].shape[0] == 1
    assert np.abs(np.sqrt(res["sqmag"]) - np.sqrt(ref["sqmag"][0])).max() < TOL


def test_peak_capacity_overflow(pkg, orc, ctx):
    # white noise, no threshold: far more local maxima than max_peaks; count reports all, the first max_peaks are stored
    x = np.random.default_rng(7).uniform(-1, 1, (1, 4096 + 256)).astype(np.float32)
    res = run_gpu(pkg, ctx, x, 4096, 256, min_db=-math.inf, max_peaks=32)
    ref = orc.stft_analyze(x, 4096, 256, min_db=-math.inf, max_peaks=32)
    assert (res["peak_count"].reshape(ref["peak_count"].shape) == ref["peak_count"]).all()
    assert (ref["peak_count"] > 32).all()
    assert (res["peak_bin"].reshape(ref["peak_bin"].shape) == ref["peak_bin"]).all()


def test_unsupported_fft_length_fails_loudly(pkg, ctx):
    with pytest.raises(pkg.CaError):
        pkg.SqMagAnalyzer(ctx, 512, 256)       # fft length 512 < 1024
    with pytest.raises(pkg.CaError):
        pkg.SqMagAnalyzer(ctx, 40000, 256)     # fft length 65536 > 32768
    with pytest.raises(pkg.CaError):
        pkg.SqMagAnalyzer(ctx, 4095, 256)      # odd window (Assert at periodicfft.cpp:274)


@pytest.mark.parametrize("window,hop,sr", [(1024, 256, 48000), (2000, 300, 48000), (8000, 3969, 44100), (16384, 1024, 48000),
                                           (30000, 5000, 48000)])
def test_other_fft_lengths_generic_path(pkg, orc, ctx, wl, window, hop, sr):
    """Windows whose FFT length is not 4096 take the fused kernels of csrc/rfftw.cu (window 8000 / stride 3969 at 44.1 kHz are the
    reference defaults: 0.1814 s and 0.09 s, source/rt.resynth.lib.cpp:945-946)."""
    x = wl.c4_streams(2, 1.5)
    res = run_gpu(pkg, ctx, x, window, hop, sample_rate=sr)
    ref = orc.stft_analyze(x, window, hop, sample_rate=sr)
    compare(res, ref)
    assert ref["peak_count"].min() >= 1


@pytest.mark.parametrize("window,hop", [(1024, 128), (2048, 511), (8192, 256), (16000, 700)])
def test_fused_other_lengths_many_frames_and_options(pkg, orc, ctx, wl, window, hop):
    """The single-pass kernels of the other FFT lengths (rfftw.cu): more frames than resident warps / CTAs, three streams of
    odd length (unaligned second and third stream), peak capacity beyond the shared-memory list, peaks-only and
    spectrum-only calls."""
    x = np.ascontiguousarray(wl.c4_streams(3, 2.0)[:, :90001])
    ref = orc.stft_analyze(x, window, hop, max_peaks=300)
    res = run_gpu(pkg, ctx, x, window, hop, max_peaks=300)
    compare(res, ref)
    ref = orc.stft_analyze(x, window, hop)
    res = run_gpu(pkg, ctx, x, window, hop, want_sqmag=False)
    assert "sqmag" not in res
    S, F = ref["peak_count"].shape
    assert (res["peak_count"].reshape(S, F) == ref["peak_count"]).all()
    m = np.arange(ref["peak_bin"].shape[-1])[None, None, :] < ref["peak_count"][..., None]
    assert (res["peak_bin"].reshape(ref["peak_bin"].shape)[m] == ref["peak_bin"][m]).all()
    res = run_gpu(pkg, ctx, x, window, hop, want_peaks=False)
    mag = np.sqrt(res["sqmag"].reshape(ref["sqmag"].shape).astype(np.float64))
    assert np.abs(mag - np.sqrt(ref["sqmag"])).max() < TOL


def test_peaks_only_and_sqmag_only(pkg, orc, ctx, wl):
    x = wl.c4_streams(2, 0.4)
    ref = orc.stft_analyze(x, 4096, 256)
    a = run_gpu(pkg, ctx, x, 4096, 256, want_sqmag=False)
    b = run_gpu(pkg, ctx, x, 4096, 256, want_peaks=False)
    assert "sqmag" not in a and "peak_bin" not in b
    m = np.arange(ref["peak_bin"].shape[-1])[None, None, :] < ref["peak_count"][..., None]
    assert (a["peak_count"].reshape(ref["peak_count"].shape) == ref["peak_count"]).all()
    assert (a["peak_bin"].reshape(ref["peak_bin"].shape)[m] == ref["peak_bin"][m]).all()
    assert np.abs(np.sqrt(b["sqmag"].reshape(ref["sqmag"].shape)) - np.sqrt(ref["sqmag"])).max() < TOL


def test_full_c2_against_oracle(pkg, orc, ctx, wl):
    """BASELINE config 2 at its full size (60 s, 11,235 frames) against the oracle: spectrum, peak bins, frequencies, dB;
    then the size-independent properties of the signal."""
    x = wl.c2_signal(60.0)[None, :]
    res = run_gpu(pkg, ctx, x, 4096, 256)
    ref = orc.stft_analyze(x, 4096, 256)
    F = 11235
    assert ref["frames"] == F and res["peak_count"].shape[0] == F
    compare(res, ref)
    assert (res["peak_count"] == 16).all()
    f = res["peak_freq"][:, :16]
    assert (np.diff(f, axis=1) > 0).all()  # ascending (invariant asserted at rt.resynth.lib.algo.cpp:14-21)
    assert np.abs(f - 220.0 * np.arange(1, 17)[None, :]).max() < 0.05
    # stationarity: the spectrum of a periodic signal is the same in every frame
    mag = np.sqrt(res["sqmag"])
    assert np.abs(mag - mag[0:1]).max() < 1e-4  # residual = leakage between partials (-85 dB), phase dependent
    # linearity: scaling the input by 0.5 scales |X| by 0.5
    res2 = run_gpu(pkg, ctx, 0.5 * x[:, :48000], 4096, 256, want_peaks=False)
    assert np.abs(np.sqrt(res2["sqmag"]) - 0.5 * mag[: res2["sqmag"].shape[0]]).max() < 1e-5


def test_host_entry_point(pkg, orc, ctx, wl):
    import ctypes as C
    x = wl.c4_streams(2, 0.3)
    S, n = x.shape
    an = pkg.SqMagAnalyzer(ctx, 4096, 256)
    F = an.num_frames(n)
    rows, P, B = S * F, an.max_peaks, an.bins
    sq = np.empty((rows, B), np.float32)
    pb = np.empty((rows, P), np.int32)
    pf = np.empty((rows, P), np.float32)
    pd = np.empty((rows, P), np.float32)
    pc = np.empty((rows,), np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = pkg.lib().ca_stft_analyze_host(ctx.handle, vp(x), n, S, n, vp(sq), B, vp(pb), vp(pf), vp(pd), vp(pc))
    assert rc == 0
    ref = orc.stft_analyze(x, 4096, 256)
    compare(dict(sqmag=sq, peak_bin=pb, peak_freq=pf, peak_db=pd, peak_count=pc), ref)


def test_two_analyzers_share_a_context(pkg, orc, ctx, wl):
    """A context holds one analysis plan (ca_stft_plan): analyzers that share it plan again when another one, or nothing but a
    round trip / vocoder call (they keep plans of their own), came in between."""
    x = wl.c4_streams(1, 0.5)
    a = pkg.SqMagAnalyzer(ctx, 4096, 256, max_peaks=32)
    b = pkg.SqMagAnalyzer(ctx, 2048, 512, max_peaks=8)
    xt = torch.from_numpy(x).cuda()
    ra = {k: v.cpu().numpy() for k, v in a(xt).items()}          # a's plan was replaced by b's: a plans again
    pkg.resynth_roundtrip(ctx, x, window=1024, hop=128, max_peaks=64)
    rb = {k: v.cpu().numpy() for k, v in b(xt).items()}
    compare(ra, orc.stft_analyze(x, 4096, 256, max_peaks=32))
    compare(rb, orc.stft_analyze(x, 2048, 512, max_peaks=8))
