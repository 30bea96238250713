"""CPU check of the PRODUCT's stateless pitch prefix (cpp.audio_b200/csrc/pitch.cuh — one source for the host form and the GPU
kernels) against the oracle's restatement of shift / harmonize / autotune (rt.resynth.lib.cpp:1687-1707): the harness
tests/cpp/prefix_host_test.cu is compiled by nvcc for the host only and run without a GPU."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "prefix_host_test")
MAJOR = [0., 2., 4., 5., 7., 9., 11., 12.]


@pytest.fixture(scope="module")
def harness(pkg):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    libdir = os.path.dirname(pkg.lib_path())
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O2", "--expt-relaxed-constexpr",
                           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "cpp.audio_b200", "csrc"),
                           os.path.join(ROOT, "tests", "cpp", "prefix_host_test.cu"), "-o", EXE, "-L" + libdir,
                           "-lcpp_audio_b200", "-Xlinker", "-rpath", "-Xlinker", libdir])
    return EXE


def run(exe, P, tcap, opt, table, peaks):
    txt = [f"{P} {tcap} {opt['shift_pre']!r} {opt['harm_pre']!r} {opt['shift_post']!r} {opt['harm_post']!r} {opt['autotune']} "
           f"{opt['max_pitch']!r} {opt['tol']!r} {opt['root']!r} {len(table)}"]
    txt.append(" ".join(repr(float(t)) for t in table))
    txt.append(str(len(peaks)))
    txt += [f"{b} {o!r} {d!r}" for b, o, d in peaks]
    out = subprocess.run([exe], input="\n".join(txt) + "\n", capture_output=True, text=True, check=True).stdout.split("\n")
    T = int(out[0])
    rows = np.array([[float(v) for v in line.split()] for line in out[1:1 + T]]).reshape(T, 5)
    return rows


def reduced_of(peaks, bin_hz=48000 / 4096):
    """well separated peaks: each one is its own interval (aggregate / reduce leave it alone)"""
    f = np.array([(b + float(np.float32(o))) * bin_hz for b, o, _ in peaks])
    pitch = 69.0 + 12.0 * np.log2(f / 440.0)
    vol = np.array([10.0 ** (float(np.float32(d)) / 20.0) for _, _, d in peaks])
    return pitch, vol


def oracle_opts(orc, opt, table):
    kw = dict(shift_pre=opt["shift_pre"], harmonize_pre=opt["harm_pre"], shift_post=opt["shift_post"],
              harmonize_post=opt["harm_post"], max_pitch=opt["max_pitch"], tolerance=opt["tol"], scale_root_pitch=opt["root"])
    if opt["autotune"] == 1:
        kw["allowed_pitches"] = np.asarray(table, dtype=np.float64)
    elif opt["autotune"] == 2:
        kw["scale"] = table
    return orc.make_pitch_options(**kw)


def random_case(rng):
    n = int(rng.integers(0, 14))
    bins = np.sort(rng.choice(np.arange(8, 1800), size=n, replace=False))
    # keep neighbours more than nearby_distance_tones (0.4) apart in pitch so that no aggregation happens
    keep = [b for i, b in enumerate(bins) if i == 0 or 12 * np.log2((b - 0.5) / (bins[i - 1] + 0.5)) > 0.45]
    peaks = [(int(b), float(np.float32(rng.uniform(-0.5, 0.5))), float(np.float32(rng.uniform(-60, 0)))) for b in keep]
    f32 = lambda v: float(np.float32(v))
    opt = dict(shift_pre=0.0, harm_pre=0.0, shift_post=0.0, harm_post=0.0, autotune=0, max_pitch=151.0, tol=0.0, root=0.0)
    table = []
    if rng.random() < 0.5: opt["shift_pre"] = f32(rng.choice([12.0, -12.0, 0.5, 7.0, -3.25]))
    if rng.random() < 0.6: opt["harm_pre"] = f32(rng.choice([12.0, -12.0, 7.0, 4.0, -5.0, 0.00005]))
    if rng.random() < 0.4: opt["shift_post"] = f32(rng.choice([12.0, -7.0, 0.25]))
    if rng.random() < 0.6: opt["harm_post"] = f32(rng.choice([12.0, -12.0, 3.0, 19.0]))
    mode = int(rng.integers(0, 3))
    opt["autotune"] = mode
    if mode == 1:
        step = float(rng.choice([1.0, 2.0, 5.0]))
        table = list(np.arange(float(rng.integers(0, 12)), 140.0, step)) if rng.random() < 0.9 else []
        opt["max_pitch"] = f32(rng.choice([151.0, 100.0, 72.0]))
        opt["tol"] = f32(rng.choice([0.3, 1.0, 6.0]))
    elif mode == 2:
        table = MAJOR
        opt["root"] = 57.0 + float(rng.integers(0, 12))
        opt["max_pitch"] = f32(rng.choice([151.0, 90.0]))
        opt["tol"] = f32(rng.choice([0.3, 1.0, 2.0]))
    return peaks, opt, table


def check_loudness_key(orc, rows, vol):
    """column 1: the key new notes are ranked by = float(volume / 60-phon equal-loudness contour at the pitch)
    (order_pitches_by_perceived_loudness, rt.resynth.lib.cpp:1716-1724)"""
    key = np.array([np.float32(v / orc.equal_loudness_db(p)) for p, v in zip(rows[:, 0], vol)], dtype=np.float64)
    assert np.allclose(rows[:, 1], key, rtol=1e-6, atol=0)   # a few float ulps: the contour is interpolated in float


def test_equal_loudness_contour_known_values(orc):
    # ISO 226:2003, 60 phon: 1 kHz is the reference (60 dB SPL by definition); the ear is most sensitive near 3-4 kHz
    p = lambda f: 69.0 + 12.0 * np.log2(f / 440.0)
    assert abs(orc.equal_loudness_db(p(1000.0)) - 60.0) < 0.05
    assert abs(orc.equal_loudness_db(p(100.0)) - 78.5) < 0.3      # published contour value
    assert orc.equal_loudness_db(p(3150.0)) < orc.equal_loudness_db(p(1000.0)) < orc.equal_loudness_db(p(200.0))
    assert orc.equal_loudness_db(p(5.0)) == orc.equal_loudness_db(p(20.0))            # clamped below the table
    mid = orc.equal_loudness_db(0.5 * (p(1000.0) + p(1250.0)))                        # linear in pitch between entries
    assert abs(mid - 0.5 * (orc.equal_loudness_db(p(1000.0)) + orc.equal_loudness_db(p(1250.0)))) < 1e-4


def test_defaults_take_the_fast_path_and_agree(orc, harness):
    peaks = [(37, 0.5, -10.0), (85, 0.25, -20.0), (400, -0.3, -35.5)]
    opt = dict(shift_pre=0.0, harm_pre=0.0, shift_post=0.0, harm_post=0.0, autotune=0, max_pitch=0.0, tol=0.0, root=0.0)
    rows = run(harness, 8, 8, opt, [], peaks)
    pitch, vol = reduced_of(peaks)
    assert np.allclose(rows[:, 0], pitch, atol=1e-12) and np.allclose(rows[:, 2], vol / np.float32(0.1), rtol=1e-6)
    check_loudness_key(orc, rows, vol)


def test_product_stages_match_the_oracle_on_random_lists(orc, harness):
    rng = np.random.default_rng(20261018)
    n_active = 0
    for _ in range(300):
        peaks, opt, table = random_case(rng)
        P = 64
        rows = run(harness, P, P, opt, table, peaks)
        pitch, vol = reduced_of(peaks)
        ep, ev = orc.pitch_stages(oracle_opts(orc, opt, table), pitch, vol)
        assert len(ep) <= P
        assert rows.shape[0] == len(ep), (opt, peaks)
        if len(ep):
            assert np.allclose(rows[:, 0], ep, rtol=0, atol=1e-11), (opt, peaks)
            assert np.allclose(rows[:, 2], ev / np.float32(0.1), rtol=1e-6, atol=0), (opt, peaks)   # event volume = gain * vol / 0.1
            check_loudness_key(orc, rows, ev)
            # event parameters of synthesize_sounds (rt.resynth.lib.cpp:264-382) follow from the final pitch
            freq = np.float32(440.0 * (2.0 ** (1 / 12)) ** (rows[:, 0] - 69.0))
            assert np.allclose(rows[:, 3], (np.float32(2) * freq / np.float32(48000)).astype(np.float64), rtol=1e-6)
        n_active += len(ep) != len(pitch) or not np.array_equal(ep, pitch)
    assert n_active > 150  # the options did something in most cases


def test_lists_are_capped_at_max_peaks(orc, harness):
    # 6 entries, two harmonisations would make 24: the product keeps the first tcap = 8 entries at every stage
    peaks = [(40 + 37 * i, 0.0, -10.0) for i in range(6)]
    opt = dict(shift_pre=0.0, harm_pre=0.7, shift_post=0.0, harm_post=0.9, autotune=0, max_pitch=0.0, tol=0.0, root=0.0)
    rows = run(harness, 8, 8, opt, [], peaks)
    assert rows.shape[0] == 8
    assert (np.diff(rows[:, 0]) > 0).all()
