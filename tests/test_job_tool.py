"""The offline job front / back end (tools/rt_resynth_job: job + preset JSON, WAV in, WAV out; reference:
source/rt.resynth.lib.params.cpp:57-386, source/rt.resynth.lib.cpp:1185-1235). CPU tests cover the host logic through
--dry-run; the GPU test checks that a job file produces exactly what ca_resynth_roundtrip_host produces."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "tools", "rt_resynth_job")


def write_wav(path, x, sr, fmt="f32"):
    x = np.atleast_2d(np.asarray(x, dtype=np.float64).T).T if np.ndim(x) == 1 else np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    ch = x.shape[1]
    if fmt == "f32":
        tag, bits, data = 3, 32, x.astype("<f4").tobytes()
    elif fmt == "i16":
        tag, bits, data = 1, 16, np.clip(np.round(x * 32768.0), -32768, 32767).astype("<i2").tobytes()
    elif fmt == "i24":
        v = np.clip(np.round(x * 8388608.0), -8388608, 8388607).astype("<i4")
        tag, bits, data = 1, 24, b"".join(struct.pack("<i", int(s))[:3] for s in v.reshape(-1))
    else:
        raise ValueError(fmt)
    hdr = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, tag, ch, sr, sr * ch * bits // 8,
                                                                               ch * bits // 8, bits) + b"data" + struct.pack("<I", len(data))
    with open(path, "wb") as f:
        f.write(hdr + data)


def read_wav_f32(path):
    b = open(path, "rb").read()
    assert b[:4] == b"RIFF" and b[8:12] == b"WAVE"
    tag, ch, sr = struct.unpack("<HHI", b[20:28])
    assert tag == 3 and b[36:40] == b"data"
    n = struct.unpack("<I", b[40:44])[0]
    return np.frombuffer(b[44:44 + n], dtype="<f4").reshape(-1, ch), sr


def preset(**floats):
    return {"bool_params": {"use_autotune": False}, "enum_params": {}, "int32_params": {}, "uint64_params": {}, "float_params": floats}


def run_tool(*args):
    return subprocess.run([TOOL, *args], capture_output=True, text=True)


@pytest.fixture(scope="module")
def tool():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tools"), "-s"])  # incremental: rebuilt when the C ABI header changed
    return TOOL


def test_dry_run_derives_the_reference_defaults(tool, tmp_path):
    sr = 48000
    t = np.arange(sr // 2) / sr
    write_wav(tmp_path / "voice.wav", np.stack([0.5 * np.sin(2 * np.pi * 440 * t), 0.25 * np.sin(2 * np.pi * 440 * t)], axis=1), sr, "i16")
    (tmp_path / "preset.json").write_text(json.dumps(preset(window_size_seconds=0.1814, window_center_stride_seconds=0.09,
                                                            min_volume=0.0001, nearby_distance_tones=0.4, max_track_pitches=1.0,
                                                            stereo_spread=1.0, env_attack_seconds=0.01, env_hold_seconds=0.0,
                                                            env_decay_seconds=0.0, env_release_seconds=0.02, env_sustain_level=1.0,
                                                            analysis_volume=0.5, pitch_shift_pre_autotune=2.0)))
    job = {"preset_file": str(tmp_path / "preset.json"), "input_voice_file": str(tmp_path / "voice.wav"), "input_carrier_file": "",
           "output_file": str(tmp_path / "out.wav"), "post": "none"}
    (tmp_path / "job.json").write_text(json.dumps(job, indent=2))
    r = run_tool("--dry-run", str(tmp_path / "job.json"))
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout)
    # getEvenWindowSizeFrames / stride of source/rt.resynth.lib.cpp:484-487, :1577-1578 at 48 kHz
    assert d["window"] == 2 * int(0.5 * np.float32(0.1814) * sr) == 8706 and d["hop"] == 4320 and d["fft_length"] == 16384
    assert d["sample_rate"] == sr and d["frames"] == sr // 2
    assert d["env"][:4] == [480, 0, 0, 960] and abs(d["gain"] - 0.5) < 1e-7
    po = d["pitch_options"]  # shift / harmonize / autotune are stages of the path (rt.resynth.lib.cpp:1687-1707)
    assert po["shift"] == [2.0, 0.0] and po["harmonize"] == [0.0, 0.0] and po["autotune"] == 0 and po["n_table"] == 0


def test_dry_run_autotune_presets(tool, tmp_path):
    """mkAutotuneFunction (rt.resynth.lib.cpp:1762-1869) as data: scale, chord (three frequency layouts), fixed intervals."""
    write_wav(tmp_path / "v.wav", 0.1 * np.ones(1000), 48000, "i16")

    def derive(enums=None, ints=None, u64=None, floats=None):
        pr = {"bool_params": {"use_autotune": True}, "enum_params": enums or {}, "int32_params": ints or {},
              "uint64_params": u64 or {}, "float_params": floats or {}}
        (tmp_path / "p.json").write_text(json.dumps(pr))
        r = run_tool("--dry-run", "--voice", str(tmp_path / "v.wav"), "--out", str(tmp_path / "o.wav"), "--preset", str(tmp_path / "p.json"))
        assert r.returncode == 0, r.stderr
        return json.loads(r.stdout)["pitch_options"]

    # defaults (:950-976): musical scale, major, root Do, tolerance 100, max pitch 150
    po = derive()
    assert po["autotune"] == 2 and po["n_table"] == 8 and po["table_first"] == 0 and po["table_last"] == 12
    assert po["scale_root_pitch"] == 69 - 9 and po["tolerance"] == 100 and po["max_pitch"] == 150
    po = derive(enums={"autotune_musical_scale_root_note": 2, "autotune_musical_scale_mode": 2}, ints={"autotune_root_note_halftones_transpose": -1})
    assert po["scale_root_pitch"] == 69 - 1 + (2 - 9)
    # chord 0b10010001 (root, major third, fifth) on Re, replicated on every octave: 11 octaves x 3 pitches around C5 + 2
    po = derive(enums={"autotune_type": 1, "autotune_chord_frequencies": 2, "autotune_musical_scale_root_note": 2},
                floats={"autotune_tolerance_pitches": 0.5})
    assert po["autotune"] == 1 and po["n_table"] == 33 and po["table_first"] == 74 - 60 and po["table_last"] == 74 + 60 + 7
    assert po["tolerance"] == 0.5
    po = derive(enums={"autotune_type": 1, "autotune_chord_frequencies": 0}, u64={"autotune_bit_chord": 5})
    assert po["n_table"] == 2 and po["table_first"] == 72 and po["table_last"] == 74
    # harmonics (the default layout): 36 harmonics, up and down, per chord note; the 36-th harmonic lies 62 semitones up
    po = derive(enums={"autotune_type": 1}, u64={"autotune_bit_chord": 1})
    assert po["n_table"] == 72 and po["table_first"] == 72 - 62 and po["table_last"] == 72 + 62
    # fixed intervals of 3 semitones through Mi + 1 = 5: 2, 5, 8 ... 149
    po = derive(enums={"autotune_type": 2, "autotune_musical_scale_root_note": 4}, ints={"autotune_factor": 3, "autotune_root_note_halftones_transpose": 1})
    assert po["autotune"] == 1 and po["table_first"] == 2 and po["table_last"] == 149 and po["n_table"] == 50


def test_dry_run_without_preset_and_error_paths(tool, tmp_path):
    write_wav(tmp_path / "v.wav", 0.1 * np.ones(1000), 44100, "i24")
    r = run_tool("--dry-run", "--voice", str(tmp_path / "v.wav"), "--out", str(tmp_path / "o.wav"))
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout)
    assert d["sample_rate"] == 44100 and d["window"] == 2 * int(0.5 * np.float32(0.1814) * 44100) and d["fft_length"] == 8192
    # bad 'post' encoding (from_string, params.cpp:170-178), missing voice file, unreadable preset
    bad = {"preset_file": "", "input_voice_file": str(tmp_path / "v.wav"), "input_carrier_file": "", "output_file": "o.wav", "post": "loud"}
    (tmp_path / "bad.json").write_text(json.dumps(bad))
    assert run_tool("--dry-run", str(tmp_path / "bad.json")).returncode != 0
    bad["post"], bad["input_voice_file"] = "none", str(tmp_path / "missing.wav")
    (tmp_path / "bad2.json").write_text(json.dumps(bad))
    assert run_tool("--dry-run", str(tmp_path / "bad2.json")).returncode != 0
    (tmp_path / "broken_preset.json").write_text("{ not json")
    bad["input_voice_file"], bad["preset_file"] = str(tmp_path / "v.wav"), str(tmp_path / "broken_preset.json")
    (tmp_path / "bad3.json").write_text(json.dumps(bad))
    r = run_tool("--dry-run", str(tmp_path / "bad3.json"))
    assert r.returncode != 0 and "Reading preset file" in r.stderr


@pytest.mark.gpu
def test_job_output_equals_the_library_call(tool, tmp_path, pkg, ctx, wl):
    sr = 48000
    x = wl.c4_streams(2, 1.0)
    for i in range(2):
        write_wav(tmp_path / f"voice{i}.wav", x[i], sr, "f32")
        (tmp_path / f"preset{i}.json").write_text(json.dumps(preset(window_size_seconds=4096 / sr + 1e-6, window_center_stride_seconds=256 / sr,
                                                                     stereo_spread=0.5, analysis_volume=0.8)))
        job = {"preset_file": str(tmp_path / f"preset{i}.json"), "input_voice_file": str(tmp_path / f"voice{i}.wav"),
               "input_carrier_file": "", "output_file": str(tmp_path / f"out{i}.wav"), "post": "none"}
        (tmp_path / f"job{i}.json").write_text(json.dumps(job))
    r = run_tool(str(tmp_path / "job0.json"), str(tmp_path / "job1.json"))   # one batch of two streams
    assert r.returncode == 0, r.stderr
    n = x.shape[1] + sr + 4096
    xp = np.zeros((2, n), np.float32)
    xp[:, :x.shape[1]] = x
    ref, _ = pkg.resynth_roundtrip(ctx, xp, window=4096, hop=256, min_db=-np.inf, max_peaks=1024, stereo_spread=0.5, gain=0.8)
    for i in range(2):
        y, sr_out = read_wav_f32(tmp_path / f"out{i}.wav")
        assert sr_out == sr and y.shape == (x.shape[1] + sr, 2)
        assert np.abs(y).max() > 0.05
        assert np.array_equal(y, ref[i, :y.shape[0]])


@pytest.mark.gpu
def test_job_with_carrier_adds_the_vocoder_mix(tool, tmp_path, pkg, ctx, wl):
    """voice = modulator, carrier file = carrier (rt.resynth.lib.cpp:1378-1420): output = resynthesis + voice_volume * voice +
    carrier_volume * carrier + vocoder_volume * vocoded, the mix going to both channels."""
    sr = 48000
    x = wl.c4_streams(1, 1.0)[0]
    car = np.random.default_rng(5).uniform(-0.4, 0.4, x.shape[0] + 5000).astype(np.float32)   # longer than the voice: truncated
    write_wav(tmp_path / "voice.wav", x, sr, "f32")
    write_wav(tmp_path / "carrier.wav", car, sr, "f32")
    p = preset(window_size_seconds=4096 / sr + 1e-6, window_center_stride_seconds=256 / sr, analysis_volume=0.5, voice_volume=0.2,
               carrier_volume=0.1, vocoder_volume=1.5, vocoder_modulator_window_size_seconds=0.04, vocoder_stride_seconds=0.01,
               vocoder_min_freq=80.0, vocoder_max_freq=6000.0)
    p["int32_params"]["vocoder_count_bands"] = 12
    (tmp_path / "preset.json").write_text(json.dumps(p))
    job = {"preset_file": str(tmp_path / "preset.json"), "input_voice_file": str(tmp_path / "voice.wav"),
           "input_carrier_file": str(tmp_path / "carrier.wav"), "output_file": str(tmp_path / "out.wav"), "post": "none"}
    (tmp_path / "job.json").write_text(json.dumps(job))
    r = run_tool(str(tmp_path / "job.json"))
    assert r.returncode == 0, r.stderr
    n = x.shape[0] + sr + 4096
    xp = np.zeros((1, n), np.float32)
    xp[0, :x.shape[0]] = x
    cp = np.zeros((1, n), np.float32)
    m = min(n, car.shape[0])
    cp[0, :m] = car[:m]
    ref, _ = pkg.resynth_roundtrip(ctx, xp, window=4096, hop=256, min_db=-np.inf, max_peaks=1024, gain=0.5)
    mix = pkg.vocoder(ctx, xp, cp, sample_rate=sr, modulator_window_seconds=0.04, stride_seconds=0.01, count_bands=12, min_freq=80.0,
                      max_freq=6000.0, vol_modulator=0.2, vol_carrier=0.1, vol_vocoder=1.5)
    y, _ = read_wav_f32(tmp_path / "out.wav")
    want = ref[0, :y.shape[0]] + mix[0, :y.shape[0], None]
    assert np.abs(mix).max() > 0.02
    assert np.array_equal(y, want.astype(np.float32))


@pytest.mark.gpu
def test_job_with_autotune_preset_equals_the_library_call(tool, tmp_path, pkg, ctx, wl):
    """use_autotune + chord (octave periodic, C major triad) + a fifth harmonised after autotune: the tool builds the same
    ca_pitch_options as pitch_options() with the allowed pitches of mkAutotuneFunction (rt.resynth.lib.cpp:1767-1800)."""
    sr = 48000
    x = wl.c4_streams(1, 1.0)[0]
    write_wav(tmp_path / "voice.wav", x, sr, "f32")
    p = preset(window_size_seconds=4096 / sr + 1e-6, window_center_stride_seconds=256 / sr, autotune_tolerance_pitches=1.5,
               pitch_harmonize_post_autotune=7.0, pitch_shift_pre_autotune=-0.25)
    p["bool_params"]["use_autotune"] = True
    p["enum_params"].update(autotune_type=1, autotune_chord_frequencies=2, autotune_musical_scale_root_note=0)
    p["int32_params"].update(autotune_max_pitch=100)
    p["uint64_params"].update(autotune_bit_chord=0b10010001)
    (tmp_path / "preset.json").write_text(json.dumps(p))
    job = {"preset_file": str(tmp_path / "preset.json"), "input_voice_file": str(tmp_path / "voice.wav"), "input_carrier_file": "",
           "output_file": str(tmp_path / "out.wav"), "post": "none"}
    (tmp_path / "job.json").write_text(json.dumps(job))
    r = run_tool(str(tmp_path / "job.json"))
    assert r.returncode == 0, r.stderr
    n = x.shape[0] + sr + 4096
    xp = np.zeros((1, n), np.float32)
    xp[0, :x.shape[0]] = x
    allowed = np.sort(np.array([72.0 + i + 12 * o for o in range(-5, 6) for i in (0, 4, 7)]))
    opts = pkg.pitch_options(shift_pre=-0.25, harmonize_post=7.0, allowed_pitches=allowed, max_pitch=100.0, tolerance=1.5)
    ref, _ = pkg.resynth_roundtrip(ctx, xp, window=4096, hop=256, min_db=-np.inf, max_peaks=1024, opts=opts)
    plain, _ = pkg.resynth_roundtrip(ctx, xp, window=4096, hop=256, min_db=-np.inf, max_peaks=1024)
    y, _ = read_wav_f32(tmp_path / "out.wav")
    assert np.abs(y).max() > 0.05
    assert np.array_equal(y, ref[0, :y.shape[0]])
    assert np.abs(ref - plain).max() > 0.05


@pytest.mark.gpu
def test_truncation_is_reported(pkg, ctx, tool, tmp_path):
    """White noise: thousands of peaks per frame. With the capacity the tool uses (n_fft/4: every local maximum fits) nothing is
    cut at the peak level; more than 256 pitches above min_volume per frame are, the library counts those frames
    (ca_resynth_truncated_frames) and the tool refuses to write a result the reference would not produce, unless it is told to."""
    rng = np.random.default_rng(9)
    x = (0.3 * rng.standard_normal(30000)).astype(np.float32)
    pkg.resynth_roundtrip(ctx, x[None, :], window=8706, hop=4320, min_db=-np.inf, max_peaks=200)
    assert pkg.truncated_frames(ctx) > 0          # 200 slots for ~2700 peaks
    pkg.resynth_roundtrip(ctx, x[None, :], window=8706, hop=4320, min_db=-np.inf, max_peaks=4096, nearby_distance_tones=0.01)
    assert pkg.truncated_frames(ctx) > 0          # every peak fits, but far more than 256 pitches are audible
    pkg.resynth_roundtrip(ctx, 1e-5 * x[None, :], window=8706, hop=4320, min_db=-np.inf, max_peaks=4096)
    assert pkg.truncated_frames(ctx) == 0         # the same peaks below min_volume: nothing reaches a capacity
    voice = tmp_path / "noise.wav"
    write_wav(str(voice), x, 48000)
    dst = tmp_path / "noise_out.wav"
    pre = tmp_path / "p.json"
    pre.write_text(json.dumps(preset(nearby_distance_tones=0.01)))
    r = subprocess.run([tool, "--voice", str(voice), "--out", str(dst), "--preset", str(pre)], capture_output=True, text=True)
    assert r.returncode == 3 and "pitches" in r.stderr and not dst.exists()
    r = subprocess.run([tool, "--allow-truncation", "--voice", str(voice), "--out", str(dst), "--preset", str(pre)], capture_output=True, text=True)
    assert r.returncode == 0 and dst.exists()
