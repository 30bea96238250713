"""bench.py's reference arm and the JSON contract, on CPU: `--impl reference` times the oracle port on the host cores (the one place
besides tests/ and smoke() that may execute oracle/), prints exactly one JSON line whose `config` is the object the GPU arm prints,
and ranks other than 0 exit without work."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench_module():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _run(extra_env=None):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                          capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)


def test_reference_arm_prints_one_line_with_the_shared_config():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    bench = _bench_module()
    assert d["impl"] == "reference" and d["metric"] == "stft_frames_per_s" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["config"] == bench.workload_config(1, 32)  # the same object as the GPU arm's (`same_config` in the driver's comparison)
    assert d["value"] > 0 and abs(d["e2e"]["value"] - d["value"]) < 1e-9 * d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "streams" in cb["sample"]


def test_reference_arm_other_ranks_exit_without_output():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ""


def test_traffic_constants_follow_the_committed_ncu_summaries():
    """profiles/traffic.json (the ncu constants bench.py quotes in roofline.traffic / roofline.ncu) is exactly what
    profiles/make_traffic.py derives from the committed ncu summaries."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    before = open(path).read()
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "make_traffic.py")], capture_output=True, text=True, timeout=120, cwd=ROOT)
        assert r.returncode == 0, r.stderr[-2000:]
        after = open(path).read()
    finally:
        with open(path, "w") as f:
            f.write(before)
    assert json.loads(after) == json.loads(before)
