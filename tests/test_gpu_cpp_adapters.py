"""Builds tests/cpp/adapters_test.cpp (plain g++, links only the C-ABI library and the oracle) and runs it on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_cpp():
    exe = os.path.join(ROOT, "tests", "cpp", "adapters_test")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", os.path.join(ROOT, "tests", "cpp", "adapters_test.cpp"), "-o", exe,
           "-L" + os.path.join(ROOT, "cpp.audio_b200"), "-lcpp_audio_b200", "-L" + os.path.join(ROOT, "oracle"), "-loracle",
           "-Wl,-rpath," + os.path.join(ROOT, "cpp.audio_b200"), "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-pthread"]
    subprocess.check_call(cmd)
    return exe


def test_adapters_compile():
    """CPU: the host adapters are valid C++17 against the public header alone (no CUDA headers needed)."""
    build_cpp()


@pytest.mark.gpu
def test_adapters_match_oracle():
    exe = build_cpp()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL OK" in r.stdout
