"""The C-ABI library loads (no GPU needed) and exports every symbol include/cpp_audio_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cpp_audio_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ca_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(pkg):
    assert declared_symbols() == sorted(pkg.DECLARED_SYMBOLS)


def test_library_exports_every_declared_symbol(pkg):
    assert os.path.exists(pkg.lib_path()), "build with __graft_entry__.build()"
    raw = ctypes.CDLL(pkg.lib_path())
    for name in declared_symbols():
        assert hasattr(raw, name), f"{name} is declared in the header but not exported"


def test_host_only_helpers(pkg):
    assert pkg.version().startswith("cpp.audio_b200")
    assert pkg.fft_length_for(4096) == 4096 and pkg.fft_length_for(4097) == 8192 and pkg.fft_length_for(8000) == 8192
    assert pkg.stft_num_frames(2880000, 4096, 256) == 11235  # BASELINE config 2
    assert pkg.stft_num_frames(4095, 4096, 256) == 0 and pkg.stft_num_frames(4096, 4096, 256) == 1


def test_no_cpu_fallback(pkg):
    """Without a B200 the context cannot be created: there is no CPU path to fall back to."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.CaError):
        pkg.Context(0)


def test_product_does_not_touch_oracle():
    pk = os.path.join(ROOT, "cpp.audio_b200")
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in src and "import oracle" not in src and "from oracle" not in src, f
