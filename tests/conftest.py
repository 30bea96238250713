import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    return g.load_package()


@pytest.fixture(scope="session")
def orc():
    import __graft_entry__ as g
    return g.load_oracle()


@pytest.fixture(scope="session")
def wl(pkg):
    import importlib
    return importlib.import_module("cpp_audio_b200.workloads")


@pytest.fixture(scope="session")
def ctx(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    c = pkg.Context(0)
    yield c
    c.close()
