import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def sim_backend():
    """Backend over the CPU kernel simulator (tests/hostsim) -- kernel logic without a GPU."""
    from tests.hostsim import NumpyBackend
    return NumpyBackend()


@pytest.fixture(scope="session")
def gpu_backend():
    from temfpy_b200.engine import TorchBackend
    return TorchBackend("cuda:0")
