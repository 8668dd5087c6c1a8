import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def cuda_backend():
    """The product library; GPU tests fail loudly when it is missing or no device is visible."""
    from evostencils_b200 import backend
    lib = backend.load_library()
    assert lib.evo_device_count() > 0, "no CUDA device visible"
    return backend
