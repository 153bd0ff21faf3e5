import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def native():
    from fibsem_optflow_b200 import _native as N
    N.lib()
    return N


@pytest.fixture(scope="session")
def gpu(native):
    if native.device_count() <= 0:
        pytest.fail("gpu-marked test run without a CUDA device: the product has no CPU fallback")
    return native
