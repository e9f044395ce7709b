import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def O():
    """the CPU oracle (checker)"""
    o = graft.import_oracle()
    o.build()
    return o


@pytest.fixture(scope="session")
def pp():
    """the product package; importing it needs the built .so but no GPU"""
    return graft.import_package()


@pytest.fixture(scope="session")
def ctx(pp):
    if pp.device_count() == 0:
        pytest.fail("GPU test selected but no sm_100 device is visible: the CUDA path must run, there is no fallback")
    c = pp.Context(0)
    yield c
    c.close()


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(1.0, np.abs(b))
