import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "hostemu")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def bmpc_lib():
    """Build (if stale) and load the CUDA library; GPU tests FAIL (not skip) if it is missing."""
    from pympc_b200 import build, _lib
    build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def osqp_port_lib():
    from oracle import osqp_port
    osqp_port.build()
    return osqp_port
