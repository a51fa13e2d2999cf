import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def _make(target):
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), target], check=True, capture_output=True)


@pytest.fixture(scope="session")
def orc():
    """The plain-C oracle (oracle/gridpp_oracle.c); built on demand (gcc only)."""
    from oracle import bindings
    if not bindings.available("oracle"):
        _make("oracle")
    return bindings.load("oracle")


@pytest.fixture(scope="session")
def ref():
    """The compiled reference (oracle/_ref/libgridpp_ref.so); only where it was built from /root/reference."""
    from oracle import bindings
    if not bindings.available("ref"):
        if os.path.exists("/root/reference/src/api/oi.cpp"):
            _make("ref")
        else:
            pytest.skip("oracle/_ref/libgridpp_ref.so not present (the reference sources are not on this box)")
    lib = bindings.load("ref")
    lib.set_omp_threads(1)
    return lib


@pytest.fixture(scope="session")
def gpp():
    """The product: gridpp_b200 over libgridpp_b200.so. Fails (not skips) when the library is missing."""
    import gridpp_b200
    return gridpp_b200
