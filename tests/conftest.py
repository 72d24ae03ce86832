import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure; oracle/rsrl_oracle.c)."""
    from oracle import pyoracle
    pyoracle.build()
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def oracle32():
    """oracle32: the fp32 device arithmetic compiled for the host (test infrastructure; oracle/oracle32.cpp)."""
    from oracle import pyoracle32
    pyoracle32.build()
    pyoracle32.lib()
    return pyoracle32


@pytest.fixture(scope="session")
def rsrl():
    """The product: ctypes view of librsrl_b200.so through the C ABI.  No fallback."""
    import rsrl_b200
    from rsrl_b200 import engine  # noqa: F401
    rsrl_b200.abi.load()
    return rsrl_b200
