import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100a device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def libgpfo():
    """Builds (if needed) and loads the C-ABI library; no compute calls are made here."""
    from gpflowopt_b200 import build, _lib
    build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def ctx(libgpfo):
    from gpflowopt_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()
