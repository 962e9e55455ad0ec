import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure; the only place besides bench.py's CPU legs that loads it)."""
    import orc as _orc
    _orc.build()
    return _orc


@pytest.fixture(scope="session")
def capi():
    from misc3d_b200 import capi as _capi
    _capi.lib()  # raises if the product library has not been built: no fallback
    return _capi


@pytest.fixture(scope="session")
def ctx(capi):
    """A GPU context.  On a box without a GPU this raises (gpu tests must not silently pass)."""
    c = capi.Context(0)
    yield c
    c.close()
