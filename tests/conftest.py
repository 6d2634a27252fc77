import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def engine_lib():
    """Build (if stale) and load the C-ABI library."""
    from theano_pyglm_b200 import build as _build
    _build.build()
    from theano_pyglm_b200 import engine
    return engine.load_library()
