import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA (sm_100a) device; run with -m gpu on a B200")


@pytest.fixture(scope="session")
def built_lib():
    """Build (if stale) and load the in-tree CUDA extension."""
    from avici_b200 import _lib
    _lib.build()
    return _lib.load()
