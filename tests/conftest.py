import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _lib_built():
    from tencirchem_b200 import _lib

    return os.path.exists(_lib.LIB_PATH)


@pytest.fixture(scope="session", autouse=True)
def _ensure_library():
    """The C-ABI library is built in-tree; build it once if a fresh checkout has no .so yet."""
    if not _lib_built():
        import __graft_entry__

        __graft_entry__.build()
    yield
