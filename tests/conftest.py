import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """The C-ABI library must exist; building it is __graft_entry__.build()'s job."""
    from frostdb_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.load()


@pytest.fixture(scope="session")
def store(built_lib):
    """One ColumnStore (one fgpu_ctx) per test session; fails loudly without a GPU."""
    from frostdb_b200.store import ColumnStore
    cs = ColumnStore(0)
    yield cs
    cs.Close()
