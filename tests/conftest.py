import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """ctypes handle on the CPU oracle (test infrastructure only)."""
    from tests import oracle_util
    return oracle_util.load_oracle()


@pytest.fixture(scope="session", autouse=True)
def _native_artifacts():
    """The tests need the in-tree native libraries, the CLI and the oracle: build whatever is missing or stale (nvcc
    cross-compiles without a GPU; a no-op when everything is up to date)."""
    import __graft_entry__ as ge
    ge.build()
