import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lh():
    import __graft_entry__ as graft

    return graft.load_package()


@pytest.fixture(scope="session")
def oracle(lh):
    """The CPU oracle behind the same ctypes harness (checker only)."""
    import __graft_entry__ as graft

    return lh.SoilLibrary(graft.build_oracle(), "lho_")


@pytest.fixture(scope="session")
def cuda(lh):
    """The product library; GPU tests fail loudly if it is missing (no fallback)."""
    return lh.cuda_library()
