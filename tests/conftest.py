import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


HOSTEMU = os.environ.get("LH_TEST_HOSTEMU") == "1"


def _load_package():
    import __graft_entry__ as graft

    lh = graft.load_package()
    if HOSTEMU and not getattr(lh, "_hostemu_patched", False):
        # TEST PROCESS ONLY (tests/test_hostemu.py starts it): the `-m gpu` tests run against the host-emulated build of the
        # product's own sources (tests/hostemu.py) instead of csrc/liblh_soil.so.  The package itself knows nothing of it.
        import hostemu

        lh.cuda_library = lambda: hostemu.library(lh)
        lh._hostemu_patched = True
    return lh


if HOSTEMU:
    _load_package()


@pytest.fixture(scope="session")
def lh():
    return _load_package()


@pytest.fixture(scope="session")
def oracle(lh):
    """The CPU oracle behind the same ctypes harness (checker only)."""
    import __graft_entry__ as graft

    return lh.SoilLibrary(graft.build_oracle(), "lho_")


@pytest.fixture(scope="session")
def cuda(lh):
    """The product library; GPU tests fail loudly if it is missing (no fallback)."""
    return lh.cuda_library()
