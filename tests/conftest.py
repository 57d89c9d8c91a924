import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """Build the CUDA library and the oracle (nvcc cross-compiles without a GPU).  make is incremental, so an up-to-date tree costs
    nothing and a stale binary can never pass for the edited sources.  Without a compiler (a box that only received the built
    files) the existing build is used."""
    import shutil

    lib = os.path.join(ROOT, "quip_b200", "libgapb200.so")
    if shutil.which("make") and (os.path.exists("/usr/local/cuda/bin/nvcc") or not os.path.exists(lib)):
        import __graft_entry__

        __graft_entry__.build()


@pytest.fixture(scope="session")
def golden():
    return GOLDEN
