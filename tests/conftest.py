import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def built_library():
    """Build libhopedg.so (and the oracle's C port) once per session if missing."""
    from hopefoam_b200 import capi
    if not capi.LIB_PATH.exists():
        import __graft_entry__
        __graft_entry__.build()
    return capi.load_library()


@pytest.fixture(scope="session")
def gpu_ctx_factory(built_library):
    from hopefoam_b200 import capi

    def make(N):
        c = capi.Context(int(os.environ.get("HDG_TEST_DEVICE", "0")))
        c.set_order(N)
        return c
    return make
