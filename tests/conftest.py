import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree CUDA library; built on demand (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    from activesetmethods_b200 import capi
    return capi.load()


@pytest.fixture(scope="session")
def gpu(built_lib):
    if built_lib.asm_device_count() < 1:
        pytest.fail("a test marked gpu ran without a CUDA device: the product path has no CPU fallback")
    return built_lib
