import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """libsrst.so must exist for every test: build it (nvcc cross-compiles without a GPU) when the
    sources are newer.  On the GPU box the prebuilt library travels with the snapshot."""
    from srgan_st_b200 import build as _b
    if _b.is_stale():
        try:
            _b.build()
        except Exception as e:  # no nvcc on the box and no prebuilt library: let the tests fail loudly
            if not os.path.exists(_b.LIB):
                pytest.exit(f"libsrst.so missing and cannot be built: {e}", returncode=2)
    yield
