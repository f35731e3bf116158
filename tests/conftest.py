import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: longer CPU test")


@pytest.fixture(scope="session", autouse=True)
def _native_library_is_built():
    """The tests exercise the in-tree libgfr_b200.so; build it when the checkout has none
    (the package itself never builds or falls back at run time)."""
    from grid_fed_rl_b200 import build
    if not os.path.exists(build.OUT):
        build.build()
    yield
