import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the product library and the oracle once per session if they are missing (no GPU needed)."""
    need = [os.path.join(ROOT, "ccv_mppi_path_tracker_b200", "libmppi_b200.so"),
            os.path.join(ROOT, "oracle", "liboracle.so"), os.path.join(ROOT, "oracle", "libtwin.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()
    yield
