import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the product library and the oracle if they are missing (both are git-ignored)."""
    import __graft_entry__ as g

    lib = os.path.join(ROOT, "fast_kinematic_simulator_b200", "libfksgpu.so")
    orc = os.path.join(ROOT, "oracle", "libfks_oracle.so")
    if not (os.path.exists(lib) and os.path.exists(orc)):
        g.build()
    yield
