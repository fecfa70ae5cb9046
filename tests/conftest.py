import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gpu():
    """Initialises the CUDA library; fails loudly (never skips) when it cannot run."""
    import sonic_b200

    sonic_b200.init(0)
    return sonic_b200
