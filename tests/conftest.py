import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_port():
    from oracle import port
    port.load()
    return port


@pytest.fixture(scope="session")
def ctx240():
    """One context for DAVIS-240C-shaped inputs, shared by the GPU tests."""
    import better_flow_b200 as bf
    c = bf.Context(180, 240, 5, max_events=1 << 21, max_slices=128, device=0)
    yield c
    c.close()
