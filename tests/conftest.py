import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "rl-agent-for-qubit-array-tuning_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def engine():
    """One Engine (qd_ctx) for the whole GPU session.  Fails loudly if libqdsim.so or the GPU is missing."""
    from qdsim import Engine
    eng = Engine(0)
    yield eng
    eng.close()
