import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # `-m gpu` tests must fail loudly, not skip, when there is no device:
    # nothing to do here.  Without a device and without `-m gpu` they are
    # deselected by the marker expression the driver passes.
    pass


@pytest.fixture(scope="session")
def small_net():
    """2-block synthetic network with random gates (fast on the CPU oracle)."""
    from dream_go_b200 import weights
    return weights.synthetic_network(seed=7, num_blocks=2, gate="random")


@pytest.fixture(scope="session")
def full_net():
    """The BASELINE configuration: 9 blocks x 128 filters."""
    from dream_go_b200 import weights
    return weights.synthetic_network(seed=20261017, num_blocks=9, gate=0.5)
