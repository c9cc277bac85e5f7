import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def window_goldens():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "window_goldens.npz"))


@pytest.fixture(scope="session")
def mednext_tiny_golden():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "mednext_tiny.npz"))


def free_port() -> int:
    """A TCP port the kernel reports as free right now (gloo rendezvous of the multi-process CPU tests)."""
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sk:
        sk.bind(("127.0.0.1", 0))
        return int(sk.getsockname()[1])
