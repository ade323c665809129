import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def repo_root():
    return ROOT


@pytest.fixture(scope="session")
def room_stl():
    return os.path.join(ROOT, "models", "room.stl")


@pytest.fixture(scope="session")
def almost_empty_stl():
    return os.path.join(ROOT, "models", "almost_empty.stl")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the CPU oracle if missing (cheap); the CUDA library is built by __graft_entry__.build()."""
    from oracle import cpu
    cpu.lib()
