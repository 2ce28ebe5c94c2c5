import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_library():
    """The in-tree shared library; built on demand so a fresh checkout can run the suite."""
    import __graft_entry__
    __graft_entry__.build()
    import descent_b200
    return descent_b200


@pytest.fixture()
def host_env(built_library):
    """Host-only Environment: builds graphs and kernel source, cannot run anything."""
    env = built_library.Environment(-1)
    yield env
    env.close()


@pytest.fixture()
def env(built_library):
    """Device Environment on cuda:0.  No fallback: a missing GPU is an error for -m gpu tests."""
    e = built_library.Environment(0)
    yield e
    e.close()
