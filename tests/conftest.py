import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Make sure the native pieces exist (no-op when they are already built)."""
    import __graft_entry__
    __graft_entry__.build()
    return True


@pytest.fixture(scope="session")
def ctx(built):
    from semiinfiniteoptimization_b200 import _abi
    return _abi.Context(0)
