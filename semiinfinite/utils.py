"""Forwards to semiinfiniteoptimization_b200.utils (same public names as the reference module)."""
from semiinfiniteoptimization_b200.utils import *  # noqa: F401,F403
from semiinfiniteoptimization_b200 import utils as _impl


def __getattr__(name):
    return getattr(_impl, name)
