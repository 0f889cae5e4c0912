"""Forwards to semiinfiniteoptimization_b200.geometryopt (same public names as the reference module)."""
from semiinfiniteoptimization_b200.geometryopt import *  # noqa: F401,F403
from semiinfiniteoptimization_b200 import geometryopt as _impl


def __getattr__(name):
    return getattr(_impl, name)
