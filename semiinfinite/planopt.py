"""Drop-in import path of the reference: `from semiinfinite import planopt` (semiinfinite/planopt.py there)."""
from semiinfiniteoptimization_b200.planopt import *  # noqa: F401,F403
from semiinfiniteoptimization_b200.planopt import arc_length_refine, planOptimizedTrajectory, planToConfig  # noqa: F401
