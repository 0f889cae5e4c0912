#!/usr/bin/env python
"""Host-side profile (cProfile) of the single-problem paths: C1 optimizeCollFree (cube vs chair) and C3
optimizeCollFreeRobot (TX90 pose), device-backed classes.  Development aid: python tools/prof_c1c3.py"""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import scenes, util
from semiinfiniteoptimization_b200 import sip

mod, g1, g2, con = scenes.c1_scene('device')
poses = scenes.c1_poses(8, seed=0)
mod.optimizeCollFree(g1, g2, poses[0], verbose=0)
t0 = time.perf_counter(); its = 0
pr = cProfile.Profile(); pr.enable()
for T in poses:
    mod.optimizeCollFree(g1, g2, T, verbose=0)
    its += mod.optimizeCollFree.last_result.num_iterations
pr.disable()
print("C1: 8 problems, %.1f ms each, %d outer iterations" % ((time.perf_counter() - t0) / 8 * 1e3, its))
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)

mod, robot, cons = scenes.c3_scene('device')
qs = scenes.c3_configs(robot, 6)
st = sip.SemiInfiniteOptimizationSettings(); st.max_iters, st.minimum_constraint_value = 5, -0.02
mod.optimizeCollFreeRobot(robot, None, qdes=qs[0], constraints=cons, settings=st, verbose=0)
t0 = time.perf_counter(); its = 0
pr = cProfile.Profile(); pr.enable()
for q in qs:
    mod.optimizeCollFreeRobot(robot, None, qdes=q, constraints=cons, settings=st, verbose=0)
    its += mod.optimizeCollFreeRobot.last_result.num_iterations
pr.disable()
print("C3: 6 problems, %.1f ms each, %d outer iterations" % ((time.perf_counter() - t0) / 6 * 1e3, its))
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
