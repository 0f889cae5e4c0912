import cProfile, pstats, sys, os
sys.path.insert(0, os.getcwd())
from tests import scenes
from semiinfiniteoptimization_b200 import sip
mod, robot, tc, con, traj, x = scenes.c4_scene('device')
con.LEVELS = 7
st = sip.SemiInfiniteOptimizationSettings(); st.max_iters, st.minimum_constraint_value = 10, -0.02
mod.optimizeCollFreeTrajectory(tc, traj, env=None, constraints=[con], settings=st, verbose=0)   # warm
pr = cProfile.Profile(); pr.enable()
mod.optimizeCollFreeTrajectory(tc, traj, env=None, constraints=[con], settings=st, verbose=0)
pr.disable()
ps = pstats.Stats(pr); ps.sort_stats('cumulative').print_stats(38)
