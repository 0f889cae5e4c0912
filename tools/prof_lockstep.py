#!/usr/bin/env python
"""Per-round profile of the device-resident lock-step C5 solve (development aid; how DESIGN section 9's per-round figures
were read off).

    python tools/prof_lockstep.py <problems> <slices> <SIO_LS_PROFILE level>
    e.g. python tools/prof_lockstep.py 8192 1 1        # one rank's shard of an 8-GPU run, every round on stderr

SIO_LS_PROFILE=1 makes sio_lockstep_pose_solve print, per round, the in-line QP launch (problems, two-row problems, ADMM
iterations, deferred problems) and every sweep (problems, culled fraction), and at the end the device time per phase;
=2 prints the call's wall time only.  Two warm-up solves come first (scratch sized, first launches done)."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from semiinfiniteoptimization_b200 import batch, workloads as W
from semiinfiniteoptimization_b200._host import se3

B, k, level = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
poses = np.array([se3.to_rowmajor12(T) for T in W.c5_poses(B, seed=3)])
solver = batch.GroupedPoseSolver(W.c5_cube(), 0.05, W.c5_scene_cloud(), groups=k)
os.environ["SIO_LS_PROFILE"] = "2"          # warm-ups: the call's wall time only
solver.solve(poses)
solver.solve(poses)
os.environ["SIO_LS_PROFILE"] = level
t0 = time.perf_counter()
res = solver.solve(poses)
print("wall %.4f s, in-line QP %.4f s, sweeps %.4f s (device time of the slower slice), %d outer iterations"
      % (time.perf_counter() - t0, res.time_qp, res.time_cons, res.outer_iterations))
