import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np
from semiinfiniteoptimization_b200 import batch, workloads as W
from semiinfiniteoptimization_b200._host import se3
B = int(sys.argv[1]); k = int(sys.argv[2])
poses = np.array([se3.to_rowmajor12(T) for T in W.c5_poses(B, seed=3)])
s = batch.GroupedPoseSolver(W.c5_cube(), 0.05, W.c5_scene_cloud(), groups=k)
os.environ["SIO_LS_PROFILE"] = "2"          # warm-ups: the call's wall time only
s.solve(poses); s.solve(poses)
os.environ["SIO_LS_PROFILE"] = sys.argv[3]
t0 = time.perf_counter(); r = s.solve(poses); print("wall", time.perf_counter() - t0, r.time_qp, r.time_cons)
