#!/usr/bin/env python
"""Does the lock-step C5 solve gain from running k independent groups of problems concurrently (one context, stream
and host thread per group)?  On a small shard (8,192 problems = one rank of an 8-GPU run) every round is bound by the
latency of its slowest QP and of ~14 small launches, which leaves most of the GPU idle; a second group fills it.

    python tools/run_c5_groups.py [--problems 8192 65536] [--groups 1 2 4]
Prints one JSON line per (problems, groups): wall seconds (best of 3) and whether the poses equal the single-group run's."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("--problems", type=int, nargs="+", default=[8192, 65536])
ap.add_argument("--groups", type=int, nargs="+", default=[1, 2, 4])
a = ap.parse_args()

from semiinfiniteoptimization_b200 import batch, workloads as W
from semiinfiniteoptimization_b200._host import se3

poses_all = np.array([se3.to_rowmajor12(T) for T in W.c5_poses(max(a.problems), seed=3)])
for B in a.problems:
    # the shard rank 0 of a (65,536 / B)-GPU run would own
    poses = poses_all[:B]
    ref = None
    for k in a.groups:
        solver = batch.GroupedPoseSolver(W.c5_cube(), 0.05, W.c5_scene_cloud(), groups=k)
        best = None
        for rep in range(4):
            t0 = time.perf_counter()
            res = solver.solve(poses)
            dt = time.perf_counter() - t0
            if rep > 0:
                best = dt if best is None else min(best, dt)
        if ref is None:
            ref = res
        same = bool(np.array_equal(res.T, ref.T) and np.array_equal(res.num_iterations, ref.num_iterations)
                    and np.array_equal(res.status_codes, ref.status_codes))
        print(json.dumps({"problems": B, "groups": k, "wall_s": best, "outer_iterations": int(res.outer_iterations),
                          "iters_per_s": res.outer_iterations / best, "same_as_one_group": same,
                          "qp_device_s": res.time_qp, "sweep_device_s": res.time_cons}), flush=True)
        solver.close()
