#!/usr/bin/env python
"""The TMA-staged shared-memory variant of the K2 bulk pass (SIO_SMEM_GRID, k_min_smem) against the default brick kernel
(k_min_f32) on a grid that fits a CTA's shared memory: the C5 geometry (cube SDF@0.05, 31^3 padded samples = 119 KB) with
--poses C5 poses x the 262,144-point scene.  Checks that both return the same minima / arg-minima and prints the bulk
kernels' device times.  A short command to put under ncu (-k regex:k_min_smem).
    python tools/run_smem.py [--poses 2048] [--iters 5]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from semiinfiniteoptimization_b200 import _abi, workloads as W
from semiinfiniteoptimization_b200._host import se3
from semiinfiniteoptimization_b200._host.geometry import fixture_geometry
ap = argparse.ArgumentParser(); ap.add_argument("--poses", type=int, default=2048); ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
ctx = _abi.Context(0)
ctx.set_host_graphs(False)          # every call is timed by the events around its bulk launch
vg = fixture_geometry("cube").convert("VolumeGrid", 0.05).getVolumeGrid()
g = ctx.grid_create(vg.values, vg.bmin, vg.bmax)
cloud = W.c5_scene_cloud()
c = ctx.cloud_create(cloud)
T = np.array([se3.to_rowmajor12(se3.inv(P)) for P in W.c5_poses(a.poses, seed=3)])
out = {"poses": a.poses, "points": len(cloud), "grid_dims": list(vg.values.shape)}
res = {}
for name, fl in (("brick_k_min_f32", _abi.SIO_F32), ("smem_tma_k_min_smem", _abi.SIO_F32 | _abi.SIO_SMEM_GRID)):
    ms = []
    for _ in range(a.iters):
        d, arg, _ = ctx.min_distance(g, c, T, flags=fl, want_under=False)
        ms.append(ctx.last_bulk_ms())
    res[name] = (d, arg)
    out[name] = {"bulk_ms_min": min(ms), "bulk_ms_median": float(np.median(ms)), "queries_per_s": a.poses * len(cloud) / (min(ms) * 1e-3)}
out["identical_results"] = bool(np.array_equal(res["brick_k_min_f32"][0], res["smem_tma_k_min_smem"][0]) and
                                np.array_equal(res["brick_k_min_f32"][1], res["smem_tma_k_min_smem"][1]))
print(json.dumps(out))
