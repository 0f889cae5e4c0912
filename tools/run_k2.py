#!/usr/bin/env python
"""Run the bench workload's K2 call a few times (optionally pruned) -- a short command to put under ncu.
    python tools/run_k2.py [--prune] [--iters 5]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from semiinfiniteoptimization_b200 import _abi
ap = argparse.ArgumentParser(); ap.add_argument("--prune", action="store_true"); ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--nobound", action="store_true")
a = ap.parse_args()
vg, cloud, T_rel, bound = bench.build_workload(0)
ctx = _abi.Context(0); g = ctx.grid_create(vg.values, vg.bmin, vg.bmax); c = ctx.cloud_create(cloud)
B = len(T_rel); dT = torch.from_numpy(T_rel).cuda(); db = torch.from_numpy(bound).cuda()
dd = torch.empty(B, dtype=torch.float64, device="cuda"); da = torch.empty(B, dtype=torch.int64, device="cuda"); dn = torch.empty(B, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
fl = _abi.SIO_F32 | (_abi.SIO_PRUNE if a.prune else 0)
for _ in range(a.iters):
    ctx.min_distance_dev([g], c, dT.data_ptr(), None if a.nobound else db.data_ptr(), B, fl, dd.data_ptr(), da.data_ptr(), d_n_under_ptr=dn.data_ptr())
ctx.sync()
print("ok", float(dd[0]), int(da[0]), ctx.last_prune_stats() if a.prune else "")
