#!/usr/bin/env python
"""Diagnostic: is the first CUDA process on a fresh GPU box slow, and if so which part?

Alternates a plain torch device copy (1 GiB) and the bench workload's K2 sweep, printing the device
time of each with the seconds since process start.  Run as the FIRST command of a gpurun call:
    python tools/first_run_probe.py [--seconds 40] [--nvml]
--nvml additionally polls NVML from a thread exactly as bench.py's ClockSampler does.
"""
import argparse, os, sys, time
T0 = time.perf_counter()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from semiinfiniteoptimization_b200 import _abi

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=40.0)
ap.add_argument("--nvml", action="store_true")
a = ap.parse_args()


def now():
    return time.perf_counter() - T0


print("%.1f s: torch imported" % now(), flush=True)
vg, cloud, T_rel, bound = bench.build_workload(0)
print("%.1f s: workload built" % now(), flush=True)
ctx = _abi.Context(0)
g = ctx.grid_create(vg.values, vg.bmin, vg.bmax)
c = ctx.cloud_create(cloud)
B = len(T_rel)
dT = torch.from_numpy(T_rel).cuda(); db = torch.from_numpy(bound).cuda()
dd = torch.empty(B, dtype=torch.float64, device="cuda"); da = torch.empty(B, dtype=torch.int64, device="cuda")
dn = torch.empty(B, dtype=torch.int32, device="cuda")
src = torch.empty(1 << 30, dtype=torch.uint8, device="cuda"); dst = torch.empty_like(src)
torch.cuda.synchronize()
print("%.1f s: device buffers ready" % now(), flush=True)
sampler = None
if a.nvml:
    sampler = bench.ClockSampler(0)
    sampler.start()
it = 0
while now() < a.seconds + 5 and it < 400:
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    w0 = time.perf_counter()
    e0.record()
    dst.copy_(src)
    e1.record()
    torch.cuda.synchronize()
    w1 = time.perf_counter()
    ctx.min_distance_dev([g], c, dT.data_ptr(), db.data_ptr(), B, _abi.SIO_F32, dd.data_ptr(), da.data_ptr(), d_n_under_ptr=dn.data_ptr())
    ctx.sync()
    w2 = time.perf_counter()
    bulk = ctx.last_bulk_ms()
    print("%.2f s it %d: copy %.3f ms dev (%.0f GB/s) %.3f ms wall | K2 wall %.3f ms, bulk kernel %.3f ms" %
          (now(), it, e0.elapsed_time(e1), 2 * (1 << 30) / e0.elapsed_time(e1) / 1e6, (w1 - w0) * 1e3, (w2 - w1) * 1e3, bulk), flush=True)
    it += 1
    if it > 20:
        time.sleep(0.2)
if sampler:
    print(sampler.stop())
