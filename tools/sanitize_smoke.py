#!/usr/bin/env python
"""One small call of every kernel family, meant to run under `compute-sanitizer --tool memcheck`
(one tool per gpurun call, B200_PROFILING.md).  Results are checked against the oracle so that a silent
out-of-bounds read would also show up as a mismatch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from semiinfiniteoptimization_b200 import _abi
from semiinfiniteoptimization_b200._host import qp
from oracle import oracle as O
from tests import util, scenes

ctx = _abi.Context(0)
rng = np.random.default_rng(0)
vg = util.random_grid(1)
g = ctx.grid_create(vg.values, vg.bmin, vg.bmax)
G = O.Grid(vg.values, vg.bmin, vg.bmax)
cloud = rng.uniform(-0.6, 1.6, (3001, 3)).astype(np.float32)
c = ctx.cloud_create(cloud)
T = util.random_poses(2, 37, rot=0.6, trans=0.3)
pts = rng.uniform(-0.6, 1.6, (50, 3))
for fl in (_abi.SIO_F32, _abi.SIO_F64):
    d, gr = ctx.query(g, T[0], pts, flags=fl)
    d0, g0 = O.query(G, T[0], pts)
    assert np.abs(d - d0).max() < 1e-5
bound = np.where(np.arange(len(T)) % 2 == 0, np.inf, -0.2)
d0, a0 = O.min_distance(G, cloud, T, bound=bound)
for fl in (_abi.SIO_F32, _abi.SIO_F64, _abi.SIO_F32 | _abi.SIO_PRUNE):
    dm, am, nu = ctx.min_distance(g, c, T, bound=bound, flags=fl)
    assert np.array_equal(am, a0) and np.abs(dm - d0).max() < 1e-12
    ctx.under_fetch()
    dm, am, _ = ctx.min_distance(g, c, T, bound=bound, flags=fl, want_under=False)
    assert np.array_equal(am, a0)
vgs = [util.sphere_grid(r=0.3 + 0.1 * k, n=20 + 4 * k, half=0.8) for k in range(3)]
gh = [ctx.grid_create(v.values, v.bmin, v.bmax) for v in vgs]
gob = rng.integers(0, 3, len(T)).astype(np.int32)
dm, am, _ = ctx.min_distance(gh, c, T, grid_of_b=gob, flags=_abi.SIO_F32 | _abi.SIO_PRUNE)
d1, a1 = O.min_distance([O.Grid(v.values, v.bmin, v.bmax) for v in vgs], cloud, T, grid_of_b=gob)
assert np.array_equal(am, a1)
ctx.pose_rows(g, T[:3], [0, 5, 5, 12], pts[:12])
# robot kernels through the classes (FK, link sweep, chain rows)
mod, robot, cons = scenes.c3_scene('device')
q = scenes.c3_configs(robot, 1)[0]
for k in cons:
    k.setx(q)
mv = cons[3].minvalue(q)
cons[3].df_dx(q, mv[1])
for k in cons:
    k.clearx()
# QP kernel
B, n = 300, 6
W = rng.uniform(0.01, 1.0, (B, 1)) * np.ones((B, n)); xdes = rng.normal(0, 0.2, (B, n))
counts = rng.integers(0, 20, B); counts[0] = 58; counts[1] = 70
offs = np.concatenate([[0], np.cumsum(counts)])
rows = rng.normal(0, 1, (offs[-1], n)); depths = rng.normal(-0.05, 0.1, offs[-1])
trs = rng.uniform(0.05, 0.5, (B, 1)) * np.ones((B, n))
dd, sd = ctx.sip_step_qp(W, xdes, offs, rows, depths, -trs, trs, reg=1e-3)
dh, sh = qp.sip_step_batch(W, xdes, offs, rows, depths, -trs, trs, reg=1e-3)
same = (sd == sh) & (sh != 2)
assert same.sum() > 200 and np.abs(dd[same] - dh[same]).max() < 1e-6
print("sanitize smoke ok")
