#!/usr/bin/env python
"""C4 (BASELINE.json configs[3]): TX90 trajectory, 67 milestones, whole-robot trajectory constraint.

    python tools/run_c4.py [--levels 8]

Times RobotTrajectoryCollisionConstraint.minvalue on the device (dense sweep of every milestone and every dyadic
time sample k/2^levels of every edge, all 6 active links, one fused FK + K2 launch) and, beside it, the reference's
milestone sweep + Lipschitz heap bisection restated on the CPU oracle (oracle/ref_geometryopt.py), plus the SIP
iterations/s of optimizeCollFreeTrajectory on both.  levels=7 is the reference's finest bisection level (parity
mode), levels=8 the 256 samples per edge BASELINE.json names.  Prints one JSON line."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
ap = argparse.ArgumentParser(); ap.add_argument("--levels", type=int, default=8); ap.add_argument("--no-cpu", action="store_true")
ap.add_argument("--shard", action="store_true", help="under torchrun: time samples in blocks across the ranks (SURVEY 8e, C4); "
                                                     "checks the agreed minimum against the unsharded dense sweep and exits")
a = ap.parse_args()
if a.shard:
    import torch, torch.distributed as tdist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        tdist.init_process_group("nccl")
from tests import scenes
from semiinfiniteoptimization_b200 import sip

out = {"config": "C4 tx90_geom_test2.xml, chair at %s, 67 milestones" % (scenes.CHAIR_IN_PATH,)}
mod, robot, tc, con, traj, x = scenes.c4_scene('device')
N = con.envs[0]._dev.ctx.cloud_size(con.envs[0]._dev.cloud_h) if hasattr(con, "envs") else None
if a.shard:
    from semiinfiniteoptimization_b200 import geometryopt as G
    res = {}
    for lv in (7, a.levels):
        con.LEVELS = lv
        G.LIPSCHITZ_SCHEDULE = False
        con.shard_time_blocks = False
        con.setx(x); ref = con.minvalue(x); con.clearx()
        con.shard_time_blocks = True
        con.setx(x); con.minvalue(x); con.clearx()              # warm-up (NCCL, scratch)
        ts = []
        for _ in range(5):
            con.setx(x)
            if world > 1: tdist.barrier()
            t0 = time.perf_counter(); got = con.minvalue(x); ts.append(time.perf_counter() - t0)
            con.clearx()
        assert got[0] == ref[0] and tuple(got[1]) == tuple(ref[1]), (got, ref)
        res["levels_%d" % lv] = {"minvalue_ms_sharded": 1e3 * min(ts), "dmin": got[0], "param": [float(v) for v in got[1]],
                                 "local_queries": con.last_sweep_queries}
    if rank == 0:
        print(json.dumps({"config": out["config"], "ranks": world, "sharded_equals_unsharded": True, **res}))
    if world > 1:
        tdist.destroy_process_group()
    sys.exit(0)
for lv in sorted({7, a.levels}):
    con.LEVELS = lv
    con.setx(x); con.minvalue(x); con.clearx()                     # warm-up
    ts = []
    for _ in range(5):
        con.setx(x)
        t0 = time.perf_counter(); mv = con.minvalue(x); ts.append(time.perf_counter() - t0)
        con.clearx()
    q = con.last_sweep_queries
    out["device_levels_%d" % lv] = {"minvalue_ms": 1e3 * min(ts), "point_queries": q, "queries_per_s": q / min(ts),
                                    "dmin": mv[0], "param": [float(v) for v in mv[1]]}
con.LEVELS = 7
st = sip.SemiInfiniteOptimizationSettings(); st.max_iters, st.minimum_constraint_value = 10, -0.02
t0 = time.perf_counter()
mod.optimizeCollFreeTrajectory(tc, traj, env=None, constraints=[con], settings=st, verbose=0)
res = mod.optimizeCollFreeTrajectory.last_result
dt = time.perf_counter() - t0
out["device_sip"] = {"iterations": res.num_iterations, "wall_s": dt, "iters_per_s": res.num_iterations / dt, "oracle_s": res.time_cons}
if not a.no_cpu:
    mod2, robot2, tc2, con2, traj2, x2 = scenes.c4_scene('restated')
    con2.setx(x2)
    t0 = time.perf_counter(); mv2 = con2.minvalue(x2); tcpu = time.perf_counter() - t0
    con2.clearx()
    out["cpu_reference_algorithm"] = {"minvalue_ms": 1e3 * tcpu, "dmin": mv2[0], "note": "milestone sweep + Lipschitz heap bisection (gopt:723-905) on the CPU oracle port"}
    from semiinfiniteoptimization_b200 import geometryopt as G
    qmin, qmax = robot2.getJointLimits()
    t0 = time.perf_counter()
    ref = sip.optimizeSemiInfinite(G.TrajectoryLengthObjective(tc2), [con2], np.asarray(x2), np.hstack([qmin] * 65), np.hstack([qmax] * 65), settings=st, verbose=0)
    dt = time.perf_counter() - t0
    out["cpu_sip"] = {"iterations": ref.num_iterations, "wall_s": dt, "iters_per_s": ref.num_iterations / dt, "oracle_s": ref.time_cons}
print(json.dumps(out))
