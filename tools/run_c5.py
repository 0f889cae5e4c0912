#!/usr/bin/env python
"""C5 (BASELINE.json configs[4]): 65,536 independent cube-vs-scene pose problems, sharded across the ranks.

    python tools/run_c5.py [--problems 65536] [--sip-problems 1024] [--max-iters 10]
    python -m torch.distributed.run --nproc-per-node N ... tools/run_c5.py ...

Part 1  one oracle sweep of this rank's block of poses x the 262,144-point scene cloud, exhaustive and with
        sub-tile culling (identical results), timed on the device (max over ranks).
Part 2  the lock-step batch SIP driver on --sip-problems problems per rank: outer iterations/s, with the split
        between oracle time, QP time and the rest.
Prints one JSON line (rank 0)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--problems", type=int, default=65536)
ap.add_argument("--sip-problems", type=int, default=65536)
ap.add_argument("--max-iters", type=int, default=10)
ap.add_argument("--qp", default="resident", help="resident: sio_lockstep_pose_solve (everything on the device); device / host: array driver with "
                "sio_sip_step_qp / hostqp.c; scalar / batch: per-problem driver")
a = ap.parse_args()

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from semiinfiniteoptimization_b200 import _abi, batch, geometryopt as G, sip, workloads as W, dist as sdist
from semiinfiniteoptimization_b200._host import se3

ctx = _abi.default_context()
cube = G.PenetrationDepthGeometry(W.c5_cube(), 0.05, None)
from semiinfiniteoptimization_b200._host.geometry import Geometry3D, PointCloud
scene = G.PenetrationDepthGeometry(Geometry3D(PointCloud(W.c5_scene_cloud())), None, 0.02)
poses = W.c5_poses(a.problems, seed=3)
lo, hi = sdist.shard_range(a.problems)
mine = poses[lo:hi]
N = ctx.cloud_size(scene._dev.cloud_h)

# ---- part 1: one sweep ------------------------------------------------------------------------------------------
T_rel = np.array([se3.to_rowmajor12(se3.inv(T)) for T in mine])          # scene frame = world
dT = torch.from_numpy(T_rel).cuda(); B = len(T_rel)
bound = torch.zeros(B, dtype=torch.float64, device="cuda")
dd = torch.empty(B, dtype=torch.float64, device="cuda"); da = torch.empty(B, dtype=torch.int64, device="cuda"); dn = torch.empty(B, dtype=torch.int32, device="cuda")
stream = torch.cuda.ExternalStream(ctx.stream)
out = {}
for name, flags in (("exhaustive", _abi.SIO_F32), ("pruned", _abi.SIO_F32 | _abi.SIO_PRUNE)):
    ms = []
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.min_distance_dev([cube._dev.grid_h], scene._dev.cloud_h, dT.data_ptr(), bound.data_ptr(), B, flags, dd.data_ptr(), da.data_ptr())      # minimum / arg-minimum only, as minvalue needs
        e1.record(stream); ctx.sync(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = torch.tensor([min(ms[1:])], dtype=torch.float64, device="cuda")
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[name] = {"ms": float(t.item()), "queries_per_s": a.problems * N / (float(t.item()) * 1e-3)}
    if name == "exhaustive":
        ref = (dd.clone(), da.clone())
    else:
        out[name]["identical_to_exhaustive"] = bool(torch.equal(ref[0], dd) and torch.equal(ref[1], da))
        ev, tot = ctx.last_prune_stats(); out[name]["fraction_evaluated"] = ev / max(tot, 1)
out["in_collision_fraction"] = float((ref[0] < 0).double().mean().item())

# ---- part 2: lock-step SIP ------------------------------------------------------------------------------------------
nb = min(a.sip_problems, len(mine))
if nb == 0:
    if rank == 0:
        print(json.dumps({"config": "C5 synthetic scene %d pts, %d problems, %d GPUs" % (N, a.problems, world), **out}))
    sys.exit(0)
st = sip.SemiInfiniteOptimizationSettings(); st.max_iters = a.max_iters
t0 = time.time()
if a.qp == "resident":
    Tm = np.array([se3.to_rowmajor12(T) for T in mine[:nb]])
    st1 = sip.SemiInfiniteOptimizationSettings(); st1.max_iters = 1
    batch.optimizeCollFreeBatchDevice(cube, scene, Tm, settings=st1)              # first launches; scratch sized for the batch
    t0 = time.time()
    res = batch.optimizeCollFreeBatchDevice(cube, scene, Tm, settings=st)
elif a.qp in ("device", "host"):
    res = batch.optimizeCollFreeBatchFast(cube, scene, mine[:nb], settings=st, qp=a.qp)
else:
    res = batch.optimizeCollFreeBatch(cube, scene, mine[:nb], settings=st, qp=a.qp)
wall = time.time() - t0
if os.environ.get("SIO_C5_DUMP"):                         # development aid: the per-problem records, for A/B comparisons
    np.save(os.environ["SIO_C5_DUMP"], np.asarray(res.records()))
rec = torch.tensor([res.outer_iterations, wall, res.time_cons, res.time_qp, res.point_queries, float((res.gx > -5e-3).sum()), nb], dtype=torch.float64, device="cuda")
gathered = None
if world > 1:
    mx = rec.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(rec, op=dist.ReduceOp.SUM)
    wall = float(mx[1].item())
    # the end-of-solve exchange of SURVEY 8e: every rank's fixed-size per-problem records, in problem order
    t1 = time.time()
    allrec = sdist.gather_records(res.records())
    gathered = {"records": int(allrec.shape[0]), "s": time.time() - t1}
    assert allrec.shape == (int(rec[6].item()), 16) and int((allrec[:, 12] > -5e-3).sum()) == int(rec[5].item())
if rank == 0 and getattr(res, "qp_iter_log", None):
    print("QP log (problems, ADMM iterations, max single, at limit):", res.qp_iter_log, file=sys.stderr)
if rank == 0 and getattr(res, "iter_log", None):
    print("iteration log (it, live, rows, qp_s):", [(a_, b_, c_, round(d_, 4)) for a_, b_, c_, d_ in res.iter_log], file=sys.stderr)
if rank == 0:
    out["sip"] = {"problems": int(rec[6].item()), "max_iters": a.max_iters, "qp": a.qp, "outer_iterations": int(rec[0].item()),
                  "wall_s": wall, "gather": gathered, "solver_s": float(getattr(res, "time", 0.0) or 0.0), "iters_per_s": float(rec[0].item()) / wall, "oracle_s": float(rec[2].item()) / world,
                  "qp_s": float(rec[3].item()) / world, "point_queries": int(rec[4].item()), "feasible": int(rec[5].item()),
                  "qp_host_fallbacks": getattr(res, "qp_host_fallbacks", 0), "qp_device_s": getattr(res, "time_qp_device", 0.0), "qp_admm_iterations": getattr(res, "qp_admm_iterations", 0),
                  "slot_overflows": getattr(res, "slot_overflows", 0), "max_params": int(max(getattr(res, "n_params", [len(p) for p in res.instantiated_params]))),
                  "mean_params": float(np.mean(getattr(res, "n_params", [len(p) for p in res.instantiated_params])))}
    print(json.dumps({"config": "C5 synthetic scene %d pts, %d problems, %d GPUs" % (N, a.problems, world), **out}))
if world > 1:
    dist.destroy_process_group()
