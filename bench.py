#!/usr/bin/env python
"""bench.py -- headline benchmark of the collision-oracle hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "C2"): the m1062 tree mesh as a signed distance grid at
gridres 0.01 (68x98x105 cell-centred float32 samples) is the moving body; the static cloud is the
declared synthetic 2^20-point upsample of bunny.pcd (point j = 4*(b[j mod 397] -
centroid) + centre(bbox(m1062)) + N(0, 0.005^2 I), numpy default_rng(1)); 64 seeded poses per GPU
(w ~ U(-0.3,0.3)^3, t ~ U(-0.1,0.1)^3).

A STEP is one oracle sweep: the K2 call (transform + trilinear SDF query of every cloud point for
every pose, minimum / arg-minimum / under-bound set with the float64 near-tie re-check) over
64 x 2^20 = 67,108,864 point-queries per GPU.  With N > 1 every rank owns its own 64 poses (weak
scaling; geometry replicated); sweeps need no collective, and the compact per-pose results reach every rank once per
reporting interval (the K timed steps) as peer stores over NVLink issued by the interval's last sweep (result boards,
include/sio_abi.h), inside that step's timed span; an NCCL all-gather of the same records verifies them after the timing.

`value`  : point-queries/s, inputs resident in HBM, timed with CUDA events on the launching stream,
           max over ranks; L2 is flushed between timed steps.
`e2e`    : the same sweep through the host-pointer C-ABI call a user of the Python API makes
           (poses + bounds H2D, results D2H inside the timed region, wall clock), issued through a prebound call
           (Context.min_distance_plan); the plain wrapper's figure is reported beside it.
`roofline`: the dominant kernel k_min_f32 against the unit that binds it, the L1 data pipe (32 B brick per query
           delivered to registers / 148 SMs x 128 B/clk x the sampled SM clock), with ncu's own utilisation of that pipe
           from the committed capture; the contract's ALGORITHMIC 48 B/query of SURVEY.md 8d against the measured HBM
           copy peak and the measured DRAM fraction are kept under `roofline.hbm`.
`modes`  : the per-point output modes of the same sweep (value only, value + gradient) with their HBM fractions.
`cpu_baseline`: the CPU oracle (oracle/, float64 brute force, all host threads) on a bounded
           sample of the same sweep, rank 0 / N=1 only.

`sip`     : BASELINE.json's second metric, SIP outer iterations/s, on configs[4] ("C5"): 65,536 independent
           cube-vs-scene optimizeCollFree problems (declared synthetic scene, 262,144 points) sharded across the
           ranks and solved in lock step entirely on the device (batch.GroupedPoseSolver: every shard as two concurrent
           slices of sio_lockstep_pose_solve), wall clock from host poses to host results -- at N > 1 including the
           end-of-solve all-gather of the per-problem records --, with the sweep / in-line QP / other split, every
           rank's solve / gather times and (`weak_scaling`) the same solve with 65,536 problems on EVERY GPU; at
           N=1 also the CPU path (restated SIP loop on the CPU oracle, one process per host core, bounded sample).
`single_problems`: C1 / C3 latency per problem, C4 minvalue and SIP iterations/s, API call latency, GPU beside CPU.

--impl reference times that CPU oracle port as the reference arm (the reference itself is
Python 2 on absent Klampt/OSQP and cannot run).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

POSES_PER_GPU = 64
CLOUD_N = 1 << 20
GRIDRES = 0.01
BYTES_PER_QUERY = 48          # 16 B float4 point + 8 x 4 B corners
GATHER_BYTES_PER_QUERY = 32
SIP_PROBLEMS = 65536
SIP_GROUPS = 2            # concurrent slices per GPU in the C5 leg (tools/run_c5_groups.py: 1 -> 2 slices = -7 % wall at 65,536 problems, -9 % at 8,192)
METRIC = "sdf_oracle_point_queries_per_s"
UNIT = "queries/s"


def build_workload(rank=0):
    from semiinfiniteoptimization_b200 import workloads
    from semiinfiniteoptimization_b200._host.geometry import fixture_geometry
    vg = fixture_geometry("m1062").convert("VolumeGrid", GRIDRES).getVolumeGrid()
    cloud, centre = workloads.c2_throughput_cloud(CLOUD_N, seed=1)
    # poses of the moving body (world <- grid) rotate about the body centre so the pose set stays in contact with the
    # cloud; cloud frame = world, T_rel = pose^-1
    T_rel = workloads.c2_throughput_poses(POSES_PER_GPU, centre, seed=1000 + rank)
    bound = np.zeros(POSES_PER_GPU)       # SIP's discovery bound min(0, min depths) (sip.py:215-219)
    return vg, cloud, T_rel, bound


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  Uses NVML (the library
    nvidia-smi itself queries -- same counters as the B200_PROFILING.md clocks line) from a thread,
    because an nvidia-smi subprocess polling every 100 ms was seen to stall CUDA calls."""
    REASONS = {0x08: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x04: "sw_power_cap"}

    def __init__(self, device, period_s=0.001):
        self.device, self.period = device, period_s
        self.sm, self.reasons, self.smax, self.power = [], set(), None, []
        self._stop = threading.Event()
        self.th = None
        self.err = None

    def _nvml_index(self, pynvml):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.device < len(ids) and ids[self.device].isdigit():
                return int(ids[self.device])
        return self.device

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._nvml_index(pynvml))
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:
            self.err = "nvml unavailable: %s" % e
            return
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def _sample(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for bit, name in self.REASONS.items():
            if r & bit:
                self.reasons.add(name)
        try:
            self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception as e:
                self.err = str(e)
                return
            time.sleep(self.period)

    def stop(self):
        if self.th is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "not sampled"]}
        self._stop.set()
        self.th.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.smax,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "power_w_max": max(self.power) if self.power else None}


def cpu_pruned(L, O, G, cloud, T_rel, bound, budget_s):
    """The hierarchy-pruned CPU arm ("pruning on"): k-d leaves of 128 points, best-first branch and bound on
    a Lipschitz lower bound (oracle/sio_oracle.c orc_min_distance_pruned) -- the host counterpart of what Klampt's
    distance_ext does and of the device's SIO_PRUNE.  Same answers as the brute-force sweep (tests)."""
    t0 = time.perf_counter()
    tree = O.CloudTree(cloud, L=L)
    lip = O.grid_lipschitz(G, L)
    build_s = time.perf_counter() - t0
    threads = L.orc_max_threads()
    npose = max(1, min(len(T_rel), threads))
    O.min_distance_pruned(G, tree, T_rel[:npose], bound=bound[:npose], use_bound=True, lip=lip, L=L)
    done, spent, t0, k = 0, 0, time.perf_counter(), 0
    while True:
        sel = [(k * npose + i) % len(T_rel) for i in range(npose)]
        _, _, ev = O.min_distance_pruned(G, tree, T_rel[sel], bound=bound[sel], use_bound=True, lip=lip, L=L)
        done += npose * len(cloud)
        spent += int(ev.sum())
        k += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s:
            break
    return {"effective_queries_per_s": done / dt, "minvalue_calls_per_s": done / dt / len(cloud), "cores": int(threads),
            "fraction_evaluated": spent / done, "hierarchy_build_s": build_s,
            "sample": "%d sweeps of %d poses x %d points, best-first over 128-point k-d leaves, float64, %.1f s; effective = "
                      "(pose, point) pairs resolved / time" % (k, npose, len(cloud), dt)}


def cpu_baseline(vg, cloud, T_rel, budget_s=12.0, native=True, bound=None):
    """CPU oracle port on a bounded sample: float64 brute force ("pruning off") and, beside it, the hierarchy-pruned arm
    ("pruning on"), all host threads."""
    from oracle import oracle as O
    L = None
    if native:
        try:
            O.build(native=True)          # tune for the GPU box's host CPU when gcc is there
            L = O.lib(prefer_native=True)
        except Exception:
            L = None
    L = L or O.lib()
    L.orc_set_threads(0)                  # every host core, whatever OMP_NUM_THREADS says
    G = O.Grid(vg.values, vg.bmin, vg.bmax)
    threads = L.orc_max_threads()
    npose = max(1, min(len(T_rel), threads))
    # one pose per thread over the whole cloud, repeated over successive pose groups until ~budget_s
    O.min_distance(G, cloud[:1 << 14], T_rel[:npose], L=L)        # warm the pages / threads
    done, t0, k = 0, time.perf_counter(), 0
    while True:
        sel = [(k * npose + i) % len(T_rel) for i in range(npose)]
        O.min_distance(G, cloud, T_rel[sel], L=L)
        done += npose * len(cloud)
        k += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s:
            break
    out = {"value": done / dt, "unit": UNIT, "cores": int(threads), "kind": "port", "pruning": "off",
           "minvalue_calls_per_s": done / dt / len(cloud),          # one call = one pose against the whole cloud
           "sample": "%d sweeps of %d poses x %d cloud points of the same workload, float64 brute force "
                     "(no hierarchy pruning), %.1f s" % (k, npose, len(cloud), dt)}
    out["pruned"] = cpu_pruned(L, O, G, cloud, T_rel, np.zeros(len(T_rel)) if bound is None else bound, budget_s / 2)
    return out


def run_reference(args):
    """Reference arm: the CPU oracle port timed on the host cores, same config/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    try:
        O.build(native=True)
        L = O.lib(prefer_native=True)
    except Exception:
        L = O.lib()
    L.orc_set_threads(0)                  # torchrun sets OMP_NUM_THREADS=1 for its workers: use every host core
    vg, cloud, T_rel, bound = build_workload(0)
    G = O.Grid(vg.values, vg.bmin, vg.bmax)
    threads = L.orc_max_threads()
    npose = max(1, min(POSES_PER_GPU, threads))
    # per step: npose poses x a slice of the cloud sized for ~2 s
    n0 = 1 << 15
    O.min_distance(G, cloud[:n0], T_rel[:npose], L=L)                 # warm the threads and pages before sizing the step
    t0 = time.perf_counter()
    O.min_distance(G, cloud[:n0], T_rel[:npose], L=L)
    rate = npose * n0 / (time.perf_counter() - t0)
    npts = int(min(CLOUD_N, max(n0, rate * 2.0 / npose)))
    for _ in range(args.warmup):
        O.min_distance(G, cloud[:npts], T_rel[:npose], L=L)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.min_distance(G, cloud[:npts], T_rel[:npose], L=L)
    dt = time.perf_counter() - t0
    val = args.steps * npose * npts / dt
    sample = "%d poses x %d cloud points per step (bounded sample of the 64 x 2^20 sweep), float64 brute force" % (npose, npts)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": int(threads), "kind": "port", "pruning": "off", "sample": sample,
                             "pruned": cpu_pruned(L, O, G, cloud, T_rel, bound, 6.0)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference is Python 2 on Klampt/OSQP (absent, un-vendored): its path cannot run; this arm is the CPU "
                    "oracle port (oracle/sio_oracle.c) on the host cores.  `value` = point queries actually evaluated per second by "
                    "the brute-force sweep (like for like with the GPU arm's exhaustive sweep); cpu_baseline.pruned = the same "
                    "sweep with hierarchy pruning, as (pose, point) pairs RESOLVED per second (like for like with the GPU arm's "
                    "`pruned` key)"}
    if not args.no_sip:
        line["sip"] = cpu_sip_baseline()
    print(json.dumps(line))


def _cpu_sip_worker(args):
    """one host process: scalar SIP (sip.py restated) on the CPU oracle for its share of C5 problems, until the budget."""
    first, step, budget = args
    import warnings
    warnings.simplefilter("ignore")
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import ref_geometryopt as R
    from semiinfiniteoptimization_b200 import geometryopt as G
    from semiinfiniteoptimization_b200 import sip, workloads
    from semiinfiniteoptimization_b200._host.geometry import Geometry3D, PointCloud
    cube = R.RefPDG(workloads.c5_cube(), 0.05, None)
    scene = R.RefPDG(Geometry3D(PointCloud(workloads.c5_scene_cloud())), None, 0.02)
    con = R.RefObjectCollisionConstraint(cube, scene)
    poses = workloads.c5_poses(SIP_PROBLEMS, seed=3)
    iters, solved, t0 = 0, 0, time.perf_counter()
    k = first
    while time.perf_counter() - t0 < budget and k < len(poses):
        res = sip.optimizeSemiInfinite(G.ObjectPoseObjective(poses[k]), [con], poses[k], verbose=0)
        iters += res.num_iterations
        solved += 1
        k += step
    return iters, solved, time.perf_counter() - t0


def cpu_sip_baseline(budget_s=8.0):
    import multiprocessing as mp
    n = os.cpu_count() or 1
    # spawn: the parent has live OpenMP (and, in the GPU arm, CUDA) state that must not be forked
    with mp.get_context("spawn").Pool(n) as pool:
        out = pool.map(_cpu_sip_worker, [(i, n, budget_s) for i in range(n)])
    iters = sum(o[0] for o in out)
    wall = max(o[2] for o in out)
    return {"value": iters / wall, "unit": "SIP outer iterations/s", "cores": n, "kind": "port",
            "sample": "%d of the 65,536 C5 problems (every %d-th), restated sip.py loop + CPU oracle port (float64 brute-force "
                      "minvalue over 262,144 points) + host QP, one process per core, %.1f s" % (sum(o[1] for o in out), n, wall)}


def sip_leg(ctx, rank, world, want_cpu):
    """C5 in lock step on this rank's shard; returns the aggregated record on rank 0."""
    import torch
    import torch.distributed as dist
    from semiinfiniteoptimization_b200 import batch, geometryopt as G, sip, workloads
    from semiinfiniteoptimization_b200 import dist as sdist
    from semiinfiniteoptimization_b200._host.geometry import Geometry3D, PointCloud
    # two concurrent slices of the shard (one context + host thread each, same GPU): one slice's QP launches overlap the
    # other's sweeps and bookkeeping launches; same answers problem by problem (tests/test_gpu_batch.py)
    # (every slice has a host thread that polls its stream: on a box with fewer host cores than 2 x ranks the slices would
    # fight over them, so such a box solves each shard in one piece)
    try:
        host_cores = len(os.sched_getaffinity(0))
    except AttributeError:
        host_cores = os.cpu_count() or 1
    groups = SIP_GROUPS if host_cores >= SIP_GROUPS * world else 1
    solver = batch.GroupedPoseSolver(workloads.c5_cube(), 0.05, workloads.c5_scene_cloud(), pcres=0.02, groups=groups, context=ctx)
    poses = workloads.c5_poses(SIP_PROBLEMS, seed=3)
    lo, hi = sdist.shard_range(SIP_PROBLEMS)
    from semiinfiniteoptimization_b200._host import se3
    poses = np.array([se3.to_rowmajor12(T) for T in poses[lo:hi]])            # this rank's poses as a (B, 12) row-major array
    lo, hi = 0, len(poses)
    warm = sip.SemiInfiniteOptimizationSettings()
    warm.max_iters = 1
    wres = solver.solve(poses, settings=warm)                                  # warm-up: first launches, scratch sized for the shard
    if world > 1:
        sdist.gather_records(wres.records(), n_total=SIP_PROBLEMS)   # ... and NCCL's first all-gather of this size (it sets up buffers: 0.2 s once)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    res = solver.solve(poses[lo:hi])                                           # host poses in, host results out
    solve_s = time.perf_counter() - t0
    t_rec = t_gat = 0.0
    if world > 1:                                                              # the path's one exchange: every rank's
        t1 = time.perf_counter()
        recs = res.records()
        t_rec = time.perf_counter() - t1
        allrec = sdist.gather_records(recs, n_total=SIP_PROBLEMS, copy=False)  # per-problem records, once, at the end of the solve
        t_gat = time.perf_counter() - t1 - t_rec
        assert allrec.shape == (SIP_PROBLEMS, 16)
    wall = time.perf_counter() - t0
    wall_local = wall
    per_rank = None
    if world > 1:                                                              # after the timed region: who waited for whom
        mine = torch.tensor([solve_s, t_rec, t_gat], dtype=torch.float64, device="cuda")
        every = torch.empty((world, 3), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(every, mine)
        per_rank = every.cpu().numpy()
    rec = torch.tensor([res.outer_iterations, wall, res.time_cons, res.time_qp, res.point_queries, float((res.gx > -5e-3).sum()),
                        float((np.asarray(res.gx0) < 0).sum())], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = rec.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(rec, op=dist.ReduceOp.SUM)
        wall = float(mx[1].item())
    weak = None
    if world > 1:
        # the same solver with the shard held FIXED (65,536 problems on every GPU, rank r's poses from seed 3 + r): what the
        # strong-scaling figure above loses is per-round latency on an 8,192-problem shard, not anything between the GPUs
        wposes = np.array([se3.to_rowmajor12(T) for T in workloads.c5_poses(SIP_PROBLEMS, seed=3 + rank)])
        solver.solve(wposes, settings=warm)                                     # scratch re-sized for the larger shard, untimed
        dist.barrier()
        t0 = time.perf_counter()
        wres = solver.solve(wposes)
        wrec = torch.tensor([wres.outer_iterations, time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        wmx = wrec.clone()
        dist.all_reduce(wmx, op=dist.ReduceOp.MAX)
        dist.all_reduce(wrec, op=dist.ReduceOp.SUM)
        weak = {"value": float(wrec[0].item()) / float(wmx[1].item()), "unit": "SIP outer iterations/s", "scaling": "weak",
                "problems_per_gpu": SIP_PROBLEMS, "outer_iterations": int(wrec[0].item()), "wall_s_max_over_ranks": float(wmx[1].item()),
                "note": "65,536 problems per GPU (seed 3 + rank), no exchange; host poses in, host results out"}
    if rank != 0:
        return None
    out = {"value": float(rec[0].item()) / wall, "unit": "SIP outer iterations/s", "scaling": "strong", "weak_scaling": weak,
           "workload": "C5: %d independent optimizeCollFree problems (cube SDF@0.05 vs synthetic 262,144-point scene, seed 2; poses "
                       "seed 3), reference default settings (max_iters 50), sharded over %d GPU(s), device-resident lock-step "
                       "solver (sio_lockstep_pose_solve), each shard solved as %d concurrent slices (batch.GroupedPoseSolver); wall clock from host poses (row-major 3x4 array) to host results, the end-of-solve "
                       "all-gather of the per-problem records included at N > 1" % (SIP_PROBLEMS, world, groups),
           "outer_iterations": int(rec[0].item()), "wall_s": wall, "sweep_device_s_per_rank": float(rec[2].item()) / world,
           "qp_device_s_per_rank": float(rec[3].item()) / world, "other_s_per_rank": wall - float(rec[2].item() + rec[3].item()) / world,
           "slices_per_gpu": groups, "host_cores": host_cores,
           "device_s_note": "sweep / qp device seconds are those of the slower slice; with more than one slice per GPU the slices' "
                            "launches overlap in time, so sweep + qp + other no longer partition the wall clock exactly",
           "library_call_s_rank0": float(res.time), "gather_s_rank0": wall_local - solve_s,
           "solve_s_per_rank": None if per_rank is None else [round(float(v), 6) for v in per_rank[:, 0]],
           "records_s_per_rank": None if per_rank is None else [round(float(v), 6) for v in per_rank[:, 1]],
           "gather_incl_wait_s_per_rank": None if per_rank is None else [round(float(v), 6) for v in per_rank[:, 2]],
           "limiting_term": "in-line QP launches (k_sip_qp): a launch lasts as long as its slowest problem under the 800-iteration "
                            "budget (0.34 ms) plus ~(problems per rank / 2,368 resident warps) x 160 us of ADMM + polish per problem (16 warps share "
                            "an SM and its L1 data pipe; profiles/r02_qp_phases.md, r02_qp_smem_operands_full.txt), so it shrinks "
                            "less than linearly with the shard: a rank of an 8-GPU run spends ~56 rounds x 0.9 ms of mostly latency "
                            "(QP 0.5, two sweeps 0.3, ~14 bookkeeping launches and 2 counter read-backs 0.1); two concurrent slices per GPU "
                            "overlap one slice's QP with the other's sweeps",
           "oracle_point_queries_resolved": int(rec[4].item()), "started_in_collision": int(rec[6].item()),
           "ended_feasible(g>-5e-3)": int(rec[5].item())}
    if want_cpu:
        out["cpu_baseline"] = cpu_sip_baseline()
    return out


REF_NAMES = dict(PDG='RefPDG', Obj='RefObjectCollisionConstraint', Kin='RefRobotKinematicsCache',
                 Link='RefRobotLinkCollisionConstraint', TC='RefRobotTrajectoryCache', Traj='RefRobotTrajectoryCollisionConstraint')


def _med(v):
    return float(np.median(v))


def single_problem_legs(family, names, device, n_c1=8, n_c3=6, c4_iters=4, reps=3):
    """BASELINE.json configs[0], [2], [3] -- single problems, where LATENCY counts -- and the API-level call latency of
    PenetrationDepthGeometry.distance(other, bound) with one pose per call.  Run once with the device-backed classes
    (family = this package's geometryopt) and once with the CPU restatement on the CPU oracle (oracle/ref_geometryopt.py):
    the same host SIP loop, the same host QP, only the oracle underneath differs."""
    from semiinfiniteoptimization_b200 import geometryopt as G, sip, workloads as W
    out = {}
    # C1: cube vs chair, optimizeCollFree defaults
    g1, g2, con = W.c1_scene(family, names)
    poses = W.c1_poses(n_c1, seed=0)

    def run_c1(T):
        return sip.optimizeSemiInfinite(G.ObjectPoseObjective(T), [con], T, settings=sip.SemiInfiniteOptimizationSettings(), verbose=0)
    run_c1(poses[0])
    ts, its, tcons = [], 0, 0.0
    for T in poses:
        t0 = time.perf_counter()
        res = run_c1(T)
        ts.append(time.perf_counter() - t0)
        its += res.num_iterations
        tcons += res.time_cons
    out["c1"] = {"ms_per_problem": 1e3 * float(np.mean(ts)), "ms_per_problem_median": 1e3 * _med(ts), "problems": n_c1,
                 "outer_iterations": int(its), "sip_iters_per_s": its / sum(ts), "oracle_fraction": tcons / sum(ts)}
    # API call latency, one pose per call (what ObjectCollisionConstraint.minvalue issues, gopt:409-412)
    g1.setTransform(poses[0])
    for _ in range(3):
        g2.distance(g1, 0.0)
    lat = []
    for k in range(60 if device else 20):
        g1.setTransform(poses[k % len(poses)])
        t0 = time.perf_counter()
        g2.distance(g1, 0.0)
        lat.append(time.perf_counter() - t0)
    npts = len(g2.pcdata.points) if hasattr(g2.pcdata, "points") else g2.pcdata.numPoints()
    out["api_call"] = {"call": "PenetrationDepthGeometry.distance(other, bound), one pose per call, C1 geometry (%d-point cloud vs "
                               "cube SDF@0.05)" % npts, "us_per_call_median": 1e6 * _med(lat), "us_per_call_min": 1e6 * float(np.min(lat))}
    pt = [0.3, 0.1, 0.2]
    lat = []
    for k in range(60 if device else 20):
        t0 = time.perf_counter()
        g1.distance(pt)
        lat.append(time.perf_counter() - t0)
    out["api_call"]["point_query_us_median"] = 1e6 * _med(lat)
    # C3: TX90 pose, 7 link constraints, max_iters 5, minimum_constraint_value -0.02 (robotposeopt.py:101-104)
    robot, cons = W.c3_scene(family, names)
    qs = W.c3_configs(robot, n_c3)
    qmin, qmax = robot.getJointLimits()
    st = sip.SemiInfiniteOptimizationSettings()
    st.max_iters, st.minimum_constraint_value = 5, -0.02

    def run_c3(q):
        return sip.optimizeSemiInfinite(G.RobotConfigObjective(robot, q), cons, [0.0] * 7, np.asarray(qmin), np.asarray(qmax), settings=st, verbose=0)
    run_c3(qs[0])
    ts, its = [], 0
    for q in qs:
        t0 = time.perf_counter()
        res = run_c3(q)
        ts.append(time.perf_counter() - t0)
        its += res.num_iterations
    out["c3"] = {"ms_per_problem": 1e3 * float(np.mean(ts)), "ms_per_problem_median": 1e3 * _med(ts), "problems": n_c3,
                 "outer_iterations": int(its), "sip_iters_per_s": its / sum(ts)}
    # C4: trajectory constraint minvalue + SIP iterations/s
    robot, tc, con4, traj, x = W.c4_scene(family, names)
    c4 = {}
    levels = (7, 8) if device else (7,)                     # the reference's bisection stops at k/128; 256 per edge is the benchmark's
    for lv in levels:
        if device:
            con4.LEVELS = lv
        con4.setx(x)
        con4.minvalue(x)
        con4.clearx()
        ts = []
        for _ in range(reps):
            con4.setx(x)
            t0 = time.perf_counter()
            mv = con4.minvalue(x)
            ts.append(time.perf_counter() - t0)
            con4.clearx()
        key = "minvalue_ms_256_per_edge" if lv == 8 else "minvalue_ms_k128"
        c4[key] = 1e3 * float(np.min(ts))
        c4["dmin_%d" % lv] = float(mv[0])
        if device:
            c4["point_queries_%d" % lv] = int(con4.last_sweep_queries)
    if device:
        con4.LEVELS = 7
    st4 = sip.SemiInfiniteOptimizationSettings()
    st4.max_iters, st4.minimum_constraint_value = c4_iters, -0.02
    xmin, xmax = np.hstack([qmin] * 65), np.hstack([qmax] * 65)
    best = None
    for _ in range(3 if device else 1):                 # the device leg is 25 ms: the first run also pays one-off host costs
        t0 = time.perf_counter()                        # (sparse-solver set-up, first touches); the fastest of three is reported
        res = sip.optimizeSemiInfinite(G.TrajectoryLengthObjective(tc), [con4], np.asarray(x), xmin, xmax, settings=st4, verbose=0)
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, res)
    dt, res = best
    c4.update({"sip_iters_per_s": res.num_iterations / dt, "sip_iterations": int(res.num_iterations), "sip_wall_s": dt,
               "sip_oracle_fraction": res.time_cons / dt, "sip_runs": 3 if device else 1})
    out["c4"] = c4
    return out


def cpu_single_problem_legs():
    """The CPU side of single_problem_legs: oracle/ref_geometryopt.py (the reference's per-call algorithms on the CPU oracle
    port), one core -- the reference itself is single-threaded Python around Klampt."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import oracle as O, ref_geometryopt as R
    O.lib().orc_set_threads(1)
    try:
        R.RefPDG.PRUNED = True                          # "pruning on": the stronger CPU arm for single problems
        try:
            pruned = single_problem_legs(R, REF_NAMES, device=False, n_c1=4, n_c3=4, c4_iters=2, reps=1)
        finally:
            R.RefPDG.PRUNED = False
        out = single_problem_legs(R, REF_NAMES, device=False, n_c1=4, n_c3=4, c4_iters=2, reps=1)
        out["pruned"] = pruned
        # C4 at the benchmark's density on the CPU: brute dense sweep, every host thread (the reference algorithm above
        # bisects down to k/128 only)
        O.lib().orc_set_threads(0)
        out["cores"] = 1
        out["kind"] = "port"
        return out
    finally:
        O.lib().orc_set_threads(0)


def workload_config(ngpu):
    return {"workload": "C2 bunny-vs-m1062 fine grid: m1062 SDF gridres 0.01 (68x98x105 f32), synthetic 2^20-point bunny "
                        "upsample (seed 1), 64 seeded poses per GPU, one K2 min/arg-min/under-bound sweep per step",
            "poses_per_gpu": POSES_PER_GPU, "cloud_points": CLOUD_N, "queries_per_step_per_gpu": POSES_PER_GPU * CLOUD_N,
            "sharding": "poses across ranks (geometry replicated), no collective inside a sweep; the per-pose results reach every "
                        "rank once per reporting interval (= the K timed steps) as peer stores over NVLink issued by the last sweep "
                        "(inside its timed step)" if ngpu > 1 else "single GPU",
            "l2": "flushed between timed steps (256 MiB write)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-sip", action="store_true", help="skip the C5 SIP iterations/s leg")
    ap.add_argument("--no-single", action="store_true", help="skip the single-problem legs (C1, C3, C4, API call latency)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from semiinfiniteoptimization_b200 import _abi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the collision oracle has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    vg, cloud, T_rel, bound = build_workload(rank)
    ctx = _abi.Context(local)
    g = ctx.grid_create(vg.values, vg.bmin, vg.bmax)
    c = ctx.cloud_create(cloud)
    B = POSES_PER_GPU
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    dT = torch.from_numpy(T_rel).cuda()
    dbound = torch.from_numpy(bound).cuda()
    # the step's outputs land in ONE packed per-rank record [dmin f64 x B | argmin i64 x B | n_under i32 x B (+pad)]: the
    # exchange is then a single all_gather_into_tensor of that buffer into a preallocated one -- no cat, no casts
    packed = torch.zeros(3 * B, dtype=torch.float64, device="cuda")
    ddmin = packed[:B]
    darg = packed[B:2 * B].view(torch.int64)
    dnu = packed[2 * B:].view(torch.int32)[:B]
    gathered = torch.empty(world * 3 * B, dtype=torch.float64, device="cuda") if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()

    # The path's only exchange: the compact per-pose results (20 B per pose) of every rank, once per reporting
    # interval (= the K timed steps).  It is done WITHOUT a collective launch: the interval's last sweep is issued with
    # SIO_PUSH_BOARDS and its tail stores this rank's records into every rank's result board over NVLink (peer memory
    # mapped through CUDA IPC, dist.ResultBoards); the barrier that ends the timed region is all the synchronisation it
    # needs.  An NCCL all-gather of the same records runs AFTER the timed region, only to verify the boards.
    boards = None
    if world > 1:
        from semiinfiniteoptimization_b200 import dist as sdist
        try:
            boards = sdist.ResultBoards(ctx, B)
        except Exception as e:                         # no peer access / IPC on this box: fall back to the NCCL exchange
            sys.stderr.write("rank %d: result boards unavailable (%s); using all_gather_into_tensor\n" % (rank, e))
            boards = None
        okf = torch.tensor([1.0 if boards is not None else 0.0], device="cuda")
        dist.all_reduce(okf, op=dist.ReduceOp.MIN)
        if okf.item() != 1.0 and boards is not None:
            dist.barrier()                             # the ranks that did map this board are past their last store (none yet)
            boards.close(collective=False)
            boards = None
        elif okf.item() != 1.0:
            dist.barrier()

    def step(push=False):
        ctx.min_distance_dev([g], c, dT.data_ptr(), dbound.data_ptr(), B, _abi.SIO_F32 | (_abi.SIO_PUSH_BOARDS if push else 0),
                             ddmin.data_ptr(), darg.data_ptr(), d_n_under_ptr=dnu.data_ptr())

    def gather():
        with torch.cuda.stream(stream):
            dist.all_gather_into_tensor(gathered, packed)

    for _ in range(args.warmup):
        step()
    if world > 1:
        gather()
    ctx.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # the bulk kernel's own launch time (roofline), from the events the ABI puts around it: a few plain steps
    bulk_ms = []
    for _ in range(8):
        with torch.cuda.stream(stream):
            flush.zero_()
        step()
        ctx.sync()
        bulk_ms.append(ctx.last_bulk_ms())
    # a step is a fixed launch sequence (2 memsets, bulk, finalize): capture it once into a CUDA graph and replay it,
    # so that host launch jitter (8 ranks share the box's cores) stays out of the device-side timing
    graph = torch.cuda.CUDAGraph()
    launches_per_step = ctx.kernel_launches
    with torch.cuda.graph(graph, stream=stream):
        step()
    launches_per_step = ctx.kernel_launches - launches_per_step
    graph_push = None
    if boards is not None:                              # the interval's last step: the same sweep + the peer stores
        graph_push = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph_push, stream=stream):
            step(push=True)
    with torch.cuda.stream(stream):
        graph.replay()
        if graph_push is not None:
            graph_push.replay()
    ctx.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()                       # evict L2 (126 MB) between timed steps
            ev[k][0].record(stream)
            (graph_push if (graph_push is not None and k == args.steps - 1) else graph).replay()
            ev[k][1].record(stream)
    gather_ms = 0.0
    if world > 1 and boards is None:                    # fallback exchange: one NCCL all-gather of the packed records, timed
        ctx.sync()
        torch.cuda.synchronize()
        dist.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        gather()
        g1.record(stream)
        g1.synchronize()
        gather_ms = g0.elapsed_time(g1)
    ctx.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()                                  # every rank's peer stores have landed on every board
    t_wall = time.perf_counter() - t_wall0
    exchange = None
    if world > 1 and boards is None:
        exchange = {"how": "NCCL all_gather_into_tensor of the packed per-rank records, once per interval (result boards unavailable)",
                    "ms": gather_ms}
    if boards is not None:
        bd, ba, bn = boards.read()
        gather()                                        # verification only, outside the timed region
        torch.cuda.synchronize()
        ref = gathered.cpu().numpy().reshape(world, 3 * B)
        ok = (np.array_equal(bd, ref[:, :B]) and np.array_equal(ba, ref[:, B:2 * B].view(np.int64))
              and np.array_equal(bn, ref[:, 2 * B:].view(np.int32)[:, :B]))
        flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        exchange = {"how": "peer stores over NVLink into every rank's result board (CUDA IPC mapped), issued by the interval's "
                           "last sweep (SIO_PUSH_BOARDS); no collective launch in the timed region",
                    "bytes_per_rank_pair": 20 * B, "verified_against_nccl_all_gather_on_every_rank": bool(flag.item() == 1.0)}
        if not ok:
            raise SystemExit("result boards differ from the NCCL all-gather of the same records")
    clocks = sampler.stop()
    launches = launches_per_step * args.steps
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(ms_steps)) + gather_ms          # K sweeps, the last one carrying the exchange (or + the fallback gather)
    per_rank = None
    if world > 1:
        mine = torch.tensor([float(sum(ms_steps)), float(ms_steps[-1])], dtype=torch.float64, device="cuda")
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)                       # diagnostics: which rank is the maximum, and why
        per_rank = [[round(float(v[0]), 3), round(float(v[1]), 3)] for v in allr]
        allc = [None] * world
        dist.all_gather_object(allc, {"sm_mhz": clocks.get("sm_mhz"), "reasons": clocks.get("reasons"), "power_w_max": clocks.get("power_w_max")})
        clocks = dict(clocks, per_rank=allc)               # every rank samples its own GPU during the timed region
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    queries_per_step = B * CLOUD_N * world
    value = queries_per_step * args.steps / (total_ms * 1e-3)

    # end to end through the host-pointer call (what PenetrationDepthGeometry.distance(other) issues)
    # ... through a prebound call (Context.min_distance_plan: the same sio_min_distance entry point with host pointers, the
    # interpreter's per-call glue done once) and, for comparison, through the plain wrapper
    for _ in range(2):
        ctx.min_distance(g, c, T_rel, bound=bound)
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.min_distance(g, c, T_rel, bound=bound)
    plain_s = time.perf_counter() - t0
    plan = ctx.min_distance_plan(g, c, B)
    for _ in range(2):
        plan(T_rel, bound)
    ctx.sync()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dmin_h, arg_h, nu_h = plan(T_rel, bound)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    n_under = int(nu_h.sum())
    e2e = {"value": queries_per_step * args.steps / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": int(B * (96 + 8)), "d2h_bytes_per_step": int(B * (8 + 8 + 4)),
           "plain_wrapper_value_this_rank": B * CLOUD_N * args.steps / plain_s,
           "call": "Context.min_distance_plan(...)(T_rel, bound) -> sio_min_distance (host pointers): poses+bounds H2D from a pinned block, dmin/argmin/n_under D2H; geometry "
                   "resident as in the reference API (PenetrationDepthGeometry builds grid+cloud once); the sorted under-bound "
                   "list itself stays on the device until sio_under_fetch asks for it"}

    # the same sweep with sub-tile culling (SIO_PRUNE): identical results, most (pose, sub-tile) pairs skipped.
    # Reported separately -- `value` above is the exhaustive sweep.
    pruned = None
    if rank == 0:
        PR = _abi.SIO_F32 | _abi.SIO_PRUNE
        d_ref = ddmin.clone()
        a_ref = darg.clone()
        for _ in range(3):
            ctx.min_distance_dev([g], c, dT.data_ptr(), dbound.data_ptr(), B, PR, ddmin.data_ptr(), darg.data_ptr(),
                                 d_n_under_ptr=dnu.data_ptr())
        ctx.sync()
        evp = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(args.steps, 50))]
        for a, b in evp:
            with torch.cuda.stream(stream):
                flush.zero_()
            a.record(stream)
            ctx.min_distance_dev([g], c, dT.data_ptr(), dbound.data_ptr(), B, PR, ddmin.data_ptr(), darg.data_ptr(),
                                 d_n_under_ptr=dnu.data_ptr())
            b.record(stream)
        ctx.sync()
        torch.cuda.synchronize()
        ms_p = float(np.mean([a.elapsed_time(b) for a, b in evp]))
        ev_sub, tot_sub = ctx.last_prune_stats()
        pruned = {"effective_queries_per_s": B * CLOUD_N / (ms_p * 1e-3), "ms_per_step": ms_p,
                  "subtiles_evaluated": ev_sub, "subtiles_total": tot_sub, "fraction_evaluated": ev_sub / max(tot_sub, 1),
                  "identical_to_exhaustive": bool(torch.equal(d_ref, ddmin) and torch.equal(a_ref, darg)),
                  "note": "SIO_PRUNE: Lipschitz lower bound per 128-point sub-tile (the GPU analogue of Klampt's bounding-hierarchy "
                          "search inside distance_ext); effective = all (pose, point) pairs resolved / time"}

    # SURVEY.md 8d also asks for the per-point output modes of the same sweep (value-only, value+gradient); they write
    # 4 B / 16 B per query to HBM, so an HBM fraction means something for them (the gradient mode reaches 0.77 of the copy
    # peak) -- what binds them is still the L1 data pipe (ncu: profiles/r02_k1_modes_full.txt)
    modes = None
    if rank == 0:
        N_ = CLOUD_N
        outg = torch.empty((B, N_, 4), dtype=torch.float32, device="cuda")
        modes = {}
        peak_hbm = 6650.0
        try:
            peak_hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
        except Exception:
            pass
        for name, ptrs, wbytes in (("value_only", (outg.data_ptr(), None), 4), ("value_and_gradient", (None, outg.data_ptr()), 16)):
            ms_m = []
            for _ in range(13):
                with torch.cuda.stream(stream):
                    flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                ctx.cloud_query_dev(g, c, dT.data_ptr(), B, _abi.SIO_F32, *ptrs)
                e1.record(stream)
                e1.synchronize()
                ms_m.append(e0.elapsed_time(e1))
            t_m = float(np.median(ms_m[3:]))
            gbs = B * N_ * (wbytes + 16.0 / 32) / (t_m * 1e-3) / 1e9          # written bytes + the cloud read once per 32-pose chunk
            modes[name] = {"queries_per_s": B * N_ / (t_m * 1e-3), "ms": t_m,
                           "hbm": {"achieved": gbs, "peak": peak_hbm, "unit": "GB/s", "frac": gbs / peak_hbm,
                                   "bytes_per_query": wbytes + 0.5},
                           # ncu (profiles/r02_k1_modes_full.txt): like the K2 bulk kernel, both modes keep the L1 data pipe 75 %
                           # busy (brick loads + the stores themselves); DRAM is 19 % / 55 % busy
                           "bound": "l1_data_pipe"}
        del outg

    sip_rec = None
    if not args.no_sip:
        sip_rec = sip_leg(ctx, rank, world, want_cpu=(world == 1 and not args.no_cpu))

    # single problems (C1, C3, C4) and the API-level call latency, device-backed classes; the CPU port beside them
    single = None
    c4_sharded = None
    if not args.no_single:
        from semiinfiniteoptimization_b200 import geometryopt as Gmod, workloads as Wmod
        if world > 1:
            # C4's time samples in contiguous blocks across the ranks: every rank sweeps its block, the global
            # deepest (link, env, t, point) is agreed on by one small gather
            robot4, tc4, con4, traj4, x4 = Wmod.c4_scene(Gmod)
            c4_sharded = {}
            for lv in (7, 8):
                con4.LEVELS = lv
                Gmod.LIPSCHITZ_SCHEDULE = False
                con4.shard_time_blocks = False
                con4.setx(x4); ref4 = con4.minvalue(x4); con4.clearx()
                con4.shard_time_blocks = True
                con4.setx(x4); con4.minvalue(x4); con4.clearx()
                ts = []
                for _ in range(5):
                    con4.setx(x4)
                    dist.barrier()
                    t0 = time.perf_counter()
                    got4 = con4.minvalue(x4)
                    ts.append(time.perf_counter() - t0)
                    con4.clearx()
                t = torch.tensor([min(ts)], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                c4_sharded["levels_%d" % lv] = {"minvalue_ms_sharded": 1e3 * float(t.item()), "ranks": world,
                                                "equals_unsharded": bool(got4[0] == ref4[0] and tuple(got4[1]) == tuple(ref4[1])),
                                                "local_point_queries": int(con4.last_sweep_queries)}
            Gmod.LIPSCHITZ_SCHEDULE = True
            con4.shard_time_blocks = False
        if rank == 0:
            single = {"gpu": single_problem_legs(Gmod, Wmod.DEVICE_NAMES, device=True)}
            if c4_sharded:
                single["gpu"]["c4"]["sharded_time_blocks"] = c4_sharded
            if world == 1 and not args.no_cpu:
                single["cpu"] = cpu_single_problem_legs()

    if rank == 0:
        bulk = float(np.mean(bulk_ms))
        q1 = B * CLOUD_N
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = q1 * BYTES_PER_QUERY / (bulk * 1e-3) / 1e9
        gather_peak = ctx.bench_gather(int(np.prod(vg.dims)), 1 << 26, iters=5)
        gathered = q1 * GATHER_BYTES_PER_QUERY / (bulk * 1e-3) / 1e9
        n_sm = torch.cuda.get_device_properties(local).multi_processor_count
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        l1_peak = n_sm * 128 * sm_mhz * 1e6 / 1e9
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
            traffic, traffic_src = tr["k_min_f32"]["dram_bytes_per_launch"], tr["k_min_f32"]["source"]
        except Exception:
            pass
        # The unit that binds k_min_f32 is the L1 data pipe (ncu, profiles/): every query pulls one 32 B brick into registers
        # and a warp-wide 256-bit load is 8 data-stage wavefronts of 128 B.  HBM does not bind -- the cloud is read once per
        # transform chunk and the bricks are cache resident (DRAM traffic per launch = `traffic`) -- so the contract's 48 B/query
        # HBM figure is kept as a side key, not as the fraction.
        roofline = {"bound": "l1_data_pipe", "kernel": "k_min_f32", "achieved": gathered, "peak": l1_peak, "unit": "GB/s",
                    "frac": gathered / l1_peak, "traffic": traffic, "traffic_source": traffic_src,
                    "bytes_per_query": GATHER_BYTES_PER_QUERY, "kernel_ms": bulk,
                    "peak_source": "%d SMs x 128 B/clk x %.0f MHz: L1TEX data-stage width x the SM clock sampled during the timed region "
                                   "(no L1 peak is in MEASURED_PEAKS.json); achieved = the 32 B brick per query delivered to registers"
                                   % (n_sm, sm_mhz),
                    "hbm": {"achieved_algorithmic": achieved, "peak": hbm_peak, "unit": "GB/s", "bytes_per_query": BYTES_PER_QUERY,
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst)" if peaks else "fallback 6650 GB/s",
                            "dram_fraction_measured": (traffic / (bulk * 1e-3) / 1e9 / hbm_peak) if traffic else None,
                            "note": "SURVEY 8d's algorithmic 48 B/query x queries / kernel time exceeds the HBM peak because those bytes "
                                    "are served by L1/L2, not DRAM: `dram_fraction_measured` = ncu DRAM bytes per launch / time / peak"},
                    "l2_gather_peak_gbs": gather_peak,
                    # what ncu's own counter says of the same pipe, all of its traffic included (bricks + cloud tiles + the
                    # transform records read from shared memory + lane groups that straddle cache lines): a committed capture
                    # of this kernel on this workload, not a live measurement
                    "pipe_busy_ncu": {"l1tex__throughput_pct_of_peak": 75.4, "binding_sub_unit": "l1tex__data_pipe_lsu_wavefronts", "issue_active_pct": 46.6,
                                      "source": "profiles/r02_ffma2_kmin_full.txt, profiles/r02_kmin_evidence.md (ncu --set full, 64 poses x 2^20 points)"}}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "ms_per_step_median": float(np.median(ms_steps)), "exchange": exchange,
                "minvalue_calls_per_s": value / CLOUD_N,            # one call = one pose against the whole cloud
                "per_rank_ms[steps_sum, last_step_with_push]": per_rank,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(world), "clocks": clocks, "e2e": e2e,
                "gpu_launches": int(launches), "roofline": roofline, "wall_ms_per_step_incl_flush": t_wall / args.steps * 1e3,
                "check": {"dmin_first": float(dmin_h[0]), "argmin_first": int(arg_h[0]), "n_under_total": n_under},
                "pruned": pruned, "modes": modes, "sip": sip_rec, "single_problems": single}
        if world == 1 and not args.no_cpu:
            cb = cpu_baseline(vg, cloud, T_rel, bound=bound)
            line["cpu_baseline"] = cb
            # like for like, all four measured in this run on this box
            line["speedup"] = {"exhaustive_e2e_over_cpu_brute": e2e["value"] / cb["value"],
                               "exhaustive_device_over_cpu_brute": value / cb["value"],
                               "pruned_device_over_cpu_pruned": (pruned["effective_queries_per_s"] / cb["pruned"]["effective_queries_per_s"])
                               if pruned else None,
                               "exhaustive_e2e_over_cpu_pruned": e2e["value"] / cb["pruned"]["effective_queries_per_s"]}
        print(json.dumps(line))
    if boards is not None:
        boards.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
