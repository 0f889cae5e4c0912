"""Latency of one ADMM iteration of the device QP kernel (csrc/sio_qp.cu), alone on the GPU and with every warp slot
busy.  A development aid: the lock-step solver's QP launches are bounded by the slowest single problem
(4000 iterations), so this number times 4000 is the floor of a launch.

  python tools/qp_latency.py [--problems 400]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semiinfiniteoptimization_b200 import _abi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--problems", type=int, default=400)
ap.add_argument("--rows", type=int, default=12)
ap.add_argument("--min-rows", type=int, default=3)
ap.add_argument("--loaded-only", type=int, default=-1, help="problem index: only the all-slots-busy launch of that problem (for ncu)")
a = ap.parse_args()

ctx = _abi.Context()
rng = np.random.default_rng(11)
B, n = a.problems, 6
W = rng.uniform(0.01, 1.0, (B, 1)) * np.ones((B, n))
xdes = rng.normal(0, 0.2, (B, n))
counts = rng.integers(max(3, a.min_rows), 2 * a.rows, B)
offs = np.concatenate([[0], np.cumsum(counts)])
rows = rng.normal(0, 1, (offs[-1], n))
depths = rng.normal(-0.1, 0.1, offs[-1])
trs = rng.uniform(0.05, 0.5, (B, 1)) * np.ones((B, n))


def one(p, reps=1):
    sl = slice(offs[p], offs[p + 1])
    o = np.concatenate([[0], np.cumsum(np.full(reps, counts[p]))])
    args = (np.repeat(W[p:p + 1], reps, 0), np.repeat(xdes[p:p + 1], reps, 0), o, np.tile(rows[sl], (reps, 1)), np.tile(depths[sl], reps),
            -np.repeat(trs[p:p + 1], reps, 0), np.repeat(trs[p:p + 1], reps, 0))
    ctx.sip_step_qp(*args, reg=1e-3)                                   # warm
    t0 = time.perf_counter()
    _, st = ctx.sip_step_qp(*args, reg=1e-3)
    dt = time.perf_counter() - t0
    tot, mx, lim = ctx.last_qp_iters(detail=True)
    return dt, tot, mx, int(st[0])


if a.loaded_only >= 0:
    print(json.dumps({"loaded": one(a.loaded_only, ctx.num_sms * 8)}))
    sys.exit(0)
recs = [one(p) for p in range(B)]
t = np.array([r[0] for r in recs]); it = np.array([r[1] for r in recs], dtype=np.float64)
A = np.stack([np.ones(B), it], 1)
(c0, c1), *_ = np.linalg.lstsq(A, t, rcond=None)
hard = int(np.argmax(it))
out = {"problems": B, "iterations_min_med_max": [float(it.min()), float(np.median(it)), float(it.max())],
       "alone": {"fixed_us": c0 * 1e6, "ns_per_admm_iteration": c1 * 1e9},
       "hardest": {"index": hard, "rows": int(counts[hard]), "iterations": float(it[hard]), "ms": t[hard] * 1e3, "status": recs[hard][3]}}
slots = ctx.num_sms * 8
for reps in (slots // 12, slots // 3, slots, 4 * slots):
    dt, tot, mx, st = one(hard, reps)
    out[f"x{reps}"] = {"ms": dt * 1e3, "ns_per_iteration_per_problem": dt * 1e9 / max(it[hard], 1), "aggregate_Miter_per_s": tot / dt / 1e6}
easy = int(np.argmin(it))
dt, tot, mx, st = one(easy, 8 * slots)
out["easiest_x%d" % (8 * slots)] = {"rows": int(counts[easy]), "iterations": float(it[easy]), "status": st, "ms": dt * 1e3,
                                    "us_per_problem_per_warp_slot": dt * 1e6 / 8}
print(json.dumps(out))
