#!/usr/bin/env python
"""Generates tests/golden/*.json by running the REFERENCE's own Python (sip.py, geometryopt.py,
constraint.py, objective.py from /root/reference, loaded by oracle/refrun with stand-in klampt / osqp
modules) on BASELINE.json's configurations.  The Klampt-side queries underneath are the CPU oracle
(a restatement: "parity unpinned" for that layer); every line of SIP / constraint logic above them is
the reference's.  Needs /root/reference, so it runs in the build container only; the fixtures it writes
are committed and are what the tests on the GPU box compare against.

    python tools/make_golden.py [--only c1,c2,c3,c4,c5,classes,oracle,planopt,paths,multi,c1targets,c3random,c4long]

Generators are deterministic: re-running one rewrites its file byte for byte (tests/test_golden_cpu.py re-runs a case of
each against the committed file whenever the reference sources are present).
"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")

from tests import scenes  # noqa: E402


def _f(v):
    """JSON-able copy: numpy scalars/arrays -> python floats/lists (repr round-trips float64 exactly)."""
    if isinstance(v, (list, tuple)):
        return [_f(x) for x in v]
    if isinstance(v, np.ndarray):
        return [_f(x) for x in v.tolist()]
    if isinstance(v, (np.floating, float)):
        return float(v)
    if isinstance(v, (np.integer, int)):
        return int(v)
    return v


def _quiet(fn, *a, **k):
    """The reference prints progress unconditionally in a few places."""
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def write(name, obj):
    os.makedirs(scenes.GOLDEN, exist_ok=True)
    fn = os.path.join(scenes.GOLDEN, name)
    obj = dict(obj)
    obj["_generated_by"] = "tools/make_golden.py: the reference package (/root/reference/semiinfinite) executed via oracle/refrun"
    json.dump(obj, open(fn, "w"))
    print("wrote %s (%.1f kB)" % (fn, os.path.getsize(fn) / 1e3))


def pose_cases(scene_fn, poses, settings_kw=None):
    mod, g1, g2, con = scene_fn('reference')
    sip = sys.modules[mod.__name__.rsplit('.', 1)[0] + '.sip']
    cases = []
    for T in poses:
        st = sip.SemiInfiniteOptimizationSettings()
        for k, v in (settings_kw or {}).items():
            setattr(st, k, v)
        out = _quiet(mod.optimizeCollFree, g1, g2, T, settings=st, verbose=0)
        cases.append({"Tinit": _f(T), "T": _f(out[0]), "trace": _f(out[1]), "params": _f(out[3])})
    return cases


def gen_c1():
    write("c1_pose.json", {"config": "C1 cube SDF@0.05 vs m797 cloud@0.02, optimizeCollFree defaults (gopt:1039-1081)",
                           "cases": pose_cases(scenes.c1_scene, scenes.c1_poses(8, seed=0))})


def gen_c2():
    write("c2_pose.json", {"config": "C2 as shipped: m1062 SDF@0.01 vs bunny.pcd (397 points), optimizeCollFree, max_iters 20",
                           "cases": pose_cases(scenes.c2_scene, scenes.c2_poses(6, seed=0), {"max_iters": 20})})


def gen_c3():
    mod, robot, cons = _quiet(scenes.c3_scene, 'reference')
    sip = sys.modules[mod.__name__.rsplit('.', 1)[0] + '.sip']
    cases = []
    for qdes in scenes.c3_configs(robot, 6, seed=3):
        st = sip.SemiInfiniteOptimizationSettings()
        st.max_iters, st.minimum_constraint_value = 5, -0.02                 # robotposeopt.py:101-104
        q, trace, times, params = _quiet(mod.optimizeCollFreeRobot, robot, None, qdes=qdes, qinit=[0.0] * 7, constraints=cons,
                                         settings=st, verbose=0)
        cases.append({"qdes": _f(qdes), "q": _f(q), "trace": _f(trace), "params": _f(params)})
    write("c3_robot.json", {"config": "C3 tx90_geom_test3.xml (synthetic link boxes), 7 link constraints, max_iters 5, "
                                      "minimum_constraint_value -0.02 (robotposeopt.py:101-104)", "cases": cases})


def gen_c4():
    mod, robot, tc, con, traj, x = _quiet(scenes.c4_scene, 'reference')
    sip = sys.modules[mod.__name__.rsplit('.', 1)[0] + '.sip']
    rng = np.random.default_rng(0)
    probes = []
    for trial in range(3):
        xa = [float(v) for v in (np.asarray(x) + (rng.normal(0, 0.05, len(x)) if trial else 0))]
        con.setx(xa)
        mv = _quiet(con.minvalue, xa)
        y = mv[1]
        row = con.df_dx(xa, y).tocoo()
        probes.append({"x": xa, "dmin": float(mv[0]), "param": _f(y), "value_at_param": float(con.value(xa, y)),
                       "row_cols": _f(row.col), "row_vals": _f(row.data),
                       "pruned_by_bound": _quiet(con.minvalue, xa, mv[0] - 1e-6) == []})
        con.clearx()
    st = sip.SemiInfiniteOptimizationSettings()
    st.max_iters, st.minimum_constraint_value = 4, -0.02
    out = _quiet(mod.optimizeCollFreeTrajectory, tc, traj, env=None, constraints=[con], settings=st, verbose=0)
    trace = [[float(v) for q in t.milestones[1:-1] for v in q] for t in out[1]]
    # a10: single link vs single environment along the trajectory (gopt:565-675)
    link_probes = []
    for li in (4, 6):
        lc = mod.RobotLinkTrajectoryCollisionConstraint(robot.link(li), con.envs[0], tc)
        for trial in range(2):
            xa = [float(v) for v in (np.asarray(x) + (rng.normal(0, 0.05, len(x)) if trial else 0))]
            lc.setx(xa)
            mv = _quiet(lc.minvalue, xa)
            rec = {"link": li, "x": xa, "empty": mv == []}
            if mv != []:
                y = mv[1]
                # df_dx right after minvalue: the reference leaves the robot at q(t_min), which is what its df_dx reads
                rec.update({"dmin": float(mv[0]), "param": _f(y), "row": _f(lc.df_dx(xa, y)), "value_at_param": float(lc.value(xa, y)),
                            "pruned_by_bound": _quiet(lc.minvalue, xa, mv[0] - 1e-6) == []})
            link_probes.append(rec)
            lc.clearx()
    write("c4_traj.json", {"link_probes": link_probes, "config": "C4 tx90_geom_test2.xml, chair moved to %s, 67 milestones (65 free), "
                                     "RobotTrajectoryCollisionConstraint, max_iters 4" % (scenes.CHAIR_IN_PATH,),
                           "x0": x, "minvalue_probes": probes, "trace": trace, "params": _f(out[3])})


def gen_c5():
    """64 of the 65,536 C5 problems through the reference's own optimizeCollFree (gopt:1039-1081 -> sip.py:795-1241),
    default settings (max_iters 50).  Per problem: final pose, outer iterations, status, f, g, instantiated parameters;
    the full pose trace for the first 8."""
    from semiinfiniteoptimization_b200 import workloads as W
    mod, cube, scene, con = scenes.c5_scene('reference')
    sip = sys.modules[mod.__name__.rsplit('.', 1)[0] + '.sip']
    allposes = frozen_c5_poses(65536, seed=3)
    cases = []
    for k, idx in enumerate(scenes.c5_sample_indices(64)):
        T = allposes[idx]
        out = _quiet(mod.optimizeCollFree, cube, scene, T, settings=sip.SemiInfiniteOptimizationSettings(), verbose=0)
        # the same run once more through optimizeSemiInfinite itself, for the fields the front-end drops (status, f, g)
        res = _quiet(sip.optimizeSemiInfinite, mod.ObjectPoseObjective(T), [con], T, settings=sip.SemiInfiniteOptimizationSettings(), verbose=0)
        assert len(res.trace) == len(out[1]) and np.array_equal(np.array(res.x[1]), np.array(out[0][1]))
        rec = {"index": idx, "Tinit": _f(T), "T": _f(out[0]), "iterations": len(out[1]) - 1, "num_iterations": int(res.num_iterations), "status": res.status,
               "fx": float(res.fx), "gx": _f(res.gx), "fx0": float(res.fx0), "gx0": _f(res.gx0), "params": _f(out[3])}
        if k < 8:
            rec["trace"] = _f(out[1])
        cases.append(rec)
    write("c5_pose.json", {"config": "C5: cube SDF@0.05 vs synthetic 262,144-point scene (seed 2), problems %s of the 65,536 (poses seed 3), "
                                     "optimizeCollFree defaults (max_iters 50)" % (scenes.c5_sample_indices(64)[:3] + ["..."],),
                           "cases": cases})


class ScriptedPlanner:
    """Stand-in planner for the planopt golden: 'finds' a fixed path on its first slice (Klampt's planners are absent)."""

    def __init__(self, path):
        self._path, self._found = [list(q) for q in path], False

    def planMore(self, n):
        self._found = True

    def getPath(self):
        return self._path if self._found else None


def planopt_inputs(WorldModel):
    """tx90_geom_test2.xml as shipped (chair clear of the arm), start / goal / planner path from robottrajopt_initial.path."""
    p0 = json.load(open(os.path.join(ROOT, "semiinfiniteoptimization_b200", "data", "paths.json")))["resources/robottrajopt_initial.path"]
    w = WorldModel()
    w.readFile("tx90_geom_test2.xml")
    robot = w.robot(0)
    robot.setConfig(p0["milestones"][0])
    return w, robot, p0


def gen_planopt():
    """The reference's planopt.py (arc_length_refine 8-33, planOptimizedTrajectory 36-245) executed with a scripted
    stand-in planner: refined trajectory, returned best path and the CSV log (time column dropped)."""
    import io
    from oracle.refrun import klampt_shim, load
    mods = load.load()
    planopt = mods["planopt"]
    from oracle.frozen.robot import WorldModel
    from oracle.frozen.trajectory import RobotTrajectory, Trajectory
    w, robot, p0 = planopt_inputs(WorldModel)
    out = {}
    rt = RobotTrajectory(robot, list(range(len(p0["milestones"]))), p0["milestones"])
    refined = planopt.arc_length_refine(rt, 12)
    plain = planopt.arc_length_refine(Trajectory(p0["times"], p0["milestones"]), 5)
    out["arc_length_refine"] = {"robot_12": {"times": _f(refined.times), "milestones": _f(refined.milestones)},
                                "plain_5": {"times": _f(plain.times), "milestones": _f(plain.milestones)}}
    klampt_shim.PLANNER_FACTORY = lambda world, rob, target, **kw: ScriptedPlanner(p0["milestones"])
    log = io.StringIO()
    robot.setConfig(p0["milestones"][0])
    best = _quiet(planopt.planOptimizedTrajectory, w, robot, p0["milestones"][-1], numRestarts=2, optimizationMaxIters=8, logFile=log)
    rows = [l.split(",") for l in log.getvalue().splitlines()]
    out["plan_optimize"] = {"args": {"numRestarts": 2, "optimizationMaxIters": 8}, "times": _f(best.times), "milestones": _f(best.milestones),
                            "length": float(best.length()),
                            "log_header": [",".join(r) for r in rows[:3]],
                            "log_rows": [[int(r[0]), int(r[1]), r[3], float(r[4])] for r in rows[3:]]}
    write("planopt.json", out)


def gen_paths():
    """The shipped `.path` / `.configs` files verbatim (data, not source) for the byte-exact round-trip test."""
    ref = os.environ.get("SIO_REFERENCE_ROOT", "/root/reference")
    files = {}
    for sub in ("resources", "output"):
        for fn in sorted(os.listdir(os.path.join(ref, sub))):
            if fn.endswith((".path", ".configs")):
                files[sub + "/" + fn] = open(os.path.join(ref, sub, fn)).read()
    write("path_files.json", {"files": files})


def gen_classes():
    """Single calls of the reference's constraint classes: value / minvalue / df_dx (gopt:394-420, 457-484)."""
    out = {}
    mod, g1, g2, con = scenes.c1_scene('reference')
    rng = np.random.default_rng(4)
    rec = []
    for T in scenes.c1_poses(5, seed=3):
        con.setx(T)
        mv = con.minvalue(T)
        pts = [[float(v) for v in (np.asarray(mv[1]) + d)] for d in rng.normal(0, 0.02, (4, 3))]
        rec.append({"T": _f(T), "dmin": float(mv[0]), "argmin_point": _f(mv[1]), "bound_prunes": con.minvalue(T, mv[0] - 1e-3) == [],
                    "points": pts, "values": [float(con.value(T, p)) for p in pts], "rows": [_f(con.df_dx(T, p)) for p in pts],
                    "pdg_distance": _f(g1.distance(pts[0])), "pdg_normal": _f(g1.normal(pts[0]))})
        con.clearx()
    out["object_constraint"] = rec
    mod, robot, cons = _quiet(scenes.c3_scene, 'reference')
    rec = []
    for q in scenes.c3_configs(robot, 3, seed=2):
        for c in cons:
            c.setx(q)
        per = []
        for c in cons:
            mv = c.minvalue(q)
            pts = [[float(v) for v in (np.asarray(mv[1]) + d)] for d in rng.normal(0, 0.03, (2, 3))]
            per.append({"dmin": float(mv[0]), "argmin_point": _f(mv[1]), "points": pts,
                        "values": [float(c.value(q, p)) for p in pts], "rows": [_f(c.df_dx(q, p)) for p in pts]})
        for c in cons:
            c.clearx()
        rec.append({"q": _f(q), "links": per})
    out["robot_link_constraints"] = rec
    write("classes.json", out)


def frozen_random_poses(seed, n, rot=0.3, trans=0.1, center=(0, 0, 0)):
    """tests/util.random_poses with the FROZEN so3 / se3 helpers: golden inputs must not move when the product's helpers are
    rewritten (they did, by one unit in the last place, when _host/so3.py went from numpy round trips to scalar arithmetic)."""
    from oracle.frozen import se3, so3
    rng = np.random.default_rng(seed)
    out = np.empty((n, 12))
    for i in range(n):
        w = rng.uniform(-rot, rot, 3)
        t = np.asarray(center) + rng.uniform(-trans, trans, 3)
        out[i] = se3.to_rowmajor12((so3.from_rotation_vector(w), list(t)))
    return out


def frozen_c5_poses(n, seed=3):
    """workloads.c5_poses with the frozen so3 (same draws, same order)."""
    from oracle.frozen import so3
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        R = so3.from_rotation_vector(rng.uniform(-np.pi / 2, np.pi / 2, 3))
        t = [float(rng.uniform(-1, 1)), float(rng.uniform(-1, 1)), float(rng.uniform(0.0, 1.0))]
        out.append((R, t))
    return out


def gen_oracle():
    """Known answers of the oracle itself on seeded arrays (regression pin for oracle/sio_oracle.c and for the
    CUDA kernels): analytic linear field (trilinear interpolation is exact on it) + seeded random grid."""
    from oracle import oracle as O
    from tests import util
    rng = np.random.default_rng(123)
    vg = util.random_grid(7)
    G = O.Grid(vg.values, vg.bmin, vg.bmax)
    T = frozen_random_poses(8, 6, rot=0.5, trans=0.2)
    pts = rng.uniform(-0.6, 1.6, (40, 3))
    d, g = O.query(G, T[0], pts)
    cloud = rng.uniform(-0.5, 1.5, (5000, 3)).astype(np.float32)
    bound = np.array([np.inf, 0.0, -0.5, 0.3, np.inf, -1.0])
    dmin, arg = O.min_distance(G, cloud, T, bound=bound)
    _, ub, ui, ud = O.under(G, cloud, T, bound)
    order = np.lexsort((ui, ub))
    ub, ui, ud = ub[order], ui[order], ud[order]
    rows_d, rows = O.pose_rows(G, T[1], pts[:9])
    out = {"grid_seed": 7, "T": _f(T), "pts": _f(pts), "query_d": _f(d), "query_grad": _f(g),
           "cloud_seed": 123, "bound": [None if not np.isfinite(b) else float(b) for b in bound],
           "dmin": _f(dmin), "argmin": _f(arg), "under_b": _f(ub), "under_idx": _f(ui), "under_d": _f(ud),
           "pose_rows_d": _f(rows_d), "pose_rows": _f(rows)}
    write("oracle_kat.json", out)


def multi_inputs(kind):
    """Two member constraints over the same pose variable: C1's cube-vs-chair pair and the same cube against the chair
    moved to (I, [1.0, 0.3, 0.1]); poses from c1_poses (seed 5); bounds None / 0.05 / -0.5 (the last prunes everything)."""
    from oracle.frozen import so3
    from semiinfiniteoptimization_b200._host.geometry import fixture_geometry
    mod, g1, g2, con_a = scenes.c1_scene(kind)
    nm = scenes._names(kind)
    g3 = getattr(mod, nm['PDG'])(fixture_geometry("m797"), 0.05, 0.02)
    g3.setTransform((so3.identity(), [1.0, 0.3, 0.1]))
    con_b = getattr(mod, nm['Obj'])(g1, g3)
    return mod, [con_a, con_b], scenes.c1_poses(4, seed=5), [None, 0.05, -0.5]


def gen_multi():
    """The reference's MultiSemiInfiniteConstraint (con:174-256) over two ObjectCollisionConstraints.  Its __init__ calls
    `sampleDomain`, a name constraint.py never imports: the loader's namespace gets
    sampleDomain = lambda d: d.sample() -- the only reading the call site allows -- and nothing else is touched."""
    mod, cons, poses, bounds = multi_inputs('reference')
    cmod = sys.modules[mod.__name__.rsplit('.', 1)[0] + '.constraint']
    cmod.sampleDomain = lambda d: d.sample()
    cases = []
    for all_negative in (True, False):
        M = cmod.MultiSemiInfiniteConstraint(cons, all_negative=all_negative)
        for T in poses:
            for bound in bounds:
                M.setx(T)
                mv = M.minvalue(T, bound)
                pairs = [mv] if len(mv) > 0 and not hasattr(mv[0], '__iter__') else mv
                rec = {"all_negative": all_negative, "T": _f(T), "bound": bound, "single": not (len(mv) == 0 or hasattr(mv[0], '__iter__')),
                       "pairs": [[float(f), _f(y)] for (f, y) in pairs],
                       # co-parameters handed back as lists: the reference's point queries accept list / tuple only (gopt:97-100)
                       "value": [float(M.value(T, [float(v) for v in y])) for (f, y) in pairs],
                       "df_dx": [_f(np.asarray(M.df_dx(T, [float(v) for v in y]))) for (f, y) in pairs]}
                M.clearx()
                cases.append(rec)
    write("multi_constraint.json", {"config": "MultiSemiInfiniteConstraint([cube vs chair@(1.2,0,0), cube vs chair@(1.0,0.3,0.1)]) (con:174-256)",
                                    "num_co_parameters": M.num_co_parameters, "cases": cases})


def gen_c4_long():
    """C4 run to its end: the reference's optimizeCollFreeTrajectory with max_iters 30 stops by itself after 14 outer
    iterations on this scene (chair across the path): final trajectory, its deepest penetration, the instantiated parameters."""
    mod, robot, tc, con, traj, x = _quiet(scenes.c4_scene, 'reference')
    sip = sys.modules[mod.__name__.rsplit('.', 1)[0] + '.sip']
    st = sip.SemiInfiniteOptimizationSettings()
    st.max_iters, st.minimum_constraint_value = 30, -0.02
    out = _quiet(mod.optimizeCollFreeTrajectory, tc, traj, env=None, constraints=[con], settings=st, verbose=0)
    xs = [float(v) for q in out[0].milestones[1:-1] for v in q]
    con.setx(xs)
    mv = _quiet(con.minvalue, xs)
    con.clearx()
    write("c4_traj_long.json", {"config": "C4 as in c4_traj.json, max_iters 30, minimum_constraint_value -0.02", "x0": x, "x": xs,
                                "iterations": len(out[1]) - 1, "dmin": float(mv[0]), "param": _f(mv[1]), "params": _f(out[3])})


def gen_c3_random():
    """optimizeCollFreeRobot with qinit = 'random-collision-free' / 'random' (gopt:1172-1207): up to 50 samples in a box growing
    from qdes to the joint limits, drawn with Python's `random` (seeded here), the first one whose constraints' minima reach
    the lower bound starts the solve.  Pins the draw order and the accept rule."""
    import random
    mod, robot, cons = _quiet(scenes.c3_scene, 'reference')
    sip = sys.modules[mod.__name__.rsplit('.', 1)[0] + '.sip']
    cases = []
    for k, qdes in enumerate(scenes.c3_configs(robot, 6, seed=3)):
        for mode in ('random-collision-free', 'random'):
            st = sip.SemiInfiniteOptimizationSettings()
            st.max_iters, st.minimum_constraint_value = 5, -0.02
            random.seed(100 + k)
            out = _quiet(mod.optimizeCollFreeRobot, robot, None, qdes=qdes, qinit=mode, constraints=cons, settings=st, verbose=0)
            q, trace = out[0], out[1]
            cases.append({"qdes": _f(qdes), "qinit": mode, "seed": 100 + k, "q": _f(q), "trace": _f(trace),
                          "draws_after": random.random()})      # the generator's state after the call: how many draws it took
    write("c3_random_init.json", {"config": "C3, optimizeCollFreeRobot with a sampled start (gopt:1172-1207), max_iters 5, "
                                            "minimum_constraint_value -0.02, random.seed(seed) before each call", "cases": cases})


def target_cases():
    """Start poses clear of the chair (x in [0, 0.3]) with targets inside it (c1_poses): the objective pulls the cube into
    contact, which is where Tdes != Tinit matters (minstep / value / the QP's xdes are all non-zero from the first iteration)."""
    from oracle.frozen import so3
    rng = np.random.default_rng(17)
    starts = [(so3.from_rotation_vector(rng.uniform(-0.3, 0.3, 3)), [float(rng.uniform(0.0, 0.3)), float(rng.uniform(-0.3, 0.5)),
                                                                     float(rng.uniform(-0.3, 0.5))]) for _ in range(12)]
    return starts, scenes.c1_poses(12, seed=23)


def gen_c1_targets():
    """optimizeCollFree(obj, env, Tinit, Tdes) with Tdes != Tinit on the C1 scene (gopt:1039-1081), reference defaults."""
    mod, g1, g2, con = scenes.c1_scene('reference')
    sip = sys.modules[mod.__name__.rsplit('.', 1)[0] + '.sip']
    starts, targets = target_cases()
    cases = []
    for T0, Td in zip(starts, targets):
        out = _quiet(mod.optimizeCollFree, g1, g2, T0, Td, settings=sip.SemiInfiniteOptimizationSettings(), verbose=0)
        res = _quiet(sip.optimizeSemiInfinite, mod.ObjectPoseObjective(Td), [con], T0, settings=sip.SemiInfiniteOptimizationSettings(), verbose=0)
        assert len(res.trace) == len(out[1])
        cases.append({"Tinit": _f(T0), "Tdes": _f(Td), "T": _f(out[0]), "iterations": len(out[1]) - 1, "num_iterations": int(res.num_iterations),
                      "status": res.status, "fx": float(res.fx), "gx": _f(res.gx), "trace": _f(out[1]), "params": _f(out[3])})
    write("c1_pose_targets.json", {"config": "C1 scene, optimizeCollFree with separate start and target poses, defaults (max_iters 50)",
                                   "cases": cases})


GENERATORS = {"c1targets": gen_c1_targets, "c3random": gen_c3_random, "c4long": gen_c4_long, "multi": gen_multi, "c1": gen_c1, "c2": gen_c2, "c3": gen_c3, "c4": gen_c4, "c5": gen_c5, "classes": gen_classes, "oracle": gen_oracle,
              "planopt": gen_planopt, "paths": gen_paths}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=",".join(GENERATORS))
    a = ap.parse_args()
    for k in a.only.split(","):
        t0 = time.time()
        GENERATORS[k]()
        print("  %s: %.1f s" % (k, time.time() - t0))
