"""SURVEY.md row f4 on the CPU: the `.path` / `.configs` wire formats against the reference's shipped files, and the
planner / optimiser alternation (planopt.py) against goldens produced by the REFERENCE's own planopt.py running through
oracle/refrun with a scripted stand-in planner (tools/make_golden.py gen_planopt)."""
import io
import json
import os
import types

import numpy as np
import pytest

from tests import scenes


class ScriptedPlanner:
    def __init__(self, path):
        self._path, self._found = [list(q) for q in path], False

    def planMore(self, n):
        self._found = True

    def getPath(self):
        return self._path if self._found else None


def _inputs():
    from semiinfiniteoptimization_b200._host.robot import WorldModel
    p0 = json.load(open(os.path.join(scenes.ROOT, "semiinfiniteoptimization_b200", "data", "paths.json")))["resources/robottrajopt_initial.path"]
    w = WorldModel()
    w.readFile("tx90_geom_test2.xml")
    robot = w.robot(0)
    robot.setConfig(p0["milestones"][0])
    return w, robot, p0


# ---- wire formats ----------------------------------------------------------------------------------------------------
def test_shipped_path_and_configs_files_round_trip_byte_for_byte():
    from semiinfiniteoptimization_b200._host import pathio
    files = scenes.load_golden("path_files.json")["files"]
    assert len(files) == 6
    shipped = json.load(open(os.path.join(scenes.ROOT, "semiinfiniteoptimization_b200", "data", "paths.json")))
    for name, text in files.items():
        if name.endswith(".configs"):
            q = pathio.configs_from_text(text)
            assert pathio.configs_to_text(q) == text
            assert len(q) == 2 and all(len(c) == 7 for c in q)
        else:
            t, m = pathio.path_from_text(text)
            assert pathio.path_to_text(t, m) == text                       # what the reference's interpreter wrote
            assert np.array_equal(np.array(m), np.array(shipped[name]["milestones"])) and np.array_equal(t, shipped[name]["times"])
            assert all(len(q) == 7 for q in m)


def test_path_io_files_methods_precision_and_errors(tmp_path):
    from semiinfiniteoptimization_b200._host import pathio
    from semiinfiniteoptimization_b200._host.trajectory import Trajectory
    rng = np.random.default_rng(0)
    times = np.cumsum(rng.uniform(0.1, 1, 9)).tolist()
    ms = rng.normal(0, 3, (9, 7)).tolist()
    ms[3][2] = 2.0
    ms[4][0] = 1e-7
    fn = str(tmp_path / "a.path")
    tr = Trajectory(times, ms)
    tr.save(fn, precise=True)                                               # repr(): every double survives
    back = Trajectory()
    back.load(fn)
    assert back.milestones == ms and np.allclose(back.times, times, atol=5e-7)   # times are "%f" in this format
    tr.save(fn)                                                             # the reference's 12 significant digits
    back.load(fn)
    assert np.allclose(back.milestones, ms, rtol=1e-11, atol=0)
    assert "2.0" in open(fn).read().split("\n")[3] and "1e-07" in open(fn).read()
    cf = str(tmp_path / "a.configs")
    pathio.save_configs(cf, ms[:3], precise=True)
    assert pathio.load_configs(cf) == ms[:3]
    assert pathio.path_from_text("") == ([], []) and pathio.configs_from_text("\n\n") == []
    with pytest.raises(ValueError):
        pathio.path_from_text("0.0\t3 1.0 2.0\n")                          # ragged: announces 3, holds 2
    with pytest.raises(ValueError):
        pathio.configs_from_text("2\t1.0 2.0 3.0\n")
    with pytest.raises(ValueError):
        pathio.path_from_text("1.0\t1 0.0\n0.5\t1 0.0\n")                  # times go backwards


# ---- arc_length_refine -----------------------------------------------------------------------------------------------
def test_arc_length_refine_equals_the_reference():
    from semiinfiniteoptimization_b200 import planopt
    from semiinfiniteoptimization_b200._host.trajectory import RobotTrajectory, Trajectory
    g = scenes.load_golden("planopt.json")["arc_length_refine"]
    w, robot, p0 = _inputs()
    got = planopt.arc_length_refine(RobotTrajectory(robot, list(range(12)), p0["milestones"]), 12)
    assert len(got.milestones) == len(g["robot_12"]["milestones"]) == 24
    assert np.allclose(got.times, g["robot_12"]["times"], atol=1e-12, rtol=0)
    assert np.allclose(got.milestones, g["robot_12"]["milestones"], atol=1e-12, rtol=0)
    got = planopt.arc_length_refine(Trajectory(p0["times"], p0["milestones"]), 5)
    assert np.allclose(got.times, g["plain_5"]["times"], atol=1e-12, rtol=0)
    assert np.allclose(got.milestones, g["plain_5"]["milestones"], atol=1e-12, rtol=0)


# ---- the alternation, restated classes on the CPU oracle ---------------------------------------------------------------
def _restated_namespace():
    from oracle import ref_geometryopt as R
    from semiinfiniteoptimization_b200 import geometryopt as G
    return types.SimpleNamespace(PenetrationDepthGeometry=R.RefPDG, RobotKinematicsCache=R.RefRobotKinematicsCache,
                                 RobotTrajectoryCache=R.RefRobotTrajectoryCache,
                                 RobotTrajectoryCollisionConstraint=R.RefRobotTrajectoryCollisionConstraint,
                                 TrajectoryLengthObjective=G.TrajectoryLengthObjective,
                                 optimizeCollFreeTrajectory=G.optimizeCollFreeTrajectory,
                                 SemiInfiniteOptimizationSettings=G.SemiInfiniteOptimizationSettings)


def check_against_golden(best, log_text, tol):
    g = scenes.load_golden("planopt.json")["plan_optimize"]
    assert len(best.milestones) == len(g["milestones"])
    assert np.allclose(best.times, g["times"], atol=1e-12, rtol=0)
    assert np.allclose(best.milestones, g["milestones"], atol=tol, rtol=0)
    assert abs(best.length() - g["length"]) <= 10 * tol
    lines = log_text.splitlines()
    assert lines[:3] == g["log_header"]
    rows = [l.split(",") for l in lines[3:]]
    assert [[int(r[0]), int(r[1]), r[3]] for r in rows] == [r[:3] for r in g["log_rows"]]
    assert np.allclose([float(r[4]) for r in rows], [r[3] for r in g["log_rows"]], rtol=2e-6)      # "%g": six digits


def test_plan_optimize_alternation_equals_the_reference(monkeypatch):
    from semiinfiniteoptimization_b200 import planopt
    monkeypatch.setattr(planopt, "geometryopt", _restated_namespace())
    w, robot, p0 = _inputs()
    log = io.StringIO()
    best = planopt.planOptimizedTrajectory(w, robot, p0["milestones"][-1], numRestarts=2, optimizationMaxIters=8, logFile=log,
                                           planner_factory=lambda world, rob, target, **kw: ScriptedPlanner(p0["milestones"]))
    check_against_golden(best, log.getvalue(), 1e-8)


# ---- the declared stand-in planners ------------------------------------------------------------------------------------
class BallWorld:
    """Feasibility callable for the planner tests: a unit box in R^3 with a ball obstacle in the middle."""

    def __init__(self):
        self.calls = 0

    def batch(self, Q):
        Q = np.asarray(Q, dtype=float).reshape(-1, 3)
        self.calls += 1
        return (np.linalg.norm(Q - 0.5, axis=1) >= 0.3) & np.all((Q >= 0) & (Q <= 1), axis=1)

    def __call__(self, q):
        return bool(self.batch([q])[0])


def test_stand_in_planners_find_feasible_paths_deterministically():
    from semiinfiniteoptimization_b200 import planopt
    f = BallWorld()
    start, goal = [0.05, 0.5, 0.5], [0.95, 0.5, 0.5]
    sl = planopt.StraightLinePlanner(start, goal, f)
    sl.planMore(50)
    assert sl.getPath() is None                                             # the ball is in the way
    sl = planopt.StraightLinePlanner(start, [0.05, 0.9, 0.5], f)
    sl.planMore(50)
    assert sl.getPath() == [start, [0.05, 0.9, 0.5]]
    paths = []
    for _ in range(2):
        rrt = planopt.RRTConnectPlanner(start, goal, BallWorld(), [0, 0, 0], [1, 1, 1], seed=4, step=0.2, resolution=0.02)
        for _ in range(40):
            rrt.planMore(50)
            if rrt.getPath():
                break
        p = rrt.getPath()
        assert p and p[0] == start and p[-1] == goal
        for a, b in zip(p[:-1], p[1:]):
            assert planopt._edge_free(f, a, b, 0.01)
        paths.append(p)
    assert paths[0] == paths[1]                                             # seeded: same path twice
    n0 = len(paths[0])
    rrt.planMore(500)                                                       # more iterations only shortcut an existing path
    assert len(rrt.getPath()) <= n0
