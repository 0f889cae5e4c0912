"""SURVEY.md row f4 on the device: the planner / optimiser alternation with the device-backed classes against the
golden produced by the reference's own planopt.py, and the declared stand-in RRT planner checking its configurations
through the collision oracle (one fused FK + cloud-vs-grid launch per edge)."""
import io

import numpy as np
import pytest

from tests import scenes
from tests.test_planopt_cpu import ScriptedPlanner, _inputs, check_against_golden

pytestmark = pytest.mark.gpu


def test_plan_optimize_alternation_on_the_device_equals_the_reference(ctx):
    from semiinfiniteoptimization_b200 import planopt
    w, robot, p0 = _inputs()
    log = io.StringIO()
    best = planopt.planOptimizedTrajectory(w, robot, p0["milestones"][-1], numRestarts=2, optimizationMaxIters=8, logFile=log,
                                           planner_factory=lambda world, rob, target, **kw: ScriptedPlanner(p0["milestones"]))
    check_against_golden(best, log.getvalue(), 1e-6)


def test_rrt_stand_in_plans_through_the_oracle_and_the_optimiser_shortens_its_path(ctx):
    """Chair across the straight joint-space segment from start to goal: the segment is infeasible, the seeded RRT finds a way around it with
    batched oracle feasibility checks, and planOptimizedTrajectory returns a shorter, still collision-free trajectory."""
    from semiinfiniteoptimization_b200 import geometryopt as G, planopt
    w, robot, p0 = _inputs()
    chair = w.rigidObject(0)
    chair.setTransform(chair.getTransform()[0], [0.59, -0.22, 0.54])      # endpoints clear by 0.18 m, the straight segment cuts the chair
    start, goal = p0["milestones"][0], p0["milestones"][-1]
    kin = G.RobotKinematicsCache(robot)
    feas = planopt.OracleFeasibility(robot, [chair], kin)
    assert feas(start) and feas(goal)
    assert not planopt._edge_free(feas, start, goal, 0.02)
    robot.setConfig(start)
    rrt = planopt.planToConfig(w, robot, goal, feasible=feas, seed=1)
    for _ in range(60):
        rrt.planMore(50)
        if rrt.getPath():
            break
    path = rrt.getPath()
    assert path and path[0] == list(start) and path[-1] == list(goal)
    for a, b in zip(path[:-1], path[1:]):
        assert planopt._edge_free(feas, a, b, 0.02)
    planned_len = sum(robot.distance(a, b) for a, b in zip(path[:-1], path[1:]))
    robot.setConfig(start)
    best = planopt.planOptimizedTrajectory(w, robot, goal, numRestarts=1, optimizationMaxIters=10, kinematicsCache=kin,
                                           plannerSettings={'type': 'rrtconnect', 'seed': 1, 'feasible': feas})
    assert best is not None and best.milestones[0] == list(start) and best.milestones[-1] == list(goal)
    assert best.length() <= planned_len + 1e-9
    cache = G.RobotTrajectoryCache(kin, len(best.milestones) - 2, qstart=list(start), qend=list(goal))
    cache.tstart, cache.tend = 0, best.times[-1]
    con = G.RobotTrajectoryCollisionConstraint([G.PenetrationDepthGeometry(chair.geometry(), None, 0.04)], cache)
    x = cache.trajectoryToState(best)
    con.setx(x)
    residual = con.minvalue(x)
    con.clearx()
    assert residual[0] >= 0
