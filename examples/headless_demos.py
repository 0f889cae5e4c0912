#!/usr/bin/env python
"""The reference's four demo drivers without their Klampt GL windows, on the drop-in classes -- what a user of
krishauser/SemiInfiniteOptimization runs after switching the import path (needs a B200; geometry comes from the
bundled fixtures, Klampt being absent from this image):

    python examples/headless_demos.py [geomopt] [robotposeopt] [robottrajopt] [robotplanopt]

  geomopt        geomopt.py:13-15, 39-43, 62: cube vs chair, gridres 0.05 / pcres 0.02, optimizeCollFree from a dragged pose
  robotposeopt   robotposeopt.py:13-14, 52, 101-104, 123: TX90 vs the tree, makeCollisionConstraints + optimizeCollFreeRobot,
                 max_iters 5, minimum_constraint_value -0.02
  robottrajopt   robottrajopt.py:13-17, 70-71, 101-131, 197: the initial path remeshed x 6, chair as a cloud only,
                 RobotTrajectoryCollisionConstraint + optimizeCollFreeTrajectory.  (With the chair moved across the path the
                 reference's own optimiser, run through oracle/refrun, stops after 14 iterations at a deepest penetration of
                 -0.0484 with |x - x0| = 0.005 and 12 instantiated parameters; so does this one: tests/golden/c4_traj_long.json.)
  robotplanopt   robotplanopt.py:67-143: planner / optimiser alternation (planOptimizedTrajectory) with the declared
                 stand-in planner

Imports go through the `semiinfinite` shim exactly as in the reference's scripts."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from semiinfinite.geometryopt import *          # noqa: F401,F403  (the reference's drivers do the same)
from semiinfinite.sip import SemiInfiniteOptimizationSettings
from semiinfiniteoptimization_b200 import workloads as W
from semiinfiniteoptimization_b200._host import so3
from semiinfiniteoptimization_b200._host.geometry import fixture_geometry
from semiinfiniteoptimization_b200._host.robot import WorldModel


def geomopt():
    gridres, pcres = 0.05, 0.02
    geom1 = PenetrationDepthGeometry(fixture_geometry("cube"), gridres, pcres)
    geom2 = PenetrationDepthGeometry(fixture_geometry("m797"), gridres, pcres)
    geom2.setTransform((so3.identity(), [1.2, 0, 0]))
    for T in W.c1_poses(3, seed=0):                       # stand-ins for the poses a user drags the cube to
        geom1.setTransform(T)
        d0 = geom2.distance(geom1)[0]
        t0 = time.time()
        Tcollfree, trace, tracetimes, cps = optimizeCollFree(geom1, geom2, T, verbose=0)
        dt = time.time() - t0
        geom1.setTransform(Tcollfree)
        print("geomopt: depth %.4f -> %.4f in %d iterations, %.1f ms, %d constraint points"
              % (d0, geom2.distance(geom1)[0], len(trace) - 1, dt * 1e3, len(cps)))


def robotposeopt():
    gridres, pcres = 0.02, 0.02
    w = WorldModel()
    w.readFile("tx90_geom_test3.xml")
    robot, obstacles = w.robot(0), [w.rigidObject(0)]
    constraints, pairs = makeCollisionConstraints(robot, obstacles, gridres, pcres)
    settings = SemiInfiniteOptimizationSettings()
    settings.max_iters = 5
    settings.minimum_constraint_value = -0.02
    for q0 in W.c3_configs(robot, 3):
        t0 = time.time()
        qcollfree, trace, times, cps = optimizeCollFreeRobot(robot, obstacles, constraints=constraints, qinit=None, qdes=q0,
                                                             verbose=0, settings=settings)
        dt = time.time() - t0
        for c in constraints:
            c.setx(qcollfree)
        dmin = min((c.minvalue(qcollfree) or [np.inf])[0] for c in constraints)
        for c in constraints:
            c.clearx()
        print("robotposeopt: moved |dq| = %.3f from the target, min clearance %.4f, %d iterations, %.1f ms"
              % (float(np.linalg.norm(np.asarray(qcollfree) - np.asarray(q0))), dmin, len(trace) - 1, dt * 1e3))


def robottrajopt():
    import semiinfinite.geometryopt as G
    robot, tc, con, traj, x = W.c4_scene(G)
    tc.set(x)
    con.setx(x)
    d0 = con.minvalue(x)[0]
    con.clearx()
    settings = SemiInfiniteOptimizationSettings()
    settings.max_iters = 30
    settings.minimum_constraint_value = -0.02
    t0 = time.time()
    trajsolved, trace, times, params = optimizeCollFreeTrajectory(tc, traj, env=None, constraints=[con], greedyStart=False,
                                                                  verbose=0, settings=settings)
    dt = time.time() - t0
    xs = [float(v) for q in trajsolved.milestones[1:-1] for v in q]
    con.setx(xs)
    mv = con.minvalue(xs)
    con.clearx()
    print("robottrajopt: deepest penetration along the path %.4f -> %s, |x - x0| = %.3f, %d iterations, %.1f ms, %d milestones, "
          "%d instantiated (link, env, t, point) parameters"
          % (d0, "%.4f" % mv[0] if len(mv) else "none", float(np.linalg.norm(np.asarray(xs) - np.asarray(x))), len(trace) - 1, dt * 1e3,
             len(trajsolved.milestones), sum(len(p) for p in params)))
    if len(mv):
        link, env, t = mv[1][:3]
        first = trajsolved.times[1]
        print("              deepest point: link %d at t = %.4f%s" % (link, t, "" if t > first else
              "  (on the FIRST segment, t <= %.4f: the reference gives such parameters an all-zero Jacobian row -- "
              "geometryopt.py:931-945, reproduced -- so the QP cannot move them)" % first))


def robotplanopt():
    from semiinfinite import planopt
    w = WorldModel()
    w.readFile("tx90_geom_test2.xml")
    robot, chair = w.robot(0), w.rigidObject(0)
    chair.setTransform(chair.getTransform()[0], list(W.CHAIR_IN_PATH))
    import json
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p0 = json.load(open(os.path.join(here, "semiinfiniteoptimization_b200", "data", "paths.json")))["resources/robottrajopt_initial.path"]
    robot.setConfig(p0["milestones"][0])
    t0 = time.time()
    best = planopt.planOptimizedTrajectory(w, robot, p0["milestones"][-1], numRestarts=2, optimizationMaxIters=8,
                                           plannerSettings={'type': 'rrtconnect', 'seed': 1}, verbose=0)
    print("robotplanopt: %s in %.2f s" % ("no collision-free path found" if best is None else
                                          "collision-free path of %d milestones, length %.3f" % (len(best.milestones), best.length()),
                                          time.time() - t0))


if __name__ == "__main__":
    demos = {"geomopt": geomopt, "robotposeopt": robotposeopt, "robottrajopt": robottrajopt, "robotplanopt": robotplanopt}
    for name in (sys.argv[1:] or ["geomopt", "robotposeopt", "robottrajopt"]):
        demos[name]()
