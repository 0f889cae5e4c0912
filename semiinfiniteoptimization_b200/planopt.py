"""Planner / optimiser alternation for robot trajectories: the role of the reference's
semiinfinite/planopt.py (`arc_length_refine` planopt.py:8-33, `planOptimizedTrajectory` planopt.py:36-245), with the
same call signature, the same restart schedule and the same CSV log rows.

What differs, and why:
  * The reference takes its sampling-based planner from Klampt (`klampt.plan.robotplanning.planToConfig`, planopt.py:114);
    Klampt is not part of this build, so the planner sits behind a two-method interface (`planMore(n)`, `getPath()`) and
    two DECLARED STAND-INS are provided: `StraightLinePlanner` and `RRTConnectPlanner` (seeded, bidirectional).  Any
    object with those two methods can be passed through `planner_factory`.
  * The stand-in planners check configurations in BATCHES through the collision oracle: all sample configurations of
    an edge go to the device in one fused FK + cloud-vs-grid launch (`sio_robot_min_distance`), instead of one
    collision query per configuration.

The optimiser side is untouched: RobotTrajectoryCache + TrajectoryLengthObjective + RobotTrajectoryCollisionConstraint +
optimizeCollFreeTrajectory from geometryopt.py, exactly as the reference wires them (planopt.py:95-101, 168-197).
"""
import time

import numpy as np

from . import geometryopt
from ._host import vectorops
from ._host.trajectory import RobotTrajectory


def arc_length_refine(traj, numNewMilestones):
    """Inserts `numNewMilestones` new milestones, spread evenly over the path's arc length (planopt.py:8-33): segment i
    receives one new milestone for every multiple of (total length / numNewMilestones) that falls strictly before its
    end, placed at equal fractions of the segment."""
    assert numNewMilestones > 0
    metric = traj.robot.distance if isinstance(traj, RobotTrajectory) else vectorops.distance
    seg = [metric(a, b) for a, b in zip(traj.milestones[:-1], traj.milestones[1:])]
    ends = np.concatenate([[0.0], np.cumsum(seg)])               # arc length at every milestone
    # the reference accumulates the running sum in a Python loop; cumsum adds in the same order
    step = float(sum(seg)) / numNewMilestones
    extra = [0] * len(seg)
    mark = step
    for i, s in enumerate(ends):
        while mark < s:
            extra[i - 1] += 1
            mark += step
    times = [traj.times[0]]
    for i, n in enumerate(extra):
        t0, t1 = traj.times[i], traj.times[i + 1]
        times += [t0 + (t1 - t0) * (float(j + 1) / (n + 1)) for j in range(n)]
        times.append(t1)
    return traj.remesh(times)[0]


# ---- planner interface and declared stand-ins -----------------------------------------------------------------------
class PlannerInterface:
    """What planOptimizedTrajectory needs from a planner (the two calls the reference makes on Klampt's MotionPlan,
    planopt.py:121-124)."""

    def planMore(self, iterations):
        raise NotImplementedError()

    def getPath(self):
        """A list of configurations from start to goal, or None."""
        raise NotImplementedError()


class OracleFeasibility:
    """Configuration-space feasibility through the collision oracle: q is feasible iff it is within the joint limits and
    every link's SDF is non-negative over every obstacle cloud.  `batch(Q)` answers a whole array of configurations with
    one fused FK + cloud-vs-grid launch per obstacle."""

    def __init__(self, robot, obstacles, kinematicsCache=None, margin=0.0, pcres=0.04):
        self.robot = robot
        self.kin = kinematicsCache or geometryopt.RobotKinematicsCache(robot)
        self.envs = [e if isinstance(e, geometryopt.PenetrationDepthGeometry) else geometryopt.PenetrationDepthGeometry(e.geometry(), None, pcres)
                     for e in obstacles]
        self.links = [i for i in range(robot.numLinks()) if self.kin.geometry[i] is not None]
        self.qmin, self.qmax = (np.asarray(v, dtype=np.float64) for v in robot.getJointLimits())
        self.margin = margin
        self.queries = 0

    def batch(self, Q):
        Q = np.asarray(Q, dtype=np.float64).reshape(-1, self.robot.numLinks())
        ok = np.all((Q >= self.qmin - 1e-12) & (Q <= self.qmax + 1e-12), axis=1)
        kin = self.kin
        for env in self.envs:
            if not ok.any():
                break
            idx = np.nonzero(ok)[0]
            dmin, _, _ = kin.context.robot_min_distance(kin.robot_h, self.robot.numLinks(), kin.link_grids, self.links,
                                                        env._dev.cloud_h, env._T12, Q[idx], flags=geometryopt._sweep_flags())
            self.queries += len(idx) * len(self.links)
            ok[idx] = dmin.min(axis=1) >= self.margin
        return ok

    def __call__(self, q):
        return bool(self.batch([q])[0])


class StraightLinePlanner(PlannerInterface):
    """DECLARED STAND-IN: succeeds iff the straight joint-space segment from start to goal is feasible at `resolution`."""

    def __init__(self, start, goal, feasible, resolution=0.02):
        self.start, self.goal, self.feasible, self.resolution = list(start), list(goal), feasible, resolution
        self.path = None
        self.checked = False

    def planMore(self, iterations):
        if not self.checked:
            self.checked = True
            if _edge_free(self.feasible, self.start, self.goal, self.resolution):
                self.path = [list(self.start), list(self.goal)]

    def getPath(self):
        return self.path


def _edge_samples(a, b, resolution):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    n = max(1, int(np.ceil(np.abs(b - a).max() / resolution)))
    u = np.arange(0, n + 1) / n
    return a[None, :] + u[:, None] * (b - a)[None, :]


def _edge_free(feasible, a, b, resolution):
    S = _edge_samples(a, b, resolution)
    if hasattr(feasible, "batch"):
        return bool(np.all(feasible.batch(S)))
    return all(feasible(list(q)) for q in S)


class RRTConnectPlanner(PlannerInterface):
    """DECLARED STAND-IN for Klampt's sampling-based planners: a seeded bidirectional RRT (extend + connect) in the joint
    box, edges validated at `resolution` through the feasibility callable (batched when it offers `batch`); once the
    trees meet the path is shortcut `shortcut_rounds` times ('shortcut': True of the reference's default settings)."""

    def __init__(self, start, goal, feasible, qmin, qmax, seed=0, step=0.4, resolution=0.05, shortcut_rounds=30):
        self.feasible, self.step, self.resolution, self.shortcut_rounds = feasible, step, resolution, shortcut_rounds
        self.qmin, self.qmax = np.asarray(qmin, dtype=np.float64), np.asarray(qmax, dtype=np.float64)
        self.rng = np.random.default_rng(seed)
        self.trees = [[(np.asarray(start, dtype=np.float64), -1)], [(np.asarray(goal, dtype=np.float64), -1)]]
        self.path = None
        self.iterations = 0
        span = self.qmax - self.qmin
        self.free = span > 0

    def _nearest(self, tree, q):
        P = np.array([n[0] for n in tree])
        return int(np.argmin(((P - q) ** 2).sum(1)))

    def _steer(self, a, b):
        d = np.linalg.norm(b - a)
        return b if d <= self.step else a + (b - a) * (self.step / d)

    def _extend(self, tree, q):
        i = self._nearest(tree, q)
        new = self._steer(tree[i][0], q)
        if np.array_equal(new, tree[i][0]) or not _edge_free(self.feasible, tree[i][0], new, self.resolution):
            return None
        tree.append((new, i))
        return len(tree) - 1

    def _branch(self, tree, i):
        out = []
        while i >= 0:
            out.append(tree[i][0])
            i = tree[i][1]
        return out

    def planMore(self, iterations):
        if self.path is not None:
            self._shortcut(max(1, iterations // 10))
            return
        for _ in range(iterations):
            self.iterations += 1
            a, b = self.trees if self.iterations % 2 else self.trees[::-1]
            q = np.where(self.free, self.rng.uniform(self.qmin, np.where(self.free, self.qmax, self.qmin + 1)), self.qmin)
            ia = self._extend(a, q)
            if ia is None:
                continue
            # connect: walk the other tree toward the new node
            target = a[ia][0]
            while True:
                ib = self._extend(b, target)
                if ib is None:
                    break
                if np.array_equal(b[ib][0], target):
                    fa, fb = self._branch(a, ia)[::-1], self._branch(b, ib)
                    path = fa + fb[1:]
                    if a is self.trees[1]:
                        path = path[::-1]
                    self.path = [list(map(float, q_)) for q_ in path]
                    self._shortcut(self.shortcut_rounds)
                    return

    def _shortcut(self, rounds):
        p = [np.asarray(q) for q in self.path]
        for _ in range(rounds):
            if len(p) <= 2:
                break
            i, j = sorted(self.rng.integers(0, len(p), 2))
            if j - i >= 2 and _edge_free(self.feasible, p[i], p[j], self.resolution):
                p = p[:i + 1] + p[j:]
        self.path = [list(map(float, q)) for q in p]

    def getPath(self):
        return self.path


def planToConfig(world, robot, target, feasible=None, type='rrtconnect', seed=0, kinematicsCache=None, **settings):
    """Stand-in for klampt.plan.robotplanning.planToConfig (planopt.py:114): a planner from robot.getConfig() to `target`
    among the world's rigid objects and terrains.  `type`: 'rrtconnect' (also chosen for Klampt's 'sbl', 'rrt', ...) or
    'straight'."""
    if feasible is None:
        obstacles = [world.rigidObject(i) for i in range(world.numRigidObjects())] + [world.terrain(i) for i in range(world.numTerrains())]
        feasible = OracleFeasibility(robot, obstacles, kinematicsCache)
    start = robot.getConfig()
    if type == 'straight':
        return StraightLinePlanner(start, target, feasible)
    qmin, qmax = robot.getJointLimits()
    kw = {k: settings[k] for k in ('step', 'resolution', 'shortcut_rounds') if k in settings}
    if settings.get('shortcut') is False:
        kw['shortcut_rounds'] = 0
    return RRTConnectPlanner(start, target, feasible, qmin, qmax, seed=seed, **kw)


# ---- the alternation ------------------------------------------------------------------------------------------------
class _Log:
    """The reference's CSV log (planopt.py:81-85 header, then `restart,iterations,time,method,path length` rows)."""

    def __init__(self, f):
        self.f = f

    def header(self, numRestarts, plannerSettings, plannerMaxIters, plannerMaxTime, optimizationMaxIters):
        if self.f:
            self.f.write("numRestarts,plannerSettings,plannerMaxIters,plannerMaxTime,optimizationMaxIters\n")
            self.f.write("%d,%s,%d,%g,%d\n" % (numRestarts, str(plannerSettings), plannerMaxIters, plannerMaxTime, optimizationMaxIters))
            self.f.write("restart,iterations,time,method,path length\n")

    def row(self, restart, iterations, t, method, cost):
        if self.f:
            self.f.write("%d,%d,%g,%s,%g\n" % (restart, iterations, t, method, cost))


def planOptimizedTrajectory(world, robot, target, maxTime=float('inf'), numRestarts=1,
                            plannerSettings={'type': 'sbl', 'shortcut': True},
                            plannerMaxIters=3000, plannerMaxTime=10.0, plannerContinueOnRestart=False,
                            optimizationMaxIters=200, kinematicsCache=None, settings=None, logFile=None,
                            planner_factory=None, verbose=0):
    """Alternates a sampling-based planner with the semi-infinite trajectory optimiser and returns the best collision-free
    RobotTrajectory found (planopt.py:36-245; arguments as there).  Per restart: (1) if a best path exists, optimise it
    for optimizationMaxIters // (2 numRestarts) iterations; (2) run the planner in slices of 50 iterations until it has a
    path (or, with plannerSettings['optimizing'], until its budget ends); (3) refine a new path to ~30 milestones by arc
    length, re-time it to [0, 1] and optimise it.  A candidate replaces the best path only if it is shorter AND its
    trajectory constraint's minimum is non-negative.

    `planner_factory(world, robot, target, **plannerSettings)` -> PlannerInterface; default: planToConfig above."""
    say = print if verbose else (lambda *a, **k: None)
    clock0 = time.time()
    if kinematicsCache is None:
        kinematicsCache = geometryopt.RobotKinematicsCache(robot)
    if settings is None:
        settings = geometryopt.SemiInfiniteOptimizationSettings()
        settings.minimum_constraint_value = -0.02
    plannerSettings = dict(plannerSettings)
    optimizing = plannerSettings.pop('optimizing', False)
    if planner_factory is None:
        planner_factory = lambda w, r, t, **kw: planToConfig(w, r, t, kinematicsCache=kinematicsCache, **kw)
    obstacles = [world.rigidObject(i) for i in range(world.numRigidObjects())] + [world.terrain(i) for i in range(world.numTerrains())]
    obstaclegeoms = [geometryopt.PenetrationDepthGeometry(obs.geometry(), None, 0.04) for obs in obstacles]
    say("Preprocessing took time", time.time() - clock0)
    log = _Log(logFile)
    log.header(numRestarts, plannerSettings, plannerMaxIters, plannerMaxTime, optimizationMaxIters)

    best = {"path": None, "cost": float('inf'), "params": None}
    startConfig = robot.getConfig()
    tstart = time.time()

    def optimise(traj0, iters, restart, label, log_feasible_only):
        """One optimizeCollFreeTrajectory run from traj0; returns (traj, cost, residual, params)."""
        t0 = time.time()
        settings.max_iters = iters
        cache = geometryopt.RobotTrajectoryCache(kinematicsCache, len(traj0.milestones) - 2, qstart=startConfig, qend=target)
        cache.tstart, cache.tend = 0, traj0.times[-1]
        constraint = geometryopt.RobotTrajectoryCollisionConstraint(obstaclegeoms, cache)
        log.row(restart, 0, t0 - tstart, label, best["cost"])
        traj, trace, times, params = geometryopt.optimizeCollFreeTrajectory(cache, traj0, obstacles, constraints=[constraint],
                                                                            settings=settings, verbose=0)

        def residual_of(tr):
            x = cache.trajectoryToState(tr)
            constraint.setx(x)
            r = constraint.minvalue(x)
            constraint.clearx()
            return r
        for i, (tr, ti) in enumerate(zip(trace, times)):
            if i == 0:
                continue
            if not log_feasible_only or (logFile and residual_of(tr)[0] >= 0):
                log.row(restart, i, t0 + ti - tstart, label, min(best["cost"], tr.length()))
        return traj, traj.length(), residual_of(traj), params

    def adopt(traj, cost, residual, params, source):
        if cost < best["cost"] and residual[0] >= 0:
            say("Got a better cost path from", source, cost, "<", best["cost"])
            best.update(path=traj, cost=cost, params=params)
            return True
        return False

    planner = None
    for restart in range(numRestarts):
        # (1) polish the incumbent
        if best["path"] is not None and optimizationMaxIters > 0:
            settings.instantiated_params = best["params"]
            traj, cost, residual, params = optimise(best["path"], optimizationMaxIters // (numRestarts * 2), restart, "optimize-best", False)
            say("Optimized cost", cost, "and residual", residual)
            adopt(traj, cost, residual, params, "optimizing the current best")
            if time.time() - tstart > maxTime:
                break
        # (2) sample
        if restart == 0 or (not plannerContinueOnRestart and planner.getPath()):
            robot.setConfig(startConfig)
            planner = planner_factory(world, robot, target, **plannerSettings)
        t0 = time.time()
        log.row(restart, 0, t0 - tstart, "sample", best["cost"])
        path, oldpath, it, t1 = None, None, 0, t0
        for it in range(0, plannerMaxIters, 50):
            oldpath = path
            planner.planMore(50)
            t1 = time.time()
            path = planner.getPath()
            if path:
                if oldpath is None:
                    say("Found a feasible path on restart", restart, "iteration", it)
                path = RobotTrajectory(robot, list(range(len(path))), path)
                cost = path.length()
                if cost < best["cost"]:
                    best.update(path=path, cost=cost)
                    log.row(restart, it + 50, t1 - tstart, "sample-optimize" if oldpath is not None else "sample", best["cost"])
                if not optimizing:
                    break
            if t1 - t0 > plannerMaxTime or t1 - tstart > maxTime:
                break
        log.row(restart, it + 50, t1 - tstart, "sample-optimize" if oldpath is not None else "sample", best["cost"])
        if t1 - tstart > maxTime:
            break
        # (3) optimise what the planner found
        if path and len(path.milestones) == 2:
            return path                                     # the straight line is feasible: nothing shorter exists
        if not path:
            say("No feasible path was found by the planner on restart", restart)
            continue
        n = len(path.milestones)
        traj0 = arc_length_refine(path, min(30 - n, n)) if n < 30 else path
        traj0.times = [float(i) / (len(traj0.milestones) - 1) for i in range(len(traj0.times))]
        if optimizationMaxIters <= 0:
            continue
        iters = optimizationMaxIters // (numRestarts * 2) if best["path"] else optimizationMaxIters // numRestarts
        # the reference tests `if bestPath:` AFTER the planner may already have made its path the best one
        settings.instantiated_params = None
        traj, cost, residual, params = optimise(traj0, iters, restart, "optimize", True)
        say("Optimized cost", cost, "and residual", residual)
        if not adopt(traj, cost, residual, params, "the optimizer"):
            say("Optimizer produced a worse cost (", cost, "vs", best["cost"], ") or negative residual, skipping")
        if time.time() - tstart > maxTime:
            break
    return best["path"]
