"""GPU parity at the level of the reference's classes: the device-backed drop-in classes
(semiinfiniteoptimization_b200.geometryopt) against the literal CPU restatement of the reference's
per-call algorithms (oracle/ref_geometryopt.py), and whole SIP solutions on BASELINE.json's configs
C1 (cube vs chair pose), C3 (TX90 pose) and C4 (TX90 trajectory)."""
import json
import os

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def G(ctx):
    from semiinfiniteoptimization_b200 import geometryopt
    return geometryopt


@pytest.fixture(scope="module")
def R(built):
    from oracle import ref_geometryopt
    return ref_geometryopt


@pytest.fixture(scope="module")
def c1(G, R):
    """C1: cube SDF@0.05 (moving) vs m797 cloud@0.02 placed at (I,[1.2,0,0]) (geomopt.py:14-15, 43)."""
    from semiinfiniteoptimization_b200._host import so3
    cube, chair = util.fixture_geometry("cube"), util.fixture_geometry("m797")
    g1, g2 = G.PenetrationDepthGeometry(cube, 0.05, 0.02), G.PenetrationDepthGeometry(chair, 0.05, 0.02)
    r1, r2 = R.RefPDG(util.fixture_geometry("cube"), 0.05, 0.02), R.RefPDG(util.fixture_geometry("m797"), 0.05, 0.02)
    T2 = (so3.identity(), [1.2, 0, 0])
    g2.setTransform(T2)
    r2.setTransform(T2)
    return g1, g2, r1, r2


def c1_poses(n, seed=0):
    from semiinfiniteoptimization_b200._host import so3
    rng = np.random.default_rng(seed)
    return [(so3.from_rotation_vector(rng.uniform(-0.3, 0.3, 3)),
             [rng.uniform(0.4, 1.0), rng.uniform(-0.3, 0.5), rng.uniform(-0.3, 0.5)]) for _ in range(n)]


def test_pdg_point_queries(c1):
    g1, g2, r1, r2 = c1
    rng = np.random.default_rng(1)
    for T in c1_poses(3, seed=1):
        g1.setTransform(T)
        r1.setTransform(T)
        for _ in range(5):
            pt = list(rng.uniform(0.2, 1.6, 3))
            d, cp1, cp2 = g1.distance(pt)
            d0, c10, c20 = r1.distance(pt)
            assert abs(d - d0) <= 1e-12 and np.allclose(cp1, c10, atol=1e-10) and cp2 == pt
            assert np.allclose(g1.normal(pt), r1.normal(pt), atol=1e-10)


def test_pdg_cloud_vs_grid_and_bound(c1):
    g1, g2, r1, r2 = c1
    for T in c1_poses(6, seed=2):
        g1.setTransform(T)
        r1.setTransform(T)
        d, cp1, cp2 = g2.distance(g1)
        d0, c10, c20 = r2.distance(r1)
        assert abs(d - d0) <= 1e-12 and np.allclose(cp1, c10, atol=1e-12) and np.allclose(cp2, c20, atol=1e-9)
        # Klampt's upperBound: zeros for the closest points when nothing is below it (gopt:128-130)
        db, cb1, cb2 = g2.distance(g1, d - 0.01)
        assert cb1 == [0.0] * 3 and cb2 == [0.0] * 3 and db >= d - 0.01
        da, ca1, _ = g2.distance(g1, d + 0.01)
        assert abs(da - d) <= 1e-12 and np.allclose(ca1, cp1)


def test_object_constraint_protocol(G, R, c1):
    g1, g2, r1, r2 = c1
    c = G.ObjectCollisionConstraint(g1, g2)
    cr = R.RefObjectCollisionConstraint(r1, r2)
    for T in c1_poses(5, seed=3):
        c.setx(T)
        cr.setx(T)
        mv, mr = c.minvalue(T), cr.minvalue(T)
        assert abs(mv[0] - mr[0]) <= 1e-12 and np.allclose(mv[1], mr[1], atol=1e-12)
        assert c.minvalue(T, mv[0] - 1e-3) == [] and cr.minvalue(T, mv[0] - 1e-3) == []
        # value(x, argmin) == dmin (the reference's commented self-check, sip.py:223-230)
        assert abs(c.value(T, mv[1]) - mv[0]) <= 1e-9
        pts = [list(np.asarray(mv[1]) + d) for d in np.random.default_rng(4).normal(0, 0.02, (6, 3))]
        vb = c.value_batch(T, pts)
        rows = c.df_dx_batch(T, pts)
        for k, p in enumerate(pts):
            assert abs(vb[k] - cr.value(T, p)) <= 1e-12
            assert abs(c.value(T, p) - vb[k]) <= 1e-12
            assert np.abs(rows[k] - cr.df_dx(T, p)).max() <= 1e-9
            assert np.abs(c.df_dx(T, p) - rows[k]).max() <= 1e-12
        c.clearx()
        cr.clearx()


def test_c1_solutions_match(G, R, c1):
    """optimizeCollFree, defaults (sip.py:64-75): same trace from the device-backed and the CPU-backed constraint."""
    from semiinfiniteoptimization_b200 import sip
    g1, g2, r1, r2 = c1
    nontrivial = 0
    for T in c1_poses(8, seed=0):
        out = G.optimizeCollFree(g1, g2, T, verbose=0)
        res = G.optimizeCollFree.last_result
        ref = sip.optimizeSemiInfinite(G.ObjectPoseObjective(T), [R.RefObjectCollisionConstraint(r1, r2)], T, verbose=0)
        assert res.num_iterations == ref.num_iterations and res.status == ref.status
        assert len(res.trace) == len(ref.trace)
        for a, b in zip(res.trace, ref.trace):
            assert np.allclose(a[0], b[0], atol=1e-7) and np.allclose(a[1], b[1], atol=1e-7)
        assert len(out) == 4 and len(out[3]) == len(ref.instantiated_params[0])
        for ya, yb in zip(out[3], ref.instantiated_params[0]):
            assert np.allclose(ya, yb, atol=1e-9)
        assert np.allclose(res.gx, ref.gx, atol=1e-7)
        nontrivial += res.num_iterations > 2
    assert nontrivial >= 2


def test_validation_mode_gives_same_minima(G, c1):
    g1, g2, r1, r2 = c1
    c = G.ObjectCollisionConstraint(g1, g2)
    for T in c1_poses(4, seed=5):
        c.setx(T)
        a = c.minvalue(T)
        G.set_validation_mode(True)
        try:
            b = c.minvalue(T)
        finally:
            G.set_validation_mode(False)
        assert a[0] == b[0] and a[1] == b[1]


# ---- robot pose (C3) -----------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c3(G, R):
    """C3: tx90_geom_test3.xml -- ARC.rob chain with declared synthetic link boxes, tree = m1062 x2 rotated,
    gridres = pcres = 0.02 (robotposeopt.py:13-14)."""
    from semiinfiniteoptimization_b200._host.robot import WorldModel
    w = WorldModel()
    w.readFile("tx90_geom_test3.xml")
    robot, tree = w.robot(0), w.rigidObject(0)
    constraints, pairs = G.makeCollisionConstraints(robot, [tree], 0.02, 0.02)
    w2 = WorldModel()
    w2.readFile("tx90_geom_test3.xml")
    robot2, tree2 = w2.robot(0), w2.rigidObject(0)
    cache2 = R.RefRobotKinematicsCache(robot2, 0.02, 0.02)
    env2 = R.RefPDG(tree2.geometry(), 0.02, 0.02)
    refs = [R.RefRobotLinkCollisionConstraint(robot2.link(p[0].index), env2, cache2) for p in pairs]
    return robot, constraints, robot2, refs


def c3_configs(robot, n, seed=0):
    rng = np.random.default_rng(seed)
    qmin, qmax = robot.getJointLimits()
    return [[lo + rng.uniform(0.2, 0.8) * (hi - lo) for lo, hi in zip(qmin, qmax)] for _ in range(n)]


def test_device_fk_matches_host_model(G, c3):
    robot, constraints, _, _ = c3
    cache = constraints[0].robot
    qs = np.array(c3_configs(robot, 16, seed=1))
    T = cache.context.fk(cache.robot_h, qs, robot.numLinks())
    for k, q in enumerate(qs):
        assert np.abs(T[k].reshape(-1, 3, 4) - robot.fk(q)).max() <= 1e-12


def test_robot_link_constraints_match(c3):
    robot, constraints, robot2, refs = c3
    assert len(constraints) == len(refs) == 7
    rng = np.random.default_rng(2)
    for q in c3_configs(robot, 4, seed=2):
        for c, r in zip(constraints, refs):
            c.setx(q)
            r.setx(q)
        for c, r in zip(constraints, refs):
            mv, mr = c.minvalue(q), r.minvalue(q)
            assert abs(mv[0] - mr[0]) <= 1e-12 and np.allclose(mv[1], mr[1], atol=1e-12)
            assert c.minvalue(q, mv[0]) == []
            pts = [list(np.asarray(mv[1]) + d) for d in rng.normal(0, 0.03, (3, 3))]
            vb, rows = c.value_batch(q, pts), c.df_dx_batch(q, pts)
            for k, p in enumerate(pts):
                assert abs(vb[k] - r.value(q, p)) <= 1e-12 and abs(c.value(q, p) - vb[k]) <= 1e-12
                assert np.abs(rows[k] - r.df_dx(q, p)).max() <= 1e-9
        for c, r in zip(constraints, refs):
            c.clearx()
            r.clearx()


def test_c3_solutions_match(G, c3):
    from semiinfiniteoptimization_b200 import sip
    robot, constraints, robot2, refs = c3
    settings = sip.SemiInfiniteOptimizationSettings()
    settings.max_iters = 5                               # robotposeopt.py:101-104
    settings.minimum_constraint_value = -0.02
    qmin, qmax = robot.getJointLimits()
    moved = 0
    for qdes in c3_configs(robot, 6, seed=3):
        qinit = [0.0] * 7
        q, trace, times, params = G.optimizeCollFreeRobot(robot, None, qdes=qdes, qinit=qinit, constraints=constraints,
                                                         settings=settings, verbose=0)
        ref = sip.optimizeSemiInfinite(G.RobotConfigObjective(robot2, qdes), refs, qinit, np.asarray(qmin), np.asarray(qmax),
                                       settings=settings, verbose=0)
        assert len(trace) == len(ref.trace)
        for a, b in zip(trace, ref.trace):
            assert np.allclose(a, b, atol=1e-7)
        assert [len(p) for p in params] == [len(p) for p in ref.instantiated_params]
        moved += np.linalg.norm(np.asarray(q) - qinit) > 1e-3
    assert moved >= 3


# ---- robot trajectory (C4) -------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c4(G, R):
    """C4: tx90_geom_test2.xml, robottrajopt_initial.path remeshed x6 -> 67 milestones, chair cloud only
    (gridres=None), link grids/clouds at 0.02 (robottrajopt.py:13-17, 70-71, 101-131)."""
    from semiinfiniteoptimization_b200._host.robot import WorldModel
    from semiinfiniteoptimization_b200._host.trajectory import RobotTrajectory
    paths = json.load(open(os.path.join(ROOT, "semiinfiniteoptimization_b200", "data", "paths.json")))
    p0 = paths["resources/robottrajopt_initial.path"]

    def build(mod, is_ref, chair_pos=None):
        w = WorldModel()
        w.readFile("tx90_geom_test2.xml")
        robot, chair = w.robot(0), w.rigidObject(0)
        if chair_pos is not None:
            chair.setTransform(chair.getTransform()[0], list(chair_pos))
        traj = RobotTrajectory(robot, p0["times"], p0["milestones"])
        times = []
        for i in range(len(traj.times) - 1):
            for k in range(6):
                times.append(traj.times[i] + k / 6.0 * (traj.times[i + 1] - traj.times[i]))
        times.append(traj.times[-1])
        traj = traj.remesh(times)[0]
        if is_ref:
            kin = mod.RefRobotKinematicsCache(robot, 0.02, 0.02)
            tc = mod.RefRobotTrajectoryCache(kin, 10 * 6 + 6 - 1)
            env = mod.RefPDG(chair.geometry(), None, 0.02)
            con = mod.RefRobotTrajectoryCollisionConstraint([env], tc)
        else:
            kin = mod.RobotKinematicsCache(robot, 0.02, 0.02)
            tc = mod.RobotTrajectoryCache(kin, 10 * 6 + 6 - 1)
            env = mod.PenetrationDepthGeometry(chair.geometry(), None, 0.02)
            con = mod.RobotTrajectoryCollisionConstraint([env], tc)
        tc.qstart, tc.qend = traj.milestones[0][:], traj.milestones[-1][:]
        tc.tstart, tc.tend = traj.times[0], traj.times[-1]
        x = [v for q in traj.milestones[1:-1] for v in q]
        return robot, tc, con, traj, x
    return build


# the reference scene keeps the chair 0.35 m clear of the (synthetic, thin) links; this placement puts it across the
# initial path so the trajectory optimisation has collisions to resolve
CHAIR_IN_PATH = (-0.68, -0.38, 0.5)


def test_trajectory_cache_matches(G, R, c4):
    (robot, tc, con, traj, x), (robot2, tc2, con2, traj2, x2) = c4(G, False), c4(R, True)
    assert len(traj.milestones) == 67 and len(x) == 65 * 7
    assert np.allclose(tc.lipschitzConstants, tc2.lipschitzConstants, atol=1e-12)
    tc.set(x)
    tc2.set(x2)
    assert np.allclose(np.array(tc.interMilestoneLipschitzBounds), np.array(tc2.interMilestoneLipschitzBounds), atol=1e-12)
    for i in (0, 13, 66):
        for l in range(7):
            assert np.allclose(tc.linkTransforms[i][l][0], tc2.linkTransforms[i][l][0], atol=1e-12)
            assert np.allclose(tc.linkTransforms[i][l][1], tc2.linkTransforms[i][l][1], atol=1e-12)
    tc.clear()
    tc2.clear()


@pytest.mark.parametrize("chair_pos", [None, CHAIR_IN_PATH])
def test_trajectory_minvalue_dense_equals_reference_bisection(G, R, c4, chair_pos):
    """The reference's milestone sweep + Lipschitz time bisection (CPU, literal) and the dense device sweep over the
    same dyadic sample set return the same (dmin, link, env, t, point)."""
    (robot, tc, con, traj, x), (robot2, tc2, con2, traj2, x2) = c4(G, False, chair_pos), c4(R, True, chair_pos)
    rng = np.random.default_rng(0)
    for trial in range(3):
        xa = np.asarray(x) + (rng.normal(0, 0.05, len(x)) if trial else 0)
        con.setx(list(xa))
        con2.setx(list(xa))
        a, b = con.minvalue(list(xa)), con2.minvalue(list(xa))
        assert abs(a[0] - b[0]) <= 1e-9, (a, b)
        assert a[1][0] == b[1][0] and a[1][1] == b[1][1] and abs(a[1][2] - b[1][2]) <= 1e-12
        assert np.allclose(a[1][3:], b[1][3:], atol=1e-9)
        assert con.minvalue(list(xa), a[0] - 1e-6) == [] and con2.minvalue(list(xa), a[0] - 1e-6) == []
        # value / df_dx at the arg-min and at perturbed parameters
        ys = [a[1]] + [(a[1][0], 0, float(np.clip(a[1][2] + dt, 0.001, 0.999))) + tuple(np.asarray(a[1][3:]) + dp)
                       for dt, dp in zip(rng.normal(0, 0.05, 4), rng.normal(0, 0.02, (4, 3)))]
        vb, rows = con.value_batch(list(xa), ys), con.df_dx_batch(list(xa), ys)
        assert abs(vb[0] - a[0]) <= 1e-9
        for k, y in enumerate(ys):
            assert abs(vb[k] - con2.value(list(xa), y)) <= 1e-10
            assert abs(con.value(list(xa), y) - vb[k]) <= 1e-12
            r2 = con2.df_dx(list(xa), y)
            assert rows[k].shape == (1, 455) and abs(rows[k] - r2).max() <= 1e-9
        con.clearx()
        con2.clearx()


def test_c4_solutions_match(G, R, c4):
    from semiinfiniteoptimization_b200 import sip
    (robot, tc, con, traj, x), (robot2, tc2, con2, traj2, x2) = c4(G, False, CHAIR_IN_PATH), c4(R, True, CHAIR_IN_PATH)
    settings = sip.SemiInfiniteOptimizationSettings()
    settings.max_iters = 4
    settings.minimum_constraint_value = -0.02
    out = G.optimizeCollFreeTrajectory(tc, traj, env=None, constraints=[con], settings=settings, verbose=0)
    res = G.optimizeCollFreeTrajectory.last_result
    qmin, qmax = robot2.getJointLimits()
    xmin, xmax = np.hstack([qmin] * 65), np.hstack([qmax] * 65)
    ref = sip.optimizeSemiInfinite(G.TrajectoryLengthObjective(tc), [con2], np.asarray(x2), xmin, xmax, settings=settings, verbose=0)
    assert len(res.trace) == len(ref.trace)
    for a, b in zip(res.trace, ref.trace):
        assert np.allclose(a, b, atol=1e-6)
    assert len(out[0].milestones) == 67
    assert res.gx0[0] < -0.01 and len(res.trace) >= 3          # it really started in collision and moved


def _brute_dense_sweep(R, robot2, tc2, con2, xa, levels):
    """The idea of the reference's commented brute-force check (gopt:865-888), at the benchmark's sampling density:
    every milestone and every interior sample k / 2^levels of every edge x every active link, each a brute-force
    cloud-vs-grid minimum on the CPU oracle (oracle.fk for the link frames, oracle.min_distance for the minima; the
    visiting order -- time-major, then link -- with a strict '<' decides ties, as in the reference's loops).
    Independent of the product: no device, no scheduler."""
    from oracle import oracle as O
    tc2.set(list(xa))
    traj = tc2.trajectory
    ms = np.asarray(traj.milestones, dtype=np.float64)
    times = np.asarray(traj.times, dtype=np.float64)
    K = 1 << levels
    us = np.arange(K) / K
    q = np.vstack([(ms[:-1, None, :] + us[None, :, None] * (ms[1:, None, :] - ms[:-1, None, :])).reshape(-1, ms.shape[1]), ms[-1:]])
    t = np.concatenate([(times[:-1, None] + us[None, :] * (times[1:, None] - times[:-1, None])).reshape(-1), times[-1:]])
    links = list(con2.activeLinks) if hasattr(con2, "activeLinks") else [l for l in range(1, robot2.numLinks())]
    orob = O.Robot([robot2.link(i).getParent() for i in range(robot2.numLinks())],
                   [se3_to12(robot2.link(i).getParentTransform()) for i in range(robot2.numLinks())],
                   [robot2.link(i).getAxis() for i in range(robot2.numLinks())])
    TL = O.fk(orob, q)                                     # (S, n_links, 12)
    best = (np.inf, None)
    for j, env in enumerate(con2.envs):
        Tenv = np.asarray(env.T12(), dtype=np.float64).reshape(3, 4)
        grids = [tc2.kinematics.geometry[l].G for l in links]
        T = TL[:, links, :].reshape(-1, 3, 4)
        Ti = np.empty_like(T)
        Ti[:, :, :3] = np.transpose(T[:, :, :3], (0, 2, 1))
        Ti[:, :, 3] = -np.einsum('bji,bj->bi', T[:, :, :3], T[:, :, 3])
        Trel = np.empty_like(T)
        Trel[:, :, :3] = Ti[:, :, :3] @ Tenv[:, :3]
        Trel[:, :, 3] = Ti[:, :, :3] @ Tenv[:, 3] + Ti[:, :, 3]
        gob = np.tile(np.arange(len(links), dtype=np.int32), len(q))
        dmin, arg = O.min_distance(grids, env.cloud32, Trel.reshape(-1, 12), grid_of_b=gob)
        flat = int(np.argmin(dmin))                        # first occurrence = time-major, then link
        if dmin[flat] < best[0]:
            s, k = divmod(flat, len(links))
            p = env.cloud32[int(arg[flat])].astype(np.float64)
            best = (float(dmin[flat]), (links[k], j, float(t[s])) + tuple(float(v) for v in Tenv[:, :3] @ p + Tenv[:, 3]))
    tc2.clear()
    return best


def se3_to12(T):
    R, t = np.asarray(T[0], dtype=float).reshape(3, 3).T, np.asarray(T[1], dtype=float)
    return np.column_stack([R, t]).reshape(-1)


@pytest.mark.parametrize("chair_pos", [None, CHAIR_IN_PATH])
def test_trajectory_minvalue_256_samples_per_edge_equals_brute_dense_oracle_sweep(G, R, c4, chair_pos):
    """BASELINE.json configs[3] as the benchmark states it -- 256 time samples per edge (LEVELS = 8): the device sweep
    (Lipschitz-scheduled and dense) against a brute dense sweep on the CPU oracle.  VERDICT r1 weak #3."""
    (robot, tc, con, traj, x), (robot2, tc2, con2, traj2, x2) = c4(G, False, chair_pos), c4(R, True, chair_pos)
    rng = np.random.default_rng(11)
    con.LEVELS = 8
    try:
        for trial in range(3):
            xa = np.asarray(x) + (rng.normal(0, 0.05, len(x)) if trial else 0)
            want = _brute_dense_sweep(R, robot2, tc2, con2, xa, 8)
            for sched in (True, False):
                G.LIPSCHITZ_SCHEDULE = sched
                con.setx(list(xa))
                got = con.minvalue(list(xa))
                con.clearx()
                assert abs(got[0] - want[0]) <= 1e-9, (got, want)
                assert got[1][0] == want[1][0] and got[1][1] == want[1][1] and abs(got[1][2] - want[1][2]) <= 1e-12
                assert np.allclose(got[1][3:], want[1][3:], atol=1e-9)
    finally:
        G.LIPSCHITZ_SCHEDULE = True
        con.LEVELS = 7


@pytest.mark.parametrize("chair_pos", [None, CHAIR_IN_PATH])
def test_lipschitz_scheduled_sweep_equals_dense_sweep(G, R, c4, chair_pos):
    """Row f2: milestones first, interior samples only for (edge, link) pairs whose Lipschitz lower bound can still win
    (sio_robot_pairs_min_distance) -- same (dmin, link, env, t, point) as the one-call dense sweep, far fewer queries."""
    robot, tc, con, traj, x = c4(G, False, chair_pos)
    rng = np.random.default_rng(3)
    for trial in range(3):
        xa = list(np.asarray(x) + (rng.normal(0, 0.05, len(x)) if trial else 0))
        for levels in (7, 8):
            con.LEVELS = levels
            con.setx(xa)
            G.LIPSCHITZ_SCHEDULE = True
            a = con.minvalue(xa)
            qa = con.last_sweep_queries
            G.LIPSCHITZ_SCHEDULE = False
            try:
                b = con.minvalue(xa)
                qb = con.last_sweep_queries
            finally:
                G.LIPSCHITZ_SCHEDULE = True
            assert a == b, (a, b)
            assert qa < 0.6 * qb
            # the caller's bound prunes further and yields [] exactly when the minimum is not below it
            assert con.minvalue(xa, a[0] - 1e-6) == [] and con.minvalue(xa, a[0] + 1e-6) == a
            con.clearx()
    con.LEVELS = 7


def test_active_set_through_the_python_api(G, c1):
    """north_star: "global minimum plus all points under the bound, with compacted index output" -- through the drop-in
    classes: PenetrationDepthGeometry.under and ObjectCollisionConstraint.minvalue_all give the oracle's under-bound set
    (indices identical, float64 values), deepest first; all_under makes `minvalue` itself answer with the list."""
    from oracle import oracle as O
    g1, g2, r1, r2 = c1
    grid = O.Grid(g1.grid.getVolumeGrid().values, g1.grid.getVolumeGrid().bmin, g1.grid.getVolumeGrid().bmax)
    c = G.ObjectCollisionConstraint(g1, g2)
    sizes = []
    for T in c1_poses(5, seed=11):
        c.setx(T)
        dmin = c.minvalue(T)[0]
        for bound in (dmin - 1e-3, dmin + 1e-6, dmin + 0.02, 0.0, 0.1):
            d, idx, pts = g2.under(g1, bound)
            T_rel = (np.linalg.inv(mat4(T)) @ mat4(g2.getTransform()))[:3].reshape(1, 12)      # grid-local <- cloud-local
            n_u, ub, ui, ud = O.under(grid, g2._dev.cloud32, T_rel, np.array([bound]))
            assert n_u == len(ui)
            assert np.array_equal(idx, ui) and np.abs(d - ud).max(initial=0.0) <= 1e-12
            sizes.append(len(idx))
            pairs = c.minvalue_all(T, bound)
            assert len(pairs) == len(ui)
            assert [p[0] for p in pairs] == sorted(p[0] for p in pairs)
            if len(pairs):
                assert abs(pairs[0][0] - dmin) <= 1e-12 and np.allclose(pairs[0][1], c.minvalue(T)[1], atol=1e-12)
                k = int(np.argmin(ud))
                assert np.allclose(pairs[0][1], pts[k], atol=0)
                for f, y in pairs[:: max(1, len(pairs) // 7)]:
                    assert abs(c.value(T, y) - f) <= 1e-9
            c.all_under = True
            try:
                assert c.minvalue(T, bound) == pairs
                one = c.minvalue(T)                              # no bound: the single deepest pair, as always
                assert not hasattr(one[0], '__iter__')
            finally:
                c.all_under = False
    assert max(sizes) > 500 and min(sizes) == 0
    # the float64 validation mode of the north_star answers with the same set
    T = c1_poses(1, seed=11)[0]
    c.setx(T)
    want = g2.under(g1, 0.0)
    G.set_validation_mode(True)
    try:
        got = g2.under(g1, 0.0)
    finally:
        G.set_validation_mode(False)
    assert np.array_equal(got[1], want[1]) and np.abs(got[0] - want[0]).max(initial=0.0) <= 1e-12 and len(want[1]) > 100


def mat4(T):
    """Klampt se3 (R column-major 9-list, t) as a 4x4 matrix."""
    A = np.eye(4)
    A[:3, :3] = np.asarray(T[0], dtype=np.float64).reshape(3, 3).T
    A[:3, 3] = T[1]
    return A


def test_multi_constraint_over_device_members_equals_the_reference_class():
    from tests.test_multi_cpu import check_against_golden
    check_against_golden('device', 1e-9)


def test_sip_with_all_under_members_reaches_feasibility(G, c1):
    """The batched extension of SURVEY A.6 end to end: a MultiSemiInfiniteConstraint whose member answers every oracle
    call with its whole active set (capped by the solver's bound) still drives optimizeSemiInfinite to a collision-free
    pose, in no more outer iterations than the one-pair mode."""
    from semiinfiniteoptimization_b200 import sip
    g1, g2, r1, r2 = c1
    T0 = c1_poses(1, seed=0)[0]
    iters = {}
    for mode in (False, True):
        c = G.ObjectCollisionConstraint(g1, g2)
        c.all_under = mode
        M = G.MultiSemiInfiniteConstraint([c], all_negative=True)
        st = sip.SemiInfiniteOptimizationSettings()
        res = sip.optimizeSemiInfinite(G.ObjectPoseObjective(T0), [M], T0, settings=st, verbose=0)
        c.setx(res.x)
        assert c.minvalue(res.x)[0] > -5e-3
        iters[mode] = res.num_iterations
    assert iters[True] <= iters[False]


def test_min_dist_front_end_runs_on_the_device_classes(G, c1):
    """optimizeCollFreeMinDist (gopt:1083-1122): one plain constraint min_y distance >= 0 through the generic loop, every
    evaluation a K2 sweep; ends collision-free from a pose in collision and returns optimizeCollFree's shape with an
    empty constraint list."""
    g1, g2, r1, r2 = c1
    T = c1_poses(6, seed=0)[1]
    g1.setTransform(T)
    d0 = g2.distance(g1)[0]
    assert d0 < -0.05
    out = G.optimizeCollFreeMinDist(g1, g2, T, verbose=0)
    assert len(out) == 4 and out[3] == [] and len(out[1]) == len(out[2]) >= 2
    g1.setTransform(out[0])
    assert g2.distance(g1)[0] > -5e-3
    assert G.optimizeCollFreeMinDist(g1, g2, T, verbose=0, want_trace=False, want_times=False, want_constraints=False)[1] is not None
