"""CPU suite: the oracle against closed-form answers and against its independent numpy restatement,
the host substrate (so3/se3, trajectories, robot FK), the QP solver, the SIP loop driven by
oracle-backed constraints, and the C-ABI surface (symbols only -- no compute without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def O(built):
    from oracle import oracle
    return oracle


# ---- C-ABI surface ----------------------------------------------------------------------------
def test_library_exports_every_declared_symbol(built):
    from semiinfiniteoptimization_b200 import _abi
    hdr = open(os.path.join(ROOT, "include", "sio_abi.h")).read()
    declared = set(re.findall(r"\b(sio_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_abi.ABI_SYMBOLS)
    lib = ctypes.CDLL(_abi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "libsio_b200.so does not export %s" % name
    assert lib.sio_abi_version() == int(re.search(r"#define SIO_ABI_VERSION (\d+)", hdr).group(1))


def test_no_cpu_fallback_without_device(built):
    """On a box without a GPU the product must fail loudly (SIO_ERR_NODEVICE), never compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from semiinfiniteoptimization_b200 import _abi
    with pytest.raises(_abi.SioError) as e:
        _abi.Context(0)
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "semiinfiniteoptimization_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn
                assert "liboracle" not in src and "orc_" not in src, fn


# ---- oracle: known answers and cross-restatement ----------------------------------------------------
def test_oracle_c_equals_numpy(O):
    vg = util.random_grid(3)
    G = O.Grid(vg.values, vg.bmin, vg.bmax)
    rng = np.random.default_rng(0)
    pts = rng.uniform(-1, 2, (5000, 3))
    for T in util.random_poses(1, 4, rot=1.0, trans=0.5):
        for norm in (True, False):
            d, g = O.query(G, T, pts, normalize=norm)
            d2, g2 = O.np_query(vg.values, vg.bmin, vg.bmax, T, pts, normalize=norm)
            assert np.abs(d - d2).max() < 1e-12
            assert np.abs(g - g2).max() < 1e-9


def test_oracle_sphere_known_answer(O):
    vg = util.sphere_grid(r=0.7, n=96)
    G = O.Grid(vg.values, vg.bmin, vg.bmax)
    rng = np.random.default_rng(1)
    pts = rng.uniform(-0.9, 0.9, (5000, 3))
    pts = pts[np.linalg.norm(pts, axis=1) > 0.2]
    d, g = O.query(G, np.eye(4)[:3].reshape(12), pts)
    r = np.linalg.norm(pts, axis=1)
    assert np.abs(d - (r - 0.7)).max() < 2e-3            # O(h^2) trilinear error, h = 2/96
    assert np.abs(g - pts / r[:, None]).max() < 5e-2
    # exact at sample positions
    h = vg.cell
    c = vg.bmin + (np.array([10, 20, 30]) + 0.5) * h
    assert abs(O.query(G, np.eye(4)[:3].reshape(12), [c], want_grad=False)[0] - vg.values[10, 20, 30]) < 1e-12


def test_oracle_outside_bbox_adds_distance(O):
    """Appendix A.2 step 2/4: d = v(clamp(p)) + |p - clamp(p)|, gradient = outward unit vector there."""
    from semiinfiniteoptimization_b200._host.geometry import analytic_sdf_grid
    vg = analytic_sdf_grid(lambda P: np.zeros(P.shape[:-1]) + 0.25, [0, 0, 0], [1, 1, 1], (4, 4, 4))
    G = O.Grid(vg.values, vg.bmin, vg.bmax)
    d, g = O.query(G, np.eye(4)[:3].reshape(12), [[2.0, 0.5, 0.5], [0.5, 0.5, 0.5], [1.0 + 3, 1.0 + 4, 0.5]])
    assert np.allclose(d, [1.25, 0.25, 5.25])
    assert np.allclose(g[0], [1, 0, 0]) and np.allclose(g[1], 0) and np.allclose(g[2], [0.6, 0.8, 0])


def test_oracle_gradient_matches_finite_differences(O):
    """The reference's own gradient check (utils.py:39-51: h=1e-4, eps=1e-3), on the raw gradient."""
    geom = util.fixture_geometry("m1062")
    vg = geom.convert("VolumeGrid", 0.02).getVolumeGrid()
    G = O.Grid(vg.values, vg.bmin, vg.bmax)
    rng = np.random.default_rng(2)
    pts = vg.bmin + rng.uniform(0.05, 0.95, (3000, 3)) * (vg.bmax - vg.bmin)
    T = util.random_poses(3, 1, rot=0.4, trans=0.02)[0]
    d, g = O.query(G, T, pts, normalize=False)
    h = 1e-6
    fd = np.stack([(O.query(G, T, pts + h * e, want_grad=False) - O.query(G, T, pts - h * e, want_grad=False)) / (2 * h)
                   for e in np.eye(3)], axis=1)
    err = np.abs(fd - g).max(axis=1)
    assert np.quantile(err, 0.98) < 1e-3           # points straddling a cell face see a kink


def test_oracle_min_distance_and_value_consistency(O):
    """sip.py:223-230's (commented) probe: value(x, argmin) == dmin."""
    vg = util.fixture_geometry("cube").convert("VolumeGrid", 0.05).getVolumeGrid()
    G = O.Grid(vg.values, vg.bmin, vg.bmax)
    cloud = util.fixture_geometry("m797").convert("PointCloud", 0.02).getPointCloud().points.astype(np.float32)
    T = util.invert12(util.random_poses(4, 6, rot=0.3, trans=0.3, center=(0.2, 0.2, 0.2)))
    dmin, arg = O.min_distance(G, cloud, T)
    dn, an = O.np_min_distance(vg.values, vg.bmin, vg.bmax, cloud, T)
    assert np.array_equal(arg, an) and np.abs(dmin - dn).max() < 1e-12
    for b in range(len(T)):
        v = O.query(G, T[b], [cloud[arg[b]].astype(np.float64)], want_grad=False)[0]
        assert abs(v - dmin[b]) < 1e-12
    # bound semantics: argmin = -1 iff dmin >= bound
    bound = np.full(len(T), np.median(dmin))
    d2, a2 = O.min_distance(G, cloud, T, bound=bound)
    assert np.array_equal(a2 >= 0, dmin < bound) and np.array_equal(d2, dmin)


def test_oracle_under_set_and_its_fingerprints(O):
    """The under-bound list (two-pass, all threads) against the numpy restatement, and the per-transform fingerprints
    the full-size GPU test compares instead of the explicit list."""
    vg = util.fixture_geometry("cube").convert("VolumeGrid", 0.05).getVolumeGrid()
    G = O.Grid(vg.values, vg.bmin, vg.bmax)
    cloud = (util.fixture_geometry("m797").convert("PointCloud", 0.02).getPointCloud().points + [0.3, 0, 0]).astype(np.float32)
    T = util.invert12(util.random_poses(10, 7, rot=0.3, trans=0.3, center=(0.2, 0.2, 0.2)))
    bound = np.array([0.01, np.inf, -0.02, 0.0, 0.05, -1.0, 0.02])
    n, ob, oi, od = O.under(G, cloud, T, bound)
    want_b, want_i, want_d = [], [], []
    for b in range(len(T)):
        d = O.cloud_query(G, cloud, T[b], want_grad=False)[0]
        sel = np.nonzero(d < bound[b])[0] if np.isfinite(bound[b]) else np.zeros(0, dtype=np.int64)
        want_b += [b] * len(sel); want_i += list(sel); want_d += list(d[sel])
    assert n == len(want_b) > 0
    assert np.array_equal(ob, want_b) and np.array_equal(oi, want_i) and np.array_equal(od, want_d)
    cnt, isum, dsum = O.under_counts(G, cloud, T, bound)
    assert np.array_equal(cnt, np.bincount(ob, minlength=len(T)))
    assert np.array_equal(isum, np.bincount(ob, weights=oi.astype(float), minlength=len(T)).astype(np.int64))
    assert np.abs(dsum - np.bincount(ob, weights=od, minlength=len(T))).max() < 1e-12
    n2, b2, _, _ = O.under(G, cloud, T, bound, cap=5)            # a capacity below the count: count exact, list truncated
    assert n2 == n and len(b2) == 5


@pytest.mark.parametrize("scene", ["cube_chair", "random_grid", "tiny"])
def test_oracle_pruned_min_distance_equals_brute_force(O, scene):
    """The hierarchy-pruned CPU arm (bench.py's "pruning on" baseline) returns what the brute-force sweep returns:
    same minimum, same lowest index, same "[]" decisions -- with fewer point queries."""
    rng = np.random.default_rng(5)
    if scene == "cube_chair":
        vg = util.fixture_geometry("cube").convert("VolumeGrid", 0.05).getVolumeGrid()
        cloud = (util.fixture_geometry("m797").convert("PointCloud", 0.02).getPointCloud().points + [0.3, 0, 0]).astype(np.float32)
        T = util.invert12(util.random_poses(4, 12, rot=0.5, trans=0.4, center=(0.2, 0.2, 0.2)))
    elif scene == "random_grid":
        vg = util.random_grid(3)
        cloud = rng.uniform(-0.6, 1.6, (5000, 3)).astype(np.float32)
        T = util.random_poses(6, 8, rot=0.5, trans=0.2)
    else:
        vg = util.sphere_grid()
        cloud = np.repeat(rng.uniform(-1.5, 1.5, (37, 3)), 3, axis=0).astype(np.float32)      # duplicates: exact ties
        T = util.random_poses(7, 5, rot=1.0, trans=0.5)
    G = O.Grid(vg.values, vg.bmin, vg.bmax)
    tree = O.CloudTree(cloud)
    d0, a0 = O.min_distance(G, cloud, T)
    d1, a1, ev = O.min_distance_pruned(G, tree, T)
    assert np.array_equal(a0, a1) and np.array_equal(d0, d1)
    if scene == "cube_chair":
        assert ev.mean() < 0.5 * len(cloud)
    bound = np.where(np.arange(len(T)) % 2 == 0, np.median(d0), np.inf)
    d2, a2 = O.min_distance(G, cloud, T, bound=bound)
    d3, a3, _ = O.min_distance_pruned(G, tree, T, bound=bound)
    assert np.array_equal(a2, a3) and np.array_equal(d2, d3)
    d4, a4, ev4 = O.min_distance_pruned(G, tree, T, bound=bound, use_bound=True)
    assert np.array_equal(a2, a4) and np.array_equal(d2[a2 >= 0], d4[a2 >= 0]) and ev4.sum() <= ev.sum()


def test_oracle_fk_and_chain_rows_vs_host_model(O):
    from semiinfiniteoptimization_b200._host.robot import load_arc_robot
    robot = load_arc_robot()
    R = O.Robot(robot.parents, robot.tparent, robot.axes, robot.jtypes)
    rng = np.random.default_rng(5)
    qs = np.array([[rng.uniform(lo, hi) if hi > lo else lo for lo, hi in zip(robot.qmin, robot.qmax)] for _ in range(10)])
    Tfk = O.fk(R, qs)
    for k, q in enumerate(qs):
        robot.setConfig(q)
        assert np.abs(Tfk[k].reshape(7, 3, 4) - robot.link_T).max() < 1e-12
    # Jacobian rows: oracle C vs the host model's getPositionJacobian
    link = 5
    vg = robot.link(link).geometry().convert("VolumeGrid", 0.02).getVolumeGrid()
    G = O.Grid(vg.values, vg.bmin, vg.bmax)
    q = qs[0]
    robot.setConfig(q)
    pts = np.array([robot.link(link).getWorldPosition(p) for p in rng.uniform(-0.2, 0.2, (20, 3))])
    d, rows = O.chain_rows(R, link, G, q, pts)
    Tl = robot.link_T[link]
    Ti = np.concatenate([Tl[:, :3].T, (-Tl[:, :3].T @ Tl[:, 3])[:, None]], axis=1).reshape(12)
    d0, g0 = O.query(G, Ti, pts)
    assert np.abs(d - d0).max() < 1e-12
    for i, p in enumerate(pts):
        Jp = np.array(robot.link(link).getPositionJacobian(robot.link(link).getLocalPosition(p)))
        assert np.abs(rows[i] - Jp.T @ (-g0[i])).max() < 1e-10
    # finite-difference check of the kinematic Jacobian itself
    lp = robot.link(link).getLocalPosition(pts[0])
    Jp = np.array(robot.link(link).getPositionJacobian(lp))
    for j in range(1, 7):
        dq = np.zeros(7)
        dq[j] = 1e-6
        robot.setConfig(q + dq)
        p1 = np.array(robot.link(link).getWorldPosition(lp))
        robot.setConfig(q - dq)
        p0 = np.array(robot.link(link).getWorldPosition(lp))
        assert np.abs((p1 - p0) / 2e-6 - Jp[:, j]).max() < 1e-6


# ---- host substrate -------------------------------------------------------------------------------------
def test_so3_se3_conventions():
    from semiinfiniteoptimization_b200._host import se3, so3
    R = so3.from_rotation_vector([0.2, -0.4, 0.7])
    M = so3.matrix(R)
    assert np.allclose(M @ M.T, np.eye(3)) and np.isclose(np.linalg.det(M), 1)
    assert np.allclose(so3.apply(R, [1, 2, 3]), M @ [1, 2, 3])          # column-major 9-list
    assert np.allclose(so3.rotation_vector(R), [0.2, -0.4, 0.7])
    T1 = (R, [0.1, 0.2, 0.3])
    T2 = (so3.from_rotation_vector([-0.1, 0.3, 0.2]), [1, -1, 0.5])
    e = se3.error(T1, T2)
    assert np.allclose(so3.matrix(so3.from_rotation_vector(e[:3])) @ so3.matrix(T2[0]), M)   # exp(e) R2 = R1
    assert np.allclose(e[3:], np.subtract(T1[1], T2[1]))
    assert np.allclose(se3.apply(se3.inv(T1), se3.apply(T1, [3, 2, 1])), [3, 2, 1])
    A = se3.to_rowmajor12(T1)
    assert np.allclose(A.reshape(3, 4)[:, :3], M) and np.allclose(se3.to_rowmajor12(se3.from_rowmajor12(A)), A)


def test_trajectory_segments_and_remesh():
    from semiinfiniteoptimization_b200._host.trajectory import Trajectory
    tr = Trajectory([0, 1, 3], [[0, 0], [1, 2], [3, 2]])
    assert tr.eval(0.5) == [0.5, 1.0] and tr.eval(2) == [2.0, 2.0]
    assert tr.eval(-1) == [0, 0] and tr.eval(9) == [3, 2]
    assert tr.getSegment(0.25) == (0, 0.25) and tr.getSegment(3) == (2, 0)
    tr2, idx = tr.remesh([0, 0.5, 1, 2, 3])
    assert tr2.times == [0, 0.5, 1, 2, 3] and tr2.milestones[3] == [2.0, 2.0] and idx == [0, 1, 2, 3, 4]


def test_mesh_conversions_are_deterministic_and_sane():
    g = util.fixture_geometry("cube")
    vg = g.convert("VolumeGrid", 0.05).getVolumeGrid()
    assert vg.dims == (30, 30, 30) and np.allclose(vg.cell, 0.05)
    # exact SDF of the unit cube at cell centres
    ax = [vg.bmin[a] + (np.arange(30) + 0.5) * 0.05 for a in range(3)]
    P = np.stack(np.meshgrid(*ax, indexing="ij"), -1)
    assert np.abs(vg.values - util.box_sdf(P - 0.5, [0.5] * 3)).max() < 1e-6
    pc = g.convert("PointCloud", 0.02).getPointCloud()
    assert 14000 < pc.numPoints() < 17000
    assert np.abs(util.box_sdf(pc.points - 0.5, [0.5] * 3)).max() < 1e-9            # samples lie on the surface
    pc2 = util.fixture_geometry("cube").convert("PointCloud", 0.02).getPointCloud()
    assert np.array_equal(pc.points, pc2.points)


# ---- QP -------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,res,pcres", [("cube", 0.05, 0.02), ("m797", 0.08, 0.05)])
def test_host_mesh_conversions_equal_the_independent_restatement(name, res, pcres):
    """Row f3's definition of mesh -> SDF and mesh -> surface cloud, host builders (hostgeom.c / geometry.py) against
    oracle/ref_meshconv.py (numpy, shares no code with them): identical samples, identical points in identical order."""
    from oracle import ref_meshconv as RM
    from semiinfiniteoptimization_b200._host import geometry as HG
    HG.set_device_converters(None, None)
    HG.set_distance_grid_builder(None)
    g = util.fixture_geometry(name)
    vg = g.convert("VolumeGrid", res).getVolumeGrid()
    h = (vg.bmax - vg.bmin) / np.array(vg.dims)
    ref = RM.mesh_sdf(g.data.verts, g.data.tris, vg.bmin, h, vg.dims)
    assert np.array_equal(ref, vg.values)
    assert (ref < 0).any() and (ref > 0).any()
    pc = g.convert("PointCloud", pcres).getPointCloud().points
    assert np.array_equal(pc, RM.surface_sample(g.data.verts, g.data.tris, pcres))


def test_qp_matches_kkt_on_random_problems():
    from semiinfiniteoptimization_b200._host import qp
    rng = np.random.default_rng(7)
    for trial in range(20):
        n, m = 6, rng.integers(1, 12)
        W = rng.uniform(0.1, 2, n)
        xdes = rng.normal(size=n)
        A = rng.normal(size=(m, n))
        b = rng.normal(size=m) * 0.3
        P = np.diag(W + 1e-3)
        q = -W * xdes
        Afull = np.vstack([A, np.eye(n)])
        l = np.concatenate([-b + 1e-3, -0.5 * np.ones(n)])
        u = np.concatenate([np.full(m, np.inf), 0.5 * np.ones(n)])
        r = qp.solve_qp(P, q, Afull, l, u)
        if r.status != "solved":
            continue
        x = r.x
        assert np.all(Afull @ x >= l - 1e-4) and np.all(Afull @ x <= u + 1e-4)
        # optimality: no feasible descent direction found by projected random probing
        f = lambda v: 0.5 * v @ P @ v + q @ v
        for _ in range(200):
            dxx = rng.normal(size=n) * 1e-3
            xn = x + dxx
            if np.all(Afull @ xn >= l) and np.all(Afull @ xn <= u):
                assert f(xn) >= f(x) - 1e-7


def test_qp_detects_infeasibility():
    from semiinfiniteoptimization_b200._host import qp
    P = np.eye(2)
    r = qp.solve_qp(P, np.zeros(2), np.array([[1.0, 0], [1.0, 0]]), np.array([1.0, -np.inf]), np.array([np.inf, 0.0]))
    assert r.status == "primal infeasible" and r.x is None


# ---- SIP loop on oracle-backed constraints (host logic) -----------------------------------------------------------
def test_sip_resolves_cube_vs_chair_penetration(O):
    from oracle.ref_geometryopt import RefObjectCollisionConstraint, RefPDG
    from semiinfiniteoptimization_b200 import sip
    from semiinfiniteoptimization_b200._host import so3
    from semiinfiniteoptimization_b200.geometryopt import ObjectPoseObjective
    cube = RefPDG(util.fixture_geometry("cube"), 0.05, 0.02)
    chair = RefPDG(util.fixture_geometry("m797"), 0.05, 0.02)
    chair.setTransform((so3.identity(), [1.2, 0, 0]))
    T = (so3.from_rotation_vector([0.1, -0.2, 0.15]), [0.7, 0.1, 0.2])
    c = RefObjectCollisionConstraint(cube, chair)
    assert c.eval_minimum(T) < -0.05                       # starts in deep penetration
    res = sip.optimizeSemiInfinite(ObjectPoseObjective(T), [c], T, verbose=0)
    assert min(res.gx) > -5e-3                             # the reference's own feasibility threshold (sip.py:1227)
    assert c.eval_minimum(res.x) > -5e-3
    assert len(res.trace) == len(res.trace_times) and res.trace[0] is T


def test_sip_reference_quirks():
    """No rows -> the QP returns xdes ignoring the trust region (sip.py:288); exclusion hash truncates toward zero."""
    from semiinfiniteoptimization_b200 import sip
    cd = sip.ConstraintGenerationData([object()])
    assert np.array_equal(cd.solve_qp(1.0, np.array([5.0, -7.0]), [], [], np.array([-.5, -.5]), np.array([.5, .5])), [5.0, -7.0])
    assert cd.instantiate(0, (0.0004, -0.0004, 1.0)) is True
    assert cd.instantiate(0, (-0.0009, 0.0009, 1.0003)) is False       # int() truncation merges across zero
    assert cd.instantiate(0, (0.0011, 0.0, 1.0)) is True


def test_interval_lipschitz_minimum():
    from oracle.ref_geometryopt import intervalLipschitzMinimum as f
    assert f(1.0, 2.0, 0) == 1.0
    assert np.isclose(f(1.0, 1.0, 2.0), 0.0)          # u = 1/2 -> a - K/2
    assert np.isclose(f(0.0, 5.0, 1.0), -0.5)         # u outside [0,1]


def test_native_batched_qp_equals_scalar_qp():
    """_host/csrc/hostqp.c (reduced n x n ADMM, OpenMP over problems) against the scalar solve_qp_osqp restatement
    (sip.py:294-378) on random SIP-step problems, slack fallbacks and row-less problems included."""
    from semiinfiniteoptimization_b200 import sip
    from semiinfiniteoptimization_b200._host import qp
    rng = np.random.default_rng(0)
    B, n = 400, 6
    W = rng.uniform(0.01, 1.0, (B, 1)) * np.ones((B, n))
    xdes = rng.normal(0, 0.2, (B, n))
    counts = rng.integers(0, 12, B)
    offs = np.concatenate([[0], np.cumsum(counts)])
    rows = rng.normal(0, 1, (offs[-1], n))
    depths = rng.normal(0.0, 0.05, offs[-1])
    trs = rng.uniform(0.05, 0.5, (B, 1)) * np.ones((B, n))
    dx, st, its = qp.sip_step_batch(W, xdes, offs, rows, depths, -trs, trs, reg=1e-3, return_iters=True)
    assert set(np.unique(st)) <= {0, 1, 2} and (st == 1).sum() > 5 and (st == 2).sum() == (counts == 0).sum()
    cd = sip.ConstraintGenerationData([None])
    cd.constraint_inflation = 1e-3
    for p in range(120):
        if counts[p] == 0:
            assert np.array_equal(dx[p], xdes[p])          # no rows: xdes, box ignored (sip.py:288)
            continue
        ref = cd.solve_qp_osqp(W[p], xdes[p], list(rows[offs[p]:offs[p + 1]]), list(depths[offs[p]:offs[p + 1]]), -trs[p], trs[p],
                               regularizationFactor=1e-3)
        assert np.abs(ref - dx[p]).max() <= 1e-8


def test_batch_result_builds_its_lists_on_first_use():
    """The device-resident driver returns arrays; the per-problem Python lists of the reference's result format
    (status strings, instantiated parameters) are derived from them on first use."""
    from semiinfiniteoptimization_b200.batch import BatchResult
    r = BatchResult(4, per_problem_lists=False)
    r.T = np.zeros((4, 12))
    r.status_codes = np.array([1, 0, 1, 0], dtype=np.int32)
    P = np.arange(4 * 3 * 3, dtype=np.float64).reshape(4, 3, 3)
    r.set_params_arrays(P, np.array([1, 0, 2, 3]))
    assert r.records().shape == (4, 16) and list(r.records()[:, 15]) == [1, 0, 2, 3]
    assert r.status == ['local optimum', 'not converged', 'local optimum', 'not converged']
    assert [len(p) for p in r.instantiated_params] == [1, 0, 2, 3] and r.instantiated_params[2][1] == list(P[2, 1])
    eager = BatchResult(2)
    assert eager.status == ['not converged'] * 2 and eager.instantiated_params == [[], []]


def _subtile_radii(P, order):
    n = len(order) // 128 * 128
    Q = P[order[:n]].reshape(-1, 128, 3).astype(np.float64)
    c = Q.mean(1, keepdims=True)
    rep = Q[np.arange(len(Q)), np.argmin(((Q - c) ** 2).sum(2), 1)][:, None, :]
    return np.sqrt(((Q - rep) ** 2).sum(2).max(1))


def test_cloud_order_is_a_permutation_with_tight_leaves():
    """sio_cloud_order (host-only part of sio_cloud_create): a permutation for every size incl. ties, deterministic, and
    -- the reason it exists -- 128-point leaves several times tighter than chunks of the plain Morton order, whose
    outliers (chunks straddling a jump of the curve) used to survive every sub-tile cull."""
    from semiinfiniteoptimization_b200 import _abi, workloads as W
    rng = np.random.default_rng(7)
    for n in (0, 1, 2, 127, 128, 129, 1000, 1025, 5000):
        P = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
        o = _abi.cloud_order(P)
        assert sorted(o.tolist()) == list(range(n))
        assert np.array_equal(o, _abi.cloud_order(P))
    same = np.tile(np.float32([0.3, -0.2, 0.9]), (777, 1))
    assert sorted(_abi.cloud_order(same).tolist()) == list(range(777))
    with pytest.raises(_abi.SioError):
        _abi.cloud_order(np.float32([[0, 0, np.nan]]))
    P = np.asarray(W.c5_scene_cloud(), dtype=np.float32)
    r = _subtile_radii(P, _abi.cloud_order(P))
    lo, hi = P.min(0), P.max(0)
    q = np.minimum(2097151.0, np.maximum(0.0, (P - lo) * (2097151.0 / (hi - lo).max()))).astype(np.uint64)

    def spread(v):
        v = v & np.uint64(0x1fffff)
        for sh, m in ((32, 0x1f00000000ffff), (16, 0x1f0000ff0000ff), (8, 0x100f00f00f00f00f), (4, 0x10c30c30c30c30c3), (2, 0x1249249249249249)):
            v = (v | (v << np.uint64(sh))) & np.uint64(m)
        return v
    key = (spread(q[:, 0]) << np.uint64(2)) | (spread(q[:, 1]) << np.uint64(1)) | spread(q[:, 2])
    rm = _subtile_radii(P, np.argsort(key, kind="stable"))
    assert r.max() < 0.5 < rm.max() and (r > 0.15).mean() < 0.05 < (rm > 0.15).mean() and np.median(r) < np.median(rm)


def test_small_qps_dense_assembly_equals_sparse_assembly_bit_for_bit():
    """sip.py assembles few-variable QPs as numpy arrays directly (no scipy.sparse round trip); handing the same rows over
    as sparse matrices forces the reference-shaped assembly.  Same matrices, same memory order, same bits -- strict and
    slack problems, with and without the box."""
    import scipy.sparse
    from semiinfiniteoptimization_b200 import sip
    rng = np.random.default_rng(21)
    cg = sip.ConstraintGenerationData([])
    slack_cases = 0
    for trial in range(60):
        n = int(rng.integers(3, 8))
        k = int(rng.integers(1, 12))
        W = rng.uniform(0.01, 1.0, n) if trial % 2 else float(rng.uniform(0.01, 1.0))
        xdes = rng.normal(0, 0.2, n)
        A = [rng.normal(0, 1, n) for _ in range(k)]
        b = list(rng.normal(0.0, 0.05, k) - (0.3 if trial % 3 == 0 else 0.0))      # deep penetrations -> slack re-solve
        box = (np.full(n, -0.3), np.full(n, 0.3)) if trial % 4 else (None, None)
        x_dense = cg.solve_qp_osqp(W, xdes, A, b, box[0], box[1], regularizationFactor=1e-3)
        A_sparse = [scipy.sparse.csr_matrix(r.reshape(1, -1)) for r in A]
        x_sparse = cg.solve_qp_osqp(W, xdes, A_sparse, b, box[0], box[1], regularizationFactor=1e-3)
        assert np.array_equal(x_dense, x_sparse), trial
        slack_cases += trial % 3 == 0
    assert slack_cases >= 15


def test_native_banded_qp_equals_the_interpreted_solver():
    """hqp_admm_banded (the trajectory QP natively: reduced banded x-update) against _host/qp.py solve_qp on random
    block-tridiagonal problems with the trajectory QP's structure -- rows spanning two neighbouring blocks, a box, strict
    and slack forms, feasible and infeasible: same status, same iteration count to within a check interval, and the same
    answer (after the polish the two agree to rounding because they identify the same active set)."""
    import scipy.sparse
    from semiinfiniteoptimization_b200._host import qp
    rng = np.random.default_rng(33)
    nb, k = 7, 12                                             # 12 blocks of 7 variables
    n = nb * k
    seen = set()
    for trial in range(24):
        t = np.zeros((k, k))
        for i in range(k):
            t[i, i] = 4 if 0 < i < k - 1 else 2
            if i:
                t[i, i - 1] = t[i - 1, i] = -2
        Px = scipy.sparse.kron(scipy.sparse.csc_matrix(t), scipy.sparse.identity(nb), format="csc") + 1e-2 * scipy.sparse.identity(n, format="csc")
        xdes = rng.normal(0, 0.3, n)
        q = -Px.dot(xdes)
        mr = int(rng.integers(2, 9))
        rows = np.zeros((mr, n))
        for r in range(mr):
            b0 = int(rng.integers(0, k - 1)) * nb
            rows[r, b0:b0 + 2 * nb] = rng.normal(0, 1, 2 * nb)
        depth = rng.normal(0.0, 0.05, mr) - (0.6 if trial % 3 == 0 else 0.0)
        lr = -depth + 1e-3
        if trial % 6 == 4:                                    # contradictory rows (strict form): primal infeasible
            rows[1] = -rows[0]
            lr[0], lr[1] = 1.0, 1.0
        lo, hi = np.full(n, -0.2), np.full(n, 0.2)
        pen = 0.0 if trial % 2 == 0 else 50.0
        R = scipy.sparse.csr_matrix(rows)
        if pen > 0:
            P = scipy.sparse.bmat([[Px, None], [None, pen * scipy.sparse.identity(mr)]], format="csc")
            A = scipy.sparse.bmat([[R, scipy.sparse.identity(mr)], [scipy.sparse.identity(n), None]], format="csc")
            qq = np.concatenate([q, np.zeros(mr)])
        else:
            P, A, qq = Px, scipy.sparse.vstack([R, scipy.sparse.identity(n)], format="csc"), q
        l, u = np.concatenate([lr, lo]), np.concatenate([np.full(mr, np.inf), hi])
        ref = qp.solve_qp(P, qq, A, l, u, eps_abs=1e-5)
        got = qp.solve_qp_structured(P, qq, A, l, u, n, mr, True, pen, eps_abs=1e-5)
        assert got is not None and got.status == ref.status, (trial, got.status, ref.status)
        assert abs(got.iters - ref.iters) <= 10
        seen.add((ref.status, pen > 0))
        if ref.status == "solved":
            assert np.abs(got.x - ref.x).max() <= 1e-7, trial
        elif ref.status == "max iterations":
            assert np.abs(got.x - ref.x).max() <= 1e-4
    assert ("solved", False) in seen and ("solved", True) in seen and any(s == "primal infeasible" for s, _ in seen)
    # a problem that is not banded is declined (the caller keeps the interpreted solver)
    dense_rows = scipy.sparse.csr_matrix(rng.normal(0, 1, (2, n)))
    A = scipy.sparse.vstack([dense_rows, scipy.sparse.identity(n)], format="csc")
    assert qp.solve_qp_structured(Px, q, A, np.concatenate([[0.0, 0.0], lo]), np.concatenate([[np.inf, np.inf], hi]), n, 2, True, 0.0) is None


def test_grid_export_formats(tmp_path):
    """grid_to_numpy / dump_grid_mat (gopt:21-36): the (m, n, p) sample array and the per-x-slice .mat file the
    reference's demo drivers write (robotposeopt.py:78)."""
    import scipy.io
    from semiinfiniteoptimization_b200 import geometryopt as G
    from semiinfiniteoptimization_b200._host.geometry import fixture_geometry
    ns = {"grid_to_numpy": G.grid_to_numpy, "dump_grid_mat": G.dump_grid_mat}
    grid = fixture_geometry("cube").convert("VolumeGrid", 0.1)
    vg = grid.getVolumeGrid()
    arr = ns["grid_to_numpy"](grid)
    assert arr.shape == tuple(vg.dims) and arr.dtype == np.float64
    assert arr[1, 2, 3] == vg.get(1, 2, 3) and arr[-1, 0, 5] == vg.get(vg.dims[0] - 1, 0, 5)
    fn = str(tmp_path / "grid.mat")
    ns["dump_grid_mat"](grid, fn)
    mat = scipy.io.loadmat(fn)
    assert sorted(k for k in mat if k.startswith("slice_")) == sorted("slice_%d" % i for i in range(vg.dims[0]))
    assert np.array_equal(mat["slice_2"], arr[2])


def test_prebound_and_fast_path_wrappers_wire_their_buffers_correctly():
    """No GPU: Context.min_distance_plan and the single-pose / single-configuration fast paths of pose_rows / chain_rows
    against a stand-in library that reads the inputs and fills the outputs THROUGH THE RAW POINTERS it is handed -- what the
    real entry points do.  Checks addresses, counts, flags, that outputs are copies where the scratch is reused, and that
    the plan's arrays are overwritten in place."""
    import ctypes as C
    from semiinfiniteoptimization_b200 import _abi

    def dbl(ptr, n):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(n,))

    class Lib:
        calls = []

        def sio_min_distance(self, h, ga, n_g, gob, cloud, T, bound, B, flags, dmin, arg, nu):
            g = np.ctypeslib.as_array(C.cast(ga, C.POINTER(C.c_int32)), shape=(n_g,))
            t = dbl(T, 12 * B)
            dbl(dmin, B)[:] = t.reshape(B, 12)[:, 3] + (dbl(bound, B) if bound else 0.0)       # something to recognise
            np.ctypeslib.as_array(C.cast(arg, C.POINTER(C.c_int64)), shape=(B,))[:] = np.arange(B) + int(g[0]) * 100 + cloud
            if nu:
                np.ctypeslib.as_array(C.cast(nu, C.POINTER(C.c_int32)), shape=(B,))[:] = flags
            self.calls.append(("min", B, flags))
            return 0

        def sio_pose_rows(self, h, grid, T, off, nT, pts, flags, d, rows):
            o = np.ctypeslib.as_array(C.cast(off, C.POINTER(C.c_int64)), shape=(nT + 1,))
            P = int(o[-1])
            p = dbl(pts, 3 * P).reshape(P, 3)
            dbl(d, P)[:] = p.sum(1) + dbl(T, 12)[3]
            if rows:
                dbl(rows, 6 * P).reshape(P, 6)[:] = np.arange(6) + p[:, :1]
            self.calls.append(("rows", nT, P, grid))
            return 0

        def sio_chain_rows(self, h, robot, lg, q, nq, cop, lop, pts, P, flags, d, rows):
            links = np.ctypeslib.as_array(C.cast(lop, C.POINTER(C.c_int32)), shape=(max(P, 1),))[:P]
            dbl(d, P)[:] = dbl(pts, 3 * P).reshape(P, 3)[:, 0] + links + dbl(q, 7)[6]
            if rows:
                dbl(rows, 7 * P).reshape(P, 7)[:] = dbl(q, 7)
            self.calls.append(("chain", nq, P, cop))
            return 0

    ctx = _abi.Context.__new__(_abi.Context)
    ctx.L, ctx.h = Lib(), None
    # the prebound K2 call
    plan = ctx.min_distance_plan(3, 4, 5, flags=9)
    T = np.arange(60.0).reshape(5, 12)
    d, a, n = plan(T, np.full(5, 0.5))
    assert np.array_equal(d, T[:, 3] + 0.5) and np.array_equal(a, np.arange(5) + 304) and np.array_equal(n, np.full(5, 9))
    d2, a2, n2 = plan(T + 1.0, np.zeros(5))
    assert d2 is d and np.array_equal(d, T[:, 3] + 1.0)                    # the plan's own arrays, overwritten in place
    nb = ctx.min_distance_plan([7, 8], 1, 2, with_bound=False, want_under=False)
    d3, a3, n3 = nb(T[:2])
    assert n3 is None and np.array_equal(d3, T[:2, 3]) and np.array_equal(a3, np.arange(2) + 701)
    with pytest.raises(ValueError):
        plan(T[:4], np.zeros(5))
    # pose_rows: fast path (one pose, few points) and general path agree; the fast path hands out copies
    pts = [[0.1, 0.2, 0.3], [1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]
    T1 = np.arange(12.0)
    d_f, r_f = ctx.pose_rows(11, T1, [0, 3], pts)
    d_k, _ = ctx.pose_rows(11, T1, [0, 2], pts[:2], want_rows=False)
    assert np.allclose(d_f, [0.6 + 3, 6 + 3, 15 + 3]) and r_f.shape == (3, 6) and np.allclose(r_f[:, 0], [0.1, 1.0, 4.0])
    assert np.allclose(d_k, d_f[:2]) and _ is None and np.allclose(d_f, [3.6, 9.0, 18.0])      # d_f survived the second call
    ctx._POSE_ROWS_CAP = 1                                                 # force the general path
    d_g, r_g = ctx.pose_rows(11, T1, [0, 3], pts)
    assert np.array_equal(d_g, d_f) and np.array_equal(r_g, r_f)
    d_e, r_e = _abi.Context.pose_rows(ctx, 11, T1, [0, 0], np.zeros((0, 3)))
    assert d_e.shape == (0,) and r_e.shape == (0, 6)
    # chain_rows likewise
    ctx._POSE_ROWS_CAP = 256
    q = [[0.0, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0]]
    lg = [1, 2, 3, 4, 5, 6, 7]
    d_c, r_c = ctx.chain_rows(0, 7, lg, q, 4, pts)
    assert np.allclose(d_c, [0.1 + 4 + 6, 1 + 4 + 6, 4 + 4 + 6]) and np.array_equal(r_c, np.tile(np.arange(7.0), (3, 1)))
    ctx._POSE_ROWS_CAP = -1
    d_cg, r_cg = ctx.chain_rows(0, 7, lg, q, 4, pts)
    assert np.array_equal(d_cg, d_c) and np.array_equal(r_cg, r_c)
    kinds = [c[0] for c in Lib.calls]
    assert kinds.count("min") == 3 and kinds.count("rows") == 4 and kinds.count("chain") == 2


def test_grouped_solver_slices_fill_one_set_of_arrays_and_errors_propagate(monkeypatch):
    """No GPU: GroupedPoseSolver.solve with the per-slice driver replaced by a stand-in -- contiguous slices whose sizes
    differ by at most one, every slice writing its rows of the SAME output arrays from its own thread, the totals summed,
    the device seconds the slowest slice's, and an exception raised in a worker thread re-raised by solve()."""
    import threading
    from semiinfiniteoptimization_b200 import batch
    seen = []
    gate = {"barrier": None}

    def fake(obj, env, Tinits, Tdess=None, settings=None, context=None, flags=None, cap=48, _out=None):
        n = len(Tinits)
        seen.append((obj, context, n, threading.get_ident(), None if Tdess is None else len(Tdess)))
        if gate["barrier"] is not None:
            gate["barrier"].wait(timeout=30)                # all slices are in flight at the same time, or this times out
        if obj == "bad":
            raise RuntimeError("slice failed")
        _out["T"][:] = np.asarray(Tinits) + 1.0
        _out["status"][:] = 1
        _out["iterations"][:] = obj
        for k in ("fx", "gx", "gx0"):
            _out[k][:] = obj
        _out["n_params"][:] = 2
        _out["params"][:] = obj
        r = batch.BatchResult(n, per_problem_lists=False)
        r.time_qp, r.time_cons, r.outer_iterations, r.point_queries, r.slot_overflows, r.qp_failures = 0.1 * obj, 0.2 * obj, n, 10 * n, 0, 0
        return r

    monkeypatch.setattr(batch, "optimizeCollFreeBatchDevice", fake)
    s = batch.GroupedPoseSolver.__new__(batch.GroupedPoseSolver)
    s.groups, s.objs, s.envs, s.ctxs = 3, [1, 2, 3], ["e"] * 3, ["c1", "c2", "c3"]
    T = np.arange(10 * 12, dtype=np.float64).reshape(10, 12)
    gate["barrier"] = threading.Barrier(3)
    res = s.solve(T, Tdess=T + 5.0, cap=4)
    gate["barrier"] = None
    sizes = sorted(c[2] for c in seen)
    assert sizes == [3, 3, 4] and sum(sizes) == 10 and all(c[4] == c[2] for c in seen)
    assert len({c[3] for c in seen}) == 3                                  # three threads
    assert np.array_equal(res.T, T + 1.0)
    assert list(res.num_iterations) == [1] * 3 + [2] * 3 + [3] * 4          # slice g owns a contiguous block, in order
    assert res.outer_iterations == 10 and res.point_queries == 100
    assert abs(res.time_qp - 0.3) < 1e-12 and abs(res.time_cons - 0.6) < 1e-12
    assert res.instantiated_params[9] == [[3.0, 3.0, 3.0]] * 2 and res.status[0] == 'local optimum'
    # fewer problems than slices: one slice per problem at most
    seen.clear()
    res = s.solve(T[:2], cap=4)
    assert sorted(c[2] for c in seen) == [1, 1]
    # a failing slice
    s.objs = [1, "bad", 3]
    with pytest.raises(RuntimeError, match="slice failed"):
        s.solve(T, cap=4)


def test_objective_combinators_equal_the_reference_classes():
    """objective.py:112-199: Add / Sub / MulConst / Mul / Div / PowConst on the same leaf objectives through this package's
    classes and through the reference's own (oracle/refrun; build container only), at random points: value, gradient,
    Hessian, minstep / integrate where the reference defines them -- including Div's '+' in the quotient rule, as is."""
    from oracle.refrun import load
    if not load.available():
        pytest.skip("reference sources are not present on this machine")
    from semiinfiniteoptimization_b200 import objective as mine
    ref = load.load()["objective"]

    class Leaf:                                             # duck-typed leaf usable by both families
        def __init__(self, A, b, c):
            self.A, self.b, self.c = A, b, c

        def value(self, x):
            return float(0.5 * x @ self.A @ x + self.b @ x + self.c)

        def gradient(self, x):
            return self.A @ x + self.b

        def hessian(self, x):
            return self.A

        def minstep(self, x):
            return -np.linalg.solve(self.A, self.gradient(x))

        def integrate(self, x, dx):
            return x + 2.0 * dx                             # recognisable: combinators integrate like their first operand

    rng = np.random.default_rng(4)
    n = 4

    def leaf():
        M = rng.normal(size=(n, n))
        return Leaf(M @ M.T + n * np.eye(n), rng.normal(size=n), 3.0 + rng.uniform())

    g, h = leaf(), leaf()
    pairs = [(mine.AddObjectiveFunction(g, h), ref.AddObjectiveFunction(g, h)),
             (mine.SubObjectiveFunction(g, h), ref.SubObjectiveFunction(g, h)),
             (mine.MulConstObjectiveFunction(g, 2.5), ref.MulConstObjectiveFunction(g, 2.5)),
             (mine.MulObjectiveFunction(g, h), ref.MulObjectiveFunction(g, h)),
             (mine.DivObjectiveFunction(g, h), ref.DivObjectiveFunction(g, h)),
             (mine.PowConstObjectiveFunction(g, 3), ref.PowConstObjectiveFunction(g, 3)),
             (mine.PowConstObjectiveFunction(g, 0.5), ref.PowConstObjectiveFunction(g, 0.5))]
    for a, b in pairs:
        assert type(a).__name__ == type(b).__name__
        for _ in range(5):
            x, dx = rng.normal(size=n), rng.normal(size=n)
            assert a.value(x) == b.value(x)
            assert np.array_equal(a.gradient(x), b.gradient(x))
            assert np.array_equal(a.integrate(x, dx), b.integrate(x, dx)) and np.array_equal(a.integrate(x, dx), x + 2.0 * dx)
            if isinstance(a, mine.DivObjectiveFunction):
                with pytest.raises(NotImplementedError):
                    a.hessian(x)
                with pytest.raises(NotImplementedError):
                    b.hessian(x)
            else:
                assert np.array_equal(a.hessian(x), b.hessian(x))
            if isinstance(a, mine.MulConstObjectiveFunction):
                assert np.array_equal(a.minstep(x), b.minstep(x))
    # the operators build the same classes (the reference's own `f * 2.0` forgets the factor and raises TypeError)
    q = mine.QuadraticObjectiveFunction(np.ones(n))
    assert isinstance(q + 1.0, mine.AddObjectiveFunction) and isinstance(1.0 + q, mine.AddObjectiveFunction)
    assert isinstance(q - q, mine.SubObjectiveFunction) and isinstance(2.0 - q, mine.SubObjectiveFunction)
    assert isinstance(q * 2.0, mine.MulConstObjectiveFunction) and isinstance(2.0 * q, mine.MulConstObjectiveFunction)
    assert isinstance(q * q, mine.MulObjectiveFunction) and isinstance(q / q, mine.DivObjectiveFunction)
    assert isinstance(q / 4.0, mine.MulConstObjectiveFunction) and (q / 4.0).h == 0.25 and isinstance(q ** 2, mine.PowConstObjectiveFunction)
    x = rng.normal(size=n)
    assert ((q + 1.0) * 2.0).value(x) == (q.value(x) + 1.0) * 2.0 and (2.0 - q).value(x) == 2.0 - q.value(x)
    with pytest.raises(NotImplementedError):
        q ** q


def test_gradient_checking_aids_equal_the_reference_functions(capsys):
    """utils.py:3-51 (numeric_gradient in its three methods, test_gradient) against the reference's own functions through
    oracle/refrun, on a smooth function of five variables -- including the one-sided methods' cumulative step, as is."""
    from oracle.refrun import load
    if not load.available():
        pytest.skip("reference sources are not present on this machine")
    from semiinfinite import utils as mine
    ref = load.load()["utils"]
    rng = np.random.default_rng(6)
    A = rng.normal(size=(5, 5))

    def f(x):
        return float(np.sin(x @ A @ x) + x[0] * x[3])

    def grad(x):
        return np.cos(x @ A @ x) * ((A + A.T) @ x) + np.array([x[3], 0, 0, x[0], 0])

    for _ in range(4):
        x = rng.normal(size=5) * 0.3
        for method in ("centered", "forward", "backward"):
            assert np.array_equal(mine.numeric_gradient(f, x, 1e-4, method), ref.numeric_gradient(f, x, 1e-4, method))
        assert np.abs(mine.numeric_gradient(f, x) - grad(x)).max() < 1e-6
        assert mine.test_gradient(f, grad, x) is True and ref.test_gradient(f, grad, x) is True
        wrong = lambda v: grad(v) + np.array([0, 0.01, 0, 0, 0])
        assert mine.test_gradient(f, wrong, x, name="w") is False
        out_mine = capsys.readouterr().out
        assert ref.test_gradient(f, wrong, x, name="w") is False
        assert out_mine == capsys.readouterr().out and "errorneous value 1" in out_mine
    with pytest.raises(ValueError):
        mine.numeric_gradient(f, np.zeros(5), method="sideways")
