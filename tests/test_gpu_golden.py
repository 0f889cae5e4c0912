"""Device-backed drop-in classes against the committed golden vectors (tests/golden/: outputs of the
REFERENCE's own Python, see tools/make_golden.py) on BASELINE.json's configurations C1-C5.

Tolerances: single oracle calls agree with the float64 goldens to 1e-9 (the device's float64
re-check makes minima and arg-min points exact; the value kernels differ only in operation order);
whole SIP traces to 1e-7 (the QP amplifies last-bit differences)."""
import numpy as np
import pytest

from tests import scenes

pytestmark = pytest.mark.gpu


def _pose_close(a, b, tol):
    return np.allclose(a[0], b[0], atol=tol, rtol=0) and np.allclose(a[1], b[1], atol=tol, rtol=0)


def _pose_cases(ctx, golden, scene_fn, settings_kw=None):
    from semiinfiniteoptimization_b200 import sip
    mod, g1, g2, con = scene_fn('device')
    nontrivial = 0
    for case in golden["cases"]:
        T = (case["Tinit"][0], case["Tinit"][1])
        Td = (case["Tdes"][0], case["Tdes"][1]) if "Tdes" in case else None
        st = sip.SemiInfiniteOptimizationSettings()
        for k, v in (settings_kw or {}).items():
            setattr(st, k, v)
        out = mod.optimizeCollFree(g1, g2, T, Td, settings=st, verbose=0)
        assert len(out) == 4 and len(out[1]) == len(case["trace"])
        for a, b in zip(out[1], case["trace"]):
            assert _pose_close(a, b, 1e-7)
        assert _pose_close(out[0], case["T"], 1e-7)
        assert len(out[3]) == len(case["params"])
        for ya, yb in zip(out[3], case["params"]):
            assert np.allclose(ya, yb, atol=1e-9, rtol=0)        # arg-min points are cloud points: identical
        nontrivial += len(case["trace"]) > 3
    return nontrivial


def test_c1_solutions_equal_the_reference(ctx):
    assert _pose_cases(ctx, scenes.load_golden("c1_pose.json"), scenes.c1_scene) >= 5


def test_c1_with_separate_targets_equals_the_reference(ctx):
    """optimizeCollFree(obj, env, Tinit, Tdes), Tdes != Tinit (tests/golden/c1_pose_targets.json), on the device classes one
    problem at a time AND through the device-resident lock-step solver with its separate target array."""
    from semiinfiniteoptimization_b200 import batch, sip
    golden = scenes.load_golden("c1_pose_targets.json")
    assert _pose_cases(ctx, golden, scenes.c1_scene) >= 10
    cases = golden["cases"]
    mod, g1, g2, con = scenes.c1_scene('device')
    res = batch.optimizeCollFreeBatchDevice(g1, g2, [(c["Tinit"][0], c["Tinit"][1]) for c in cases],
                                            [(c["Tdes"][0], c["Tdes"][1]) for c in cases], settings=sip.SemiInfiniteOptimizationSettings())
    want_T = np.array([_se3_to12(c["T"]) for c in cases])
    assert [s for s in res.status] == [c["status"] for c in cases]
    assert np.array_equal(np.asarray(res.num_iterations), np.array([c["num_iterations"] for c in cases]))
    assert np.abs(res.T - want_T).max() <= 1e-6
    assert np.allclose(res.fx, [c["fx"] for c in cases], atol=1e-6, rtol=0)
    for p, c in enumerate(cases):
        assert np.allclose(np.array(res.instantiated_params[p]), np.array(c["params"]), atol=1e-9, rtol=0)


def test_c2_solutions_equal_the_reference(ctx):
    assert _pose_cases(ctx, scenes.load_golden("c2_pose.json"), scenes.c2_scene, {"max_iters": 20}) >= 4


def _se3_to12(T):
    R, t = np.asarray(T[0], dtype=float), np.asarray(T[1], dtype=float)
    return np.concatenate([np.column_stack([R.reshape(3, 3).T, t]).reshape(-1)])       # Klampt R is column-major


def test_c5_solutions_equal_the_reference(ctx):
    """BASELINE.json configs[4]: 64 of the 65,536 problems, reference's own optimizeCollFree (tests/golden/c5_pose.json)
    against (a) the drop-in optimizeCollFree on the device-backed classes, one problem at a time, and (b) the
    device-resident lock-step solver optimizeCollFreeBatchDevice: status, outer iterations, final pose <= 1e-6,
    instantiated parameters identical."""
    from semiinfiniteoptimization_b200 import batch, sip
    golden = scenes.load_golden("c5_pose.json")
    cases = golden["cases"]
    assert len(cases) >= 64 and sum(c["iterations"] >= 5 for c in cases) >= 20
    mod, cube, scene, con = scenes.c5_scene('device', context=ctx)
    poses = [(c["Tinit"][0], c["Tinit"][1]) for c in cases]
    # (a) the reference's call shape, problem by problem (every 4th case: each is up to 50 x 2 sweeps of 262,144 points)
    for case in cases[::4]:
        T = (case["Tinit"][0], case["Tinit"][1])
        out = mod.optimizeCollFree(cube, scene, T, settings=sip.SemiInfiniteOptimizationSettings(), verbose=0)
        assert len(out[1]) - 1 == case["iterations"]
        assert _pose_close(out[0], case["T"], 1e-6)
        assert len(out[3]) == len(case["params"])
        for ya, yb in zip(out[3], case["params"]):
            assert np.allclose(ya, yb, atol=1e-9, rtol=0)
        if "trace" in case:
            for a, b in zip(out[1], case["trace"]):
                assert _pose_close(a, b, 1e-6)
    # (b) all 64 at once on the device
    res = batch.optimizeCollFreeBatchDevice(cube, scene, poses, settings=sip.SemiInfiniteOptimizationSettings(), context=ctx)
    assert res.slot_overflows == 0
    want_T = np.array([_se3_to12(c["T"]) for c in cases])
    want_it = np.array([c["num_iterations"] for c in cases])          # res.num_iterations of the reference (sip.py:847, 1217)
    want_status = [c["status"] for c in cases]
    same_status = np.array([a == b for a, b in zip(res.status, want_status)])
    same_it = np.asarray(res.num_iterations) == want_it
    close = np.abs(res.T - want_T).max(1) <= 1e-6
    assert same_status.all(), [(i, res.status[i], want_status[i]) for i in np.nonzero(~same_status)[0]]
    assert same_it.all(), [(i, int(res.num_iterations[i]), int(want_it[i])) for i in np.nonzero(~same_it)[0]]
    assert close.all(), np.abs(res.T - want_T).max(1)[~close]
    n_par = np.array([len(c["params"]) for c in cases])
    assert np.array_equal(np.asarray(res.n_params), n_par)
    for p, c in enumerate(cases):
        if n_par[p]:
            assert np.allclose(np.array(res.instantiated_params[p]), np.array(c["params"]), atol=1e-9, rtol=0)
    gx = np.array([min(c["gx"]) if isinstance(c["gx"], list) else c["gx"] for c in cases], dtype=float)
    fin = np.isfinite(gx)
    assert np.allclose(np.asarray(res.gx)[fin], gx[fin], atol=1e-6, rtol=0)


def test_c3_solutions_equal_the_reference(ctx):
    from semiinfiniteoptimization_b200 import sip
    golden = scenes.load_golden("c3_robot.json")
    mod, robot, cons = scenes.c3_scene('device')
    for case in golden["cases"]:
        st = sip.SemiInfiniteOptimizationSettings()
        st.max_iters, st.minimum_constraint_value = 5, -0.02
        q, trace, times, params = mod.optimizeCollFreeRobot(robot, None, qdes=case["qdes"], qinit=[0.0] * 7, constraints=cons,
                                                            settings=st, verbose=0)
        assert len(trace) == len(case["trace"])
        assert np.allclose(np.array(trace), np.array(case["trace"]), atol=1e-7, rtol=0)
        assert np.allclose(q, case["q"], atol=1e-7, rtol=0)
        assert [len(p) for p in params] == [len(p) for p in case["params"]]
        for pa, pb in zip(params, case["params"]):
            assert np.allclose(np.array(pa).reshape(-1), np.array(pb).reshape(-1), atol=1e-9, rtol=0)


def test_c3_sampled_start_equals_the_reference(ctx):
    """optimizeCollFreeRobot(qinit='random-collision-free' | 'random') (gopt:1172-1207) against the reference's own run with
    the same `random.seed`: same number of draws (the generator's next value agrees), same start, same solution."""
    import random
    from semiinfiniteoptimization_b200 import sip
    golden = scenes.load_golden("c3_random_init.json")
    mod, robot, cons = scenes.c3_scene('device')
    took_several = 0
    for case in golden["cases"]:
        st = sip.SemiInfiniteOptimizationSettings()
        st.max_iters, st.minimum_constraint_value = 5, -0.02
        random.seed(case["seed"])
        out = mod.optimizeCollFreeRobot(robot, None, qdes=case["qdes"], qinit=case["qinit"], constraints=cons, settings=st, verbose=0)
        assert random.random() == case["draws_after"]
        assert len(out[1]) == len(case["trace"])
        assert np.allclose(np.array(out[1]), np.array(case["trace"]), atol=1e-7, rtol=0) and np.allclose(out[0], case["q"], atol=1e-7, rtol=0)
        took_several += len(case["trace"]) > 1
    assert took_several >= 2


def test_c4_trajectory_constraint_and_solution_equal_the_reference(ctx):
    from semiinfiniteoptimization_b200 import sip
    golden = scenes.load_golden("c4_traj.json")
    mod, robot, tc, con, traj, x = scenes.c4_scene('device')
    for probe in golden["minvalue_probes"]:
        xa = probe["x"]
        con.setx(xa)
        mv = con.minvalue(xa)
        assert abs(mv[0] - probe["dmin"]) <= 1e-9
        assert mv[1][0] == probe["param"][0] and mv[1][1] == probe["param"][1] and abs(mv[1][2] - probe["param"][2]) <= 1e-12
        assert np.allclose(mv[1][3:], probe["param"][3:], atol=1e-9, rtol=0)
        y = tuple(probe["param"])
        assert abs(con.value(xa, y) - probe["value_at_param"]) <= 1e-9
        row = con.df_dx(xa, y).tocoo()
        assert list(row.col) == probe["row_cols"] and np.allclose(row.data, probe["row_vals"], atol=1e-9, rtol=0)
        assert (con.minvalue(xa, mv[0] - 1e-6) == []) == probe["pruned_by_bound"]
        con.clearx()
    # single-link trajectory constraint (gopt:565-675): the dense device sweep covers the reference's k/64 bisection
    for probe in golden["link_probes"]:
        lc = mod.RobotLinkTrajectoryCollisionConstraint(robot.link(probe["link"]), con.envs[0], tc)
        xa = probe["x"]
        lc.setx(xa)
        mv = lc.minvalue(xa)
        assert (mv == []) == probe["empty"]
        if not probe["empty"]:
            assert abs(mv[0] - probe["dmin"]) <= 1e-9 and abs(mv[1][0] - probe["param"][0]) <= 1e-12
            assert np.allclose(mv[1][1:], probe["param"][1:], atol=1e-9, rtol=0)
            y = tuple(probe["param"])
            assert abs(lc.value(xa, y) - probe["value_at_param"]) <= 1e-9
            assert np.abs(np.asarray(lc.df_dx(xa, y)) - np.array(probe["row"])).max() <= 1e-9
            assert (lc.minvalue(xa, mv[0] - 1e-6) == []) == probe["pruned_by_bound"]
        lc.clearx()
    st = sip.SemiInfiniteOptimizationSettings()
    st.max_iters, st.minimum_constraint_value = 4, -0.02
    out = mod.optimizeCollFreeTrajectory(tc, traj, env=None, constraints=[con], settings=st, verbose=0)
    res = mod.optimizeCollFreeTrajectory.last_result
    assert len(res.trace) == len(golden["trace"])
    assert np.allclose(np.array(res.trace), np.array(golden["trace"]), atol=1e-6, rtol=0)
    assert len(out[0].milestones) == 67


def test_c4_run_to_its_end_equals_the_reference(ctx):
    """tests/golden/c4_traj_long.json: optimizeCollFreeTrajectory with max_iters 30 on the device classes stops where the
    reference's own run stops -- 14 outer iterations, the same trajectory (1e-6) and the same deepest (link, env, t, point)."""
    from semiinfiniteoptimization_b200 import sip
    golden = scenes.load_golden("c4_traj_long.json")
    mod, robot, tc, con, traj, x = scenes.c4_scene('device')
    st = sip.SemiInfiniteOptimizationSettings()
    st.max_iters, st.minimum_constraint_value = 30, -0.02
    out = mod.optimizeCollFreeTrajectory(tc, traj, env=None, constraints=[con], settings=st, verbose=0)
    res = mod.optimizeCollFreeTrajectory.last_result
    assert res.num_iterations == golden["iterations"]
    xs = [float(v) for q in out[0].milestones[1:-1] for v in q]
    assert np.abs(np.array(xs) - np.array(golden["x"])).max() <= 1e-6
    con.setx(xs)
    mv = con.minvalue(xs)
    con.clearx()
    assert abs(mv[0] - golden["dmin"]) <= 1e-6 and int(mv[1][0]) == golden["param"][0] and abs(mv[1][2] - golden["param"][2]) <= 1e-6
    assert len(out[3][0]) == len(golden["params"][0]) == 12
    assert np.allclose(np.array(out[3][0]), np.array(golden["params"][0]), atol=1e-6, rtol=0)


def test_class_level_goldens(ctx):
    g = scenes.load_golden("classes.json")
    mod, g1, g2, con = scenes.c1_scene('device')
    for rec in g["object_constraint"]:
        T = (rec["T"][0], rec["T"][1])
        con.setx(T)
        mv = con.minvalue(T)
        assert abs(mv[0] - rec["dmin"]) <= 1e-12 and np.allclose(mv[1], rec["argmin_point"], atol=1e-12, rtol=0)
        assert (con.minvalue(T, mv[0] - 1e-3) == []) == rec["bound_prunes"]
        vb, rows = con.value_batch(T, rec["points"]), con.df_dx_batch(T, rec["points"])
        for k, (p, v, row) in enumerate(zip(rec["points"], rec["values"], rec["rows"])):
            assert abs(con.value(T, p) - v) <= 1e-12 and abs(vb[k] - v) <= 1e-12
            assert np.abs(con.df_dx(T, p) - np.array(row)).max() <= 1e-9 and np.abs(rows[k] - np.array(row)).max() <= 1e-9
        d, cp1, cp2 = g1.distance(rec["points"][0])
        assert abs(d - rec["pdg_distance"][0]) <= 1e-12 and np.allclose(cp1, rec["pdg_distance"][1], atol=1e-10, rtol=0)
        assert np.allclose(g1.normal(rec["points"][0]), rec["pdg_normal"], atol=1e-10, rtol=0)
        con.clearx()
    mod, robot, cons = scenes.c3_scene('device')
    for rec in g["robot_link_constraints"]:
        q = rec["q"]
        for c in cons:
            c.setx(q)
        for c, per in zip(cons, rec["links"]):
            mv = c.minvalue(q)
            assert abs(mv[0] - per["dmin"]) <= 1e-12 and np.allclose(mv[1], per["argmin_point"], atol=1e-12, rtol=0)
            for p, v, row in zip(per["points"], per["values"], per["rows"]):
                assert abs(c.value(q, p) - v) <= 1e-12 and np.abs(c.df_dx(q, p) - np.array(row)).max() <= 1e-9
        for c in cons:
            c.clearx()


def test_oracle_known_answers_on_device(ctx):
    """K1 (float64), K2 min/arg-min/under-bound and K3 pose rows against oracle_kat.json."""
    from tests import util
    g = scenes.load_golden("oracle_kat.json")
    vg = util.random_grid(g["grid_seed"])
    gh = ctx.grid_create(vg.values, vg.bmin, vg.bmax)
    T = np.array(g["T"])
    pts = np.array(g["pts"])
    d, grad = ctx.query(gh, T[0], pts, flags=1)                 # SIO_F64
    assert np.abs(d - np.array(g["query_d"])).max() <= 1e-12 and np.abs(grad - np.array(g["query_grad"])).max() <= 1e-9
    rng = np.random.default_rng(g["cloud_seed"])
    rng.uniform(-0.6, 1.6, (40, 3))
    cloud = rng.uniform(-0.5, 1.5, (5000, 3)).astype(np.float32)
    ch = ctx.cloud_create(cloud)
    bound = np.array([np.inf if b is None else b for b in g["bound"]])
    for flags in (0, 1, 8):                                      # float32 bulk + float64 finalize, all-float64, pruned
        dmin, arg, n_under = ctx.min_distance(gh, ch, T, bound=bound, flags=flags)
        assert np.abs(dmin - np.array(g["dmin"])).max() <= 1e-12 and np.array_equal(arg, np.array(g["argmin"]))
        ub, ui, ud = ctx.under_fetch()
        assert np.array_equal(ub, np.array(g["under_b"])) and np.array_equal(ui, np.array(g["under_idx"]))
        assert np.abs(ud - np.array(g["under_d"])).max() <= 1e-12
        assert np.array_equal(n_under, np.bincount(np.array(g["under_b"], dtype=np.int64), minlength=len(T)))
    rd, rows = ctx.pose_rows(gh, T[1], [0, 9], pts[:9])
    assert np.abs(rd - np.array(g["pose_rows_d"])).max() <= 1e-12 and np.abs(rows - np.array(g["pose_rows"])).max() <= 1e-9
    ctx.cloud_destroy(ch)
    ctx.grid_destroy(gh)
