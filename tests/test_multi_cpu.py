"""MultiSemiInfiniteConstraint (constraint.py:174-256) against the reference's own class, run through oracle/refrun on two
ObjectCollisionConstraints (tools/make_golden.py gen_multi -> tests/golden/multi_constraint.json).  The member constraints
here are the CPU restatement's (oracle/ref_geometryopt.py); tests/test_gpu_classes.py repeats it with the device classes."""
import numpy as np
import pytest

from tests import scenes


def check_against_golden(kind, tol):
    from semiinfiniteoptimization_b200.constraint import MultiSemiInfiniteConstraint
    from tools.make_golden import multi_inputs
    golden = scenes.load_golden("multi_constraint.json")
    mod, cons, poses, bounds = multi_inputs(kind)
    Ms = {flag: MultiSemiInfiniteConstraint(cons, all_negative=flag) for flag in (True, False)}
    assert Ms[True].num_co_parameters == golden["num_co_parameters"]
    for rec in golden["cases"]:
        M = Ms[rec["all_negative"]]
        T = (rec["T"][0], rec["T"][1])
        M.setx(T)
        mv = M.minvalue(T, rec["bound"])
        single = len(mv) > 0 and not hasattr(mv[0], '__iter__')
        assert single == rec["single"]
        pairs = [mv] if single else mv
        assert len(pairs) == len(rec["pairs"])
        for (f, y), (f0, y0), v0, r0 in zip(pairs, rec["pairs"], rec["value"], rec["df_dx"]):
            assert abs(f - f0) <= tol and np.abs(np.asarray(y) - np.asarray(y0)).max() <= tol
            assert int(y[0]) == int(y0[0])
            assert abs(M.value(T, [float(v) for v in y]) - v0) <= tol
            assert np.abs(np.asarray(M.df_dx(T, [float(v) for v in y])) - np.asarray(r0)).max() <= max(tol, 1e-9)
            assert abs(M(T, [float(v) for v in y]) - v0) <= tol          # __call__ brackets the value with setx / clearx
            M.setx(T)
        M.clearx()


def test_multi_constraint_equals_the_reference_class():
    check_against_golden('restated', 1e-12)


def test_multi_constraint_over_list_returning_members():
    """all_negative mode with members that answer with several pairs (the protocol's third form, con:141-151): every
    pair comes back tagged with its member's index and padded to the common co-parameter length; the best-only mode
    keeps the deepest."""
    from semiinfiniteoptimization_b200.constraint import MultiSemiInfiniteConstraint, SemiInfiniteConstraintInterface, SetDomain

    class Member(SemiInfiniteConstraintInterface):
        def __init__(self, pairs, width):
            self.pairs, self.width = pairs, width

        def domain(self):
            return SetDomain([tuple([0.0] * self.width)])

        def minvalue(self, x, bound=None):
            out = [(f, y) for (f, y) in self.pairs if bound is None or f < bound]
            return out[0] if len(out) == 1 else out

        def value(self, x, y):
            return dict((tuple(p), f) for f, p in self.pairs)[tuple(float(v) for v in y)]

    a = Member([(-0.3, (1.0, 2.0)), (-0.1, (3.0, 4.0)), (0.2, (5.0, 6.0))], 2)
    b = Member([(-0.5, (7.0, 8.0, 9.0))], 3)
    M = MultiSemiInfiniteConstraint([a, b], all_negative=True)
    assert M.num_co_parameters == [2, 3] and M.max_co_parameters == 3
    got = M.minvalue(None, -1.0)                      # a negative bound is raised to 0: everything below zero comes back
    assert [f for f, _ in got] == [-0.3, -0.1, -0.5]
    assert [list(y) for _, y in got] == [[0, 1, 2, 0], [0, 3, 4, 0], [1, 7, 8, 9]]
    assert M.value(None, got[1][1]) == -0.1 and M.value(None, got[2][1]) == -0.5
    best = MultiSemiInfiniteConstraint([a, b], all_negative=False).minvalue(None, None)
    assert best[0] == -0.5 and list(best[1]) == [1, 7, 8, 9]
    assert MultiSemiInfiniteConstraint([a, b], all_negative=False).minvalue(None, -0.6) == []


def test_multi_golden_is_what_the_reference_produces_now():
    from oracle.refrun import load
    if not load.available():
        pytest.skip("reference sources are not present on this machine")
    import sys
    from tools.make_golden import multi_inputs
    golden = scenes.load_golden("multi_constraint.json")
    mod, cons, poses, bounds = multi_inputs('reference')
    cmod = sys.modules[mod.__name__.rsplit('.', 1)[0] + '.constraint']
    cmod.sampleDomain = lambda d: d.sample()
    M = cmod.MultiSemiInfiniteConstraint(cons, all_negative=True)
    rec = golden["cases"][1]
    T = (rec["T"][0], rec["T"][1])
    M.setx(T)
    mv = M.minvalue(T, rec["bound"])
    assert [float(f) for f, _ in mv] == [p[0] for p in rec["pairs"]]
    assert all(np.array_equal(np.asarray(y), np.asarray(p[1])) for (_, y), p in zip(mv, rec["pairs"]))


def test_minimum_constraint_adaptor_drives_the_generic_loop_to_a_collision_free_pose():
    """MinimumConstraintAdaptor (constraint.py:259-277, made to run) + optimizeStandard, the body of optimizeCollFreeMinDist
    (gopt:1083-1122), on the CPU restatement of the object constraint: value = the oracle's minimum, gradient = the row at
    the minimising point; from poses in collision the loop ends collision-free, and agrees with minvalue / df_dx of the
    constraint it wraps."""
    from semiinfiniteoptimization_b200 import sip
    from semiinfiniteoptimization_b200.constraint import MinimumConstraintAdaptor
    from semiinfiniteoptimization_b200.geometryopt import ObjectPoseObjective
    mod, g1, g2, con = scenes.c1_scene('restated')
    ad = MinimumConstraintAdaptor(con)
    solved = 0
    for T in scenes.c1_poses(6, seed=0):
        ad.setx(T)
        d0, y0 = con.minvalue(T)
        assert ad.value(T) == d0 and np.array_equal(ad.df_dx(T), con.df_dx(T, y0)) and list(ad.minparam) == list(y0)
        ad.clearx()
        assert ad(T) == d0                                  # __call__ brackets value with setx / clearx
        if d0 > -0.01:
            continue
        res = sip.optimizeStandard(ObjectPoseObjective(T), [ad], T, settings=sip.SemiInfiniteOptimizationSettings(), verbose=0)
        con.setx(res.x)
        d1 = con.minvalue(res.x)[0]
        con.clearx()
        assert d1 > -5e-3 and d1 > d0 + 0.005, (d0, d1)
        solved += 1
    assert solved >= 3
