"""GPU: the lock-step batch driver (C5) against the scalar drop-in path, problem by problem."""
import numpy as np
import pytest

from tests import util
from tests.test_gpu_classes import c1_poses

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene(ctx):
    from semiinfiniteoptimization_b200 import geometryopt as G
    from semiinfiniteoptimization_b200._host import so3
    cube = G.PenetrationDepthGeometry(util.fixture_geometry("cube"), 0.05, 0.02)
    chair = G.PenetrationDepthGeometry(util.fixture_geometry("m797"), 0.05, 0.02)
    chair.setTransform((so3.identity(), [1.2, 0, 0]))
    return G, cube, chair


def test_batch_equals_scalar_problem_by_problem(scene):
    from semiinfiniteoptimization_b200 import batch
    G, cube, chair = scene
    poses = c1_poses(10, seed=11)
    res = batch.optimizeCollFreeBatch(cube, chair, poses, qp='scalar', keep_trace=True)
    assert res.T.shape == (10, 12) and res.point_queries > 0
    moved = 0
    for p, T in enumerate(poses):
        G.optimizeCollFree(cube, chair, T, verbose=0)
        ref = G.optimizeCollFree.last_result
        assert res.status[p] == ref.status and res.num_iterations[p] == ref.num_iterations
        assert len(res.trace[p]) == len(ref.trace)
        for a, b in zip(res.trace[p], ref.trace):
            assert np.allclose(a[0], b[0], atol=1e-8) and np.allclose(a[1], b[1], atol=1e-8)
        assert len(res.instantiated_params[p]) == len(ref.instantiated_params[0])
        assert np.isclose(res.gx[p], ref.gx[0], atol=1e-8) or (np.isinf(res.gx[p]) and np.isinf(ref.gx[0]))
        moved += ref.num_iterations > 2
    assert moved >= 3


def test_batch_qp_reaches_feasibility(scene):
    """The lock-step ADMM variant is not bit-identical to the scalar QP, but must resolve the same collisions."""
    from semiinfiniteoptimization_b200 import batch
    G, cube, chair = scene
    poses = c1_poses(16, seed=12)
    res = batch.optimizeCollFreeBatch(cube, chair, poses, qp='batch')
    ref = batch.optimizeCollFreeBatch(cube, chair, poses, qp='scalar')
    c = G.ObjectCollisionConstraint(cube, chair)
    from semiinfiniteoptimization_b200._host import se3
    feasible = 0
    for p in range(len(poses)):
        got = c.eval_minimum(se3.from_rowmajor12(res.T[p]))
        want = c.eval_minimum(se3.from_rowmajor12(ref.T[p]))
        # -5e-3 is the reference's own feasibility threshold (sip.py:1227); a start at the SDF's medial axis
        # (zero gradient) defeats the reference's method too, so only demand what the scalar QP achieves
        assert got > -5e-3 or want <= -5e-3, (p, got, want)
        feasible += got > -5e-3
    assert feasible >= 12
    rec = res.records()
    assert rec.shape == (16, 16) and np.allclose(rec[:, :12], res.T)


def test_array_driver_equals_scalar_driver(scene):
    """optimizeCollFreeBatchFast (padded arrays + native batched QP + pruned sweeps) walks every problem through the
    same iterations as the per-problem driver; poses agree to the QP's rounding."""
    from semiinfiniteoptimization_b200 import batch
    G, cube, chair = scene
    poses = c1_poses(48, seed=21)
    ref = batch.optimizeCollFreeBatch(cube, chair, poses, qp='scalar', keep_trace=True)
    fast = batch.optimizeCollFreeBatchFast(cube, chair, poses, keep_trace=True)
    assert np.array_equal(ref.num_iterations, fast.num_iterations) and ref.status == fast.status
    assert ref.outer_iterations == fast.outer_iterations > 200
    assert np.abs(ref.T - fast.T).max() <= 1e-6
    for p in range(len(poses)):
        assert len(ref.trace[p]) == len(fast.trace[p])
        assert [len(y) for y in ref.instantiated_params[p]] == [len(y) for y in fast.instantiated_params[p]]
        assert np.allclose(np.array(ref.instantiated_params[p]).reshape(-1), np.array(fast.instantiated_params[p]).reshape(-1), atol=1e-12)
    assert np.allclose(ref.gx, fast.gx, atol=1e-6) and np.allclose(ref.fx, fast.fx, atol=1e-6)
    assert fast.time_qp > 0 and fast.point_queries == ref.point_queries
