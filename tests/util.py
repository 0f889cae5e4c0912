"""Shared helpers for the tests: seeded geometry, poses, and oracle wrappers."""
import numpy as np

from semiinfiniteoptimization_b200._host import se3, so3
from semiinfiniteoptimization_b200._host.geometry import analytic_sdf_grid, fixture_geometry


def sphere_grid(r=0.7, n=48, half=1.0):
    return analytic_sdf_grid(lambda P: np.linalg.norm(P, axis=-1) - r, [-half] * 3, [half] * 3, (n, n + 3, n - 5))


def box_sdf(P, half):
    q = np.abs(P) - np.asarray(half)
    return np.linalg.norm(np.maximum(q, 0), axis=-1) + np.minimum(q.max(-1), 0)


def random_grid(seed, dims=(17, 23, 19)):
    rng = np.random.default_rng(seed)
    from semiinfiniteoptimization_b200._host.geometry import VolumeGrid
    return VolumeGrid(rng.standard_normal(dims).astype(np.float32), [-0.3, -0.5, 0.1], [0.9, 0.7, 1.4])


def random_poses(seed, n, rot=0.3, trans=0.1, center=(0, 0, 0)):
    """Row-major 3x4 poses T = (exp(w), t): w ~ U(-rot,rot)^3, t ~ center + U(-trans,trans)^3."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, 12))
    for i in range(n):
        w = rng.uniform(-rot, rot, 3)
        t = np.asarray(center) + rng.uniform(-trans, trans, 3)
        out[i] = se3.to_rowmajor12((so3.from_rotation_vector(w), list(t)))
    return out


def invert12(T):
    T = np.asarray(T, dtype=np.float64).reshape(-1, 3, 4)
    R, t = T[:, :, :3], T[:, :, 3]
    out = np.empty_like(T)
    out[:, :, :3] = np.transpose(R, (0, 2, 1))
    out[:, :, 3] = -np.einsum('bji,bj->bi', R, t)
    return out.reshape(-1, 12)
