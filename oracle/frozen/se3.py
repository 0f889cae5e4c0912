"""FROZEN COPY (oracle/frozen, round 2) of semiinfiniteoptimization_b200/_host/se3.py as of commit 92f6a08 -- TEST INFRASTRUCTURE.

The golden generator (tools/make_golden.py -> oracle/refrun) answers the reference's Klampt / OSQP calls with THIS file,
not with the product's module, so a change to the product can no longer rewrite tests/golden/.  Do not edit it to follow
the product: the product is tested AGAINST it (tests/test_frozen_cpu.py) at a stated tolerance.

Original module docstring follows.

Rigid transforms in Klampt's convention: T = (R, t), R a column-major 9-list
(`klampt.math.se3`; used at geometryopt.py:177, 383-386).  `to_rowmajor12` is the conversion to
the C-ABI's row-major 3x4 `double[12]` (SURVEY.md 8b "Data conventions")."""
import numpy as np

from . import so3, vectorops


def identity():
    return (so3.identity(), [0., 0., 0.])


def apply(T, p):
    return vectorops.add(so3.apply(T[0], p), T[1])


def inv(T):
    Rinv = so3.inv(T[0])
    return (Rinv, [-v for v in so3.apply(Rinv, T[1])])


def mul(T1, T2):
    return (so3.mul(T1[0], T2[0]), apply(T1, T2[1]))


def error(T1, T2):
    """[so3.error(R1,R2), t1 - t2] as a 6-list."""
    return so3.error(T1[0], T2[0]) + vectorops.sub(T1[1], T2[1])


def homogeneous(T):
    H = np.eye(4)
    H[:3, :3] = so3.matrix(T[0])
    H[:3, 3] = T[1]
    return H


def from_homogeneous(H):
    H = np.asarray(H, dtype=np.float64)
    return (so3.from_matrix(H[:3, :3]), [float(v) for v in H[:3, 3]])


def to_rowmajor12(T):
    """(R,t) -> float64[12] = row-major [R | t]."""
    out = np.empty((3, 4), dtype=np.float64)
    out[:, :3] = so3.matrix(T[0])
    out[:, 3] = T[1]
    return out.reshape(12)


def from_rowmajor12(a):
    a = np.asarray(a, dtype=np.float64).reshape(3, 4)
    return (so3.from_matrix(a[:, :3]), [float(v) for v in a[:, 3]])


def rel12(Ta12, Tb12):
    """Row-major 3x4 of Ta^-1 * Tb (both given as row-major 3x4, possibly batched on Ta)."""
    A = np.asarray(Ta12, dtype=np.float64).reshape(-1, 3, 4)
    B = np.asarray(Tb12, dtype=np.float64).reshape(-1, 3, 4)
    Ra, ta = A[:, :, :3], A[:, :, 3]
    Rb, tb = B[:, :, :3], B[:, :, 3]
    RaT = np.transpose(Ra, (0, 2, 1))
    out = np.empty((max(len(A), len(B)), 3, 4), dtype=np.float64)
    out[:, :, :3] = RaT @ Rb
    out[:, :, 3] = np.einsum('bij,bj->bi', RaT, tb - ta)
    return out.reshape(-1, 12)
