"""FROZEN COPY (oracle/frozen, round 2) of semiinfiniteoptimization_b200/_host/so3.py as of commit 92f6a08 -- TEST INFRASTRUCTURE.

The golden generator (tools/make_golden.py -> oracle/refrun) answers the reference's Klampt / OSQP calls with THIS file,
not with the product's module, so a change to the product can no longer rewrite tests/golden/.  Do not edit it to follow
the product: the product is tested AGAINST it (tests/test_frozen_cpu.py) at a stated tolerance.

Original module docstring follows.

Rotation helpers in Klampt's convention: a rotation is a 9-list in COLUMN-major order
(`klampt.math.so3`; used at geometryopt.py:91, 391, 415).  Only the entry points the hot path's
callers need are provided."""
import math

import numpy as np


def identity():
    return [1., 0., 0., 0., 1., 0., 0., 0., 1.]


def matrix(R):
    """3x3 numpy matrix of the column-major 9-list."""
    return np.asarray(R, dtype=np.float64).reshape(3, 3).T.copy()


def from_matrix(M):
    return [float(v) for v in np.asarray(M, dtype=np.float64).T.reshape(9)]


def inv(R):
    return [R[0], R[3], R[6], R[1], R[4], R[7], R[2], R[5], R[8]]


def mul(R1, R2):
    return from_matrix(matrix(R1) @ matrix(R2))


def apply(R, v):
    return [R[0] * v[0] + R[3] * v[1] + R[6] * v[2],
            R[1] * v[0] + R[4] * v[1] + R[7] * v[2],
            R[2] * v[0] + R[5] * v[1] + R[8] * v[2]]


def from_rotation_vector(w):
    """exp([w]x) by Rodrigues' formula."""
    w = np.asarray(w, dtype=np.float64)
    th = float(np.linalg.norm(w))
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=np.float64)
    if th < 1e-12:
        M = np.eye(3) + K + 0.5 * (K @ K)
    else:
        M = np.eye(3) + (math.sin(th) / th) * K + ((1 - math.cos(th)) / (th * th)) * (K @ K)
    return from_matrix(M)


def from_axis_angle(aa):
    axis, angle = aa
    a = np.asarray(axis, dtype=np.float64)
    n = np.linalg.norm(a)
    return from_rotation_vector(a / n * angle) if n > 0 else identity()


def rotation_vector(R):
    """log map: the vector w with exp([w]x) = R."""
    M = matrix(R)
    c = (np.trace(M) - 1.0) * 0.5
    c = min(1.0, max(-1.0, c))
    th = math.acos(c)
    v = np.array([M[2, 1] - M[1, 2], M[0, 2] - M[2, 0], M[1, 0] - M[0, 1]])
    if th < 1e-8:
        return [float(x) for x in 0.5 * v]
    if abs(th - math.pi) < 1e-5:
        # near pi the antisymmetric part vanishes; recover the axis from R + I
        A = (M + np.eye(3)) * 0.5
        k = int(np.argmax(np.diag(A)))
        axis = A[:, k] / math.sqrt(max(A[k, k], 1e-300))
        if np.dot(axis, v) < 0:
            axis = -axis
        return [float(x) for x in axis * th]
    return [float(x) for x in v * (th / (2.0 * math.sin(th)))]


def error(R1, R2):
    """Rotation vector of R1*R2^T (the Lie-derivative 'difference' used by se3.error)."""
    return rotation_vector(mul(R1, inv(R2)))


def quaternion(R):
    """(w,x,y,z)."""
    M = matrix(R)
    tr = np.trace(M)
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        return [0.25 * s, (M[2, 1] - M[1, 2]) / s, (M[0, 2] - M[2, 0]) / s, (M[1, 0] - M[0, 1]) / s]
    i = int(np.argmax(np.diag(M)))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = math.sqrt(M[i, i] - M[j, j] - M[k, k] + 1.0) * 2
    q = [0.0] * 4
    q[0] = (M[k, j] - M[j, k]) / s
    q[1 + i] = 0.25 * s
    q[1 + j] = (M[j, i] + M[i, j]) / s
    q[1 + k] = (M[k, i] + M[i, k]) / s
    return q


def rotation(axis, angle):
    return from_axis_angle((axis, angle))
