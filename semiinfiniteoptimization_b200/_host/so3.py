"""Rotation helpers in Klampt's convention: a rotation is a 9-list in COLUMN-major order
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
    """R1 R2 on the column-major 9-lists, scalar arithmetic (a numpy round trip costs 11 us for these 27 products; the SIP
    loop of a single problem spends a fifth of its time in this module)."""
    a0, a1, a2, a3, a4, a5, a6, a7, a8 = R1
    b0, b1, b2, b3, b4, b5, b6, b7, b8 = R2
    return [a0 * b0 + a3 * b1 + a6 * b2, a1 * b0 + a4 * b1 + a7 * b2, a2 * b0 + a5 * b1 + a8 * b2,
            a0 * b3 + a3 * b4 + a6 * b5, a1 * b3 + a4 * b4 + a7 * b5, a2 * b3 + a5 * b4 + a8 * b5,
            a0 * b6 + a3 * b7 + a6 * b8, a1 * b6 + a4 * b7 + a7 * b8, a2 * b6 + a5 * b7 + a8 * b8]


def apply(R, v):
    return [R[0] * v[0] + R[3] * v[1] + R[6] * v[2],
            R[1] * v[0] + R[4] * v[1] + R[7] * v[2],
            R[2] * v[0] + R[5] * v[1] + R[8] * v[2]]


def from_rotation_vector(w):
    """exp([w]x) by Rodrigues' formula: I + s K + c K^2 with K = [w]x, s = sin(th)/th, c = (1 - cos th)/th^2."""
    x, y, z = float(w[0]), float(w[1]), float(w[2])
    xx, yy, zz = x * x, y * y, z * z
    th = math.sqrt(xx + yy + zz)
    if th < 1e-12:
        s, c = 1.0, 0.5
    else:
        s, c = math.sin(th) / th, (1 - math.cos(th)) / (th * th)
    xy, xz, yz = x * y, x * z, y * z
    # K^2 = w w' - |w|^2 I;  column-major: [m00, m10, m20, m01, m11, m21, m02, m12, m22]
    return [1.0 + c * (-(yy + zz)), s * z + c * xy, -s * y + c * xz,
            -s * z + c * xy, 1.0 + c * (-(xx + zz)), s * x + c * yz,
            s * y + c * xz, -s * x + c * yz, 1.0 + c * (-(xx + yy))]


def from_axis_angle(aa):
    axis, angle = aa
    a = np.asarray(axis, dtype=np.float64)
    n = np.linalg.norm(a)
    return from_rotation_vector(a / n * angle) if n > 0 else identity()


def rotation_vector(R):
    """log map: the vector w with exp([w]x) = R."""
    c = ((R[0] + R[4] + R[8]) - 1.0) * 0.5
    c = min(1.0, max(-1.0, c))
    th = math.acos(c)
    vx, vy, vz = R[5] - R[7], R[6] - R[2], R[1] - R[3]        # M[2,1]-M[1,2], M[0,2]-M[2,0], M[1,0]-M[0,1]
    if th < 1e-8:
        return [0.5 * vx, 0.5 * vy, 0.5 * vz]
    if abs(th - math.pi) >= 1e-5:
        k = th / (2.0 * math.sin(th))
        return [vx * k, vy * k, vz * k]
    M = matrix(R)
    v = np.array([vx, vy, vz])
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
