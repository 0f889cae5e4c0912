"""CPU ORACLE for the collision-oracle hot path -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the arithmetic restated here belongs to Klampt 0.8.x / KrisLibrary (un-vendored,
unpinned; README.md:54 of the reference), which is absent from /root/reference and from this
image, and the reference ships no golden vector or test for it.  The restatement
follows SURVEY.md Appendix A.2-A.4 and is pinned only by closed-form known answers (tests/).

Two independent restatements live here so they can check each other:
  * `liboracle.so` (oracle/sio_oracle.c, float64, OpenMP) through ctypes -- the checker proper and
    the CPU timing baseline;
  * `np_*` functions -- a vectorised numpy version written separately from the C one.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_c = ctypes
_lib = None


class _Grid(_c.Structure):
    _fields_ = [("v", _c.c_void_p), ("nx", _c.c_int32), ("ny", _c.c_int32), ("nz", _c.c_int32),
                ("bmin", _c.c_double * 3), ("bmax", _c.c_double * 3)]


class _Robot(_c.Structure):
    _fields_ = [("n", _c.c_int32), ("parent", _c.c_void_p), ("jtype", _c.c_void_p),
                ("tparent", _c.c_void_p), ("axis", _c.c_void_p)]


def build(native=False):
    """Compile oracle/sio_oracle.c (the recipe is oracle/Makefile)."""
    subprocess.check_call(["make", "-s", "-C", _HERE] + (["native"] if native else []))


def lib(prefer_native=False):
    global _lib
    if _lib is not None and not prefer_native:
        return _lib
    path = os.path.join(_HERE, "_build", "liboracle.so")
    if prefer_native:
        npath = os.path.join(_HERE, "_build", "liboracle_native.so")
        if os.path.exists(npath):
            path = npath
    if not os.path.exists(path):
        build()
    L = _c.CDLL(path)
    P = _c.c_void_p
    I64 = _c.c_int64
    L.orc_max_threads.restype = _c.c_int
    L.orc_set_threads.argtypes = [_c.c_int]
    L.orc_set_threads.restype = None
    L.orc_query.argtypes = [P, P, P, I64, _c.c_int, P, P]
    L.orc_cloud_query.argtypes = [P, P, I64, P, I64, _c.c_int, P, P]
    L.orc_min_distance.argtypes = [P, P, P, I64, P, P, I64, P, P]
    L.orc_min_distance_1.argtypes = [P, P, I64, P, P, P]
    L.orc_under.argtypes = [P, P, P, I64, P, P, I64, I64, P, P, P]
    L.orc_under.restype = I64
    L.orc_under_counts.argtypes = [P, P, P, I64, P, P, I64, P, P, P]
    L.orc_under_counts.restype = I64
    L.orc_pose_rows.argtypes = [P, P, P, I64, _c.c_int, P, P]
    L.orc_fk.argtypes = [P, P, I64, P]
    L.orc_chain_rows.argtypes = [P, _c.c_int32, P, P, P, I64, _c.c_int, P, P]
    L.orc_min_distance_f32.argtypes = [P, P, I64, P, I64, P, P]
    L.orc_tree_build.argtypes = [P, I64]
    L.orc_tree_build.restype = P
    L.orc_tree_free.argtypes = [P]
    L.orc_tree_free.restype = None
    L.orc_grid_lipschitz.argtypes = [P]
    L.orc_grid_lipschitz.restype = _c.c_double
    L.orc_min_distance_pruned.argtypes = [P, _c.c_double, P, P, P, I64, _c.c_int, P, P, P]
    L.orc_min_distance_pruned.restype = None
    if not prefer_native:
        _lib = L
    return L


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Grid:
    """float32 samples (nx,ny,nz) at cell centres + bbox; keeps the ctypes struct alive."""

    def __init__(self, values, bmin, bmax):
        self.values = np.ascontiguousarray(values, dtype=np.float32)
        assert self.values.ndim == 3
        self.bmin = _f64(bmin).copy()
        self.bmax = _f64(bmax).copy()
        self.dims = self.values.shape

    def struct(self):
        s = _Grid()
        s.v = self.values.ctypes.data
        s.nx, s.ny, s.nz = (int(d) for d in self.dims)
        s.bmin[:] = list(self.bmin)
        s.bmax[:] = list(self.bmax)
        return s


def _grid_array(grids):
    arr = (_Grid * len(grids))()
    for i, g in enumerate(grids):
        arr[i] = g.struct()
    return arr


class Robot:
    def __init__(self, parent, tparent, axis, jtype=None):
        self.n = len(parent)
        self.parent = np.ascontiguousarray(parent, dtype=np.int32)
        self.tparent = _f64(tparent).reshape(self.n, 12)
        self.axis = _f64(axis).reshape(self.n, 3)
        self.jtype = np.zeros(self.n, dtype=np.int32) if jtype is None else np.ascontiguousarray(jtype, dtype=np.int32)

    def struct(self):
        s = _Robot()
        s.n = self.n
        s.parent = self.parent.ctypes.data
        s.jtype = self.jtype.ctypes.data
        s.tparent = self.tparent.ctypes.data
        s.axis = self.axis.ctypes.data
        return s


# ---------------------------------------------------------------------------------------------
# C-backed oracle entry points
# ---------------------------------------------------------------------------------------------
def query(grid, T, pts, normalize=True, want_grad=True, L=None):
    """grid.distance_point (geometryopt.py:104-108,153) for P points given in the frame that
    `T` (row-major 3x4: grid-local <- that frame) maps from.  Returns d[P], grad[P,3] (w.r.t. the
    input point)."""
    L = L or lib()
    pts = _f64(pts).reshape(-1, 3)
    T = _f64(T).reshape(12)
    d = np.empty(len(pts))
    g = np.empty((len(pts), 3)) if want_grad else None
    s = grid.struct()
    L.orc_query(_c.byref(s), T.ctypes.data, pts.ctypes.data, len(pts), int(normalize), d.ctypes.data,
                g.ctypes.data if want_grad else None)
    return (d, g) if want_grad else d


def cloud_query(grid, cloud, T, normalize=True, want_grad=True):
    cloud = np.ascontiguousarray(cloud, dtype=np.float32).reshape(-1, 3)
    T = _f64(T).reshape(-1, 12)
    B, N = len(T), len(cloud)
    d = np.empty((B, N))
    g = np.empty((B, N, 3)) if want_grad else None
    s = grid.struct()
    lib().orc_cloud_query(_c.byref(s), cloud.ctypes.data, N, T.ctypes.data, B, int(normalize), d.ctypes.data,
                          g.ctypes.data if want_grad else None)
    return (d, g) if want_grad else d


def min_distance(grids, cloud, T, bound=None, grid_of_b=None, L=None):
    """pc.distance_ext(grid, upperBound) (geometryopt.py:117-127), brute force, for B transforms."""
    L = L or lib()
    if isinstance(grids, Grid):
        grids = [grids]
    cloud = np.ascontiguousarray(cloud, dtype=np.float32).reshape(-1, 3)
    T = _f64(T).reshape(-1, 12)
    B = len(T)
    dmin = np.empty(B)
    arg = np.empty(B, dtype=np.int64)
    ga = _grid_array(grids)
    gob = None if grid_of_b is None else np.ascontiguousarray(grid_of_b, dtype=np.int32)
    bnd = None if bound is None else _f64(np.broadcast_to(bound, (B,)))
    L.orc_min_distance(ga, None if gob is None else gob.ctypes.data, cloud.ctypes.data, len(cloud),
                       T.ctypes.data, None if bnd is None else bnd.ctypes.data, B,
                       dmin.ctypes.data, arg.ctypes.data)
    return dmin, arg


def under(grids, cloud, T, bound, grid_of_b=None, cap=1 << 22):
    if isinstance(grids, Grid):
        grids = [grids]
    cloud = np.ascontiguousarray(cloud, dtype=np.float32).reshape(-1, 3)
    T = _f64(T).reshape(-1, 12)
    B = len(T)
    bnd = _f64(np.broadcast_to(bound, (B,)))
    ob = np.empty(cap, dtype=np.int64)
    oi = np.empty(cap, dtype=np.int64)
    od = np.empty(cap)
    ga = _grid_array(grids)
    gob = None if grid_of_b is None else np.ascontiguousarray(grid_of_b, dtype=np.int32)
    n = lib().orc_under(ga, None if gob is None else gob.ctypes.data, cloud.ctypes.data, len(cloud), T.ctypes.data,
                        bnd.ctypes.data, B, cap, ob.ctypes.data, oi.ctypes.data, od.ctypes.data)
    m = min(n, cap)
    return n, ob[:m], oi[:m], od[:m]


def under_counts(grids, cloud, T, bound, grid_of_b=None):
    """Per transform: #{i : d < bound}, the sum of those indices and the sum of those distances (fingerprints of the
    under-bound set at sizes where the explicit list is too long to keep)."""
    if isinstance(grids, Grid):
        grids = [grids]
    cloud = np.ascontiguousarray(cloud, dtype=np.float32).reshape(-1, 3)
    T = _f64(T).reshape(-1, 12)
    B = len(T)
    bnd = _f64(np.broadcast_to(bound, (B,)))
    cnt = np.empty(B, dtype=np.int64)
    isum = np.empty(B, dtype=np.int64)
    dsum = np.empty(B)
    ga = _grid_array(grids)
    gob = None if grid_of_b is None else np.ascontiguousarray(grid_of_b, dtype=np.int32)
    lib().orc_under_counts(ga, None if gob is None else gob.ctypes.data, cloud.ctypes.data, len(cloud), T.ctypes.data,
                           bnd.ctypes.data, B, cnt.ctypes.data, isum.ctypes.data, dsum.ctypes.data)
    return cnt, isum, dsum


def pose_rows(grid, Tpose, pts, normalize=True):
    pts = _f64(pts).reshape(-1, 3)
    Tpose = _f64(Tpose).reshape(12)
    d = np.empty(len(pts))
    rows = np.empty((len(pts), 6))
    s = grid.struct()
    lib().orc_pose_rows(_c.byref(s), Tpose.ctypes.data, pts.ctypes.data, len(pts), int(normalize),
                        d.ctypes.data, rows.ctypes.data)
    return d, rows


def fk(robot, q):
    q = _f64(q).reshape(-1, robot.n)
    out = np.empty((len(q), robot.n, 12))
    s = robot.struct()
    lib().orc_fk(_c.byref(s), q.ctypes.data, len(q), out.ctypes.data)
    return out


def chain_rows(robot, link, grid, q, pts, normalize=True):
    pts = _f64(pts).reshape(-1, 3)
    q = _f64(q).reshape(robot.n)
    d = np.empty(len(pts))
    rows = np.empty((len(pts), robot.n))
    rs, gs = robot.struct(), grid.struct()
    lib().orc_chain_rows(_c.byref(rs), int(link), _c.byref(gs), q.ctypes.data, pts.ctypes.data, len(pts),
                         int(normalize), d.ctypes.data, rows.ctypes.data)
    return d, rows


def min_distance_f32(grid, cloud, T, L=None):
    """Single-precision brute force: CPU TIMING BASELINE only."""
    L = L or lib()
    cloud = np.ascontiguousarray(cloud, dtype=np.float32).reshape(-1, 3)
    T = _f64(T).reshape(-1, 12)
    B = len(T)
    dmin = np.empty(B, dtype=np.float32)
    arg = np.empty(B, dtype=np.int64)
    s = grid.struct()
    L.orc_min_distance_f32(_c.byref(s), cloud.ctypes.data, len(cloud), T.ctypes.data, B, dmin.ctypes.data, arg.ctypes.data)
    return dmin, arg


class CloudTree:
    """k-d leaves of 128 points with bounding spheres: the hierarchy of the pruned CPU arm (orc_min_distance_pruned)."""

    def __init__(self, cloud, L=None):
        self.L = L or lib()
        self.cloud = np.ascontiguousarray(cloud, dtype=np.float32).reshape(-1, 3)
        self.h = self.L.orc_tree_build(self.cloud.ctypes.data, len(self.cloud))
        if not self.h:
            raise MemoryError("orc_tree_build failed")

    def __del__(self):
        try:
            if self.h:
                self.L.orc_tree_free(self.h)
                self.h = None
        except Exception:
            pass


def grid_lipschitz(grid, L=None):
    s = grid.struct()
    return float((L or lib()).orc_grid_lipschitz(_c.byref(s)))


def min_distance_pruned(grid, tree, T, bound=None, use_bound=False, lip=None, L=None):
    """The hierarchy-pruned CPU arm ("pruning on"): same (dmin, argmin) as min_distance, plus the point queries spent per
    transform.  use_bound=True lets the caller's bound prune too (Klampt's upperBound): dmin is then exact only below it."""
    L = L or tree.L
    T = _f64(T).reshape(-1, 12)
    B = len(T)
    dmin = np.empty(B)
    arg = np.empty(B, dtype=np.int64)
    ev = np.empty(B, dtype=np.int64)
    bnd = None if bound is None else _f64(np.broadcast_to(bound, (B,)))
    s = grid.struct()
    if lip is None:
        lip = grid_lipschitz(grid, L)
    L.orc_min_distance_pruned(_c.byref(s), float(lip), tree.h, T.ctypes.data, None if bnd is None else bnd.ctypes.data, B,
                              int(bool(use_bound)), dmin.ctypes.data, arg.ctypes.data, ev.ctypes.data)
    return dmin, arg, ev


# ---------------------------------------------------------------------------------------------
# independent numpy restatement (Appendix A.2), used to cross-check the C oracle
# ---------------------------------------------------------------------------------------------
def np_query_local(values, bmin, bmax, p):
    """p: (...,3) grid-local float64.  Returns d, grad_local (unnormalised)."""
    v = np.asarray(values, dtype=np.float64)
    dims = np.array(v.shape, dtype=np.float64)
    bmin = _f64(bmin)
    bmax = _f64(bmax)
    h = (bmax - bmin) / dims
    c = np.clip(p, bmin, bmax)
    off = p - c
    dout = np.sqrt((off * off).sum(-1))
    u = (c - bmin) / h - 0.5
    interior = (u > 0) & (u < dims - 1)
    uc = np.clip(u, 0, dims - 1)
    i0 = np.minimum(np.floor(uc), dims - 1).astype(np.int64)
    i1 = np.minimum(i0 + 1, (dims - 1).astype(np.int64))
    f = uc - i0
    fx, fy, fz = f[..., 0], f[..., 1], f[..., 2]

    def V(a, b, c_):
        return v[a[..., 0], b[..., 1], c_[..., 2]]
    v000, v001 = V(i0, i0, i0), V(i0, i0, i1)
    v010, v011 = V(i0, i1, i0), V(i0, i1, i1)
    v100, v101 = V(i1, i0, i0), V(i1, i0, i1)
    v110, v111 = V(i1, i1, i0), V(i1, i1, i1)
    c00 = v000 * (1 - fz) + v001 * fz
    c01 = v010 * (1 - fz) + v011 * fz
    c10 = v100 * (1 - fz) + v101 * fz
    c11 = v110 * (1 - fz) + v111 * fz
    c0 = c00 * (1 - fy) + c01 * fy
    c1 = c10 * (1 - fy) + c11 * fy
    val = c0 * (1 - fx) + c1 * fx
    gx = c1 - c0
    gy = (c01 - c00) * (1 - fx) + (c11 - c10) * fx
    gz = ((v001 - v000) * (1 - fy) + (v011 - v010) * fy) * (1 - fx) + ((v101 - v100) * (1 - fy) + (v111 - v110) * fy) * fx
    g = np.stack([gx, gy, gz], axis=-1) / h
    g = np.where(interior, g, 0.0)
    with np.errstate(invalid="ignore", divide="ignore"):
        outward = np.where(dout[..., None] > 0, off / dout[..., None], 0.0)
    return val + dout, g + outward


def np_query(values, bmin, bmax, T, pts, normalize=True):
    T = _f64(T).reshape(3, 4)
    pts = _f64(pts).reshape(-1, 3)
    p = pts @ T[:, :3].T + T[:, 3]
    d, gl = np_query_local(values, bmin, bmax, p)
    g = gl @ T[:, :3]
    if normalize:
        n = np.linalg.norm(g, axis=-1, keepdims=True)
        g = np.where(n > 0, g / np.where(n > 0, n, 1), 0.0)
    return d, g


def np_min_distance(values, bmin, bmax, cloud, T):
    T = _f64(T).reshape(-1, 3, 4)
    cloud = np.asarray(cloud, dtype=np.float32).astype(np.float64).reshape(-1, 3)
    dmin = np.empty(len(T))
    arg = np.empty(len(T), dtype=np.int64)
    for b in range(len(T)):
        p = cloud @ T[b, :, :3].T + T[b, :, 3]
        d, _ = np_query_local(values, bmin, bmax, p)
        arg[b] = int(np.argmin(d))
        dmin[b] = d[arg[b]]
    return dmin, arg
