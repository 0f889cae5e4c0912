"""Host-side geometry containers and one-off conversions.

In the reference these objects are Klampt's (`Geometry3D`, `PointCloud`, `VolumeGrid`;
geometryopt.py:38-94) and Klampt builds the SDF grid / surface cloud once on the host.  Klampt is
not available to this build, so this module supplies the small subset of that interface the
hot path's callers touch, plus the conversions:

  mesh -> VolumeGrid : cell-centred signed distance samples, z fastest (C helper hostgeom.c)
  mesh -> PointCloud : vertices + a barycentric lattice per triangle with spacing <= res

Nothing here is evaluated per SIP iteration: the results are uploaded once through
`sio_grid_create` / `sio_cloud_create` and stay resident on the device.
"""
import ctypes
import os

import numpy as np

from . import se3, so3

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
_DATA = os.path.join(_PKG, "data")

# Named conversion constants (SURVEY.md Appendix A.1 -- Klampt-side choices kept switchable)
SDF_PAD_CELLS = 5            # cells of padding added around the mesh AABB on every side
AUTO_GRID_CELLS = 32         # gridres == 0 ("auto"): longest AABB side / AUTO_GRID_CELLS
AUTO_PC_DIVS = 48            # pcres == 0 ("auto"): longest AABB side / AUTO_PC_DIVS

_hostgeom = None


def _lib():
    global _hostgeom
    if _hostgeom is None:
        path = os.path.join(_PKG, "lib", "libsio_hostgeom.so")
        if not os.path.exists(path):
            raise RuntimeError("libsio_hostgeom.so is not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = ctypes.CDLL(path)
        lib.hg_mesh_sdf.restype = ctypes.c_int
        lib.hg_mesh_sdf.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.hg_sdf_sign_pass.restype = ctypes.c_int
        lib.hg_sdf_sign_pass.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.hg_points_mesh_dist.restype = ctypes.c_int
        lib.hg_points_mesh_dist.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                            ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        _hostgeom = lib
    return _hostgeom


class TriangleMesh:
    def __init__(self, verts, tris):
        self.verts = np.ascontiguousarray(verts, dtype=np.float64).reshape(-1, 3)
        self.tris = np.ascontiguousarray(tris, dtype=np.int32).reshape(-1, 3)

    def bbox(self):
        return self.verts.min(0), self.verts.max(0)

    def area(self):
        a, b, c = (self.verts[self.tris[:, i]] for i in range(3))
        return 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1).sum()


class PointCloud:
    """Surface samples; `vertices` is the flat xyz list Klampt exposes (read at geometryopt.py:273)."""

    def __init__(self, points=None):
        self.points = np.zeros((0, 3)) if points is None else np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)

    @property
    def vertices(self):
        return self.points.reshape(-1)

    def numPoints(self):
        return len(self.points)

    def getPoint(self, i):
        return [float(v) for v in self.points[i]]

    def addPoint(self, p):
        self.points = np.vstack([self.points, np.asarray(p, dtype=np.float64)[None, :3]])
        return len(self.points) - 1

    def bbox(self):
        return self.points.min(0), self.points.max(0)


class VolumeGrid:
    """Axis-aligned box [bmin,bmax] with dims cells; values at CELL CENTRES, z fastest (App. A.1)."""

    def __init__(self, values, bmin, bmax):
        self.values = np.ascontiguousarray(values, dtype=np.float32)
        assert self.values.ndim == 3
        self.dims = tuple(int(d) for d in self.values.shape)
        self.bmin = np.asarray(bmin, dtype=np.float64).copy()
        self.bmax = np.asarray(bmax, dtype=np.float64).copy()

    @property
    def cell(self):
        return (self.bmax - self.bmin) / np.asarray(self.dims, dtype=np.float64)

    def get(self, i, j, k):
        return float(self.values[i, j, k])

    def bbox(self):
        return self.bmin, self.bmax


def sample_mesh_surface(mesh, res):
    """Vertices plus a lattice on every triangle so that no surface point is farther than ~res
    from a sample (the dispersion contract of Klampt's mesh->PointCloud conversion).

    Per triangle the lattice is spanned by the two SHORTER edges (apex = vertex opposite the
    longest edge) with n1 x n2 subdivisions, n = ceil(edge / res), so thin triangles are not
    over-sampled."""
    V, T = mesh.verts, mesh.tris
    pts = [V]
    if np.isfinite(res) and res > 0 and len(T):
        P3 = V[T]                                        # (nt,3,3)
        e = np.stack([np.linalg.norm(P3[:, 2] - P3[:, 1], axis=1),      # edge opposite vertex 0
                      np.linalg.norm(P3[:, 0] - P3[:, 2], axis=1),
                      np.linalg.norm(P3[:, 1] - P3[:, 0], axis=1)], axis=1)
        apex = np.argmax(e, axis=1)
        idx = np.arange(len(T))
        a = P3[idx, apex]
        b = P3[idx, (apex + 1) % 3]
        c = P3[idx, (apex + 2) % 3]
        n1 = np.maximum(1, np.ceil(np.linalg.norm(b - a, axis=1) / res - 1e-9).astype(np.int64))
        n2 = np.maximum(1, np.ceil(np.linalg.norm(c - a, axis=1) / res - 1e-9).astype(np.int64))
        pairs = np.unique(np.stack([n1, n2], axis=1), axis=0)
        for (m1, m2) in pairs:
            if m1 == 1 and m2 == 1:
                continue
            sel = np.nonzero((n1 == m1) & (n2 == m2))[0]
            ii, jj = np.meshgrid(np.arange(m1 + 1), np.arange(m2 + 1), indexing="ij")
            u = ii / m1
            v = jj / m2
            keep = (u + v) <= 1.0 + 1e-12
            corner = ((ii == 0) & (jj == 0)) | ((ii == m1) & (jj == 0)) | ((ii == 0) & (jj == m2))
            keep &= ~corner
            u = u[keep][None, :, None]
            v = v[keep][None, :, None]
            P = a[sel][:, None, :] * (1 - u - v) + b[sel][:, None, :] * u + c[sel][:, None, :] * v
            pts.append(P.reshape(-1, 3))
    P = np.vstack(pts)
    # merge the duplicates created along shared edges (quantise to 1e-9 m)
    key = np.round(P * 1e9).astype(np.int64)
    _, first = np.unique(key, axis=0, return_index=True)
    return P[np.sort(first)]


# optional accelerator for the distance part of mesh_to_sdf: fn(verts, tris, bmin, h, dims) -> float64 (nx,ny,nz)
# unsigned distances (geometryopt installs the device builder, sio_mesh_distance_grid; values identical to the host's)
_distance_grid_fn = None


def set_distance_grid_builder(fn):
    global _distance_grid_fn
    _distance_grid_fn = fn


# the whole conversion on the device (row f3): fn(verts, tris, bmin, h, dims) -> float32 (nx,ny,nz) SIGNED samples
# (sio_mesh_sdf_grid) and fn(verts, tris, res) -> float64 (n,3) surface points (sio_mesh_surface_sample).  geometryopt
# installs both; the host implementations below remain what the CPU-only tests and the golden generator run.
_sdf_grid_fn = None
_surface_sample_fn = None


def set_device_converters(sdf_fn, sample_fn):
    global _sdf_grid_fn, _surface_sample_fn
    _sdf_grid_fn, _surface_sample_fn = sdf_fn, sample_fn


def mesh_to_sdf(mesh, res, pad_cells=SDF_PAD_CELLS):
    lo, hi = mesh.bbox()
    dims = np.maximum(1, np.ceil((hi - lo) / res - 1e-9).astype(np.int64)) + 2 * pad_cells
    bmin = lo - pad_cells * res
    h = np.array([res, res, res], dtype=np.float64)
    bmax = bmin + dims * h
    out = np.empty(tuple(int(d) for d in dims), dtype=np.float32)
    dims32 = np.ascontiguousarray(dims, dtype=np.int32)
    bmin_c = np.ascontiguousarray(bmin)
    if _sdf_grid_fn is not None:
        return VolumeGrid(np.ascontiguousarray(_sdf_grid_fn(mesh.verts, mesh.tris, bmin_c, h, dims32), dtype=np.float32), bmin, bmax)
    if _distance_grid_fn is not None:
        dist = np.ascontiguousarray(_distance_grid_fn(mesh.verts, mesh.tris, bmin_c, h, dims32), dtype=np.float64)
        rc = _lib().hg_sdf_sign_pass(mesh.verts.ctypes.data, mesh.tris.ctypes.data, len(mesh.tris), bmin_c.ctypes.data,
                                     h.ctypes.data, dims32.ctypes.data, dist.ctypes.data, out.ctypes.data)
    else:
        rc = _lib().hg_mesh_sdf(mesh.verts.ctypes.data, len(mesh.verts), mesh.tris.ctypes.data, len(mesh.tris),
                                bmin_c.ctypes.data, h.ctypes.data, dims32.ctypes.data, out.ctypes.data)
    if rc != 0:
        raise RuntimeError("mesh -> SDF failed with status %d" % rc)
    return VolumeGrid(out, bmin, bmax)


def analytic_sdf_grid(fn, bmin, bmax, dims):
    """Sample a closed-form SDF `fn(P[...,3]) -> d` at cell centres (for known-answer tests)."""
    bmin = np.asarray(bmin, dtype=np.float64)
    bmax = np.asarray(bmax, dtype=np.float64)
    dims = tuple(int(d) for d in dims)
    h = (bmax - bmin) / np.asarray(dims)
    ax = [bmin[a] + (np.arange(dims[a]) + 0.5) * h[a] for a in range(3)]
    X, Y, Z = np.meshgrid(*ax, indexing="ij")
    vals = fn(np.stack([X, Y, Z], axis=-1))
    return VolumeGrid(vals.astype(np.float32), bmin, bmax)


class Geometry3D:
    """The slice of Klampt's `Geometry3D` the reference's hot path calls
    (geometryopt.py:52-94): a typed payload + a current rigid transform + `convert`."""

    def __init__(self, data=None):
        self.data = data
        self._T = se3.identity()

    # -- payload --
    def type(self):
        if isinstance(self.data, TriangleMesh):
            return "TriangleMesh"
        if isinstance(self.data, PointCloud):
            return "PointCloud"
        if isinstance(self.data, VolumeGrid):
            return "VolumeGrid"
        return ""

    def empty(self):
        if self.data is None:
            return True
        if isinstance(self.data, TriangleMesh):
            return len(self.data.tris) == 0
        if isinstance(self.data, PointCloud):
            return self.data.numPoints() == 0
        return False

    def numElements(self):
        if isinstance(self.data, TriangleMesh):
            return len(self.data.tris)
        if isinstance(self.data, PointCloud):
            return self.data.numPoints()
        if isinstance(self.data, VolumeGrid):
            return int(np.prod(self.data.dims))
        return 0

    def clone(self):
        import copy
        g = Geometry3D(copy.deepcopy(self.data))
        g._T = (list(self._T[0]), list(self._T[1]))
        return g

    def getPointCloud(self):
        assert isinstance(self.data, PointCloud)
        return self.data

    def getVolumeGrid(self):
        assert isinstance(self.data, VolumeGrid)
        return self.data

    def getTriangleMesh(self):
        assert isinstance(self.data, TriangleMesh)
        return self.data

    # -- transforms --
    def setCurrentTransform(self, R, t):
        self._T = (list(R), list(t))

    def getCurrentTransform(self):
        return self._T

    def transform(self, R, t):
        """Bake a rigid transform (and, with `scale`, a scaling) into the LOCAL data, as Klampt's
        world loader does for `<geometry rotateX= scale=>` attributes."""
        M = so3.matrix(R)
        tt = np.asarray(t, dtype=np.float64)
        if isinstance(self.data, TriangleMesh):
            self.data.verts = self.data.verts @ M.T + tt
        elif isinstance(self.data, PointCloud):
            self.data.points = self.data.points @ M.T + tt
        else:
            raise NotImplementedError("transform() of a %s" % self.type())

    def scale(self, s):
        if isinstance(self.data, TriangleMesh):
            self.data.verts = self.data.verts * float(s)
        elif isinstance(self.data, PointCloud):
            self.data.points = self.data.points * float(s)
        else:
            raise NotImplementedError("scale() of a %s" % self.type())

    def getBBTight(self):
        lo, hi = self.data.bbox()
        return [float(v) for v in lo], [float(v) for v in hi]

    # -- conversion (geometryopt.py:60, 65, 79) --
    def convert(self, kind, param=0):
        lo, hi = self.data.bbox()
        span = float(np.max(np.asarray(hi) - np.asarray(lo)))
        if kind == "VolumeGrid":
            if isinstance(self.data, VolumeGrid):
                out = Geometry3D(self.data)
            elif isinstance(self.data, TriangleMesh):
                res = float(param) if param and param > 0 else span / AUTO_GRID_CELLS
                out = Geometry3D(mesh_to_sdf(self.data, res))
            else:
                raise ValueError("cannot convert a %s to a VolumeGrid" % self.type())
        elif kind == "PointCloud":
            if isinstance(self.data, PointCloud):
                out = Geometry3D(self.data)      # an existing cloud is not densified
            elif isinstance(self.data, TriangleMesh):
                res = float(param) if param and param > 0 else span / AUTO_PC_DIVS
                if _surface_sample_fn is not None and np.isfinite(res):
                    out = Geometry3D(PointCloud(np.asarray(_surface_sample_fn(self.data.verts, self.data.tris, res), dtype=np.float64)))
                else:
                    out = Geometry3D(PointCloud(sample_mesh_surface(self.data, res)))
            else:
                raise ValueError("cannot convert a %s to a PointCloud" % self.type())
        else:
            raise ValueError("unsupported conversion target %r" % (kind,))
        out._T = (list(self._T[0]), list(self._T[1]))
        return out

    def loadFile(self, fn):
        g = load_geometry(fn)
        self.data = g.data
        return True


# ---------------------------------------------------------------------------------------------
# loaders
# ---------------------------------------------------------------------------------------------
def load_off(fn):
    toks = open(fn).read().split()
    if toks[0] != "OFF":
        raise ValueError("%s: not an OFF file" % fn)
    nv, nf = int(toks[1]), int(toks[2])
    k = 4
    V = np.array(toks[k:k + 3 * nv], dtype=np.float64).reshape(nv, 3)
    k += 3 * nv
    tris = []
    for _ in range(nf):
        c = int(toks[k])
        idx = [int(t) for t in toks[k + 1:k + 1 + c]]
        k += 1 + c
        for j in range(1, c - 1):
            tris.append((idx[0], idx[j], idx[j + 1]))
    return TriangleMesh(V, np.array(tris, dtype=np.int32).reshape(-1, 3))


def load_pcd(fn):
    lines = open(fn).read().splitlines()
    for i, l in enumerate(lines):
        if l.startswith("DATA"):
            if "ascii" not in l:
                raise ValueError("%s: only ASCII PCD is supported" % fn)
            body = lines[i + 1:]
            break
    else:
        raise ValueError("%s: no DATA line" % fn)
    return PointCloud(np.array([[float(t) for t in l.split()[:3]] for l in body if l.strip()]))


def load_geometry(fn):
    ext = os.path.splitext(fn)[1].lower()
    if ext == ".off":
        return Geometry3D(load_off(fn))
    if ext == ".pcd":
        return Geometry3D(load_pcd(fn))
    raise ValueError("unsupported geometry file %r" % fn)


_fixture_cache = None


def fixture_geometry(name):
    """Geometry shipped under data/geometry.npz: 'cube', 'm797', 'm1062', 'sphere' (meshes) and
    'bunny' (397-point cloud) -- the inputs BASELINE.json's configs name."""
    global _fixture_cache
    if _fixture_cache is None:
        _fixture_cache = dict(np.load(os.path.join(_DATA, "geometry.npz")))
    if name == "bunny":
        return Geometry3D(PointCloud(_fixture_cache["bunny_pts"].copy()))
    return Geometry3D(TriangleMesh(_fixture_cache[name + "_verts"].copy(), _fixture_cache[name + "_tris"].copy()))


def unit_box_mesh(lo, hi):
    """Axis-aligned box as a 12-triangle mesh (synthetic stand-in geometry)."""
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    V = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
    T = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1],
                  [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]], dtype=np.int32)
    return TriangleMesh(V, T)
