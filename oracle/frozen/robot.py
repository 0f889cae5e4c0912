"""FROZEN COPY (oracle/frozen, round 2) of semiinfiniteoptimization_b200/_host/robot.py as of commit 92f6a08 -- TEST INFRASTRUCTURE.

The golden generator (tools/make_golden.py -> oracle/refrun) answers the reference's Klampt / OSQP calls with THIS file,
not with the product's module, so a change to the product can no longer rewrite tests/golden/.  Do not edit it to follow
the product: the product is tested AGAINST it (tests/test_frozen_cpu.py) at a stated tolerance.

Original module docstring follows.

Host-side robot model: the slice of Klampt's `RobotModel` / `RobotModelLink` /
`WorldModel` that the reference's hot-path callers use (geometryopt.py:187-366, 444-484,
694-699, 923-926; robotposeopt.py:25-52; robottrajopt.py:36-71).

Klampt is not available to this build, so kinematic STRUCTURE (parents, TParent, axes, limits) is
parsed here from the `.rob` fixture and handed to the device once (`sio_robot_create`).  The host
model keeps a float64 FK for bookkeeping calls such as `link.getTransform()`; every query on the
oracle path does its kinematics on the device (`sio_fk`, `sio_chain_rows`,
`sio_robot_min_distance`).

TX90L link meshes are missing from the reference checkout (data/ARC.rob:20 points into an absent
Klampt-examples tree), so `load_arc_robot` builds DECLARED SYNTHETIC link geometry: one box per
link spanning the offset to its child, plus the `mount 6` cube the file itself specifies
(SURVEY.md 8d, C3).
"""
import json
import math
import os

import numpy as np

from . import se3, so3
from semiinfiniteoptimization_b200._host.geometry import Geometry3D, TriangleMesh, fixture_geometry, unit_box_mesh

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "semiinfiniteoptimization_b200", "data")


def _rot_axis(axis, q):
    x, y, z = axis
    c, s = math.cos(q), math.sin(q)
    C = 1.0 - c
    return np.array([[c + x * x * C, x * y * C - z * s, x * z * C + y * s],
                     [y * x * C + z * s, c + y * y * C, y * z * C - x * s],
                     [z * x * C - y * s, z * y * C + x * s, c + z * z * C]])


class RobotModelLink:
    def __init__(self, robot, index):
        self._robot = robot
        self.index = index

    def robot(self):
        return self._robot

    def getName(self):
        return self._robot.link_names[self.index]

    def getIndex(self):
        return self.index

    def getParent(self):
        return int(self._robot.parents[self.index])

    def getParentTransform(self):
        return se3.from_rowmajor12(self._robot.tparent[self.index])

    def getAxis(self):
        return [float(v) for v in self._robot.axes[self.index]]

    def geometry(self):
        return self._robot.geometries[self.index]

    def getTransform(self):
        return se3.from_rowmajor12(self._robot.link_T[self.index].reshape(12))

    def getLocalPosition(self, pworld):
        T = self._robot.link_T[self.index]
        return [float(v) for v in T[:, :3].T @ (np.asarray(pworld[:3], dtype=np.float64) - T[:, 3])]

    def getWorldPosition(self, plocal):
        T = self._robot.link_T[self.index]
        return [float(v) for v in T[:, :3] @ np.asarray(plocal[:3], dtype=np.float64) + T[:, 3]]

    def getPositionJacobian(self, plocal):
        """3 x n list-of-lists: d(world position of the link-fixed point)/dq."""
        r = self._robot
        p = np.asarray(self.getWorldPosition(plocal))
        J = np.zeros((3, r.n))
        j = self.index
        while j >= 0:
            T = r.link_T[j]
            w = T[:, :3] @ r.axes[j]
            J[:, j] = w if r.jtypes[j] == 1 else np.cross(w, p - T[:, 3])
            j = int(r.parents[j])
        return [list(map(float, row)) for row in J]


class RobotModel:
    def __init__(self, name, parents, tparent12, axes, qmin, qmax, jtypes=None, geometries=None, link_names=None):
        self.name = name
        self.n = len(parents)
        self.parents = np.asarray(parents, dtype=np.int32)
        self.tparent = np.asarray(tparent12, dtype=np.float64).reshape(self.n, 12)
        self.axes = np.asarray(axes, dtype=np.float64).reshape(self.n, 3)
        self.jtypes = np.zeros(self.n, dtype=np.int32) if jtypes is None else np.asarray(jtypes, dtype=np.int32)
        self.qmin = [float(v) for v in qmin]
        self.qmax = [float(v) for v in qmax]
        self.geometries = geometries if geometries is not None else [Geometry3D() for _ in range(self.n)]
        self.link_names = link_names or ["link%d" % i for i in range(self.n)]
        self.q = [0.0] * self.n
        self.link_T = np.zeros((self.n, 3, 4))
        self.setConfig(self.q)

    def getName(self):
        return self.name

    def numLinks(self):
        return self.n

    def link(self, i):
        if isinstance(i, str):
            i = self.link_names.index(i)
        return RobotModelLink(self, int(i))

    def getJointLimits(self):
        return list(self.qmin), list(self.qmax)

    def getConfig(self):
        return list(self.q)

    def setConfig(self, q):
        assert len(q) == self.n, "configuration has %d entries, robot has %d links" % (len(q), self.n)
        self.q = [float(v) for v in q]
        self.link_T = self.fk(self.q)
        for i, g in enumerate(self.geometries):
            if g is not None and not g.empty():
                g.setCurrentTransform(*se3.from_rowmajor12(self.link_T[i].reshape(12)))

    def fk(self, q):
        """float64 forward kinematics -> (n,3,4) world <- link (same recurrence as the device's k_fk)."""
        out = np.zeros((self.n, 3, 4))
        for l in range(self.n):
            Tp = self.tparent[l].reshape(3, 4)
            if self.jtypes[l] == 1:
                J = np.eye(4)[:3]
                J = J.copy()
                J[:, 3] = self.axes[l] * q[l]
            else:
                J = np.zeros((3, 4))
                J[:, :3] = _rot_axis(self.axes[l], q[l])
            L = np.zeros((3, 4))
            L[:, :3] = Tp[:, :3] @ J[:, :3]
            L[:, 3] = Tp[:, :3] @ J[:, 3] + Tp[:, 3]
            p = int(self.parents[l])
            if p < 0:
                out[l] = L
            else:
                out[l, :, :3] = out[p, :, :3] @ L[:, :3]
                out[l, :, 3] = out[p, :, :3] @ L[:, 3] + out[p, :, 3]
        return out

    # Klampt's RobotModel.interpolate / interpolateDeriv / distance for plain (non-spin) joints
    def interpolate(self, a, b, u):
        return [x + u * (y - x) for x, y in zip(a, b)]

    def interpolateDeriv(self, a, b):
        return [y - x for x, y in zip(a, b)]

    def distance(self, a, b):
        return math.sqrt(sum((x - y) ** 2 for x, y in zip(a, b)))

    def randomizeConfig(self, rng=None):
        rng = rng or np.random.default_rng()
        self.setConfig([lo + rng.uniform() * (hi - lo) for lo, hi in zip(self.qmin, self.qmax)])


class RigidObjectModel:
    def __init__(self, name, geom):
        self.name = name
        self._geom = geom

    def getName(self):
        return self.name

    def geometry(self):
        return self._geom

    def getTransform(self):
        return self._geom.getCurrentTransform()

    def setTransform(self, R, t):
        self._geom.setCurrentTransform(R, t)


class WorldModel:
    """Enough of Klampt's WorldModel to load the two TX90 scenes BASELINE.json names."""

    def __init__(self):
        self.robots = []
        self.rigidObjects = []
        self.terrains = []

    def numRobots(self):
        return len(self.robots)

    def numRigidObjects(self):
        return len(self.rigidObjects)

    def numTerrains(self):
        return len(self.terrains)

    def robot(self, i):
        return self.robots[i]

    def rigidObject(self, i):
        return self.rigidObjects[i]

    def readFile(self, fn):
        """Accepts the fixture names 'tx90_geom_test2.xml' / 'tx90_geom_test3.xml' (any directory)."""
        key = os.path.basename(fn)
        worlds = json.load(open(os.path.join(_DATA, "worlds.json")))
        if key not in worlds:
            raise IOError("world %r is not among the shipped fixtures %s" % (fn, sorted(worlds)))
        w = worlds[key]
        self.robots.append(load_arc_robot())
        for ro in w["rigidObjects"]:
            g = fixture_geometry(os.path.splitext(ro["mesh"])[0])
            # Klampt applies scale, then the rotateX/Y/Z attributes, to the mesh itself
            if ro.get("scale", 1) != 1:
                g.scale(ro["scale"])
            if ro.get("rotateX", 0):
                g.transform(so3.rotation([1, 0, 0], ro["rotateX"]), [0, 0, 0])
            g.setCurrentTransform(so3.identity(), list(ro["position"]))
            self.rigidObjects.append(RigidObjectModel(ro["name"], g))
        return True


def _oriented_box(p0, p1, half):
    """Box mesh around the segment p0->p1 with square cross-section `half` (link-local frame)."""
    p0 = np.asarray(p0, dtype=np.float64)
    p1 = np.asarray(p1, dtype=np.float64)
    d = p1 - p0
    L = np.linalg.norm(d)
    if L < 1e-9:
        return unit_box_mesh(p0 - half, p0 + half)
    z = d / L
    a = np.array([1.0, 0, 0]) if abs(z[0]) < 0.9 else np.array([0, 1.0, 0])
    x = np.cross(a, z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    base = unit_box_mesh([-half, -half, -half], [half, half, L + half])
    R = np.stack([x, y, z], axis=1)
    return TriangleMesh(base.verts @ R.T + p0, base.tris)


def load_arc_robot(synthetic_half_width=0.05):
    """data/ARC.rob (Staubli TX90L kinematics).  TParent is row-major R then t per link
    (ARC.rob:2-8); all axes z; link 0 fixed (qMin = qMax = 0)."""
    rob = json.load(open(os.path.join(_DATA, "arc_rob.json")))
    n = len(rob["parents"])
    tp = np.asarray(rob["TParent"], dtype=np.float64).reshape(n, 12)
    tparent = np.zeros((n, 3, 4))
    tparent[:, :, :3] = tp[:, :9].reshape(n, 3, 3)
    tparent[:, :, 3] = tp[:, 9:]
    axes = np.asarray(rob["axis"], dtype=np.float64).reshape(n, 3)
    qmin = np.radians(rob["qMinDeg"])
    qmax = np.radians(rob["qMaxDeg"])
    parents = rob["parents"]
    # synthetic link geometry: box from the link origin to each child's origin (link-local)
    geoms = []
    for l in range(n):
        children = [c for c in range(n) if parents[c] == l]
        if children:
            meshes = [_oriented_box([0, 0, 0], tparent[c][:, 3], synthetic_half_width) for c in children]
            V = np.vstack([m.verts for m in meshes])
            off = np.cumsum([0] + [len(m.verts) for m in meshes[:-1]])
            T = np.vstack([m.tris + o for m, o in zip(meshes, off)])
            geoms.append(Geometry3D(TriangleMesh(V, T)))
        else:
            geoms.append(Geometry3D())
    # `mount 6 cube.off  <3x3 scale/rotation (column-major rows as written)> <t>` (ARC.rob:21)
    m = rob.get("mount")
    if m is not None:
        M = np.asarray(m["T"][:9], dtype=np.float64).reshape(3, 3)
        t = np.asarray(m["T"][9:], dtype=np.float64)
        cube = fixture_geometry("cube").data
        geoms[m["link"]] = Geometry3D(TriangleMesh(cube.verts @ M.T + t, cube.tris))
    names = ["tx90_link%d" % i for i in range(n)]
    return RobotModel("tx90ball", parents, tparent.reshape(n, 12), axes, qmin, qmax, geometries=geoms, link_names=names)
