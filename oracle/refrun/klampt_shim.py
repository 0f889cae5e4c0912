"""Stand-in `klampt` package for running the reference's own Python (see oracle/refrun/__init__.py).

Data model: geometry containers and file loaders come from this repository's host substrate
(`_host/geometry.py`: Geometry3D, PointCloud, VolumeGrid -- inputs, identical on both sides of every
comparison); the MATH (RobotModel FK and Jacobians, Trajectory, so3/se3/vectorops with Klampt's
conventions) comes from the frozen copies under oracle/frozen/.  The two geometric queries the
reference issues (geometryopt.py:104-108, 124-127, 139, 153) are answered by the CPU oracle:

    Geometry3D.distance_point(pt[, settings])   -> oracle.query      
    Geometry3D.distance_ext(other, settings)    -> oracle.min_distance (Appendix A.3)

so the Klampt-side arithmetic stays a restatement; the result objects carry the fields the reference
reads (d, hasClosestPoints, cp1, cp2, hasGradients, grad1).
"""
import sys
import types

import numpy as np

from oracle import oracle as O
from semiinfiniteoptimization_b200._host import geometry as _hg       # data containers + file loaders only (inputs)
from oracle.frozen import robot as _hr                                 # FK / Jacobians / so3 / se3 / trajectories: frozen copies,
from oracle.frozen import se3, so3, trajectory as _ht, vectorops       # so product changes cannot move the goldens


class DistanceQuerySettings:
    def __init__(self):
        self.relErr = 0.0
        self.absErr = 0.0
        self.upperBound = float('inf')


class DistanceQueryResult:
    def __init__(self):
        self.d = float('inf')
        self.hasClosestPoints = False
        self.hasGradients = False
        self.cp1 = [0.0] * 3
        self.cp2 = [0.0] * 3
        self.grad1 = [0.0] * 3
        self.grad2 = [0.0] * 3
        self.elem1 = -1
        self.elem2 = -1


class GeometricPrimitive:
    def __init__(self):
        self.type = ""
        self.properties = []

    def setPoint(self, pt):
        self.type = "Point"
        self.properties = list(pt)


def _inv12(T):
    T = np.asarray(se3.to_rowmajor12(T), dtype=np.float64).reshape(3, 4)
    out = np.empty((3, 4))
    out[:, :3] = T[:, :3].T
    out[:, 3] = -T[:, :3].T @ T[:, 3]
    return out


def _mul34(A, B):
    out = np.empty((3, 4))
    out[:, :3] = A[:, :3] @ B[:, :3]
    out[:, 3] = A[:, :3] @ B[:, 3] + A[:, 3]
    return out


def _oracle_grid(geom):
    vg = geom.getVolumeGrid()
    g = getattr(vg, "_oracle_grid", None)
    if g is None:
        g = O.Grid(vg.values, vg.bmin, vg.bmax)
        vg._oracle_grid = g
    return g


def _cloud32(geom):
    pc = geom.getPointCloud()
    c = getattr(pc, "_cloud32", None)
    if c is None or len(c) != pc.numPoints():
        c = np.ascontiguousarray(pc.points, dtype=np.float32)
        pc._cloud32 = c
    return c


def distance_point(self, pt, settings=None):
    """Klampt Geometry3D.distance_point on a VolumeGrid (gopt:104-108, 139, 153), Appendix A.2."""
    assert self.type() == "VolumeGrid", "distance_point is only restated for VolumeGrid geometries"
    pt = [float(v) for v in pt[:3]]
    d, g = O.query(_oracle_grid(self), _inv12(self._T).reshape(12), [pt])
    res = DistanceQueryResult()
    res.d = float(d[0])
    res.hasClosestPoints = True
    res.hasGradients = True
    # cp1 = surface point of the grid object, cp2 = the query point, grad1 = -gradient
    res.cp1 = [pt[k] - float(g[0, k]) * res.d for k in range(3)]
    res.cp2 = list(pt)
    res.grad1 = [-float(v) for v in g[0]]
    res.grad2 = [float(v) for v in g[0]]
    return res


def distance_ext(self, other, settings=None):
    """Klampt Geometry3D.distance_ext, PointCloud (self) vs VolumeGrid (other) (gopt:124-127), Appendix A.3;
    or the flipped roles."""
    if self.type() == "VolumeGrid" and other.type() == "PointCloud":
        r = distance_ext(other, self, settings)
        r.cp1, r.cp2 = r.cp2, r.cp1
        return r
    assert self.type() == "PointCloud" and other.type() == "VolumeGrid"
    bound = None
    if settings is not None and settings.upperBound < float('inf'):
        bound = [float(settings.upperBound)]
    Tpc = np.asarray(se3.to_rowmajor12(self._T), dtype=np.float64).reshape(3, 4)
    T_rel = _mul34(_inv12(other._T), Tpc)
    cloud = _cloud32(self)
    dmin, arg = O.min_distance(_oracle_grid(other), cloud, T_rel.reshape(1, 12), bound=bound)
    res = DistanceQueryResult()
    res.d = float(dmin[0])
    if arg[0] >= 0:
        p = cloud[int(arg[0])].astype(np.float64)
        cp1 = [float(v) for v in Tpc[:, :3] @ p + Tpc[:, 3]]
        res.hasClosestPoints = True
        res.cp1 = cp1
        res.cp2 = distance_point(other, cp1).cp1
        res.elem1 = int(arg[0])
    return res


Geometry3D = _hg.Geometry3D
PLANNER_FACTORY = None          # callable(world, robot, target, **settings) -> object with planMore(n) / getPath()


def install():
    """Registers the stand-in modules under the names the reference imports."""
    if "klampt" in sys.modules and getattr(sys.modules["klampt"], "_sio_shim", False):
        return
    # the two Klampt queries become methods of the host Geometry3D for the lifetime of this (test) process
    _hg.Geometry3D.distance_point = distance_point
    _hg.Geometry3D.distance_ext = distance_ext
    k = types.ModuleType("klampt")
    k._sio_shim = True
    k.__path__ = []
    for name, obj in [("Geometry3D", Geometry3D), ("DistanceQuerySettings", DistanceQuerySettings),
                      ("DistanceQueryResult", DistanceQueryResult), ("GeometricPrimitive", GeometricPrimitive),
                      ("PointCloud", _hg.PointCloud), ("VolumeGrid", _hg.VolumeGrid), ("TriangleMesh", _hg.TriangleMesh),
                      ("RobotModel", _hr.RobotModel), ("RobotModelLink", _hr.RobotModelLink),
                      ("RigidObjectModel", _hr.RigidObjectModel), ("WorldModel", _hr.WorldModel)]:
        setattr(k, name, obj)
    k.__all__ = [n for n in dir(k) if not n.startswith("_")]
    km = types.ModuleType("klampt.math")
    km.__path__ = []
    km.vectorops, km.so3, km.se3 = vectorops, so3, se3
    kmodel = types.ModuleType("klampt.model")
    kmodel.__path__ = []
    ktraj = types.ModuleType("klampt.model.trajectory")
    ktraj.Trajectory, ktraj.RobotTrajectory = _ht.Trajectory, _ht.RobotTrajectory
    kmodel.trajectory = ktraj
    k.math, k.model = km, kmodel
    # klampt.plan.robotplanning.planToConfig (planopt.py:2, 114): Klampt's planners are absent; the reference's
    # planner / optimiser alternation is run with whatever stand-in the caller puts into PLANNER_FACTORY
    kplan = types.ModuleType("klampt.plan")
    kplan.__path__ = []
    krp = types.ModuleType("klampt.plan.robotplanning")

    def planToConfig(world, robot, target, **settings):
        if PLANNER_FACTORY is None:
            raise RuntimeError("klampt.plan.robotplanning is a stand-in: set oracle.refrun.klampt_shim.PLANNER_FACTORY")
        return PLANNER_FACTORY(world, robot, target, **settings)
    krp.planToConfig = planToConfig
    kplan.robotplanning = krp
    k.plan = kplan
    sys.modules.update({"klampt.plan": kplan, "klampt.plan.robotplanning": krp})
    sys.modules.update({"klampt": k, "klampt.math": km, "klampt.math.vectorops": vectorops, "klampt.math.so3": so3,
                        "klampt.math.se3": se3, "klampt.model": kmodel, "klampt.model.trajectory": ktraj})
