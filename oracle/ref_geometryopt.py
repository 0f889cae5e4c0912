"""CPU restatement of the reference's collision classes -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (see oracle/oracle.py): Klampt's arithmetic is restated, not linked.

These classes follow the reference's per-call algorithms literally (one oracle call per parameter,
per milestone and per bisection sample; heap-based best-first time bisection with Lipschitz
pruning), on top of the CPU oracle (oracle/sio_oracle.c).  They exist so tests can drive the SAME
host SIP loop with a CPU-backed and a device-backed constraint and compare values, arg-min
parameters, Jacobian rows and whole solution traces.  Nothing in the product imports this module.

Followed reference code (file:line under /root/reference/semiinfinite/geometryopt.py):
    PenetrationDepthGeometry.distance / normal          95-168
    RobotKinematicsCache                                187-207
    RobotTrajectoryCache                                209-366
    ObjectCollisionConstraint                           394-420
    RobotLinkCollisionConstraint                        457-484
    intervalLipschitzMinimum                            554-563
    RobotLinkTrajectoryCollisionConstraint              565-675
    RobotTrajectoryCollisionConstraint                  678-956
"""
import heapq

import numpy as np
import scipy.sparse

from oracle import oracle as O
from semiinfiniteoptimization_b200._host import se3, so3, vectorops
from semiinfiniteoptimization_b200._host.geometry import Geometry3D
from semiinfiniteoptimization_b200._host.trajectory import RobotTrajectory
from semiinfiniteoptimization_b200.constraint import SemiInfiniteConstraintInterface


def _inv12(T12):
    T = np.asarray(T12, dtype=np.float64).reshape(3, 4)
    out = np.empty((3, 4))
    out[:, :3] = T[:, :3].T
    out[:, 3] = -T[:, :3].T @ T[:, 3]
    return out.reshape(12)


def _mul12(A12, B12):
    A = np.asarray(A12, dtype=np.float64).reshape(3, 4)
    B = np.asarray(B12, dtype=np.float64).reshape(3, 4)
    out = np.empty((3, 4))
    out[:, :3] = A[:, :3] @ B[:, :3]
    out[:, 3] = A[:, :3] @ B[:, 3] + A[:, 3]
    return out.reshape(12)


class RefPDG:
    """gopt:38-168 on the CPU oracle.  Shares nothing with the device."""

    PRUNED = False        # True: cloud-vs-grid minima by the hierarchy-pruned search (bench.py's "pruning on" CPU legs)

    def __init__(self, geom, gridres=0, pcres=0):
        if isinstance(geom, RefPDG):
            self.geom, self.grid, self.pc, self.pcdata = geom.geom, geom.grid, geom.pc, geom.pcdata
            self.G, self.cloud32 = geom.G, geom.cloud32
            self.T = geom.T
        else:
            assert isinstance(geom, Geometry3D)
            self.geom = geom
            self.grid = self.pc = self.pcdata = None
            self.G = self.cloud32 = None
            self.T = geom.getCurrentTransform()
            if gridres is not None:
                self.grid = geom.convert('VolumeGrid', gridres)
                vg = self.grid.getVolumeGrid()
                self.G = O.Grid(vg.values, vg.bmin, vg.bmax)
            if pcres is not None:
                self.pc = geom.convert('PointCloud', pcres)
                self.pcdata = self.pc.getPointCloud()
                self.cloud32 = np.ascontiguousarray(self.pcdata.points, dtype=np.float32)

    def setTransform(self, T):
        self.T = (list(T[0]), list(T[1]))

    def getTransform(self):
        return self.T

    def T12(self):
        return se3.to_rowmajor12(self.T)

    def distance(self, other, bound=None):
        if isinstance(other, (list, tuple, np.ndarray)):
            pt = [float(v) for v in other[:3]]
            d, g = O.query(self.G, _inv12(self.T12()), [pt])
            d = float(d[0])
            return d, [pt[k] - g[0, k] * d for k in range(3)], pt
        assert isinstance(other, RefPDG) and other.G is not None
        T_rel = _mul12(_inv12(other.T12()), self.T12())
        if RefPDG.PRUNED:
            # the hierarchy-pruned CPU arm (what Klampt's distance_ext does in spirit): same answer wherever it is below the
            # bound, which is all this method reads
            tree = getattr(self, "_tree", None)
            if tree is None:
                tree = self._tree = O.CloudTree(self.cloud32)
            lip = getattr(other, "_lip", None)
            if lip is None:
                lip = other._lip = O.grid_lipschitz(other.G)
            dmin, arg, _ = O.min_distance_pruned(other.G, tree, T_rel, bound=None if bound is None else [bound],
                                                 use_bound=bound is not None, lip=lip)
        else:
            dmin, arg = O.min_distance(other.G, self.cloud32, T_rel, bound=None if bound is None else [bound])
        d = float(dmin[0])
        if arg[0] < 0 or (bound is not None and d >= bound):
            return d, [0.0] * 3, [0.0] * 3
        p = self.cloud32[int(arg[0])].astype(np.float64)
        T = self.T12().reshape(3, 4)
        cp1 = [float(v) for v in T[:, :3] @ p + T[:, 3]]
        return d, cp1, other.distance(cp1)[1]

    def normal(self, pt):
        d, g = O.query(self.G, _inv12(self.T12()), [[float(v) for v in pt[:3]]])
        return [-float(v) for v in g[0]]


class RefObjectCollisionConstraint(SemiInfiniteConstraintInterface):
    """gopt:394-420."""

    def __init__(self, obj, env):
        self.objgeom = RefPDG(obj)
        self.env = RefPDG(env)

    def setx(self, T):
        self.objgeom.setTransform(T)

    def value(self, T, pt):
        return self.objgeom.distance(pt)[0]

    def minvalue(self, T, bound=None):
        dmin, cp1, cp2 = self.env.distance(self.objgeom, bound)
        if bound is not None and dmin >= bound:
            return []
        return dmin, cp1

    def df_dx(self, T, pt):
        worlddir = self.objgeom.normal(pt)
        cporiented = so3.apply(T[0], vectorops.sub(pt[:3], T[1]))     # R (pt - t), as written (gopt:415)
        return np.array(list(vectorops.cross(cporiented, worlddir)) + list(worlddir))

    def domain(self):
        return RefPenetrationGeometryDomain(self.env)


class RefPenetrationGeometryDomain:
    """gopt:171-183: the world positions of a geometry's cloud points."""

    def __init__(self, g):
        self.g = g

    def sample(self):
        import random
        return se3.apply(self.g.getTransform(), self.g.pcdata.getPoint(random.randint(0, self.g.pcdata.numPoints() - 1)))

    def extrema(self):
        return [se3.apply(self.g.getTransform(), self.g.pcdata.getPoint(i)) for i in range(self.g.pcdata.numPoints())]


class RefRobotKinematicsCache:
    """gopt:187-207."""

    def __init__(self, robot, gridres=0.05, pcres=0.02):
        self.robot = robot
        self.geometry = []
        for i in range(robot.numLinks()):
            g = robot.link(i).geometry()
            self.geometry.append(None if (g is None or g.empty()) else RefPDG(g, gridres, pcres))
        self.dirty = True

    def set(self, q):
        if self.dirty:
            self.robot.setConfig(q)
            for i in range(self.robot.numLinks()):
                if self.geometry[i] is not None:
                    self.geometry[i].setTransform(self.robot.link(i).getTransform())
            self.dirty = False

    def clear(self):
        self.dirty = True


class RefRobotLinkCollisionConstraint(SemiInfiniteConstraintInterface):
    """gopt:457-484."""

    def __init__(self, link, env, robotcache):
        self.link = link
        self.robot = robotcache
        self.env = RefPDG(env)

    def setx(self, q):
        self.robot.set(q)

    def clearx(self):
        self.robot.clear()

    def value(self, q, pt):
        return self.robot.geometry[self.link.index].distance(pt)[0]

    def minvalue(self, q, bound=None):
        dmin, cp1, cp2 = self.env.distance(self.robot.geometry[self.link.index], bound)
        if bound is not None and dmin >= bound:
            return []
        return dmin, cp1

    def df_dx(self, q, pt):
        localpt = self.link.getLocalPosition(pt[:3])
        worlddir = self.robot.geometry[self.link.index].normal(pt)
        Jp = self.link.getPositionJacobian(localpt)
        return np.dot(np.array(Jp).T, worlddir)


class RefRobotTrajectoryCache:
    """gopt:209-366 (host FK per milestone)."""

    def __init__(self, kinematics, numPoints, times=False, qstart=None, qend=None):
        self.robot = kinematics.robot
        self.kinematics = kinematics
        self.numPoints = numPoints
        self.times = times
        self.qstart, self.qend = qstart, qend
        self.tstart, self.tend = 0, 1
        self.trajectory = None
        self.interMilestoneLipschitzBounds = []
        n = self.robot.numLinks()
        self.lipschitzConstants = np.zeros((n, n))
        radii = np.zeros((n, n))
        ancestors = [[] for _ in range(n)]
        for l in range(n):
            p = self.robot.link(l).getParent()
            if p >= 0:
                ancestors[l] = [p] + ancestors[p]
        for l in range(n):
            link = self.robot.link(l)
            p = link.getParent()
            if p >= 0:
                T = link.getParentTransform()
                for a in ancestors[l]:
                    radii[a, l] = radii[a, p] + vectorops.norm(T[1])
                radii[p, l] += vectorops.norm(T[1])
            axis = link.getAxis()
            g = self.kinematics.geometry[l]
            Kgeom = 0
            if g is not None:
                for v in g.pcdata.points:
                    if p < 0:
                        Kgeom = max(Kgeom, vectorops.norm(vectorops.cross(v, axis)))
                    else:
                        Kgeom = max(Kgeom, vectorops.norm(v))
            for a in ancestors[l]:
                self.lipschitzConstants[a, l] = radii[a, l] + Kgeom
            self.lipschitzConstants[l, l] = Kgeom
        self.dirty = True

    def set(self, x):
        if self.dirty:
            self.trajectory = self.stateToTrajectory(x)
            m = len(self.trajectory.milestones)
            self.interMilestoneLipschitzBounds = [
                self.workspaceMovementBounds(self.trajectory.milestones[i], self.trajectory.milestones[i + 1]) for i in range(m - 1)]
            self.linkTransforms = []
            for i in range(m):
                self.robot.setConfig(self.trajectory.milestones[i])
                self.linkTransforms.append([self.robot.link(j).getTransform() for j in range(self.robot.numLinks())])
            self.dirty = False

    def clear(self):
        self.dirty = True

    def workspaceMovementBounds(self, q1, q2):
        dq = np.abs(self.robot.interpolateDeriv(q1, q2))
        return np.dot(self.lipschitzConstants.T, dq)

    def stateToTrajectory(self, x):
        n = self.robot.numLinks()
        m = self.numPoints
        x = list(x)
        assert self.times is False and len(x) == m * n
        milestones = []
        if self.qstart:
            milestones.append(list(self.qstart))
        for i in range(m):
            milestones.append(x[i * n:(i + 1) * n])
        if self.qend:
            milestones.append(list(self.qend))
        M = len(milestones)
        times = [self.tstart + float(i) / (M - 1) * (self.tend - self.tstart) for i in range(M)]
        return RobotTrajectory(self.robot, times, milestones)

    def trajectoryToState(self, traj):
        """gopt:343-356 (times is False here)."""
        if self.qstart is not None:
            assert list(traj.milestones[0]) == list(self.qstart), "Trajectory needs to match the start configuration"
        if self.qend is not None:
            assert list(traj.milestones[-1]) == list(self.qend), "Trajectory needs to match the end configuration"
        sshift = 1 if self.qstart is not None else 0
        eshift = 1 if self.qend is not None else 0
        assert len(traj.milestones) == self.numPoints + sshift + eshift
        assert traj.times[0] == self.tstart and traj.times[-1] == self.tend
        return [v for q in traj.milestones[sshift:len(traj.milestones) - eshift] for v in q]

    def straightLineState(self):
        """gopt:357-366."""
        assert self.qstart is not None, "Need a start config for a straight line state"
        qend = self.qend if self.qend is not None else self.qstart
        sshift = 1
        eshift = 1 if self.qend is not None else 0
        us = [float(i + sshift) / (self.numPoints + sshift + eshift - 1) for i in range(self.numPoints)]
        qs = [self.robot.interpolate(self.qstart, qend, u) for u in us]
        ts = [] if self.times is not True else [self.tstart + u * (self.tend - self.tstart) for u in us]
        return ts + [v for q in qs for v in q]


def intervalLipschitzMinimum(a, b, K):
    """gopt:554-563."""
    if K == 0:
        return min(a, b)
    u = (a - b + K) / (2 * K)
    if u < 0 or u > 1:
        return min(a - 0.5 * K, b - 0.5 * K)
    return a - u * K


class RefRobotTrajectoryCollisionConstraint(SemiInfiniteConstraintInterface):
    """gopt:678-956: milestone sweep, Lipschitz edge bounds, best-first time bisection."""

    def __init__(self, envs, trajectorycache):
        self.robot = trajectorycache.robot
        self.envs = envs
        self.trajectory = trajectorycache
        self.epsilon = 1e-2
        qmin, qmax = self.robot.getJointLimits()
        self.activeLinks = [i for i in range(self.robot.numLinks()) if qmin[i] < qmax[i]]
        self.numchecks = 0

    def setx(self, x):
        self.trajectory.set(x)

    def clearx(self):
        self.trajectory.clear()

    def value(self, x, y):
        link, env, t = int(y[0]), int(y[1]), y[2]
        q = self.trajectory.trajectory.eval(t)
        self.trajectory.kinematics.set(q)
        res = self.trajectory.kinematics.geometry[link].distance(y[3:6])
        self.trajectory.kinematics.clear()
        return res[0]

    def minvalue(self, x, bound=None):
        inflation = 1e-3
        tc = self.trajectory
        traj = tc.trajectory
        kin = tc.kinematics
        robot = tc.robot
        dmin = float('inf') if bound is None else bound
        linkmin = envmin = tmin = qmin = None
        M = len(traj.milestones)
        active_milestones = [dict() for _ in range(M)]
        for link in self.activeLinks[::-1]:
            geom = kin.geometry[link]
            if geom is None:
                continue
            for i in range(M):
                geom.setTransform(tc.linkTransforms[i][link])
                dbnd = 1000
                if i > 0:
                    dbnd = tc.interMilestoneLipschitzBounds[i - 1][link]
                if i + 1 < M:
                    dbnd = min(dbnd, tc.interMilestoneLipschitzBounds[i][link])
                dimin = dmin + dbnd * 0.5 + inflation
                for j, env in enumerate(self.envs):
                    dij = env.distance(geom, dimin)[0]
                    self.numchecks += 1
                    if dij < dimin:
                        dimin = dij
                        active_milestones[i][link, j] = dij
                    if dij < dmin:
                        dmin, tmin, qmin, envmin, linkmin = dij, traj.times[i], traj.milestones[i], j, link
        edgequeue = []
        tie = 0
        for i in range(M - 1):
            di, dn = active_milestones[i], active_milestones[i + 1]
            dbnd = tc.interMilestoneLipschitzBounds[i]
            active = dict()
            for (link, env), d in di.items():
                if (link, env) in dn:
                    dlow = intervalLipschitzMinimum(d, dn[link, env], dbnd[link])
                else:
                    dlow = d - dbnd[link] * 0.5
                if dlow < dmin:
                    active[link, env] = dlow
            for (link, env), d in dn.items():
                if (link, env) not in di:
                    dlow = d - dbnd[link] * 0.5
                    if dlow < dmin:
                        active[link, env] = dlow
            if len(active) > 0:
                heapq.heappush(edgequeue, (min(active.values()), tie, active, (i, 0, 1, di, dn)))
                tie += 1
        while len(edgequeue) > 0:
            dlowmin, _, dlow, (edge, u, v, du, dv) = heapq.heappop(edgequeue)
            if dlowmin >= dmin:
                continue
            midpt = (u + v) * 0.5
            duration = v - u
            q = traj.interpolate(traj.milestones[edge], traj.milestones[edge + 1], midpt, traj.times[edge + 1] - traj.times[edge])
            kin.set(q)
            active = dict()
            for (link, env), d in dlow.items():
                if d >= dmin:
                    continue
                geom = kin.geometry[link]
                dbnd = tc.interMilestoneLipschitzBounds[edge][link]
                dij = self.envs[env].distance(geom, dmin + dbnd * duration * 0.5 + inflation)[0]
                self.numchecks += 1
                if dij < dmin + dbnd * duration * 0.5:
                    active[link, env] = dij
                if dij < dmin:
                    dmin = dij
                    tmin = traj.times[edge] + midpt * (traj.times[edge + 1] - traj.times[edge])
                    qmin, linkmin, envmin = q, link, env
            kin.clear()
            duration *= 0.5
            if len(active) > 0 and duration > self.epsilon:
                dbnd = tc.interMilestoneLipschitzBounds[edge]
                c1, c2 = dict(), dict()
                for (link, env), d in active.items():
                    if (link, env) in du:
                        dl = intervalLipschitzMinimum(du[link, env], d, duration * dbnd[link])
                    else:
                        dl = d - dbnd[link] * duration * 0.5
                    if dl < dmin:
                        c1[link, env] = dl
                    if (link, env) in dv:
                        dl = intervalLipschitzMinimum(d, dv[link, env], duration * dbnd[link])
                    else:
                        dl = d - dbnd[link] * duration * 0.5
                    if dl < dmin:
                        c2[link, env] = dl
                if len(c1) > 0:
                    heapq.heappush(edgequeue, (min(c1.values()), tie, c1, (edge, u, midpt, du, active)))
                    tie += 1
                if len(c2) > 0:
                    heapq.heappush(edgequeue, (min(c2.values()), tie, c2, (edge, midpt, v, active, dv)))
                    tie += 1
        if tmin is None:
            return []
        robot.setConfig(qmin)
        geom = kin.geometry[linkmin]
        geom.setTransform(robot.link(linkmin).getTransform())
        d, cp1, cp2 = self.envs[envmin].distance(geom)
        return dmin, (linkmin, envmin, tmin) + tuple(cp1)

    def df_dx(self, x, y):
        link, env, t = int(y[0]), int(y[1]), y[2]
        pt = y[3:6]
        tc = self.trajectory
        robot = self.robot
        q = tc.trajectory.eval(t)
        tc.kinematics.set(q)
        worlddir = tc.kinematics.geometry[link].normal(pt)
        tc.kinematics.clear()
        localpt = robot.link(link).getLocalPosition(pt[:3])
        Jp = robot.link(link).getPositionJacobian(localpt)
        dq = np.dot(np.array(Jp).T, worlddir)
        i, u = tc.trajectory.getSegment(t)
        m = tc.numPoints
        n = robot.numLinks()
        sshift = 1 if tc.qstart is not None else 0
        i -= sshift
        qs, inds = [], []
        if i >= 0 and i + 1 <= m:
            qs.append((1 - u) * dq)
            inds += list(range(i * n, i * n + n))
        if i >= 0 and i + 1 < m:
            qs.append(u * dq)
            inds += list(range(i * n + n, i * n + 2 * n))
        if len(qs) == 0:
            return scipy.sparse.csr_matrix((1, m * n))
        return scipy.sparse.csr_matrix((np.hstack(qs), inds, [0, len(inds)]), shape=(1, m * n))
