"""Seeded synthetic inputs for BASELINE.json's configurations -- shared by bench.py,
the tools and the tests so that every number is quoted on the same arrays.

Declared stand-ins: `data/scene2_1.pcd` is missing from the reference checkout, so C5's scene is the
seeded cloud below; the C2 throughput cloud is the seeded 2^20-point upsample of bunny.pcd.
"""
import numpy as np

from ._host import se3, so3
from ._host.geometry import fixture_geometry


def c2_throughput_cloud(n=1 << 20, seed=1):
    """Point j = 4 (b[j mod 397] - centroid(b)) + centre(bbox(m1062)) + N(0, 0.005^2 I)."""
    tree = fixture_geometry("m1062")
    b = fixture_geometry("bunny").getPointCloud().points
    rng = np.random.default_rng(seed)
    lo, hi = tree.data.bbox()
    centre = 0.5 * (lo + hi)
    j = np.arange(n) % len(b)
    cloud = 4.0 * (b[j] - b.mean(0)) + centre + rng.normal(0.0, 0.005, (n, 3))
    return cloud.astype(np.float32), centre


def c2_throughput_poses(n, centre, seed):
    """Row-major 3x4 T_rel = pose^-1 of n poses of the moving tree: w ~ U(-0.3,0.3)^3 about the body centre,
    t ~ U(-0.1,0.1)^3 (the cloud frame is the world)."""
    prng = np.random.default_rng(seed)
    T_rel = np.empty((n, 12))
    for i in range(n):
        w = prng.uniform(-0.3, 0.3, 3)
        t = prng.uniform(-0.1, 0.1, 3)
        R = so3.from_rotation_vector(w)
        tt = [float(v) for v in (centre - so3.matrix(R) @ centre + t)]
        T_rel[i] = se3.to_rowmajor12(se3.inv((R, tt)))
    return T_rel


def c5_scene_cloud(n=262144, seed=2):
    """Synthetic stand-in for scene2_1.pcd: a table plane z = 0 over [-1,1]^2 (half of the points) and the surfaces
    of 12 random boxes standing on it inside [-1,1]^2 x [0,1] (the other half, area-weighted)."""
    rng = np.random.default_rng(seed)
    n_plane = n // 2
    plane = np.column_stack([rng.uniform(-1, 1, n_plane), rng.uniform(-1, 1, n_plane), np.zeros(n_plane)])
    lo = np.column_stack([rng.uniform(-0.9, 0.5, 12), rng.uniform(-0.9, 0.5, 12), np.zeros(12)])
    size = np.column_stack([rng.uniform(0.1, 0.4, 12), rng.uniform(0.1, 0.4, 12), rng.uniform(0.1, 0.9, 12)])
    hi = np.minimum(lo + size, [1.0, 1.0, 1.0])
    size = hi - lo
    # faces: +-x, +-y, +z (the bottom sits on the table)
    areas = np.column_stack([size[:, 1] * size[:, 2]] * 2 + [size[:, 0] * size[:, 2]] * 2 + [size[:, 0] * size[:, 1]])
    p = (areas / areas.sum()).reshape(-1)
    n_box = n - n_plane
    pick = rng.choice(len(p), size=n_box, p=p)
    box, face = pick // 5, pick % 5
    u, v = rng.uniform(0, 1, n_box), rng.uniform(0, 1, n_box)
    pts = np.empty((n_box, 3))
    fx = face < 2
    pts[fx, 0] = np.where(face[fx] == 0, lo[box[fx], 0], hi[box[fx], 0])
    pts[fx, 1] = lo[box[fx], 1] + u[fx] * size[box[fx], 1]
    pts[fx, 2] = lo[box[fx], 2] + v[fx] * size[box[fx], 2]
    fy = (face >= 2) & (face < 4)
    pts[fy, 1] = np.where(face[fy] == 2, lo[box[fy], 1], hi[box[fy], 1])
    pts[fy, 0] = lo[box[fy], 0] + u[fy] * size[box[fy], 0]
    pts[fy, 2] = lo[box[fy], 2] + v[fy] * size[box[fy], 2]
    fz = face == 4
    pts[fz, 2] = hi[box[fz], 2]
    pts[fz, 0] = lo[box[fz], 0] + u[fz] * size[box[fz], 0]
    pts[fz, 1] = lo[box[fz], 1] + v[fz] * size[box[fz], 1]
    return np.vstack([plane, pts]).astype(np.float32)


def c5_poses(n, seed=3):
    """Initial poses of the moving cube: uniform in the scene volume [-1,1]^2 x [0,1], rotations
    w ~ U(-pi/2,pi/2)^3 ("uniform in the scene volume with random rotations").  Klampt se3 tuples."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        R = so3.from_rotation_vector(rng.uniform(-np.pi / 2, np.pi / 2, 3))
        t = [float(rng.uniform(-1, 1)), float(rng.uniform(-1, 1)), float(rng.uniform(0.0, 1.0))]
        out.append((R, t))
    return out


def c5_cube():
    """The moving body of C5: cube.off (unit cube), converted to an SDF at gridres 0.05 by the caller."""
    return fixture_geometry("cube")


# ---- single-problem configurations (C1, C3, C4) for bench.py's latency legs ------------------------------------------
# `family` is a module with the reference's class names (this package's geometryopt) or the CPU restatement's
# (`names` maps the roles); the builders only assemble inputs, they do not import either.
DEVICE_NAMES = dict(PDG='PenetrationDepthGeometry', Obj='ObjectCollisionConstraint', Kin='RobotKinematicsCache',
                    Link='RobotLinkCollisionConstraint', TC='RobotTrajectoryCache', Traj='RobotTrajectoryCollisionConstraint')
CHAIR_IN_PATH = (-0.68, -0.38, 0.5)       # puts the chair across the initial TX90 path (the shipped placement never touches it)


def c1_poses(n, seed=0):
    """C1's headless poses: w ~ U(-0.3,0.3)^3, t in a box overlapping the chair."""
    rng = np.random.default_rng(seed)
    return [(so3.from_rotation_vector(rng.uniform(-0.3, 0.3, 3)),
             [float(rng.uniform(0.4, 1.0)), float(rng.uniform(-0.3, 0.5)), float(rng.uniform(-0.3, 0.5))]) for _ in range(n)]


def c1_scene(family, names=DEVICE_NAMES):
    """cube SDF@0.05 (moving) vs m797 cloud@0.02 at (I,[1.2,0,0]) (geomopt.py:14-15, 43)."""
    PDG = getattr(family, names['PDG'])
    g1, g2 = PDG(fixture_geometry("cube"), 0.05, 0.02), PDG(fixture_geometry("m797"), 0.05, 0.02)
    g2.setTransform((so3.identity(), [1.2, 0, 0]))
    return g1, g2, getattr(family, names['Obj'])(g1, g2)


def c3_scene(family, names=DEVICE_NAMES):
    """tx90_geom_test3.xml, declared synthetic link boxes, tree = m1062 x2 rotated; gridres = pcres = 0.02."""
    from ._host.robot import WorldModel
    w = WorldModel()
    w.readFile("tx90_geom_test3.xml")
    robot, tree = w.robot(0), w.rigidObject(0)
    cache = getattr(family, names['Kin'])(robot, 0.02, 0.02)
    env = getattr(family, names['PDG'])(tree.geometry(), 0.02, 0.02)
    cons = [getattr(family, names['Link'])(robot.link(i), env, cache) for i in range(robot.numLinks())
            if not robot.link(i).geometry().empty()]
    return robot, cons


def c3_configs(robot, n, seed=3):
    rng = np.random.default_rng(seed)
    qmin, qmax = robot.getJointLimits()
    return [[float(lo + rng.uniform(0.2, 0.8) * (hi - lo)) for lo, hi in zip(qmin, qmax)] for _ in range(n)]


def c4_scene(family, names=DEVICE_NAMES, chair_pos=CHAIR_IN_PATH):
    """tx90_geom_test2.xml, robottrajopt_initial.path remeshed x6 -> 67 milestones (65 free, x in R^455), chair cloud
    only (gridres=None), link grids/clouds at 0.02 (robottrajopt.py:13-17, 70-71, 101-131)."""
    import json
    import os
    from ._host.robot import WorldModel
    from ._host.trajectory import RobotTrajectory
    here = os.path.dirname(os.path.abspath(__file__))
    p0 = json.load(open(os.path.join(here, "data", "paths.json")))["resources/robottrajopt_initial.path"]
    w = WorldModel()
    w.readFile("tx90_geom_test2.xml")
    robot, chair = w.robot(0), w.rigidObject(0)
    if chair_pos is not None:
        chair.setTransform(chair.getTransform()[0], list(chair_pos))
    traj = RobotTrajectory(robot, p0["times"], p0["milestones"])
    times = []
    for i in range(len(traj.times) - 1):
        for k in range(6):
            times.append(traj.times[i] + k / 6.0 * (traj.times[i + 1] - traj.times[i]))
    times.append(traj.times[-1])
    traj = traj.remesh(times)[0]
    kin = getattr(family, names['Kin'])(robot, 0.02, 0.02)
    tc = getattr(family, names['TC'])(kin, 10 * 6 + 6 - 1)
    env = getattr(family, names['PDG'])(chair.geometry(), None, 0.02)
    con = getattr(family, names['Traj'])([env], tc)
    tc.qstart, tc.qend = traj.milestones[0][:], traj.milestones[-1][:]
    tc.tstart, tc.tend = traj.times[0], traj.times[-1]
    x = [float(v) for q in traj.milestones[1:-1] for v in q]
    return robot, tc, con, traj, x
