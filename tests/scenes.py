"""Scene builders shared by tools/make_golden.py (which drives the REFERENCE's own classes through
oracle/refrun) and the golden tests (which drive the CPU restatement and the device-backed classes).

`kind` selects the class family:  'reference' = the reference package loaded by oracle/refrun,
'restated' = oracle/ref_geometryopt.py on the CPU oracle, 'device' = semiinfiniteoptimization_b200."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CHAIR_IN_PATH = (-0.68, -0.38, 0.5)       # puts the chair across the initial TX90 path (see test_gpu_classes.py)


def family(kind):
    if kind == 'reference':
        from oracle.refrun import load
        return load.load()['geometryopt']
    if kind == 'restated':
        from oracle import ref_geometryopt
        return ref_geometryopt
    from semiinfiniteoptimization_b200 import geometryopt
    return geometryopt


def _robot_modules(kind):
    """The reference's code runs on the FROZEN kinematics (oracle/frozen), the restatement and the product on the
    product's host substrate: the goldens pin the latter against the former."""
    if kind == 'reference':
        from oracle.frozen.robot import WorldModel
        from oracle.frozen.trajectory import RobotTrajectory
    else:
        from semiinfiniteoptimization_b200._host.robot import WorldModel
        from semiinfiniteoptimization_b200._host.trajectory import RobotTrajectory
    return WorldModel, RobotTrajectory


def _names(kind):
    if kind == 'restated':
        return dict(PDG='RefPDG', Obj='RefObjectCollisionConstraint', Kin='RefRobotKinematicsCache',
                    Link='RefRobotLinkCollisionConstraint', TC='RefRobotTrajectoryCache', Traj='RefRobotTrajectoryCollisionConstraint')
    return dict(PDG='PenetrationDepthGeometry', Obj='ObjectCollisionConstraint', Kin='RobotKinematicsCache',
                Link='RobotLinkCollisionConstraint', TC='RobotTrajectoryCache', Traj='RobotTrajectoryCollisionConstraint')


def c1_poses(n, seed=0):
    from oracle.frozen import so3          # inputs are generated with the frozen helpers: goldens do not move with the product
    rng = np.random.default_rng(seed)
    return [(so3.from_rotation_vector(rng.uniform(-0.3, 0.3, 3)),
             [float(rng.uniform(0.4, 1.0)), float(rng.uniform(-0.3, 0.5)), float(rng.uniform(-0.3, 0.5))]) for _ in range(n)]


def c1_scene(kind):
    """C1: cube SDF@0.05 (moving) vs m797 cloud@0.02 at (I,[1.2,0,0]) (geomopt.py:14-15, 43)."""
    from oracle.frozen import so3          # inputs are generated with the frozen helpers: goldens do not move with the product
    from semiinfiniteoptimization_b200._host.geometry import fixture_geometry
    mod, nm = family(kind), _names(kind)
    PDG = getattr(mod, nm['PDG'])
    g1, g2 = PDG(fixture_geometry("cube"), 0.05, 0.02), PDG(fixture_geometry("m797"), 0.05, 0.02)
    g2.setTransform((so3.identity(), [1.2, 0, 0]))
    return mod, g1, g2, getattr(mod, nm['Obj'])(g1, g2)


TREE_INTERIOR = (0.41, 0.06, 0.42)       # deepest interior sample of the m1062 SDF at gridres 0.01 (value -0.019: thin branches)


def c2_scene(kind, gridres=0.01):
    """C2 as shipped: m1062 SDF@gridres (moving) vs the 397-point bunny cloud; the cloud's centroid is placed at the
    deepest interior point of the tree so the poses below start in contact."""
    from oracle.frozen import so3          # inputs are generated with the frozen helpers: goldens do not move with the product
    from semiinfiniteoptimization_b200._host.geometry import fixture_geometry
    mod, nm = family(kind), _names(kind)
    PDG = getattr(mod, nm['PDG'])
    tree, bunny = fixture_geometry("m1062"), fixture_geometry("bunny")
    b = bunny.getPointCloud().points
    g1, g2 = PDG(tree, gridres, None), PDG(bunny, None, 0.005)
    g2.setTransform((so3.identity(), [float(v) for v in np.asarray(TREE_INTERIOR) - b.mean(0)]))
    return mod, g1, g2, getattr(mod, nm['Obj'])(g1, g2)


def c2_poses(n, seed=0):
    """Poses of the tree: rotation w ~ U(-0.3,0.3)^3 about TREE_INTERIOR, then a shift t ~ U(-0.03,0.03)^3."""
    from oracle.frozen import so3          # inputs are generated with the frozen helpers: goldens do not move with the product
    rng = np.random.default_rng(seed)
    c = np.asarray(TREE_INTERIOR)
    out = []
    for _ in range(n):
        R = so3.from_rotation_vector(rng.uniform(-0.3, 0.3, 3))
        t = c - so3.matrix(R) @ c + rng.uniform(-0.03, 0.03, 3)
        out.append((R, [float(v) for v in t]))
    return out


def c3_scene(kind):
    """C3: tx90_geom_test3.xml, declared synthetic link boxes, tree = m1062 x2 rotated; gridres = pcres = 0.02."""
    WorldModel, _ = _robot_modules(kind)
    mod, nm = family(kind), _names(kind)
    w = WorldModel()
    w.readFile("tx90_geom_test3.xml")
    robot, tree = w.robot(0), w.rigidObject(0)
    if kind == 'restated':
        cache = mod.RefRobotKinematicsCache(robot, 0.02, 0.02)
        env = mod.RefPDG(tree.geometry(), 0.02, 0.02)
        cons = [mod.RefRobotLinkCollisionConstraint(robot.link(i), env, cache) for i in range(robot.numLinks())
                if not robot.link(i).geometry().empty()]
    else:
        cons, _ = mod.makeCollisionConstraints(robot, [tree], 0.02, 0.02)
    return mod, robot, cons


def c3_configs(robot, n, seed=3):
    rng = np.random.default_rng(seed)
    qmin, qmax = robot.getJointLimits()
    return [[float(lo + rng.uniform(0.2, 0.8) * (hi - lo)) for lo, hi in zip(qmin, qmax)] for _ in range(n)]


def c4_scene(kind, chair_pos=CHAIR_IN_PATH):
    """C4: tx90_geom_test2.xml, robottrajopt_initial.path remeshed x6 -> 67 milestones (65 free, x in R^455), chair
    cloud only (gridres=None), link grids/clouds at 0.02 (robottrajopt.py:13-17, 70-71, 101-131)."""
    WorldModel, RobotTrajectory = _robot_modules(kind)
    mod, nm = family(kind), _names(kind)
    p0 = json.load(open(os.path.join(ROOT, "semiinfiniteoptimization_b200", "data", "paths.json")))["resources/robottrajopt_initial.path"]
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
    kin = getattr(mod, nm['Kin'])(robot, 0.02, 0.02)
    tc = getattr(mod, nm['TC'])(kin, 10 * 6 + 6 - 1)
    env = getattr(mod, nm['PDG'])(chair.geometry(), None, 0.02)
    con = getattr(mod, nm['Traj'])([env], tc)
    tc.qstart, tc.qend = traj.milestones[0][:], traj.milestones[-1][:]
    tc.tstart, tc.tend = traj.times[0], traj.times[-1]
    x = [float(v) for q in traj.milestones[1:-1] for v in q]
    return mod, robot, tc, con, traj, x


def c5_scene(kind, context=None):
    """C5: cube SDF@0.05 (moving) vs the declared synthetic 262,144-point scene cloud (scene2_1.pcd is missing from the
    reference checkout), scene at the identity -- the geometry of BASELINE.json configs[4]."""
    from semiinfiniteoptimization_b200 import workloads as W
    from semiinfiniteoptimization_b200._host.geometry import Geometry3D, PointCloud
    mod, nm = family(kind), _names(kind)
    PDG = getattr(mod, nm['PDG'])
    kw = {"context": context} if (kind == 'device' and context is not None) else {}
    cube = PDG(W.c5_cube(), 0.05, None, **kw)
    scene = PDG(Geometry3D(PointCloud(W.c5_scene_cloud())), None, 0.02, **kw)
    return mod, cube, scene, getattr(mod, nm['Obj'])(cube, scene)


def c5_sample_indices(n=64, total=65536):
    """Every (total / n)-th of the 65,536 C5 problems (seed 3)."""
    return list(range(0, total, total // n))[:n]


def load_golden(name):
    return json.load(open(os.path.join(GOLDEN, name)))
