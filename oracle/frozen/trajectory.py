"""FROZEN COPY (oracle/frozen, round 2) of semiinfiniteoptimization_b200/_host/trajectory.py as of commit 92f6a08 -- TEST INFRASTRUCTURE.

The golden generator (tools/make_golden.py -> oracle/refrun) answers the reference's Klampt / OSQP calls with THIS file,
not with the product's module, so a change to the product can no longer rewrite tests/golden/.  Do not edit it to follow
the product: the product is tested AGAINST it (tests/test_frozen_cpu.py) at a stated tolerance.

Original module docstring follows.

Piecewise-linear trajectories with the interface the reference uses from
`klampt.model.trajectory` (Trajectory / RobotTrajectory: eval, getSegment, interpolate, remesh,
duration -- call sites geometryopt.py:341, 364, 589, 635, 712, 809, 916, 926; robottrajopt.py:101-118).
Host-side helper; time-sample interpolation for the device sweeps is done in bulk with numpy by
the trajectory cache."""
import bisect


class Trajectory:
    def __init__(self, times=None, milestones=None):
        self.times = list(times) if times is not None else []
        self.milestones = [list(m) for m in milestones] if milestones is not None else []

    def startTime(self):
        return self.times[0]

    def endTime(self):
        return self.times[-1]

    def duration(self):
        return self.times[-1] - self.times[0]

    def getSegment(self, t):
        """(i, u): t lies on segment i (between milestones i and i+1) at fraction u.
        Before the start: (-1, 0); at/after the end: (len-1, 0)."""
        if len(self.times) == 0:
            raise ValueError("empty trajectory")
        if len(self.times) == 1:
            return (-1, 0)
        if t >= self.times[-1]:
            return (len(self.milestones) - 1, 0)
        if t <= self.times[0]:
            return (0, 0)
        i = bisect.bisect_right(self.times, t)
        p = i - 1
        u = (t - self.times[p]) / (self.times[i] - self.times[p])
        return (p, u)

    def interpolate(self, a, b, u, dt=None):
        return [x + u * (y - x) for x, y in zip(a, b)]

    def eval(self, t):
        i, u = self.getSegment(t)
        if i < 0:
            return list(self.milestones[0])
        if i + 1 >= len(self.milestones):
            return list(self.milestones[-1])
        return self.interpolate(self.milestones[i], self.milestones[i + 1], u, self.times[i + 1] - self.times[i])

    def remesh(self, newtimes):
        """Returns (trajectory with the union of old and new times, indices of newtimes in it)."""
        alltimes = sorted(set(self.times) | set(newtimes))
        ms = [self.eval(t) for t in alltimes]
        idx = [alltimes.index(t) for t in newtimes]
        return self.__class__.__new__(self.__class__)._init_like(self, alltimes, ms), idx

    def _init_like(self, other, times, milestones):
        Trajectory.__init__(self, times, milestones)
        return self

    def length(self, metric=None):
        """Klampt's Trajectory.length (added for the reference's planopt.py, which the round-1 copy did not run)."""
        if metric is None:
            metric = lambda a, b: sum((x - y) ** 2 for x, y in zip(a, b)) ** 0.5
        return sum(metric(a, b) for a, b in zip(self.milestones[:-1], self.milestones[1:]))


class RobotTrajectory(Trajectory):
    def __init__(self, robot, times=None, milestones=None):
        Trajectory.__init__(self, times, milestones)
        self.robot = robot

    def interpolate(self, a, b, u, dt=None):
        return self.robot.interpolate(a, b, u)

    def length(self, metric=None):
        return Trajectory.length(self, metric or self.robot.distance)

    def _init_like(self, other, times, milestones):
        RobotTrajectory.__init__(self, other.robot, times, milestones)
        return self


def load_path(fn):
    """Klampt linear path format: one `time<TAB>n q1 ... qn` line per milestone."""
    times, ms = [], []
    for line in open(fn).read().splitlines():
        t = line.split()
        if not t:
            continue
        n = int(t[1])
        times.append(float(t[0]))
        ms.append([float(v) for v in t[2:2 + n]])
    return Trajectory(times, ms)


def save_path(traj, fn):
    with open(fn, "w") as f:
        for t, m in zip(traj.times, traj.milestones):
            f.write("%f\t%d %s\n" % (t, len(m), " ".join(repr(float(v)) for v in m)))
