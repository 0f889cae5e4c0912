"""Objective-function protocol of the SIP solver -- same method names and meaning as the
reference's semiinfinite/objective.py:5-76: value / minstep / gradient / hessian / integrate.
Objectives are O(n) host arithmetic and stay on the host (SURVEY.md component 7); only the
protocol and the combinators the collision front-ends need are provided."""
import numpy as np
import scipy.sparse


class ObjectiveFunctionInterface:
    def __init__(self):
        pass

    def value(self, x):
        raise NotImplementedError()

    def minstep(self, x):
        """dx minimising f(x+dx) to second order (= -H^-1 g)."""
        raise NotImplementedError()

    def gradient(self, x):
        raise NotImplementedError()

    def hessian(self, x):
        """scalar (multiple of I), 1-D array (diagonal), 2-D array or sparse matrix."""
        return 1

    def integrate(self, x, dx):
        return np.asarray(x) + np.asarray(dx)

    def __add__(self, rhs):
        return _Sum(self, _as_objective(rhs), 1.0)

    def __radd__(self, lhs):
        return _Sum(_as_objective(lhs), self, 1.0)

    def __sub__(self, rhs):
        return _Sum(self, _as_objective(rhs), -1.0)

    def __mul__(self, s):
        if isinstance(s, ObjectiveFunctionInterface):
            raise NotImplementedError("products of objectives are outside the collision hot path")
        return _Scaled(self, float(s))

    __rmul__ = __mul__

    def __truediv__(self, s):
        return _Scaled(self, 1.0 / float(s))


class ConstantObjectiveFunction(ObjectiveFunctionInterface):
    def __init__(self, c):
        self.c = c

    def value(self, x):
        return self.c

    def minstep(self, x):
        return np.zeros(np.shape(x))

    def gradient(self, x):
        return np.zeros(np.shape(x))

    def hessian(self, x):
        return 0


class LinearObjectiveFunction(ObjectiveFunctionInterface):
    """f(x) = c^T x; the step length is set by |c| (objective.py:93-110)."""

    def __init__(self, c):
        self.c = np.asarray(c, dtype=np.float64)
        self.hess = scipy.sparse.csc_matrix(np.outer(self.c, self.c) / float(self.c @ self.c))

    def value(self, x):
        return float(self.c @ np.asarray(x))

    def minstep(self, x):
        return -self.c

    def gradient(self, x):
        return self.c

    def hessian(self, x):
        return self.hess


class QuadraticObjectiveFunction(ObjectiveFunctionInterface):
    """f(x) = weight * |x - xdes|^2."""

    def __init__(self, xdes, weight=1.0):
        self.xdes = np.asarray(xdes, dtype=np.float64)
        self.weight = weight

    def value(self, x):
        d = np.asarray(x) - self.xdes
        return self.weight * float(d @ d)

    def minstep(self, x):
        return self.xdes - np.asarray(x)

    def gradient(self, x):
        return 2 * self.weight * (np.asarray(x) - self.xdes)

    def hessian(self, x):
        return self.weight


def _as_objective(v):
    return v if isinstance(v, ObjectiveFunctionInterface) else ConstantObjectiveFunction(v)


class _Sum(ObjectiveFunctionInterface):
    def __init__(self, g, h, sign):
        self.g, self.h, self.sign = g, h, sign

    def value(self, x):
        return self.g.value(x) + self.sign * self.h.value(x)

    def gradient(self, x):
        return self.g.gradient(x) + self.sign * self.h.gradient(x)

    def hessian(self, x):
        return self.g.hessian(x) + self.sign * self.h.hessian(x)

    def integrate(self, x, dx):
        return self.g.integrate(x, dx)


class _Scaled(ObjectiveFunctionInterface):
    def __init__(self, g, s):
        self.g, self.s = g, s

    def value(self, x):
        return self.g.value(x) * self.s

    def gradient(self, x):
        return self.g.gradient(x) * self.s

    def hessian(self, x):
        return self.g.hessian(x) * self.s

    def minstep(self, x):
        return self.g.minstep(x)

    def integrate(self, x, dx):
        return self.g.integrate(x, dx)
