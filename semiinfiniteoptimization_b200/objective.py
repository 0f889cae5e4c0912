"""Objective-function protocol of the SIP solver -- same method names and meaning as the
reference's semiinfinite/objective.py:5-76: value / minstep / gradient / hessian / integrate.
Objectives are O(n) host arithmetic and stay on the host (SURVEY.md component 7).  The arithmetic combinators carry the
reference's class names (objective.py:112-199) and semantics, checked against its own classes (tests/test_host_cpu.py)."""
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
        return AddObjectiveFunction(self, _as_objective(rhs))

    def __radd__(self, lhs):
        return AddObjectiveFunction(_as_objective(lhs), self)

    def __sub__(self, rhs):
        return SubObjectiveFunction(self, _as_objective(rhs))

    def __rsub__(self, lhs):
        return SubObjectiveFunction(_as_objective(lhs), self)

    def __mul__(self, rhs):
        if isinstance(rhs, ObjectiveFunctionInterface):
            return MulObjectiveFunction(self, rhs)
        return MulConstObjectiveFunction(self, rhs)        # the reference forgets the factor here (objective.py:58): TypeError

    def __rmul__(self, lhs):
        return MulConstObjectiveFunction(self, lhs)

    def __truediv__(self, rhs):
        if isinstance(rhs, ObjectiveFunctionInterface):
            return DivObjectiveFunction(self, rhs)
        return MulConstObjectiveFunction(self, 1.0 / rhs)

    __div__ = __truediv__                                  # the reference is Python 2

    def __pow__(self, rhs):
        if isinstance(rhs, ObjectiveFunctionInterface):
            raise NotImplementedError("Can't raise a function to a function power")
        return PowConstObjectiveFunction(self, rhs)


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


class AddObjectiveFunction(ObjectiveFunctionInterface):
    """f(x) = g(x) + h(x); integrates like g (objective.py:112-123)."""

    def __init__(self, g, h):
        self.g, self.h = g, h

    def value(self, x):
        return self.g.value(x) + self.h.value(x)

    def gradient(self, x):
        return self.g.gradient(x) + self.h.gradient(x)

    def hessian(self, x):
        return self.g.hessian(x) + self.h.hessian(x)

    def integrate(self, x, dx):
        return self.g.integrate(x, dx)


class SubObjectiveFunction(AddObjectiveFunction):
    """f(x) = g(x) - h(x) (objective.py:126-137)."""

    def value(self, x):
        return self.g.value(x) - self.h.value(x)

    def gradient(self, x):
        return self.g.gradient(x) - self.h.gradient(x)

    def hessian(self, x):
        return self.g.hessian(x) - self.h.hessian(x)


class MulConstObjectiveFunction(ObjectiveFunctionInterface):
    """f(x) = g(x) * h with h a number; the minimising step is g's (objective.py:140-153)."""

    def __init__(self, g, h):
        self.g, self.h = g, h

    def value(self, x):
        return self.g.value(x) * self.h

    def gradient(self, x):
        return self.g.gradient(x) * self.h

    def hessian(self, x):
        return self.g.hessian(x) * self.h

    def minstep(self, x):
        return self.g.minstep(x)

    def integrate(self, x, dx):
        return self.g.integrate(x, dx)


class MulObjectiveFunction(ObjectiveFunctionInterface):
    """f(x) = g(x) * h(x), product rule to second order (objective.py:156-167)."""

    def __init__(self, g, h):
        self.g, self.h = g, h

    def value(self, x):
        return self.g.value(x) * self.h.value(x)

    def gradient(self, x):
        return self.g.gradient(x) * self.h.value(x) + self.g.value(x) * self.h.gradient(x)

    def hessian(self, x):
        gg, hg = self.g.gradient(x), self.h.gradient(x)
        return self.g.hessian(x) * self.h.value(x) + np.outer(gg, hg) + self.g.value(x) * self.h.hessian(x) + np.outer(hg, gg)

    def integrate(self, x, dx):
        return self.g.integrate(x, dx)


class DivObjectiveFunction(ObjectiveFunctionInterface):
    """f(x) = g(x) / h(x) (objective.py:170-182)."""

    def __init__(self, g, h):
        self.g, self.h = g, h

    def value(self, x):
        return self.g.value(x) / self.h.value(x)

    def gradient(self, x):
        hval = self.h.value(x)
        return (self.g.gradient(x) * hval + self.g.value(x) * self.h.gradient(x)) / (hval * hval)   # quirk: '+' as in the reference (the quotient rule has '-')

    def hessian(self, x):
        raise NotImplementedError("Can't do hessian of function/function division yet")

    def integrate(self, x, dx):
        return self.g.integrate(x, dx)


class PowConstObjectiveFunction(ObjectiveFunctionInterface):
    """f(x) = g(x) ** h with h a number (objective.py:185-199)."""

    def __init__(self, g, h):
        self.g, self.h = g, h

    def value(self, x):
        return pow(self.g.value(x), self.h)

    def gradient(self, x):
        return self.h * pow(self.g.value(x), self.h - 1) * self.g.gradient(x)

    def hessian(self, x):
        gval, ggrad = self.g.value(x), self.g.gradient(x)
        return self.h * (self.h - 1) * pow(gval, self.h - 2) * np.outer(ggrad, ggrad) + self.h * pow(gval, self.h - 1) * self.g.hessian(x)

    def integrate(self, x, dx):
        return self.g.integrate(x, dx)
