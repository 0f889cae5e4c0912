"""Constraint plug-in protocol -- same names, arguments and return conventions as the reference's
semiinfinite/constraint.py:4-43 (ConstraintInterface), :46-93 (domains) and :96-171
(SemiInfiniteConstraintInterface).  This protocol is the drop-in boundary: the SIP solver only ever
calls setx / clearx / value / minvalue / df_dx on a constraint (sip.py:141-169, 196-275, 853-899).

Two OPTIONAL batched methods extend the protocol; the solver uses them when a constraint provides
them (the device-backed collision constraints do) and falls back to the per-parameter calls
otherwise, so user-defined constraints written against the reference keep working:

    value_batch(x, ys)  -> sequence of f(x, y) for y in ys
    df_dx_batch(x, ys)  -> sequence of rows (np.ndarray(n) or scipy.sparse 1 x n)
"""
import itertools
import random

import numpy as np


def numeric_gradient(f, x, h=1e-4):
    """Centred differences of a scalar function of a vector (the reference's debugging aid,
    utils.py:3-37, restricted to plain vectors)."""
    x = np.asarray(x, dtype=np.float64)
    g = np.zeros(len(x))
    for i in range(len(x)):
        e = np.zeros(len(x))
        e[i] = h
        g[i] = (f(x + e) - f(x - e)) / (2 * h)
    return g


class ConstraintInterface:
    """f(x) >= 0 with x in R^n; call pattern setx(x); value(x); clearx()."""

    def __init__(self):
        pass

    def dims(self):
        return 0

    def __call__(self, x):
        self.setx(x)
        fx = self.value(x)
        self.clearx()
        return fx

    def setx(self, x):
        pass

    def clearx(self):
        pass

    def value(self, x):
        raise NotImplementedError()

    def df_dx(self, x):
        raise NotImplementedError()

    def df_dx_numeric(self, x, h=1e-4):
        self.clearx()
        res = numeric_gradient(lambda v: self.__call__(v), x, h)
        self.setx(x)
        return res


class DomainInterface:
    def sample(self):
        raise NotImplementedError()

    def extrema(self):
        raise NotImplementedError()


class IntervalDomain(DomainInterface):
    def __init__(self, vmin, vmax):
        self.interval = (vmin, vmax)
        self.vmin, self.vmax = vmin, vmax

    def sample(self):
        return random.uniform(self.vmin, self.vmax)

    def extrema(self):
        return [self.vmin, self.vmax]


class SetDomain(DomainInterface):
    def __init__(self, items):
        self.items = list(items)

    def sample(self):
        return random.choice(self.items)

    def extrema(self):
        return self.items


class CartesianProductDomain(DomainInterface):
    def __init__(self, *domains):
        # the reference's constructor takes one list but its call sites pass positionals
        # (geometryopt.py:674, 953); accept both
        if len(domains) == 1 and isinstance(domains[0], (list, tuple)):
            domains = domains[0]
        self.domains = list(domains)

    def sample(self):
        return np.hstack([d.sample() for d in self.domains])

    def extrema(self):
        return [np.hstack(e) for e in itertools.product(*[d.extrema() for d in self.domains])]


class UnionDomain(DomainInterface):
    def __init__(self, domains):
        self.domains = domains

    def sample(self):
        return random.choice(self.domains).sample()

    def extrema(self):
        return sum([d.extrema() for d in self.domains], [])


class SemiInfiniteConstraintInterface:
    """f(x, y) >= 0 for all y in Y.  Call pattern setx(x); value(x,y) / minvalue(x,bound) /
    df_dx(x,y); clearx().  `minvalue` returns [] (nothing under the bound), one (fmin, param)
    pair, or a list of such pairs (constraint.py:141-151)."""

    def __init__(self):
        pass

    def dims(self):
        return 0

    def __call__(self, x, y):
        self.setx(x)
        res = self.value(x, y)
        self.clearx()
        return res

    def eval_minimum(self, x):
        """min_y f(x,y) without caching (constraint.py:118-127)."""
        self.setx(x)
        newpts = self.minvalue(x)
        self.clearx()
        if len(newpts) > 0 and not hasattr(newpts[0], '__iter__'):
            return newpts[0]
        if len(newpts) > 0:
            return min(dmin for (dmin, param) in newpts)
        raise ValueError("minvalue does not appear to return a minimum value...")

    def setx(self, x):
        pass

    def clearx(self):
        pass

    def value(self, x, y):
        raise NotImplementedError()

    def minvalue(self, x, bound=None):
        raise NotImplementedError()

    def df_dx(self, x, y):
        raise NotImplementedError()

    def df_dy(self, x, y):
        raise NotImplementedError()

    def domain(self):
        raise NotImplementedError()

    def df_dx_numeric(self, x, y, h=1e-4):
        self.clearx()
        res = numeric_gradient(lambda v: self.__call__(v, y), x, h)
        self.setx(x)
        return res

    def df_dy_numeric(self, x, y, h=1e-4):
        return numeric_gradient(lambda v: self.value(x, v), y, h)


class SemiInfiniteConstraintAdaptor(SemiInfiniteConstraintInterface):
    """Presents a plain ConstraintInterface as a semi-infinite one with a single dummy
    parameter, so mixed constraint lists go through one solver (sip.py:811-813)."""

    def __init__(self, c):
        self.c = c

    def setx(self, x):
        self.c.setx(x)

    def clearx(self):
        self.c.clearx()

    def value(self, x, y):
        return self.c.value(x)

    def minvalue(self, x, bound=None):
        v = self.c.value(x)
        if bound is not None and v >= bound:
            return []
        return v, (0,)

    def df_dx(self, x, y):
        return self.c.df_dx(x)

    def domain(self):
        return SetDomain([(0,)])


class MinimumConstraintAdaptor(ConstraintInterface):
    """The plain constraint f(x) = min_y f(x, y) >= 0 of a semi-infinite one (constraint.py:259-277): value = the oracle's
    minimum at x, gradient = df/dx at the minimising parameter -- the classical "minimum distance" formulation the
    reference's optimizeCollFreeMinDist compares the exchange method with.
    The reference's class cannot run as written (`self.f.setx()` without its argument, an undefined `self.minvalue`,
    `df_dy` where the x-gradient is meant); this is what its docstring and call site (gopt:1108-1110) describe."""

    def __init__(self, semi_infinite_constraint):
        self.f = semi_infinite_constraint
        self.minval = None
        self.minparam = None

    def dims(self):
        return self.f.dims()

    def setx(self, x):
        self.f.setx(x)
        res = self.f.minvalue(x)
        if len(res) > 0 and hasattr(res[0], '__iter__'):   # a list of pairs: the deepest one
            res = min(res, key=lambda pair: pair[0])
        self.minval, self.minparam = res

    def clearx(self):
        self.f.clearx()

    def value(self, x):
        return self.minval

    def df_dx(self, x):
        return self.f.df_dx(x, self.minparam)


class MultiSemiInfiniteConstraint(SemiInfiniteConstraintInterface):
    """Several scalar semi-infinite constraints c_0 .. c_{n-1} presented as one:
        min_y f(x, y) = min_i min_{y_i} c_i(x, y_i),      y = (i, y_i, zero padding to a common length)
    (constraint.py:174-256).  all_negative=True: `minvalue` hands back every pair its members report under
    min(bound, 0) -- with members that answer with their whole active set (ObjectCollisionConstraint.all_under) one
    oracle call instantiates all violated points at once; False: the single deepest pair.
    The reference sizes the padding with `sampleDomain(c.domain())`, a name it never imports; here the length of a
    co-parameter comes from `c.domain().sample()`."""

    def __init__(self, constraints, all_negative=True):
        self.constraints = list(constraints)
        self.all_negative = all_negative
        self.num_co_parameters = []
        for c in self.constraints:
            assert c.dims() == 0
            self.num_co_parameters.append(len(np.atleast_1d(c.domain().sample())))
        self.max_co_parameters = max(self.num_co_parameters)

    def _member(self, y):
        index = int(y[0])
        assert 0 <= index < len(self.constraints)
        return index, self.constraints[index], y[1:1 + self.num_co_parameters[index]]

    def _pad(self, i, yi):
        return np.hstack(([i], yi, [0] * (self.max_co_parameters - len(yi))))

    def __call__(self, x, y):
        _, c, yi = self._member(y)
        c.setx(x)
        res = c.value(x, yi)
        c.clearx()
        return res

    def setx(self, x):
        for c in self.constraints:
            c.setx(x)

    def clearx(self):
        for c in self.constraints:
            c.clearx()

    def value(self, x, y):
        _, c, yi = self._member(y)
        return c.value(x, yi)

    @staticmethod
    def _pairs(mv):
        return [mv] if len(mv) > 0 and not hasattr(mv[0], '__iter__') else mv

    def minvalue(self, x, bound=None):
        if self.all_negative:
            res = []
            if bound is not None and bound < 0:              # members are asked for everything below zero  (constraint.py:226-227)
                bound = 0
            for i, c in enumerate(self.constraints):
                for (fi, yi) in self._pairs(c.minvalue(x, bound)):
                    res.append((fi, self._pad(i, yi)))
            return res
        best = None
        for i, c in enumerate(self.constraints):
            for (fi, yi) in self._pairs(c.minvalue(x, bound)):
                if bound is None or fi < bound:
                    bound = fi
                    best = self._pad(i, yi)
        if best is None:
            return []
        return (bound, best)

    def df_dx(self, x, y):
        _, c, yi = self._member(y)
        return c.df_dx(x, yi)

    def df_dy(self, x, y):
        _, c, yi = self._member(y)
        return np.hstack(([0], c.df_dy(x, yi)))

    def domain(self):
        raise NotImplementedError("TODO: Can't inspect domain of a MultiSemiInfiniteConstraint yet")
