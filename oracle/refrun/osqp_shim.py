"""Stand-in `osqp` module (OSQP is absent from this image) on the FROZEN solver oracle/frozen/qp.py -- a
restatement of OSQP's published algorithm that no longer follows the product's solvers (round-1 verdict:
"goldens move with the product").  The product's QP paths (host numpy, native hostqp.c, device k_sip_qp)
are tested against the goldens this produces, at a stated tolerance.
Interface used by the reference: sip.py:302, 327-331, 355-357."""
import sys
import types

import numpy as np

from oracle.frozen import qp as _qp


class _Info:
    def __init__(self, status, iters, pri, dua):
        self.status, self.iter, self.pri_res, self.dua_res = status, iters, pri, dua


class _Results:
    def __init__(self, x, y, info):
        self.x, self.y, self.info = x, y, info


class OSQP:
    def setup(self, P=None, q=None, A=None, l=None, u=None, **settings):
        self.P, self.q, self.A, self.l, self.u = P, np.asarray(q, dtype=np.float64), A, np.asarray(l, dtype=np.float64), np.asarray(u, dtype=np.float64)
        self.settings = settings

    def solve(self):
        kw = {}
        if "eps_abs" in self.settings:
            kw["eps_abs"] = self.settings["eps_abs"]
        P, A = self.P, self.A
        if P.shape[0] <= _qp.DENSE_MAX_VARS:          # dense KKT factorisation for the small problems, sparse LU above
            P, A = P.toarray(), A.toarray()
        r = _qp.solve_qp(P, self.q, A, self.l, self.u, **kw)
        n = len(self.q)
        if r.x is not None:          # solved, or the last iterate when the iteration limit was hit (as OSQP does)
            x = np.asarray(r.x, dtype=np.float64)
        else:
            x = np.array([None] * n, dtype=object)      # what OSQP hands back for an infeasible problem
        return _Results(x, r.y, _Info(r.status, r.iters, r.pri_res, r.dua_res))


def install():
    if "osqp" in sys.modules and getattr(sys.modules["osqp"], "_sio_shim", False):
        return
    m = types.ModuleType("osqp")
    m._sio_shim = True
    m.OSQP = OSQP
    sys.modules["osqp"] = m
