"""Exchange-method semi-infinite programming solver: the CALLER of the collision oracle.

Mirrors the public surface of the reference's semiinfinite/sip.py -- `SemiInfiniteOptimizationSettings`
(sip.py:46-75), `SemiInfiniteOptimizationResult` (:80-111), `score` (:134-172),
`ConstraintGenerationData` (:183-378) and `optimizeSemiInfinite` (:795-1241) -- for the DEFAULT flag
path only (trust region, 'most violating all' oracle, 'minimum' scoring metric, constraint
dropping, OSQP-style QP).  Solver logic stays on the host (BASELINE.json north_star); what changes
is how it talks to constraints: wherever the reference loops over instantiated parameters issuing
one Klampt call each (sip.py:152, 201, 893-897), this solver hands the whole parameter list to the
constraint's optional `value_batch` / `df_dx_batch` methods, which the device-backed collision
constraints implement with one kernel launch.  Constraints without those methods are driven through
the per-parameter protocol exactly as before.

Reference quirks that change results are reproduced and marked `# quirk`.
"""
import math
import time

import numpy as np
import scipy.sparse

from .constraint import *          # noqa: F401,F403  (the reference re-exports these, sip.py:13)
from .objective import *           # noqa: F401,F403
from ._host import qp as _qp

# flag surface of the reference (sip.py:22-44); only the default values are implemented
DEBUG_GRADIENTS = False
DEBUG_TRAJECTORY_INITIALIZATION = False
STEP_SIZE_METHOD = 'trust region'
DETECT_NEW_COLLISIONS_IN_STEP = True
ORACLE_METHOD = 'most violating all'
REMOVE_CONSTRAINTS = True
QP_SOLVE_METHOD = 'osqp'
SCORING_METRIC = 'minimum'
VERBOSE_DROPS = False              # the reference prints every dropped parameter unconditionally


class SemiInfiniteOptimizationSettings:
    """Same attributes and defaults as sip.py:64-75."""

    def __init__(self):
        self.max_iters = 50
        self.xepsilon = 1e-4
        self.constraint_inflation = 1e-3
        self.constraint_drop_value = 0.1
        self.minimum_constraint_value = -float('inf')
        self.initial_objective_score_weight = 1e-1
        self.initial_penalty_parameter = 1e-1
        self.minimum_objective_score_weight = 1e-5
        self.parameter_exclusion_distance = 1e-3
        self.qp_solver_params = {'regularizationFactor': 1e-3}


class SemiInfiniteOptimizationResult:
    """Same members as sip.py:98-111."""

    def __init__(self):
        self.x0 = None
        self.fx0 = None
        self.gx0 = None
        self.x = None
        self.fx = None
        self.gx = None
        self.instantiated_params = []
        self.trace = []
        self.trace_times = []
        self.num_iterations = 0
        self.status = 'not converged'
        self.time = 0
        self.time_cons = 0
        self.time_qp = 0


def scoring_metric(v):
    """'minimum' metric: |most negative entry| (sip.py:123-126)."""
    r = min(v)
    return 0 if r >= 0 else abs(r)


def scoring_metric2(v):
    v = [d for d in v if d < 0]
    return 0 if len(v) == 0 else scoring_metric(v)


# ---- batched access to the constraint protocol ------------------------------------------------
def constraint_values(c, x, ys):
    """[c.value(x,y) for y in ys] -- through value_batch (one device launch) when offered."""
    if len(ys) == 0:
        return []
    vb = getattr(c, 'value_batch', None)
    if vb is not None:
        return [float(v) for v in vb(x, ys)]
    return [c.value(x, y) for y in ys]


def constraint_rows(c, x, ys):
    if len(ys) == 0:
        return []
    rb = getattr(c, 'df_dx_batch', None)
    if rb is not None:
        return list(rb(x, ys))
    return [c.df_dx(x, y) for y in ys]


def _as_pairs(newpts):
    """minvalue may return [], one (f, param) pair, or a list of pairs (constraint.py:141-151)."""
    if len(newpts) > 0 and not hasattr(newpts[0], '__iter__'):
        return [newpts]
    return newpts


def score(objective, constraints, x, constraint_generation_data, objScoreWeight, discoverNew=False):
    """Merit value at x: sum over constraints of the scoring metric of the negative parameter
    values, plus objScoreWeight * f(x).  With discoverNew the oracle is also asked for a deeper
    parameter, which is instantiated as a side effect even if the step is later rejected  # quirk
    Returns (score, fx, gx).  (sip.py:134-172)"""
    for c in constraints:
        c.setx(x)
    cdata = constraint_generation_data
    fx = objective.value(x)
    gx = np.zeros(len(constraints))
    gscore = 0.0
    for i, c in enumerate(constraints):
        if len(cdata.instantiated_params[i]) == 0:
            gx[i] = float('inf')
            gvals = []
        else:
            gvals = constraint_values(c, x, cdata.instantiated_params[i])
            gx[i] = min(gvals)
            gvals = [g for g in gvals if g < 0]
        if discoverNew:
            for (dmin, newparam) in _as_pairs(c.minvalue(x, gx[i])):
                if dmin < gx[i]:
                    gx[i] = dmin
                    if cdata.instantiate(i, newparam):
                        if dmin < 0:
                            gvals.append(dmin)
        if len(gvals) > 0:
            gscore += scoring_metric(gvals)
    for c in constraints:
        c.clearx()
    return (gscore + objScoreWeight * fx, fx, gx)


def _score(objectiveVal, constraintVals, objScoreWeight):
    return sum(scoring_metric2(cs) for cs in constraintVals) + objScoreWeight * objectiveVal


class ConstraintGenerationData:
    """Index-set bookkeeping + the QP step (sip.py:183-378)."""

    def __init__(self, constraints):
        self.constraints = constraints
        self.instantiated_params = [[] for c in constraints]
        self.visited_points = [set() for c in constraints]
        self.exclusion_distance = 1e-3
        self.constraint_inflation = 1e-5
        self.constraint_drop_value = float('inf')
        self.qp_time = 0.0

    def _key(self, y):
        # Python int() truncates toward zero                                            # quirk
        return tuple(int(v / self.exclusion_distance) for v in y)

    def instantiate(self, cindex, y):
        """Adds parameter y to constraint cindex unless one within the exclusion grid exists."""
        k = self._key(y)
        if k not in self.visited_points[cindex]:
            self.instantiated_params[cindex].append(y)
            self.visited_points[cindex].add(k)
            return True
        return False

    def oracle(self, x, update_x=True, method='most violating all'):
        """Per constraint: values of the instantiated parameters (dropping those at or above
        constraint_drop_value), then -- unless method is None -- one minvalue query bounded by
        min(0, current minimum) whose result is instantiated if strictly deeper
        (sip.py:193-276, 'most violating all' / None branches)."""
        if method is not None and method != 'most violating all':
            raise NotImplementedError("only the reference's default oracle method is implemented")
        if not update_x:
            for c in self.constraints:
                c.setx(x)
        depths = []
        for i, c in enumerate(self.constraints):
            cdepths = constraint_values(c, x, self.instantiated_params[i])
            if method is not None and REMOVE_CONSTRAINTS:
                keep = [v < self.constraint_drop_value for v in cdepths]
                if not all(keep):
                    if VERBOSE_DROPS:
                        print("Dropping", len(keep) - sum(keep), "instantiated parameters of constraint", i)
                    self.instantiated_params[i] = [y for k, y in zip(keep, self.instantiated_params[i]) if k]
                    cdepths = [v for k, v in zip(keep, cdepths) if k]
                    self.visited_points[i] = set(self._key(y) for y in self.instantiated_params[i])
            olddmin = 0 if len(cdepths) == 0 else min(0, min(cdepths))
            if method is not None:
                for (dmin, newparam) in _as_pairs(c.minvalue(x, olddmin)):
                    if dmin < olddmin and self.instantiate(i, newparam):
                        cdepths.append(dmin)
            depths.append(cdepths)
        if not update_x:
            for c in self.constraints:
                c.clearx()
        return depths

    def solve_qp(self, W, xdes, A, b, xmin=None, xmax=None, **kwargs):
        if len(A) == 0:
            return xdes                # trust region and bounds are ignored with no rows   # quirk
        t0 = time.time()
        res = self.solve_qp_osqp(W, xdes, A, b, xmin, xmax, **kwargs)
        self.qp_time += time.time() - t0
        return res

    def solve_qp_osqp(self, W, xdes, A, b, xmin=None, xmax=None, verbose=0, regularizationFactor=1e-2,
                      slack_penalty='auto'):
        """min ||x - xdes||_W^2 + reg ||x||^2  s.t.  A x + b >= inflation, xmin <= x <= xmax, with
        the reference's slack re-solve when the strict problem is infeasible or its solution is
        out of scale (sip.py:294-378).  OSQP is replaced by the restated ADMM in _host/qp.py."""
        xdes = np.asarray(xdes, dtype=np.float64)
        n = len(xdes)
        m = len(b)
        if n <= _qp.DENSE_MAX_VARS and not scipy.sparse.issparse(W) and np.ndim(W) <= 1 \
                and not any(scipy.sparse.issparse(r) for r in A):
            return self._solve_qp_dense(W, xdes, A, b, xmin, xmax, regularizationFactor, slack_penalty)
        if isinstance(W, (int, float)):
            P = scipy.sparse.diags([float(W)] * n, format='csc')
        elif scipy.sparse.issparse(W):
            P = scipy.sparse.csc_matrix(W)
        else:
            W = np.asarray(W, dtype=np.float64)
            P = scipy.sparse.diags(W, format='csc') if W.ndim == 1 else scipy.sparse.csc_matrix(W)
        q = -P.dot(xdes)
        if regularizationFactor != 0:
            P = P + scipy.sparse.diags([regularizationFactor] * n, format='csc')
        rows = [r if scipy.sparse.issparse(r) else scipy.sparse.csr_matrix(np.asarray(r, dtype=np.float64).reshape(1, -1))
                for r in A]
        Amat = scipy.sparse.vstack(rows, format='csc')
        b = np.asarray(b, dtype=np.float64)
        l = -b + np.ones(m) * self.constraint_inflation
        u = np.full(m, np.inf)
        if xmin is not None:
            assert xmax is not None
            Amat = scipy.sparse.vstack((Amat, scipy.sparse.identity(n, format='csc')), format='csc')
            l = np.hstack((l, xmin))
            u = np.hstack((u, xmax))

        def run(Pm, qv, Am, pen=0.0):
            if Pm.shape[0] > _qp.DENSE_MAX_VARS:
                if _qp.NATIVE_BANDED:                       # opt-in native loop for banded problems (trajectories)
                    res = _qp.solve_qp_structured(Pm, qv, Am, l, u, n, m, xmin is not None, pen, eps_abs=1e-5)
                    if res is not None:
                        return res
                return _qp.solve_qp(Pm, qv, Am, l, u, eps_abs=1e-5)
            return _qp.solve_qp(Pm.toarray(), qv, Am.toarray(), l, u, eps_abs=1e-5)

        bnorm = 1
        if any(v < 0 for v in b):
            bnorm = np.linalg.norm([v for v in b if v < 0])
        results = run(P, q, Amat)
        if results.x is None or np.linalg.norm(results.x) * regularizationFactor > bnorm:
            if slack_penalty == 'auto':
                bmin = np.min(b)
                slack_penalty = (1 + len([v for v in b if v < 0])) / (abs(bmin) + 1e-3)
            Ps = scipy.sparse.bmat([[P, None], [None, scipy.sparse.diags([slack_penalty] * m, format='csc')]], format='csc')
            qs = np.hstack((q, np.zeros(m)))
            Aslack = scipy.sparse.identity(m, format='csc')
            if xmin is not None:
                Aslack = scipy.sparse.vstack((Aslack, scipy.sparse.csc_matrix((n, m))), format='csc')
            As = scipy.sparse.hstack((Amat, Aslack), format='csc')
            results = run(Ps, qs, As, float(slack_penalty))
            if results.x is None:
                raise RuntimeError("QP step failed even with slack variables")
        return results.x[:n]


    def _solve_qp_dense(self, W, xdes, A, b, xmin, xmax, regularizationFactor, slack_penalty):
        """solve_qp_osqp for a few variables with a diagonal weight (object poses, robot configurations): the very same
        matrices, built as numpy arrays directly.  Going through scipy.sparse (diags, vstack, bmat, toarray) cost 3 ms
        per QP -- more than solving it -- for identical entries."""
        n, m = len(xdes), len(b)
        w = np.full(n, float(W)) if np.ndim(W) == 0 else np.asarray(W, dtype=np.float64)
        q = -(w * xdes)
        P = np.diag(w + regularizationFactor) if regularizationFactor != 0 else np.diag(w)
        Amat = np.vstack([np.asarray(r, dtype=np.float64).reshape(1, -1) for r in A])
        b = np.asarray(b, dtype=np.float64)
        l = -b + np.ones(m) * self.constraint_inflation
        u = np.full(m, np.inf)
        if xmin is not None:
            assert xmax is not None
            Amat = np.vstack((Amat, np.eye(n)))
            l = np.hstack((l, xmin))
            u = np.hstack((u, xmax))
        bnorm = 1
        if any(v < 0 for v in b):
            bnorm = np.linalg.norm([v for v in b if v < 0])
        # column-major, as csc_matrix.toarray() hands them over: the memory order picks the BLAS kernel of A x and A'y,
        # hence the rounding, and the traces are compared with the reference's to the last bit
        P, Amat = np.asfortranarray(P), np.asfortranarray(Amat)
        results = _qp.solve_qp(P, q, Amat, l, u, eps_abs=1e-5)
        if results.x is None or np.linalg.norm(results.x) * regularizationFactor > bnorm:
            if slack_penalty == 'auto':
                bmin = np.min(b)
                slack_penalty = (1 + len([v for v in b if v < 0])) / (abs(bmin) + 1e-3)
            Ps = np.zeros((n + m, n + m), order='F')
            Ps[:n, :n] = P
            Ps[n:, n:] = np.diag([float(slack_penalty)] * m)
            qs = np.hstack((q, np.zeros(m)))
            As = np.zeros((Amat.shape[0], n + m), order='F')
            As[:, :n] = Amat
            As[:m, n:] = np.eye(m)
            results = _qp.solve_qp(Ps, qs, As, l, u, eps_abs=1e-5)
            if results.x is None:
                raise RuntimeError("QP step failed even with slack variables")
        return results.x[:n]


def optimizeSemiInfinite(objective, constraints, xinit, xmin=None, xmax=None, settings=None, verbose=1):
    """min f(x)  s.t.  c(x,y) >= 0 for all y in domain(c)  [and xmin <= x <= xmax]
    by constraint generation with a trust-region QP step (sip.py:795-1241, default flags).
    Returns a SemiInfiniteOptimizationResult."""
    if not all(isinstance(c, SemiInfiniteConstraintInterface) for c in constraints):
        constraints = [c if isinstance(c, SemiInfiniteConstraintInterface) else SemiInfiniteConstraintAdaptor(c)
                       for c in constraints]
    res = SemiInfiniteOptimizationResult()
    res.x0 = xinit
    x = xinit
    res.trace = [xinit]
    res.trace_times = [0]
    if settings is None:
        settings = SemiInfiniteOptimizationSettings()
    objScoreWeight = settings.initial_objective_score_weight
    res.fx0 = objective.value(x)
    res.gx0 = np.zeros(len(constraints))
    res.x = x
    res.fx = res.fx0
    res.gx = np.zeros(len(constraints))
    cdata = ConstraintGenerationData(constraints)
    cdata.exclusion_distance = settings.parameter_exclusion_distance
    cdata.constraint_inflation = settings.constraint_inflation
    cdata.constraint_drop_value = settings.constraint_drop_value
    res.instantiated_params = cdata.instantiated_params
    minimum_constraint_value = settings.minimum_constraint_value
    xepsilon2 = settings.xepsilon ** 2 * len(x)      # len of an se3 pose (R,t) is 2              # quirk
    update_oracle = True
    trust_region_size = 0.5
    tstart = time.time()
    tgeom = 0.0
    iters = 0
    for iters in range(settings.max_iters):
        dxdes = np.asarray(objective.minstep(x))
        for c in constraints:
            c.setx(x)
        t1 = time.time()
        # new parameters are looked for here only on iterations 0, 1 and multiples of 100;        # quirk
        # otherwise discovery happens in score() at the proposed step (sip.py:855-870)
        cdepths = cdata.oracle(x, update_x=False, method=ORACLE_METHOD if update_oracle else None)
        update_oracle = True
        if iters % 100 != 0:
            update_oracle = False
        res.gx = np.zeros(len(constraints))
        depths = []
        rows = []
        for i, c in enumerate(constraints):
            res.gx[i] = min(cdepths[i]) if len(cdepths[i]) > 0 else float('inf')
            if iters == 0:
                res.gx0[i] = res.gx[i]
                if res.gx0[i] < settings.minimum_constraint_value:
                    if verbose >= 1:
                        print("WARNING: initial constraint value %d is below the minimum %g < %g" %
                              (i, res.gx0[i], settings.minimum_constraint_value))
                    minimum_constraint_value = res.gx0[i]
            # one row per instantiated parameter (sip.py:893-897), fetched in one batch
            rws = constraint_rows(c, x, cdata.instantiated_params[i])
            depths.extend(cdepths[i])
            rows.extend(rws)
        tgeom += time.time() - t1
        for c in constraints:
            c.clearx()
        if iters == 0 and verbose >= 1:
            print("Beginning collision-free optimization at f(x) =", res.fx0, "g(x) =", res.gx0)

        # box trust region intersected with the variable bounds (sip.py:945-953)
        n = len(dxdes)
        if xmin is None:
            dxmax = np.ones(n) * trust_region_size
            dxmin = -dxmax
        else:
            dxmin = np.maximum(xmin - x, -np.ones(n) * trust_region_size)
            dxmax = np.minimum(xmax - x, np.ones(n) * trust_region_size)
        H = objective.hessian(x)
        weight = objScoreWeight * H
        dx = cdata.solve_qp(W=weight, xdes=dxdes, A=rows, b=depths, xmin=dxmin, xmax=dxmax,
                            verbose=verbose, **settings.qp_solver_params)
        dx = np.asarray(dx, dtype=np.float64)
        dxnorm2 = float(np.dot(dx, dx))
        if dxnorm2 < xepsilon2:
            res.status = 'local optimum'
            break
        dxnorm = math.sqrt(dxnorm2)
        xnext = objective.integrate(x, dx)
        sorig = _score(res.fx, cdepths, objScoreWeight)
        num_instantiations0 = sum(len(p) for p in cdata.instantiated_params)
        t0 = time.time()
        snew, fxnew, gxnew = score(objective, constraints, xnext, cdata, objScoreWeight,
                                   discoverNew=DETECT_NEW_COLLISIONS_IN_STEP)
        tgeom += time.time() - t0
        accept_step = True
        reject_reason = None
        if snew >= sorig:
            accept_step = False
            reject_reason = 'increase in score'
        if np.min(gxnew) < minimum_constraint_value:
            reject_reason = 'under minimum constraint value'   # only labels the step; does not reject  # quirk
        if DETECT_NEW_COLLISIONS_IN_STEP:
            num_instantiations = sum(len(p) for p in cdata.instantiated_params)
            if num_instantiations > num_instantiations0 and np.min(gxnew) < 0:
                update_oracle = False
                accept_step = False
                reject_reason = 'new deepest collision'
        if not accept_step:
            xnext = x
            fxnew = res.fx
            gxnew = res.gx
            trust_region_size *= 0.5
            if trust_region_size > dxnorm:
                trust_region_size = np.max(np.abs(dx)) * 0.5
            if verbose >= 1:
                print("  Step length %g rejected due to %s, shrinking trust region to %g" %
                      (dxnorm, reject_reason, trust_region_size))
        else:
            trust_region_size *= 5.0 / 2.0
            if verbose >= 1:
                print("  Step length %g accepted, growing trust region to %g" % (dxnorm, trust_region_size))
        x = xnext
        res.trace.append(x)
        res.trace_times.append(time.time() - tstart)
        res.fx = fxnew
        res.gx = gxnew
        # objective weight schedule: shrink while violating, grow while feasible (sip.py:1210-1215)
        maxviolation = min(res.gx)
        scale = 0.75 * ((math.atan(maxviolation * 3) * 2 / math.pi + 0.5) * 1.75 + 0.25)
        objScoreWeight *= scale
        if objScoreWeight < settings.minimum_objective_score_weight:
            objScoreWeight = settings.minimum_objective_score_weight

    tfinish = time.time()
    res.x = x
    res.num_iterations = iters
    res.time = tfinish - tstart
    res.time_cons = tgeom
    res.time_qp = cdata.qp_time
    if verbose >= 1:
        print("Completed in time", res.time, "and", iters, "iterations")
        print("   geometry time", tgeom)
        print("   final values f(x) =", res.fx, "g(x) =", res.gx)
    return res


def optimizeStandard(objective, constraints, xinit, xmin=None, xmax=None, settings=None, verbose=1):
    """The reference's plain-NLP loop (sip.py:483-761) is a non-default path outside the collision
    hot path; plain constraints are served by the same exchange loop through the adaptor."""
    return optimizeSemiInfinite(objective, [SemiInfiniteConstraintAdaptor(c) for c in constraints], xinit, xmin, xmax,
                                settings, verbose)
