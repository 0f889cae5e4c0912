"""Lock-step solver for MANY independent object-pose problems (BASELINE.json config 5: 65,536
cube-vs-scene `optimizeCollFree` problems).

Every problem is exactly the reference's `optimizeCollFree(obj, env, Tinit, Tdes)` (gopt:1039-1081 on
sip.py:795-1241, default flags); what changes is the execution order.  The reference would run the
problems one after another, each issuing a handful of single Klampt calls per iteration.  Here all
problems advance through the SAME outer iteration together, so each oracle request of an iteration
becomes one batched device call over all live problems:

    values + rows of every instantiated parameter of every problem   -> sio_pose_rows (1 launch)
    minvalue of every problem (B poses x N cloud points)              -> sio_min_distance (K2 sweep)

Problems are independent, so they shard across GPUs by contiguous blocks with no collective during
the solve (dist.py gathers the per-problem records at the end).

The per-problem arithmetic (trust region, merit score, weight schedule, exclusion hash, quirks) is
the scalar solver's, statement for statement; tests run both and compare traces.  The QP is solved
per problem on the host (`qp='scalar'`, bit-identical to the scalar solver) or, for throughput, for
all problems of one size at once with the lock-step ADMM of _host/qp.py (`qp='batch'`).
"""
import math
import time

import numpy as np

from . import _abi, dist
from . import sip as _sip
from ._host import qp as _qp
from .geometryopt import BULK_FLAGS, POINT_FLAGS, PenetrationDepthGeometry, _inv12, _mul12


def _log_so3_batch(M):
    """Rotation vectors of a batch of rotation matrices (B,3,3), same branches as _host/so3."""
    from ._host import so3
    return np.array([so3.rotation_vector(so3.from_matrix(m)) for m in M])


def _exp_so3_batch(W):
    from ._host import so3
    return np.array([so3.matrix(so3.from_rotation_vector(w)) for w in W])


class BatchResult:
    def __init__(self, B, per_problem_lists=True):
        """per_problem_lists=False (the device-resident driver): the per-problem Python lists (`status`,
        `instantiated_params`) are built from the returned arrays on first use -- 65,536 fresh lists cost tens of
        milliseconds of allocation and garbage-collector scans, a tenth of the whole solve."""
        self.B = B
        self.T = None                    # (B,12) final poses, row-major 3x4
        self._status = ['not converged'] * B if per_problem_lists else None
        self.status_codes = None         # (B,) int32: 1 = local optimum, 0 = not converged, 2 = QP without a solution
        self.num_iterations = np.zeros(B, dtype=np.int64)
        self.fx = np.zeros(B)
        self.gx = np.full(B, np.inf)
        self.gx0 = np.full(B, np.inf)
        self._params = [[] for _ in range(B)] if per_problem_lists else None
        self._params_arrays = None
        self.trace = None
        self.time = 0.0
        self.time_cons = 0.0
        self.time_qp = 0.0
        self.outer_iterations = 0        # sum over problems of completed outer iterations
        self.point_queries = 0           # (pose, cloud point) pairs evaluated by K2 sweeps

    @property
    def status(self):
        if self._status is None:
            names = {1: 'local optimum', 0: 'not converged', 2: 'qp failed'}          # 2: the reference raises there (sip.py:358-378)
            self._status = [names.get(int(c), 'not converged') for c in self.status_codes]
        return self._status

    @status.setter
    def status(self, v):
        self._status = v

    @property
    def instantiated_params(self):
        """Per problem, the list of instantiated world points (built on first use from the device arrays)."""
        if self._params_arrays is not None:
            P, n = self._params_arrays
            self._params = [[list(P[p, k]) for k in range(n[p])] for p in range(len(n))]
            self._params_arrays = None
        return self._params

    @instantiated_params.setter
    def instantiated_params(self, v):
        self._params, self._params_arrays = v, None

    def set_params_arrays(self, P, n):
        self._params_arrays = (P, n)
        self.n_params = np.asarray(n)

    def records(self):
        """Fixed-size per-problem records for the end-of-solve gather: pose(12) gx fx iters nparams."""
        B = self.B
        rec = np.zeros((B, 16))
        rec[:, :12] = self.T
        rec[:, 12] = self.gx
        rec[:, 13] = self.fx
        rec[:, 14] = self.num_iterations
        rec[:, 15] = self.n_params if self._params_arrays is not None else [len(p) for p in self.instantiated_params]
        return rec


def optimizeCollFreeBatch(obj, env, Tinits, Tdess=None, settings=None, qp='scalar', keep_trace=False, context=None):
    """Solves B independent `optimizeCollFree` problems in lock step.

    obj / env: PenetrationDepthGeometry (obj owns the SDF grid and moves; env owns the cloud).
    Tinits / Tdess: sequences of Klampt se3 poses (R column-major 9-list, t)."""
    from ._host import se3, so3
    assert isinstance(obj, PenetrationDepthGeometry) and isinstance(env, PenetrationDepthGeometry)
    ctx = context or obj.context
    settings = settings or _sip.SemiInfiniteOptimizationSettings()
    B = len(Tinits)
    if Tdess is None:
        Tdess = Tinits
    R = np.array([so3.matrix(T[0]) for T in Tinits])
    t = np.array([T[1] for T in Tinits], dtype=np.float64)
    Rdes = np.array([so3.matrix(T[0]) for T in Tdess])
    tdes = np.array([T[1] for T in Tdess], dtype=np.float64)
    Tenv = env._T12.reshape(3, 4)
    cloud_world = env._dev.cloud32.astype(np.float64) @ Tenv[:, :3].T + Tenv[:, 3]
    grid_h, cloud_h = obj._dev.grid_h, env._dev.cloud_h
    N = len(cloud_world)

    res = BatchResult(B)
    params = res.instantiated_params
    visited = [set() for _ in range(B)]
    excl = settings.parameter_exclusion_distance
    w = np.full(B, settings.initial_objective_score_weight)
    trs = np.full(B, 0.5)
    update_oracle = np.ones(B, dtype=bool)
    live = np.ones(B, dtype=bool)
    min_cv = np.full(B, settings.minimum_constraint_value)
    xeps2 = settings.xepsilon ** 2 * 2          # len((R,t)) == 2 in the reference        # quirk
    if keep_trace:
        res.trace = [[(so3.from_matrix(R[p]), list(t[p]))] for p in range(B)]

    def pose12(idx):
        T = np.empty((len(idx), 3, 4))
        T[:, :, :3] = R[idx]
        T[:, :, 3] = t[idx]
        return T.reshape(-1, 12)

    def objective(idx, Rm, tm):
        rv = _log_so3_batch(np.einsum('bij,bkj->bik', Rdes[idx], Rm))
        dt = tdes[idx] - tm
        return (dt * dt).sum(1) + (rv * rv).sum(1), np.hstack([rv, dt])

    def key(y):
        return tuple(int(v / excl) for v in y)

    def instantiate(p, y):
        k = key(y)
        if k not in visited[p]:
            params[p].append(y)
            visited[p].add(k)
            return True
        return False

    def values(idx, Rm, tm, want_rows):
        """values (and rows) of every instantiated parameter of problems idx at poses (Rm, tm)."""
        counts = [len(params[p]) for p in idx]
        offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        if offs[-1] == 0:
            return offs, np.zeros(0), np.zeros((0, 6))
        pts = np.array([y for p in idx for y in params[p]], dtype=np.float64)
        T = np.empty((len(idx), 3, 4))
        T[:, :, :3] = Rm
        T[:, :, 3] = tm
        d, rows = ctx.pose_rows(grid_h, T.reshape(-1, 12), offs, pts, flags=POINT_FLAGS, want_rows=want_rows)
        return offs, d, rows

    def minvalue(idx, Rm, tm, bound):
        """(dmin, arg) of every problem in idx: one K2 sweep of len(idx) poses x N points."""
        T = np.empty((len(idx), 3, 4))
        T[:, :, :3] = Rm
        T[:, :, 3] = tm
        Tinv = np.empty_like(T)
        Tinv[:, :, :3] = np.transpose(Rm, (0, 2, 1))
        Tinv[:, :, 3] = -np.einsum('bji,bj->bi', Rm, tm)
        # T_rel = T_obj^-1 * T_env
        Trel = np.empty_like(T)
        Trel[:, :, :3] = Tinv[:, :, :3] @ Tenv[:, :3]
        Trel[:, :, 3] = Tinv[:, :, :3] @ Tenv[:, 3] + Tinv[:, :, 3]
        dmin, arg, _ = ctx.min_distance(grid_h, cloud_h, Trel.reshape(-1, 12), bound=bound, flags=BULK_FLAGS, want_under=False)
        res.point_queries += len(idx) * N
        return dmin, arg

    fx0, _ = objective(np.arange(B), R, t)
    res.fx = fx0.copy()
    tstart = time.time()
    for it in range(settings.max_iters):
        idx = np.nonzero(live)[0]
        if len(idx) == 0:
            break
        res.num_iterations[idx] = it
        _, dxdes = objective(idx, R[idx], t[idx])
        # ---- oracle at x (sip.py:193-276) ---------------------------------------------------
        t0 = time.time()
        offs, d, _ = values(idx, R[idx], t[idx], False)
        depths = []
        for k, p in enumerate(idx):
            cd = list(d[offs[k]:offs[k + 1]])
            if update_oracle[p]:
                keep = [v < settings.constraint_drop_value for v in cd]
                if not all(keep):
                    params[p][:] = [y for kk, y in zip(keep, params[p]) if kk]
                    cd = [v for kk, v in zip(keep, cd) if kk]
                    visited[p] = set(key(y) for y in params[p])
            depths.append(cd)
        sel = [k for k, p in enumerate(idx) if update_oracle[p]]
        if sel:
            olddmin = np.array([0.0 if len(depths[k]) == 0 else min(0.0, min(depths[k])) for k in sel])
            pi = idx[sel]
            dmin, arg = minvalue(pi, R[pi], t[pi], olddmin)
            for j, k in enumerate(sel):
                if arg[j] >= 0 and dmin[j] < olddmin[j]:
                    if instantiate(idx[k], list(cloud_world[arg[j]])):
                        depths[k].append(float(dmin[j]))
        update_oracle[idx] = (it % 100 == 0)
        offs, _, rows = values(idx, R[idx], t[idx], True)
        res.time_cons += time.time() - t0
        gx = np.array([min(cd) if cd else np.inf for cd in depths])
        if it == 0:
            res.gx0[idx] = gx
            low = gx < min_cv[idx]
            min_cv[idx[low]] = gx[low]
        # ---- QP step (sip.py:945-964) -----------------------------------------------------------
        t0 = time.time()
        dx = np.zeros((len(idx), 6))
        cd_helper = _sip.ConstraintGenerationData([None])
        cd_helper.constraint_inflation = settings.constraint_inflation
        for k, p in enumerate(idx):
            A = rows[offs[k]:offs[k + 1]]
            if len(A) == 0:
                dx[k] = dxdes[k]                     # no rows: xdes, trust region ignored        # quirk
                continue
            if qp == 'scalar':
                dx[k] = cd_helper.solve_qp_osqp(w[p] * np.ones(6), dxdes[k], list(A), depths[k],
                                                -np.ones(6) * trs[p], np.ones(6) * trs[p], **settings.qp_solver_params)
        if qp == 'batch':
            _solve_batch(dx, idx, rows, offs, depths, dxdes, w, trs, settings, cd_helper)
        res.time_qp += time.time() - t0
        dxn2 = (dx * dx).sum(1)
        conv = dxn2 < xeps2
        for k in np.nonzero(conv)[0]:
            res.status[idx[k]] = 'local optimum'
        live[idx[conv]] = False
        res.gx[idx[conv]] = gx[conv]
        go = ~conv
        if not go.any():
            continue
        idx, dx, dxn2, gx = idx[go], dx[go], dxn2[go], gx[go]
        depths = [cd for cd, g_ in zip(depths, go) if g_]
        dxnorm = np.sqrt(dxn2)
        # ---- proposed step and merit (sip.py:1146-1156, 134-172) ------------------------------------
        Rn = np.einsum('bij,bjk->bik', _exp_so3_batch(dx[:, :3]), R[idx])
        tn = t[idx] + dx[:, 3:]
        sorig = np.array([_sip._score(res.fx[p], [cd], w[p]) for p, cd in zip(idx, depths)])
        ninst0 = np.array([len(params[p]) for p in idx])
        t0 = time.time()
        fxn, _ = objective(idx, Rn, tn)
        offs, d, _ = values(idx, Rn, tn, False)
        gxn = np.array([d[offs[k]:offs[k + 1]].min() if offs[k + 1] > offs[k] else np.inf for k in range(len(idx))])
        gneg = [[v for v in d[offs[k]:offs[k + 1]] if v < 0] for k in range(len(idx))]
        dmin, arg = minvalue(idx, Rn, tn, gxn)
        for k, p in enumerate(idx):
            if arg[k] >= 0 and dmin[k] < gxn[k]:
                gxn[k] = dmin[k]
                if instantiate(p, list(cloud_world[arg[k]])) and dmin[k] < 0:
                    gneg[k].append(float(dmin[k]))
        res.time_cons += time.time() - t0
        snew = np.array([(_sip.scoring_metric(g) if g else 0.0) for g in gneg]) + w[idx] * fxn
        for k, p in enumerate(idx):
            accept = not (snew[k] >= sorig[k])
            if len(params[p]) > ninst0[k] and gxn[k] < 0:
                update_oracle[p] = False
                accept = False
            if accept:
                R[p], t[p] = Rn[k], tn[k]
                res.fx[p], res.gx[p] = fxn[k], gxn[k]
                trs[p] *= 2.5
            else:
                res.gx[p] = gx[k]
                trs[p] *= 0.5
                if trs[p] > dxnorm[k]:
                    trs[p] = np.max(np.abs(dx[k])) * 0.5
            if keep_trace:
                res.trace[p].append((so3.from_matrix(R[p]), list(t[p])))
            scale = 0.75 * ((math.atan(res.gx[p] * 3) * 2 / math.pi + 0.5) * 1.75 + 0.25)
            w[p] = max(w[p] * scale, settings.minimum_objective_score_weight)
            res.outer_iterations += 1
    res.T = pose12(np.arange(B))
    res.time = time.time() - tstart
    return res


def _solve_batch(dx, idx, rows, offs, depths, dxdes, w, trs, settings, cd_helper):
    """All QPs of one row count at once with the lock-step ADMM; problems whose strict solution
    trips the reference's slack criterion are re-solved individually."""
    reg = settings.qp_solver_params.get('regularizationFactor', 1e-2)
    counts = np.array([offs[k + 1] - offs[k] for k in range(len(idx))])
    for m in np.unique(counts):
        if m == 0:
            continue
        ks = np.nonzero(counts == m)[0]
        nb = len(ks)
        P = np.zeros((nb, 6, 6))
        P[:, np.arange(6), np.arange(6)] = (w[idx[ks]] + reg)[:, None]
        q = -w[idx[ks]][:, None] * dxdes[ks]
        A = np.zeros((nb, m + 6, 6))
        l = np.zeros((nb, m + 6))
        u = np.full((nb, m + 6), _qp.INF)
        for j, k in enumerate(ks):
            A[j, :m] = rows[offs[k]:offs[k + 1]]
            l[j, :m] = -np.asarray(depths[k]) + settings.constraint_inflation
        A[:, m:, :] = np.eye(6)
        l[:, m:] = -trs[idx[ks]][:, None]
        u[:, m:] = trs[idx[ks]][:, None]
        x, rp, rd = _qp.solve_qp_batch(P, q, A, l, u, None)
        for j, k in enumerate(ks):
            b = np.asarray(depths[k])
            bnorm = np.linalg.norm(b[b < 0]) if (b < 0).any() else 1.0
            if rp[j] > 1e-4 or rd[j] > 1e-4 or np.linalg.norm(x[j]) * reg > bnorm:
                p = idx[k]
                dx[k] = cd_helper.solve_qp_osqp(w[p] * np.ones(6), dxdes[k], list(rows[offs[k]:offs[k + 1]]), depths[k],
                                                -np.ones(6) * trs[p], np.ones(6) * trs[p], **settings.qp_solver_params)
            else:
                dx[k] = x[j]


def solve_sharded(obj, env, Tinits, Tdess=None, settings=None, qp='scalar', driver='scalar'):
    """C5 across ranks: contiguous blocks of problems per rank, no collective during the solve, one
    gather of the fixed-size per-problem records at the end.  driver: 'scalar'
    (optimizeCollFreeBatch), 'fast' (optimizeCollFreeBatchFast) or 'device' (optimizeCollFreeBatchDevice).
    Returns (records of ALL problems [pose(12) gx fx iterations nparams], this rank's BatchResult)."""
    lo, hi = dist.shard_range(len(Tinits))
    Ti = Tinits[lo:hi]
    Td = None if Tdess is None else Tdess[lo:hi]
    if driver == 'device':
        local = optimizeCollFreeBatchDevice(obj, env, Ti, Td, settings)
    elif driver == 'fast':
        local = optimizeCollFreeBatchFast(obj, env, Ti, Td, settings)
    else:
        local = optimizeCollFreeBatch(obj, env, Ti, Td, settings, qp)
    return dist.gather_records(local.records(), n_total=len(Tinits)), local


# =================================================================================================
# array-at-a-time lock-step driver: no per-problem Python, native batched QP
# =================================================================================================
def _log_so3_vec(M):
    """Vectorised _host/so3.rotation_vector (same three branches) for a batch (B,3,3)."""
    c = np.clip((np.trace(M, axis1=1, axis2=2) - 1.0) * 0.5, -1.0, 1.0)
    th = np.arccos(c)
    v = np.stack([M[:, 2, 1] - M[:, 1, 2], M[:, 0, 2] - M[:, 2, 0], M[:, 1, 0] - M[:, 0, 1]], axis=1)
    out = np.empty_like(v)
    small = th < 1e-8
    nearpi = ~small & (np.abs(th - math.pi) < 1e-5)
    gen = ~small & ~nearpi
    out[small] = 0.5 * v[small]
    out[gen] = v[gen] * (th[gen] / (2.0 * np.sin(th[gen])))[:, None]
    for k in np.nonzero(nearpi)[0]:                      # measure-zero branch: scalar code
        from ._host import so3
        out[k] = so3.rotation_vector(so3.from_matrix(M[k]))
    return out


def _exp_so3_vec(W):
    """Vectorised _host/so3.from_rotation_vector (Rodrigues) for a batch (B,3) -> (B,3,3)."""
    th = np.linalg.norm(W, axis=1)
    K = np.zeros((len(W), 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -W[:, 2], W[:, 1]
    K[:, 1, 0], K[:, 1, 2] = W[:, 2], -W[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -W[:, 1], W[:, 0]
    K2 = K @ K
    small = th < 1e-12
    ths = np.where(small, 1.0, th)
    a = np.where(small, 1.0, np.sin(ths) / ths)
    b = np.where(small, 0.5, (1 - np.cos(ths)) / (ths * ths))
    return np.eye(3)[None] + a[:, None, None] * K + b[:, None, None] * K2


def optimizeCollFreeBatchFast(obj, env, Tinits, Tdess=None, settings=None, context=None, flags=None, keep_trace=False, qp='device'):
    """The same lock-step algorithm as optimizeCollFreeBatch, written array-at-a-time: instantiated
    parameters live in padded (B, cap, 3) arrays, the exclusion hash is an integer key array, every
    oracle request is one device call over all live problems and all QPs of an outer iteration are
    solved by ONE call: qp='device' -> sio_sip_step_qp (one warp per problem; the few problems that need the
    reference's slack re-solve go to the host solver), qp='host' -> _host/csrc/hostqp.c (OpenMP).  Per-problem
    arithmetic and quirks are those of sip.py:795-1241 / the scalar driver above; the QP answers differ
    from the scalar path only in rounding (reduced n x n system instead of the KKT LU)."""
    from ._host import se3, so3
    assert isinstance(obj, PenetrationDepthGeometry) and isinstance(env, PenetrationDepthGeometry)
    ctx = context or obj.context
    settings = settings or _sip.SemiInfiniteOptimizationSettings()
    if flags is None:
        flags = BULK_FLAGS | _abi.SIO_PRUNE
    B = len(Tinits)
    if Tdess is None:
        Tdess = Tinits
    R = np.array([so3.matrix(T[0]) for T in Tinits])
    t = np.array([T[1] for T in Tinits], dtype=np.float64)
    Rdes = np.array([so3.matrix(T[0]) for T in Tdess])
    tdes = np.array([T[1] for T in Tdess], dtype=np.float64)
    Tenv = env._T12.reshape(3, 4)
    cloud_world = env._dev.cloud32.astype(np.float64) @ Tenv[:, :3].T + Tenv[:, 3]
    grid_h, cloud_h = obj._dev.grid_h, env._dev.cloud_h
    N = len(cloud_world)
    res = BatchResult(B)
    cap = 8
    P = np.zeros((B, cap, 3))
    keys = np.zeros((B, cap, 3), dtype=np.int64)
    cnt = np.zeros(B, dtype=np.int64)
    excl = settings.parameter_exclusion_distance
    w = np.full(B, settings.initial_objective_score_weight)
    trs = np.full(B, 0.5)
    update_oracle = np.ones(B, dtype=bool)
    live = np.ones(B, dtype=bool)
    min_cv = np.full(B, settings.minimum_constraint_value)
    xeps2 = settings.xepsilon ** 2 * 2          # len((R,t)) == 2 in the reference        # quirk
    reg = settings.qp_solver_params.get('regularizationFactor', 1e-2)
    status = np.zeros(B, dtype=bool)             # True = 'local optimum'
    if keep_trace:
        res.trace = [[(so3.from_matrix(R[p]), list(t[p]))] for p in range(B)]

    def objective(idx, Rm, tm):
        rv = _log_so3_vec(np.einsum('bij,bkj->bik', Rdes[idx], Rm))
        dt = tdes[idx] - tm
        return (dt * dt).sum(1) + (rv * rv).sum(1), np.hstack([rv, dt])

    def valid_mask(idx):
        return np.arange(cap)[None, :] < cnt[idx][:, None]

    def values(idx, Rm, tm, want_rows):
        """flat values (and rows) of every instantiated parameter of problems idx, problem-major."""
        mask = valid_mask(idx)
        counts = cnt[idx]
        offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        if offs[-1] == 0:
            return mask, offs, np.zeros(0), np.zeros((0, 6))
        T = np.empty((len(idx), 3, 4))
        T[:, :, :3] = Rm
        T[:, :, 3] = tm
        r_, c_ = np.nonzero(mask)                          # problem-major: gather only the live parameters
        d, rows = ctx.pose_rows(grid_h, T.reshape(-1, 12), offs, P[idx[r_], c_], flags=POINT_FLAGS, want_rows=want_rows)
        return mask, offs, d, rows

    def padded(mask, d):
        D = np.full(mask.shape, np.inf)
        D[mask] = d
        return D

    def minvalue(idx, Rm, tm, bound):
        Tinv_R = np.transpose(Rm, (0, 2, 1))
        Trel = np.empty((len(idx), 3, 4))
        Trel[:, :, :3] = Tinv_R @ Tenv[:, :3]
        Trel[:, :, 3] = np.einsum('bij,j->bi', Tinv_R, Tenv[:, 3]) - np.einsum('bji,bj->bi', Rm, tm)
        dmin, arg, _ = ctx.min_distance(grid_h, cloud_h, Trel.reshape(-1, 12), bound=bound, flags=flags, want_under=False)
        res.point_queries += len(idx) * N
        return dmin, arg

    def instantiate(pidx, pts):
        """appends pts[k] to problem pidx[k] unless its exclusion key is present; returns the added mask."""
        nonlocal P, keys, cap
        if len(pidx) == 0:
            return np.zeros(0, dtype=bool)
        k = np.trunc(pts / excl).astype(np.int64)                                   # int() truncates toward zero (sip.py:280)
        seen = ((keys[pidx] == k[:, None, :]).all(-1) & valid_mask(pidx)).any(1)
        add = ~seen
        tgt = pidx[add]
        if len(tgt) and cnt[tgt].max() >= cap:
            P = np.concatenate([P, np.zeros_like(P)], axis=1)
            keys = np.concatenate([keys, np.zeros_like(keys)], axis=1)
            cap *= 2
        P[tgt, cnt[tgt]] = pts[add]
        keys[tgt, cnt[tgt]] = k[add]
        cnt[tgt] += 1
        return add

    fx0, _ = objective(np.arange(B), R, t)
    res.fx = fx0.copy()
    res.iter_log = []                            # (iteration, live problems, QP rows, QP seconds)
    tstart = time.time()
    for it in range(settings.max_iters):
        idx = np.nonzero(live)[0]
        if len(idx) == 0:
            break
        res.num_iterations[idx] = it
        _, dxdes = objective(idx, R[idx], t[idx])
        # ---- oracle at x (sip.py:193-276) ---------------------------------------------------
        t0 = time.time()
        uo = update_oracle[idx]
        # the oracle only drops / discovers parameters on iterations 0, 1, 100k (sip.py:855-870): on every other
        # iteration nothing changes between the values and the rows, so one device call returns both
        rows = None
        if uo.any():
            mask, offs, d, _ = values(idx, R[idx], t[idx], False)
        else:
            mask, offs, d, rows = values(idx, R[idx], t[idx], True)
        D = padded(mask, d)
        drop = mask & (D >= settings.constraint_drop_value) & uo[:, None]
        if drop.any():
            rows_ = np.nonzero(drop.any(1))[0]
            order = np.argsort(drop[rows_] | ~mask[rows_], axis=1, kind='stable')     # kept entries first, original order
            pi = idx[rows_]
            P[pi] = np.take_along_axis(P[pi], order[:, :, None], axis=1)
            keys[pi] = np.take_along_axis(keys[pi], order[:, :, None], axis=1)
            D[rows_] = np.take_along_axis(D[rows_], order, axis=1)
            cnt[pi] -= drop[rows_].sum(1)
            mask = valid_mask(idx)
            D[~mask] = np.inf
        sel = np.nonzero(uo)[0]
        if len(sel):
            olddmin = np.minimum(0.0, D[sel].min(1))
            pi = idx[sel]
            dmin, arg = minvalue(pi, R[pi], t[pi], olddmin)
            cond = (arg >= 0) & (dmin < olddmin)
            if cond.any():
                slot = cnt[pi[cond]].copy()
                added = instantiate(pi[cond], cloud_world[arg[cond]])
                if cap > D.shape[1]:
                    D = np.concatenate([D, np.full((len(D), cap - D.shape[1]), np.inf)], axis=1)
                rr = sel[cond][added]
                D[rr, slot[added]] = dmin[cond][added]
        update_oracle[idx] = (it % 100 == 0)
        if rows is None:
            mask, offs, _, rows = values(idx, R[idx], t[idx], True)
        res.time_cons += time.time() - t0
        gx = D.min(1)
        if it == 0:
            res.gx0[idx] = gx
            low = gx < min_cv[idx]
            min_cv[idx[low]] = gx[low]
        # ---- QP step (sip.py:945-964): every problem in one native call -----------------------------
        t0 = time.time()
        ones6 = np.ones((len(idx), 6))
        Wq, lo_q, hi_q, dep = w[idx][:, None] * ones6, -trs[idx][:, None] * ones6, trs[idx][:, None] * ones6, D[mask]
        if qp == 'device':
            dx, qst = ctx.sip_step_qp(Wq, dxdes, offs, rows, dep, lo_q, hi_q, reg=reg, inflation=settings.constraint_inflation)
            res.time_qp_device = getattr(res, 'time_qp_device', 0.0) + time.time() - t0
            qi = ctx.last_qp_iters(detail=True)
            res.qp_admm_iterations = getattr(res, 'qp_admm_iterations', 0) + qi[0]
            res.qp_iter_log = getattr(res, 'qp_iter_log', []) + [(len(idx), qi[0], qi[1], qi[2])]
            todo = np.nonzero(qst == 3)[0]
            if len(todo):                            # slack re-solve (sip.py:332-357) on the host
                cn = (offs[1:] - offs[:-1])[todo]
                sub = np.concatenate([[0], np.cumsum(cn)]).astype(np.int64)
                take = np.concatenate([np.arange(offs[k], offs[k + 1]) for k in todo]) if sub[-1] else np.zeros(0, dtype=np.int64)
                dx[todo], _ = _qp.sip_step_batch(Wq[todo], dxdes[todo], sub, rows[take], dep[take], lo_q[todo], hi_q[todo], reg=reg,
                                                 inflation=settings.constraint_inflation)
                res.qp_host_fallbacks = getattr(res, 'qp_host_fallbacks', 0) + len(todo)
        else:
            dx, _ = _qp.sip_step_batch(Wq, dxdes, offs, rows, dep, lo_q, hi_q, reg=reg, inflation=settings.constraint_inflation)
        res.time_qp += time.time() - t0
        res.iter_log.append((it, len(idx), int(offs[-1]), time.time() - t0))
        dxn2 = (dx * dx).sum(1)
        conv = dxn2 < xeps2
        status[idx[conv]] = True
        live[idx[conv]] = False
        res.gx[idx[conv]] = gx[conv]
        go = ~conv
        if not go.any():
            continue
        idx, dx, dxn2, gx, D = idx[go], dx[go], dxn2[go], gx[go], D[go]
        dxnorm = np.sqrt(dxn2)
        # ---- proposed step and merit (sip.py:1146-1156, 134-172) ------------------------------------
        Rn = _exp_so3_vec(dx[:, :3]) @ R[idx]
        tn = t[idx] + dx[:, 3:]
        sorig = np.maximum(0.0, -D.min(1)) + w[idx] * res.fx[idx]
        ninst0 = cnt[idx].copy()
        t0 = time.time()
        fxn, _ = objective(idx, Rn, tn)
        mask, offs, d, _ = values(idx, Rn, tn, False)
        Dn = padded(mask, d)
        gxn = Dn.min(1)
        metric = np.maximum(0.0, -gxn)
        dmin, arg = minvalue(idx, Rn, tn, gxn)
        cond = (arg >= 0) & (dmin < gxn)
        if cond.any():
            gxn[cond] = dmin[cond]
            added = instantiate(idx[cond], cloud_world[arg[cond]])
            hit = np.nonzero(cond)[0][added & (dmin[cond] < 0)]
            metric[hit] = np.maximum(metric[hit], -dmin[hit])
        res.time_cons += time.time() - t0
        snew = metric + w[idx] * fxn
        accept = ~(snew >= sorig)
        newcol = (cnt[idx] > ninst0) & (gxn < 0)
        update_oracle[idx[newcol]] = False
        accept &= ~newcol
        ia, ir = idx[accept], idx[~accept]
        R[ia], t[ia] = Rn[accept], tn[accept]
        res.fx[ia], res.gx[ia] = fxn[accept], gxn[accept]
        trs[ia] *= 2.5
        res.gx[ir] = gx[~accept]
        trs[ir] *= 0.5
        shrink = trs[ir] > dxnorm[~accept]
        trs[ir[shrink]] = np.max(np.abs(dx[~accept][shrink]), axis=1) * 0.5
        if keep_trace:
            for p in idx:
                res.trace[p].append((so3.from_matrix(R[p]), list(t[p])))
        scale = 0.75 * ((np.arctan(res.gx[idx] * 3) * 2 / math.pi + 0.5) * 1.75 + 0.25)
        w[idx] = np.maximum(w[idx] * scale, settings.minimum_objective_score_weight)
        res.outer_iterations += len(idx)
    T = np.empty((B, 3, 4))
    T[:, :, :3] = R
    T[:, :, 3] = t
    res.T = T.reshape(B, 12)
    res.status = ['local optimum' if s else 'not converged' for s in status]
    res.instantiated_params = [[list(P[p, k]) for k in range(cnt[p])] for p in range(B)]
    res.time = time.time() - tstart
    return res


def optimizeCollFreeBatchDevice(obj, env, Tinits, Tdess=None, settings=None, context=None, flags=None, cap=48, _out=None):
    """The lock-step algorithm with ALL per-problem state on the device (sio_lockstep_pose_solve, csrc/sio_lockstep.cu):
    the host sequences kernel launches and reads two counters per outer iteration.  Same iterations per problem as
    optimizeCollFreeBatchFast (tests compare them); `cap` = parameter slots per problem (parameters beyond it are
    counted in result.slot_overflows and not instantiated)."""
    from ._host import se3, so3
    assert isinstance(obj, PenetrationDepthGeometry) and isinstance(env, PenetrationDepthGeometry)
    ctx = context or obj.context
    settings = settings or _sip.SemiInfiniteOptimizationSettings()
    if flags is None:
        flags = BULK_FLAGS | _abi.SIO_PRUNE
    B = len(Tinits)
    if Tdess is None:
        Tdess = Tinits
    def rowmajor(Ts):            # Klampt se3 (R column-major 9-list, t) -> row-major 3x4, vectorised
        if isinstance(Ts, np.ndarray):                      # already row-major 3x4 per problem: (B, 12) or (B, 3, 4)
            return np.ascontiguousarray(Ts, dtype=np.float64).reshape(len(Ts), 12)
        Rc = np.array([T[0] for T in Ts], dtype=np.float64).reshape(-1, 3, 3)          # [b][col][row]
        out = np.empty((len(Ts), 3, 4))
        out[:, :, :3] = np.transpose(Rc, (0, 2, 1))
        out[:, :, 3] = np.array([T[1] for T in Ts], dtype=np.float64)
        return out.reshape(-1, 12)
    Ti = rowmajor(Tinits)
    Td = Ti if Tdess is Tinits else rowmajor(Tdess)
    # the environment cloud in world coordinates: 3 ms of numpy for 262,144 points, the same array for every call with an
    # unmoved environment -- kept on the geometry, keyed by its pose
    key = env._T12.tobytes()
    cached = getattr(env, "_cloud_world_cache", None)
    if cached is None or cached[0] != key or cached[1] is not env._dev.cloud32:
        Tenv = env._T12.reshape(3, 4)
        cached = (key, env._dev.cloud32, np.ascontiguousarray(env._dev.cloud32.astype(np.float64) @ Tenv[:, :3].T + Tenv[:, 3]))
        env._cloud_world_cache = cached
    cloud_world = cached[2]
    t0 = time.time()
    out = ctx.lockstep_pose_solve(obj._dev.grid_h, env._dev.cloud_h, env._T12, cloud_world, Ti, Td, settings, cap=cap, sweep_flags=flags, out=_out)
    res = BatchResult(B, per_problem_lists=False)
    res.time = time.time() - t0
    res.T = out["T"]
    res.status_codes = out["status"]
    res.num_iterations = out["iterations"].astype(np.int64)
    res.fx, res.gx, res.gx0 = out["fx"], out["gx"], out["gx0"]
    res.set_params_arrays(out["params"], out["n_params"])
    res.time_qp, res.time_cons = out["stats"]["qp_device_s"], out["stats"]["sweep_device_s"]
    res.outer_iterations = out["stats"]["outer_iterations"]
    res.point_queries = out["stats"]["point_queries"]
    res.slot_overflows = out["stats"]["slot_overflows"]
    res.qp_failures = out["stats"]["qp_failures"]          # problems stopped with status 2 ('qp failed')
    return res


class GroupedPoseSolver:
    """optimizeCollFreeBatchDevice over `groups` concurrent slices of the batch: one device context (stream, scratch,
    resident copy of the two geometries) and one host thread per slice, all on the same GPU.

    Why: a round of the lock-step solver is a chain of dependent launches; with few problems per GPU (8,192 = one rank of
    an 8-GPU C5 run) every QP launch lasts as long as its slowest problem and the small bookkeeping launches run one
    after the other on a mostly idle device.  Independent slices have no such dependence on each other, so their chains
    interleave on the SMs.  The problems are independent (each keeps its own iteration index, no quantity is shared
    between problems), so the answers are those of the single-context solve, problem by problem (tests).

    obj / env: Geometry3D (as PenetrationDepthGeometry takes them) -- obj converted to an SDF at `gridres`, env to a
    cloud at `pcres`; env_T: the environment's pose (Klampt se3), default as loaded."""

    def __init__(self, obj, gridres, env, pcres=0.02, groups=2, context=None, env_T=None):
        from ._host.geometry import Geometry3D, PointCloud
        first = context or _abi.default_context()
        device = first.device
        if isinstance(env, np.ndarray):
            env = Geometry3D(PointCloud(env))
        self.groups = max(1, int(groups))
        self.ctxs, self.objs, self.envs = [], [], []
        for k in range(self.groups):
            ctx = first if k == 0 else _abi.Context(device)
            self.ctxs.append(ctx)
            self.objs.append(PenetrationDepthGeometry(obj, gridres, None, context=ctx))
            e = PenetrationDepthGeometry(env, None, pcres, context=ctx)
            if env_T is not None:
                e.setTransform(env_T)
            self.envs.append(e)
        # constructing a PenetrationDepthGeometry points the host substrate's mesh converters at ITS context; leave them on
        # the caller's, which outlives close()
        from . import geometryopt as _G
        if self.groups > 1 and _G.DEVICE_CONVERSION:
            _G._hostgeometry.set_device_converters(first.mesh_sdf_grid, first.mesh_surface_sample)

    def solve(self, Tinits, Tdess=None, settings=None, cap=48):
        import threading
        B = len(Tinits)
        k = min(self.groups, max(1, B))
        if k == 1:
            return optimizeCollFreeBatchDevice(self.objs[0], self.envs[0], Tinits, Tdess, settings=settings, context=self.ctxs[0], cap=cap)
        cuts = [(B * g) // k for g in range(k + 1)]
        parts, errs = [None] * k, [None] * k
        whole = _abi.lockstep_outputs(B, cap)                 # every slice fills its rows of the same arrays: nothing is merged afterwards

        def run(g):
            try:
                lo, hi = cuts[g], cuts[g + 1]
                parts[g] = optimizeCollFreeBatchDevice(self.objs[g], self.envs[g], Tinits[lo:hi], None if Tdess is None else Tdess[lo:hi],
                                                       settings=settings, context=self.ctxs[g], cap=cap,
                                                       _out={n: (a[lo:hi] if a is not None else None) for n, a in whole.items()})
            except BaseException as e:                        # re-raised on the calling thread
                errs[g] = e
        t0 = time.time()
        threads = [threading.Thread(target=run, args=(g,)) for g in range(1, k)]
        for t in threads:
            t.start()
        run(0)
        for t in threads:
            t.join()
        for e in errs:
            if e is not None:
                raise e
        res = BatchResult(B, per_problem_lists=False)
        res.time = time.time() - t0
        res.T, res.status_codes = whole["T"], whole["status"]
        res.num_iterations = whole["iterations"].astype(np.int64)
        res.fx, res.gx, res.gx0 = whole["fx"], whole["gx"], whole["gx0"]
        res.set_params_arrays(whole["params"], whole["n_params"])
        # device seconds of the slices overlap in time: the per-slice maxima are what a wall clock can see
        res.time_qp, res.time_cons = max(p.time_qp for p in parts), max(p.time_cons for p in parts)
        res.outer_iterations = sum(p.outer_iterations for p in parts)
        res.point_queries = sum(p.point_queries for p in parts)
        res.slot_overflows = sum(p.slot_overflows for p in parts)
        res.qp_failures = sum(p.qp_failures for p in parts)
        return res

    def close(self):
        for ctx in self.ctxs[1:]:
            ctx.close()
        self.ctxs, self.objs, self.envs = [], [], []
