"""FROZEN QP SOLVER of the golden generator (oracle/frozen, round 2) -- TEST INFRASTRUCTURE.

What it is: the interpreted (numpy / scipy only) ADMM + polish of semiinfiniteoptimization_b200/_host/qp.py as of commit
92f6a08, i.e. a restatement of the published OSQP algorithm (Stellato et al., "OSQP: an operator splitting solver for
quadratic programs"): KKT system [[P + sigma I, A'], [A, -1/rho]], over-relaxation alpha = 1.6, rho = 0.1 adapted from
the residual ratio, sigma = 1e-6, residual test every 10 iterations, primal-infeasibility certificate, max_iter 4000.
oracle/refrun/osqp_shim.py answers the reference's `osqp.OSQP().setup/solve` calls (sip.py:302, 327-331, 355-357) with
THIS file, so a change to the product's solvers (host numpy, native hostqp.c, device k_sip_qp) can no longer rewrite
tests/golden/; the product's solvers are tested AGAINST it at a stated tolerance (tests/test_frozen_cpu.py).

Named deviations from what the reference actually runs (OSQP's defaults with eps_abs = 1e-5, sip.py:327), all switchable:
    EPS_REL = 1e-5   OSQP default 1e-3
    POLISH  = True   OSQP default False
    no Ruiz equilibration (OSQP default: 10 scaling iterations); rho adapts on a fixed 50-iteration schedule (OSQP's
    default adapts on a fraction of its set-up TIME, which makes its iterates machine dependent)
The QPs of this path are strictly convex (P = 2 (W + reg I) > 0), so they have ONE minimiser; OSQP at eps_rel = 1e-3
returns a point within its tolerance of that minimiser and this solver returns the minimiser itself (to ~1e-10 after
polish).  Real OSQP is absent from this image: "parity unpinned" for its iterates; tests/test_frozen_cpu.py compares with
it automatically when `import osqp` works, and against an independent active-set solver (`exact_qp` below) always.
"""
import itertools

import numpy as np
import scipy.linalg
import scipy.sparse
import scipy.sparse.linalg

EPS_REL = 1e-5
POLISH = True
INF = 1e20
DENSE_MAX_VARS = 64          # callers factor the KKT system densely up to this many variables, sparsely above


class QPResult:
    def __init__(self, x, y, status, iters, pri_res, dua_res):
        self.x = x
        self.y = y
        self.status = status          # 'solved' | 'primal infeasible' | 'max iterations'
        self.iters = iters
        self.pri_res = pri_res
        self.dua_res = dua_res


def _kkt_solver(P, A, sigma, rho_vec, sparse):
    n = P.shape[0]
    if sparse:
        # K = [[P + sigma I, A'], [A, -diag(1 / rho)]] assembled from its triplets (bmat builds the same matrix through
        # several intermediate matrices and takes four times as long; entry values and pattern are identical)
        m = A.shape[0]
        Pc, Ac = P.tocoo(), A.tocoo()
        dn, dm = np.arange(n), np.arange(n, n + m)
        rows = np.concatenate([Pc.row, dn, Ac.col, Ac.row + n, dm])
        cols = np.concatenate([Pc.col, dn, Ac.row + n, Ac.col, dm])
        data = np.concatenate([Pc.data, np.full(n, sigma), Ac.data, Ac.data, -1.0 / rho_vec])
        K = scipy.sparse.csc_matrix((data, (rows, cols)), shape=(n + m, n + m))
        lu = scipy.sparse.linalg.splu(K)
        return lu.solve
    m = A.shape[0]
    K = np.zeros((n + m, n + m))
    K[:n, :n] = P + sigma * np.eye(n)
    K[:n, n:] = A.T
    K[n:, :n] = A
    K[n:, n:] = -np.diag(1.0 / rho_vec)
    lu, piv = scipy.linalg.lu_factor(K)
    getrs = scipy.linalg.lapack.dgetrs                     # what lu_solve calls, without its 15 us of argument checking per call
    return lambda b: getrs(lu, piv, b)[0]


def solve_qp(P, q, A, l, u, eps_abs=1e-5, eps_rel=None, max_iter=4000, rho=0.1, sigma=1e-6, alpha=1.6,
             polish=None, check_every=10, adapt_every=50):
    eps_rel = EPS_REL if eps_rel is None else eps_rel
    polish = POLISH if polish is None else polish
    sparse = scipy.sparse.issparse(P) or scipy.sparse.issparse(A)
    if sparse:
        P = scipy.sparse.csc_matrix(P)
        A = scipy.sparse.csc_matrix(A)
    else:
        P = np.asarray(P, dtype=np.float64)
        A = np.asarray(A, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    l = np.maximum(np.asarray(l, dtype=np.float64), -INF)
    u = np.minimum(np.asarray(u, dtype=np.float64), INF)
    n, m = len(q), len(l)
    eq = (u - l) < 1e-12

    def rho_vector(r):
        v = np.full(m, r)
        v[eq] = 1e3 * r
        return v

    rho_vec = rho_vector(rho)
    solve = _kkt_solver(P, A, sigma, rho_vec, sparse)
    At = A.T                                                # built once: the residual tests multiply by it every ten iterations
    x = np.zeros(n)
    z = np.zeros(m)
    y = np.zeros(m)
    status = "max iterations"
    r_prim = r_dual = np.inf
    it = 0
    # The loop below is the textbook iteration
    #     rhs = [sigma x - q ; z - y / rho];  [xt ; nu] = K^-1 rhs;  zt = z + (nu - y) / rho
    #     x <- alpha xt + (1 - alpha) x;  zr = alpha zt + (1 - alpha) z;  z <- clip(zr + y / rho, l, u);  y <- y + rho (zr - z)
    # written with preallocated buffers (same operations in the same order, so the same bits; a third fewer
    # temporaries and ufunc calls, which is what a 455-variable problem's iteration costs besides the KKT solve)
    rhs = np.empty(n + m)
    yr, zt, zr, t1, t2 = np.empty(m), np.empty(m), np.empty(m), np.empty(m), np.empty(m)
    xn, tx = np.empty(n), np.empty(n)
    zn, yn = np.empty(m), np.empty(m)
    one_minus_alpha = 1 - alpha
    dy = np.zeros(m)
    for it in range(1, max_iter + 1):
        np.multiply(sigma, x, out=tx)
        np.subtract(tx, q, out=rhs[:n])
        np.divide(y, rho_vec, out=yr)                       # y / rho, used twice
        np.subtract(z, yr, out=rhs[n:])
        sol = solve(rhs)
        xt, nu = sol[:n], sol[n:]
        np.subtract(nu, y, out=t1)
        np.divide(t1, rho_vec, out=t1)
        np.add(z, t1, out=zt)
        np.multiply(alpha, xt, out=xn)
        np.multiply(one_minus_alpha, x, out=tx)
        np.add(xn, tx, out=xn)
        np.multiply(alpha, zt, out=zr)
        np.multiply(one_minus_alpha, z, out=t2)
        np.add(zr, t2, out=zr)
        np.add(zr, yr, out=zn)
        np.maximum(zn, l, out=zn)
        np.minimum(zn, u, out=zn)
        np.subtract(zr, zn, out=t1)
        np.multiply(rho_vec, t1, out=t1)
        np.add(y, t1, out=yn)
        checking = not (it % check_every and it != max_iter)
        if checking:
            dy = yn - y
        x, xn = xn, x
        z, zn = zn, z
        y, yn = yn, y
        if it % check_every and it != max_iter:
            continue
        Ax = A @ x
        Px = P @ x
        Aty = At @ y
        r_prim = np.abs(Ax - z).max() if m else 0.0
        r_dual = np.abs(Px + q + Aty).max()
        nAx = max(np.abs(Ax).max() if m else 0.0, np.abs(z).max() if m else 0.0)
        nd = max(np.abs(Px).max(), np.abs(Aty).max() if m else 0.0, np.abs(q).max())
        if r_prim <= eps_abs + eps_rel * nAx and r_dual <= eps_abs + eps_rel * nd:
            status = "solved"
            break
        # primal infeasibility certificate (OSQP paper, sec. 3.4)
        ndy = np.abs(dy).max() if m else 0.0
        if ndy > 1e-12:
            dyn = dy / ndy
            if np.abs(At @ dyn).max() <= 1e-4 and (u @ np.maximum(dyn, 0) + l @ np.minimum(dyn, 0)) < -1e-4:
                status = "primal infeasible"
                break
        if it % adapt_every == 0 and r_prim > 0 and r_dual > 0:
            scale = np.sqrt((r_prim / max(nAx, 1e-12)) / (r_dual / max(nd, 1e-12)))
            if scale > 5 or scale < 0.2:
                rho = float(np.clip(rho * scale, 1e-6, 1e6))
                rho_vec = rho_vector(rho)
                solve = _kkt_solver(P, A, sigma, rho_vec, sparse)
    if status == "solved" and polish and m:
        xp = _polish(P, q, A, l, u, x, y, sparse)
        if xp is not None:
            x = xp
    return QPResult(x if status != "primal infeasible" else None, y, status, it, r_prim, r_dual)


def _polish(P, q, A, l, u, x, y, sparse, delta=1e-10, tol=1e-7):
    """Solve the equality-constrained QP on the active set the ADMM iterate identifies; keep the
    result only if it is feasible and its multipliers have the signs of a minimiser (then it IS the minimiser)."""
    Ax = A @ x
    lo = (y < -1e-9) | (Ax - l < 1e-7)
    up = (y > 1e-9) | (u - Ax < 1e-7)
    lo &= l > -INF / 2
    up &= u < INF / 2
    act = np.nonzero(lo | up)[0]          # may be empty: the candidate is then the unconstrained minimiser
    rhs_c = np.where(lo[act], l[act], u[act])
    n = len(q)
    try:
        if sparse:
            Aa = A[act]
            na = len(act)
            Pc, Ac = P.tocoo(), Aa.tocoo()                  # same matrix as bmat([[P + delta I, Aa'], [Aa, -delta I]]), from triplets
            dn, da = np.arange(n), np.arange(n, n + na)
            rows = np.concatenate([Pc.row, dn, Ac.col, Ac.row + n, da])
            cols = np.concatenate([Pc.col, dn, Ac.row + n, Ac.col, da])
            data = np.concatenate([Pc.data, np.full(n, delta), Ac.data, Ac.data, np.full(na, -delta)])
            K = scipy.sparse.csc_matrix((data, (rows, cols)), shape=(n + na, n + na))
            sol = scipy.sparse.linalg.splu(K).solve(np.concatenate([-q, rhs_c]))
        else:
            Aa = A[act]
            K = np.block([[P + delta * np.eye(n), Aa.T], [Aa, -delta * np.eye(len(act))]])
            sol = np.linalg.solve(K, np.concatenate([-q, rhs_c]))
    except Exception:
        return None
    xp = sol[:n]
    lam = sol[n:]
    if not np.all(np.isfinite(sol)):
        return None
    Axp = A @ xp
    if np.any(Axp < l - tol) or np.any(Axp > u + tol):
        return None
    # KKT test instead of an objective comparison (the ADMM iterate is slightly infeasible, so its objective is LOWER than
    # the minimiser's and the comparison used to reject most polishes): stationarity holds by construction
    # (P xp + q + Aa' lam = 0), xp is feasible, so xp is the minimiser iff no multiplier has the wrong sign --
    # lam <= 0 on a row held at its lower bound, lam >= 0 at its upper bound
    lamtol = 1e-9 * (1.0 + (np.abs(lam).max() if len(lam) else 0.0))
    only_lo = lo[act] & ~up[act]
    only_up = up[act] & ~lo[act]
    if np.any(lam[only_lo] > lamtol) or np.any(lam[only_up] < -lamtol):
        return None
    return xp


def exact_qp(P, q, A, l, u, tol=1e-9):
    """INDEPENDENT check solver for small dense strictly convex QPs (n <= 8, m <= 14): enumerate active sets in order of
    size, solve each equality-constrained problem by its KKT system and return the first point that is primal feasible
    with multipliers of the right sign -- the unique minimiser, with no iteration and no tolerance to tune.
    Returns x, or None if the problem is infeasible.  Exponential in m: tests only."""
    P = np.asarray(P, dtype=np.float64)
    A = np.asarray(A, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    l = np.maximum(np.asarray(l, dtype=np.float64), -INF)
    u = np.minimum(np.asarray(u, dtype=np.float64), INF)
    n, m = len(q), len(l)
    sides = []
    for i in range(m):
        s = []
        if l[i] > -INF / 2:
            s.append(-1)
        if u[i] < INF / 2:
            s.append(+1)
        sides.append(s)
    for k in range(0, min(m, n) + 1):
        for rows in itertools.combinations(range(m), k):
            for sg in itertools.product(*[sides[i] for i in rows]):
                rows_a = list(rows)
                if k:
                    Aa = A[rows_a]
                    rhs = np.array([l[i] if s < 0 else u[i] for i, s in zip(rows_a, sg)])
                    K = np.block([[P, Aa.T], [Aa, np.zeros((k, k))]])
                    try:
                        sol = np.linalg.solve(K, np.concatenate([-q, rhs]))
                    except np.linalg.LinAlgError:
                        continue
                    x, lam = sol[:n], sol[n:]
                    # stationarity P x + q + Aa' lam = 0: lam <= 0 on lower bounds, >= 0 on upper bounds
                    if any((s < 0 and v > tol) or (s > 0 and v < -tol) for s, v in zip(sg, lam)):
                        continue
                else:
                    x = np.linalg.solve(P, -q)
                Ax = A @ x
                if np.all(Ax >= l - tol) and np.all(Ax <= u + tol):
                    return x
    return None
