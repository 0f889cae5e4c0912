"""Host QP solver for the SIP step (the role OSQP plays for the reference, sip.py:294-378).

OSQP is not available to this build, so its published algorithm (Stellato et al., "OSQP: an
operator splitting solver for quadratic programs") is restated here: ADMM on

    min 1/2 x'Px + q'x   s.t.  l <= Ax <= u

with the indirect KKT system [[P + sigma I, A'], [A, -1/rho]], over-relaxation alpha, adaptive
rho, the standard primal/dual residual test and the primal-infeasibility certificate, followed by
an optional polish (equality-constrained solve on the identified active set).  The QP is small
(n = 6 / 7 / 455 variables) and stays on the host per BASELINE.json's north_star.

Iteration limit 4000 = OSQP's default `max_iter` (the reference passes only `eps_abs=1e-5`, sip.py:327); a
problem that reaches it returns the last iterate, as OSQP does.

`solve_qp` handles one problem (dense or sparse); `solve_qp_batch` runs many small dense problems
of identical shape in lock-step with numpy (the C5 batch).
"""
import numpy as np
import scipy.linalg
import scipy.sparse
import scipy.sparse.linalg

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


def solve_qp(P, q, A, l, u, eps_abs=1e-5, eps_rel=1e-5, max_iter=4000, rho=0.1, sigma=1e-6, alpha=1.6,
             polish=True, check_every=10, adapt_every=50):
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
    if NATIVE_DENSE and not sparse and m and check_every == 10 and adapt_every == 50:
        d = np.diagonal(P)
        if not (P - np.diag(d)).any():                      # diagonal quadratic term: the native loop applies
            res = _solve_qp_native_diag(P, d, q, A, l, u, eps_abs, eps_rel, max_iter, rho, sigma, alpha, polish)
            if res is not None:
                return res
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


def solve_qp_batch(P, q, A, l, u, nrows, iters=400, rho=0.1, sigma=1e-6, alpha=1.6):
    """Lock-step ADMM for many small dense QPs of one padded shape.

    P: (B,n,n), q: (B,n), A: (B,m,n), l/u: (B,m); rows >= nrows[b] of problem b are padding and
    must be given as 0 <= 0*x <= 0-free rows (A row 0, l=-INF, u=INF).  Fixed iteration count;
    returns x (B,n) and the final primal/dual residuals."""
    P = np.asarray(P, dtype=np.float64)
    A = np.asarray(A, dtype=np.float64)
    B, m, n = A.shape
    l = np.maximum(l, -INF)
    u = np.minimum(u, INF)
    K = np.zeros((B, n + m, n + m))
    K[:, :n, :n] = P + sigma * np.eye(n)
    K[:, :n, n:] = np.transpose(A, (0, 2, 1))
    K[:, n:, :n] = A
    idx = np.arange(m)
    K[:, n + idx, n + idx] = -1.0 / rho
    Kinv = np.linalg.inv(K)
    x = np.zeros((B, n))
    z = np.zeros((B, m))
    y = np.zeros((B, m))
    for _ in range(iters):
        rhs = np.concatenate([sigma * x - q, z - y / rho], axis=1)
        sol = np.einsum('bij,bj->bi', Kinv, rhs)
        xt, nu = sol[:, :n], sol[:, n:]
        zt = z + (nu - y) / rho
        x = alpha * xt + (1 - alpha) * x
        zr = alpha * zt + (1 - alpha) * z
        z_new = np.clip(zr + y / rho, l, u)
        y = y + rho * (zr - z_new)
        z = z_new
    Ax = np.einsum('bij,bj->bi', A, x)
    r_prim = np.abs(Ax - z).max(axis=1)
    r_dual = np.abs(np.einsum('bij,bj->bi', P, x) + q + np.einsum('bji,bj->bi', A, y)).max(axis=1)
    return x, r_prim, r_dual


# ---------------------------------------------------------------------------------------------
# one large banded problem natively (csrc/hostqp.c hqp_admm_banded): the trajectory QP
# ---------------------------------------------------------------------------------------------
# The native loop takes the x-update in its reduced, banded form, so its ADMM iterates differ from solve_qp's in the last
# bits; the polish then lands both on the same active set and the same answer (the golden trajectory traces, compared at
# 1e-12, are unchanged -- tests/test_golden_cpu.py runs the reference through solve_qp and this package through the native
# loop).  A QP that stops at the iteration limit (no polish) can differ by rounding.  False = interpreted loop everywhere.
NATIVE_BANDED = True
# the same for small dense problems with a diagonal quadratic term (poses, configurations): ADMM natively, polish in numpy
NATIVE_DENSE = True
# ... and their polish too (hostqp.c's own Gaussian elimination instead of numpy's block matrix + LAPACK solve: the same
# rule and the same system, answers equal to rounding -- 1e-15 relative on the polished x; the golden traces, compared at
# 1e-8, do not move).  C1: 24 of 147 ms of eight optimizeCollFree runs were this numpy code.
NATIVE_POLISH = True
MAX_BANDWIDTH = 64


def _solve_qp_native_diag(P, d, q, A, l, u, eps_abs, eps_rel, max_iter, rho, sigma, alpha, polish):
    import ctypes as C
    n, m = len(q), len(l)
    Ac = np.ascontiguousarray(A, dtype=np.float64)
    dd, qq = np.ascontiguousarray(d, dtype=np.float64), np.ascontiguousarray(q, dtype=np.float64)
    ll, uu = np.ascontiguousarray(l), np.ascontiguousarray(u)
    x, y = np.zeros(n), np.zeros(m)
    iters = C.c_int(0)
    r_prim, r_dual = C.c_double(0.0), C.c_double(0.0)
    if NATIVE_POLISH:
        st = _native_lib().hqp_solve_diag(n, m, dd.ctypes.data, qq.ctypes.data, Ac.ctypes.data, ll.ctypes.data, uu.ctypes.data, eps_abs, eps_rel,
                                          int(max_iter), rho, sigma, alpha, 1 if polish else 0, x.ctypes.data, y.ctypes.data, C.byref(iters),
                                          C.byref(r_prim), C.byref(r_dual))
    else:
        st = _native_lib().hqp_admm_diag(n, m, dd.ctypes.data, qq.ctypes.data, Ac.ctypes.data, ll.ctypes.data, uu.ctypes.data, eps_abs, eps_rel,
                                         int(max_iter), rho, sigma, alpha, x.ctypes.data, y.ctypes.data, C.byref(iters), C.byref(r_prim),
                                         C.byref(r_dual))
    if st < 0:
        return None
    name = ("solved", "max iterations", "primal infeasible")[st]
    if name == "solved" and polish and not NATIVE_POLISH:
        xp = _polish(P, q, A, l, u, x, y, False)
        if xp is not None:
            x = xp
    return QPResult(x if name != "primal infeasible" else None, y, name, iters.value, r_prim.value, r_dual.value)


def solve_qp_structured(P, q, A, l, u, n, mr, has_box, pen, eps_abs=1e-5, eps_rel=1e-5, max_iter=4000, rho=0.1, sigma=1e-6,
                        alpha=1.6, polish=True):
    """The QP solve_qp(P, q, A, l, u) solves, for the structure sip.py builds: variables [x (n) | one slack per general row
    (pen > 0 only)], P = blockdiag(Px, pen I), A = [[R, I], [I_n, 0]] with mr general rows R and, if has_box, the identity
    rows of the box.  Returns a QPResult, or None when Px + R'R is not banded within MAX_BANDWIDTH (caller falls back)."""
    import ctypes as C
    P = scipy.sparse.csc_matrix(P)
    A = scipy.sparse.csr_matrix(A)
    Px = P[:n, :n].tocoo()
    R = A[:mr, :n].tocsr()
    R.sort_indices()
    bw = int(np.abs(Px.row - Px.col).max()) if Px.nnz else 0
    if mr:
        counts = np.diff(R.indptr)
        if (counts == 0).any():
            return None
        first = R.indices[R.indptr[:-1]]
        last = R.indices[R.indptr[1:] - 1]
        bw = max(bw, int((last - first).max()))
    if bw > MAX_BANDWIDTH:
        return None
    Pb = np.zeros((bw + 1, n))
    low = Px.row >= Px.col
    np.add.at(Pb, (Px.row[low] - Px.col[low], Px.col[low]), Px.data[low])
    q = np.asarray(q, dtype=np.float64)
    l = np.asarray(l, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    qx = np.ascontiguousarray(q[:n])
    lr, ur = np.ascontiguousarray(l[:mr]), np.ascontiguousarray(u[:mr])
    xmin = np.ascontiguousarray(l[mr:mr + n]) if has_box else None
    xmax = np.ascontiguousarray(u[mr:mr + n]) if has_box else None
    m = mr + (n if has_box else 0)
    x, sv, y = np.zeros(n), np.zeros(max(mr, 1)), np.zeros(max(m, 1))
    rp = np.ascontiguousarray(R.indptr, dtype=np.int32)
    ci = np.ascontiguousarray(R.indices, dtype=np.int32)
    rvals = np.ascontiguousarray(R.data, dtype=np.float64)
    status, iters = C.c_int(0), C.c_int(0)
    r_prim, r_dual = C.c_double(0.0), C.c_double(0.0)
    L = _native_lib()
    rc = L.hqp_admm_banded(n, bw, Pb.ctypes.data, qx.ctypes.data, mr, rp.ctypes.data, ci.ctypes.data, rvals.ctypes.data,
                           lr.ctypes.data, ur.ctypes.data, 1 if has_box else 0, None if xmin is None else xmin.ctypes.data,
                           None if xmax is None else xmax.ctypes.data, float(pen), eps_abs, eps_rel, int(max_iter), rho, sigma, alpha,
                           x.ctypes.data, sv.ctypes.data, y.ctypes.data, C.byref(status), C.byref(iters), C.byref(r_prim),
                           C.byref(r_dual))
    if rc != 0 or status.value < 0:
        return None
    xf = np.concatenate([x, sv[:mr]]) if pen > 0 else x
    yv = y[:m]
    name = ("solved", "max iterations", "primal infeasible")[status.value]
    if name == "solved" and polish and m:
        xp = _polish(P, q, scipy.sparse.csc_matrix(A), np.maximum(l, -INF), np.minimum(u, INF), xf, yv, True)
        if xp is not None:
            xf = xp
    return QPResult(xf if name != "primal infeasible" else None, yv, name, iters.value, r_prim.value, r_dual.value)


# ---------------------------------------------------------------------------------------------
# native batched SIP step (csrc/hostqp.c): every problem's QP, slack fallback included, in one call
# ---------------------------------------------------------------------------------------------
_native = None


def _native_lib():
    global _native
    if _native is None:
        import ctypes as C
        import os
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lib", "libsio_hostqp.so")
        if not os.path.exists(path):
            from .. import build
            build.build_hostqp()
        L = C.CDLL(path)
        P = C.c_void_p
        L.hqp_sip_step_batch.argtypes = [C.c_int64, C.c_int, P, P, P, P, P, P, P, C.c_double, C.c_double, C.c_double, C.c_int, P, P, P]
        L.hqp_sip_step_batch.restype = C.c_int
        L.hqp_max_threads.restype = C.c_int
        L.hqp_admm_banded.argtypes = [C.c_int, C.c_int, P, P, C.c_int, P, P, P, P, P, C.c_int, P, P, C.c_double, C.c_double, C.c_double,
                                      C.c_int, C.c_double, C.c_double, C.c_double, P, P, P, P, P, P, P]
        L.hqp_admm_banded.restype = C.c_int
        L.hqp_admm_diag.argtypes = [C.c_int, C.c_int, P, P, P, P, P, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double,
                                    P, P, P, P, P]
        L.hqp_admm_diag.restype = C.c_int
        L.hqp_solve_diag.argtypes = [C.c_int, C.c_int, P, P, P, P, P, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double,
                                     C.c_int, P, P, P, P, P]
        L.hqp_solve_diag.restype = C.c_int
        _native = L
    return _native


def sip_step_batch(W, xdes, offs, rows, depths, xmin=None, xmax=None, reg=1e-2, inflation=1e-3, eps_abs=1e-5, max_iter=4000,
                   return_iters=False):
    """dx of `len(W)` problems at once: min (x-xdes)'W(x-xdes) + reg|x|^2  s.t.  rows x + depths >= inflation,
    xmin <= x <= xmax, with the reference's slack re-solve (sip.py:294-378).  W, xdes, xmin, xmax: (B,n);
    offs: (B+1,) row ranges into rows (total,n) / depths (total,).  Returns dx (B,n), status (B,) int32
    (0 strict, 1 slack, 2 no rows -> xdes, -1 failed)."""
    W = np.ascontiguousarray(W, dtype=np.float64)
    xdes = np.ascontiguousarray(xdes, dtype=np.float64)
    B, n = W.shape
    offs = np.ascontiguousarray(offs, dtype=np.int64)
    rows = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, n)
    depths = np.ascontiguousarray(depths, dtype=np.float64)
    assert len(offs) == B + 1 and offs[-1] == len(rows) == len(depths)
    if xmin is not None:
        xmin = np.ascontiguousarray(xmin, dtype=np.float64)
        xmax = np.ascontiguousarray(xmax, dtype=np.float64)
    dx = np.empty((B, n))
    status = np.empty(B, dtype=np.int32)
    iters = np.empty(B, dtype=np.int32) if return_iters else None
    rc = _native_lib().hqp_sip_step_batch(B, n, W.ctypes.data, xdes.ctypes.data, None if xmin is None else xmin.ctypes.data,
                                          None if xmax is None else xmax.ctypes.data, offs.ctypes.data, rows.ctypes.data,
                                          depths.ctypes.data, float(reg), float(inflation), float(eps_abs), int(max_iter),
                                          dx.ctypes.data, status.ctypes.data, None if iters is None else iters.ctypes.data)
    if rc:
        raise RuntimeError("hqp_sip_step_batch failed with code %d" % rc)
    return (dx, status, iters) if return_iters else (dx, status)
