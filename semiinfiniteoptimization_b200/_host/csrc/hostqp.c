/* hostqp.c -- batched host solver for the small QP of the SIP step.
 *
 * The reference solves one QP per outer iteration with OSQP (sip.py:294-378):
 *
 *     min (x - xdes)' W (x - xdes) + reg |x|^2   s.t.  A x + b >= inflation,  xmin <= x <= xmax
 *
 * and, when the strict problem is infeasible or its solution is out of scale, a second QP with one
 * slack variable per row (sip.py:332-357).  With 65,536 problems in lock step (BASELINE.json configs[4])
 * that is 65,536 QPs per outer iteration, so they are solved here natively, one OpenMP thread per problem.
 *
 * Algorithm = the one in _host/qp.py (OSQP's ADMM: over-relaxation alpha, adaptive rho, the same residual
 * test every 10 iterations, the primal-infeasibility certificate, the active-set polish), restated for
 * dense problems with a DIAGONAL quadratic term.  Because P is diagonal the KKT step is taken in its reduced
 * form  (P + sigma I + A' diag(rho) A) x~ = sigma x - q + A'(rho z - y),  z~ = A x~  -- an n x n Cholesky
 * instead of an (n+m) x (n+m) LU; same iterates up to rounding.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define HQP_INF 1e20
enum { HQP_SOLVED = 0, HQP_MAXIT = 1, HQP_INFEASIBLE = 2 };

typedef struct {
    double eps_abs, eps_rel, rho, sigma, alpha;
    int max_iter, check_every, adapt_every, polish;
} hqp_opts;

/* in-place Cholesky of the n x n SPD matrix K (row-major, lower part used); returns 0 on success */
static int chol(double* K, int n) {
    for (int j = 0; j < n; ++j) {
        double d = K[j * n + j];
        for (int k = 0; k < j; ++k) d -= K[j * n + k] * K[j * n + k];
        if (!(d > 0.0)) return -1;
        d = sqrt(d);
        K[j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = K[i * n + j];
            for (int k = 0; k < j; ++k) s -= K[i * n + k] * K[j * n + k];
            K[i * n + j] = s / d;
        }
    }
    return 0;
}
static void chol_solve(const double* L, int n, double* b) {
    for (int i = 0; i < n; ++i) {
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= L[i * n + k] * b[k];
        b[i] = s / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * b[k];
        b[i] = s / L[i * n + i];
    }
}
/* Gaussian elimination with partial pivoting, N x N row-major, solves M x = b in place; 0 on success */
static int lu_solve(double* M, double* b, int N) {
    for (int c = 0; c < N; ++c) {
        int p = c;
        double best = fabs(M[c * N + c]);
        for (int r = c + 1; r < N; ++r) if (fabs(M[r * N + c]) > best) { best = fabs(M[r * N + c]); p = r; }
        if (!(best > 0.0)) return -1;
        if (p != c) {
            for (int k = 0; k < N; ++k) { double t = M[c * N + k]; M[c * N + k] = M[p * N + k]; M[p * N + k] = t; }
            double t = b[c]; b[c] = b[p]; b[p] = t;
        }
        for (int r = c + 1; r < N; ++r) {
            double f = M[r * N + c] / M[c * N + c];
            if (f == 0.0) continue;
            for (int k = c; k < N; ++k) M[r * N + k] -= f * M[c * N + k];
            b[r] -= f * b[c];
        }
    }
    for (int r = N - 1; r >= 0; --r) {
        double s = b[r];
        for (int k = r + 1; k < N; ++k) s -= M[r * N + k] * b[k];
        b[r] = s / M[r * N + r];
    }
    return 0;
}

static void build_K(int n, int m, const double* Pd, const double* A, const double* rv, double sigma, double* K) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = (i == j) ? Pd[i] + sigma : 0.0;
            for (int r = 0; r < m; ++r) s += rv[r] * A[r * n + i] * A[r * n + j];
            K[i * n + j] = s;
        }
}

/* work: doubles, at least n*n + 8*m + 4*n + (n+m)*(n+m+1) */
static int admm(int n, int m, const double* Pd, const double* q, const double* A, const double* l0, const double* u0,
                const hqp_opts* o, double* x, double* pri, double* dua, int* iters_out, double* work, double* y_out) {
    double* K = work;               work += n * n;
    double* l = work;               work += m;
    double* u = work;               work += m;
    double* z = work;               work += m;
    double* y = work;               work += m;
    double* rv = work;              work += m;
    double* zt = work;              work += m;
    double* dy = work;              work += m;
    double* Ax = work;              work += m;
    double* xt = work;              work += n;
    double* Px = work;              work += n;
    double* Aty = work;             work += n;
    double* tmpn = work;            work += n;
    double* kkt = work;             /* (n+m)*(n+m+1) for the polish */
    double rho = o->rho;
    for (int r = 0; r < m; ++r) {
        l[r] = l0[r] > -HQP_INF ? l0[r] : -HQP_INF;
        u[r] = u0[r] < HQP_INF ? u0[r] : HQP_INF;
        rv[r] = (u[r] - l[r] < 1e-12) ? 1e3 * rho : rho;
        z[r] = 0.0; y[r] = 0.0;
    }
    for (int i = 0; i < n; ++i) x[i] = 0.0;
    build_K(n, m, Pd, A, rv, o->sigma, K);
    if (chol(K, n)) return HQP_MAXIT;
    int status = HQP_MAXIT, it = 0;
    double r_prim = INFINITY, r_dual = INFINITY;
    for (it = 1; it <= o->max_iter; ++it) {
        for (int i = 0; i < n; ++i) xt[i] = o->sigma * x[i] - q[i];
        for (int r = 0; r < m; ++r) {
            double w = rv[r] * z[r] - y[r];
            for (int i = 0; i < n; ++i) xt[i] += A[r * n + i] * w;
        }
        chol_solve(K, n, xt);
        double ndy = 0.0;
        for (int r = 0; r < m; ++r) {
            double s = 0.0;
            for (int i = 0; i < n; ++i) s += A[r * n + i] * xt[i];
            double zr = o->alpha * s + (1.0 - o->alpha) * z[r];
            double c = zr + y[r] / rv[r];
            double zn = c < l[r] ? l[r] : (c > u[r] ? u[r] : c);
            double yn = y[r] + rv[r] * (zr - zn);
            dy[r] = yn - y[r];
            if (fabs(dy[r]) > ndy) ndy = fabs(dy[r]);
            z[r] = zn; y[r] = yn;
        }
        for (int i = 0; i < n; ++i) x[i] = o->alpha * xt[i] + (1.0 - o->alpha) * x[i];
        if ((it % o->check_every) && it != o->max_iter) continue;
        double nAx = 0.0, nd = 0.0;
        r_prim = 0.0; r_dual = 0.0;
        for (int i = 0; i < n; ++i) { Px[i] = Pd[i] * x[i]; Aty[i] = 0.0; }
        for (int r = 0; r < m; ++r) {
            double s = 0.0;
            for (int i = 0; i < n; ++i) { s += A[r * n + i] * x[i]; Aty[i] += A[r * n + i] * y[r]; }
            Ax[r] = s;
            if (fabs(s - z[r]) > r_prim) r_prim = fabs(s - z[r]);
            if (fabs(s) > nAx) nAx = fabs(s);
            if (fabs(z[r]) > nAx) nAx = fabs(z[r]);
        }
        for (int i = 0; i < n; ++i) {
            double g = Px[i] + q[i] + Aty[i];
            if (fabs(g) > r_dual) r_dual = fabs(g);
            if (fabs(Px[i]) > nd) nd = fabs(Px[i]);
            if (fabs(Aty[i]) > nd) nd = fabs(Aty[i]);
            if (fabs(q[i]) > nd) nd = fabs(q[i]);
        }
        if (r_prim <= o->eps_abs + o->eps_rel * nAx && r_dual <= o->eps_abs + o->eps_rel * nd) { status = HQP_SOLVED; break; }
        if (ndy > 1e-12) {                       /* primal infeasibility certificate (OSQP paper, sec. 3.4) */
            double nAt = 0.0, sup = 0.0;
            for (int i = 0; i < n; ++i) tmpn[i] = 0.0;
            for (int r = 0; r < m; ++r) {
                double d = dy[r] / ndy;
                for (int i = 0; i < n; ++i) tmpn[i] += A[r * n + i] * d;
                sup += u[r] * (d > 0 ? d : 0.0) + l[r] * (d < 0 ? d : 0.0);
            }
            for (int i = 0; i < n; ++i) if (fabs(tmpn[i]) > nAt) nAt = fabs(tmpn[i]);
            if (nAt <= 1e-4 && sup < -1e-4) { status = HQP_INFEASIBLE; break; }
        }
        if (it % o->adapt_every == 0 && r_prim > 0 && r_dual > 0) {
            double scale = sqrt((r_prim / (nAx > 1e-12 ? nAx : 1e-12)) / (r_dual / (nd > 1e-12 ? nd : 1e-12)));
            if (scale > 5 || scale < 0.2) {
                rho *= scale;
                if (rho < 1e-6) rho = 1e-6;
                if (rho > 1e6) rho = 1e6;
                for (int r = 0; r < m; ++r) rv[r] = (u[r] - l[r] < 1e-12) ? 1e3 * rho : rho;
                build_K(n, m, Pd, A, rv, o->sigma, K);
                if (chol(K, n)) break;
            }
        }
    }
    if (it > o->max_iter) it = o->max_iter;
    if (status == HQP_SOLVED && o->polish && m > 0) {
        /* active set: rows at (or pushed against) a finite bound; equality-constrained solve on it */
        int na = 0;
        int* act = (int*)dy;                     /* dy is free now: reuse as the index list (m ints fit in m doubles) */
        double* rhs_c = zt;
        for (int r = 0; r < m; ++r) {
            int lo = ((y[r] < -1e-9) || (Ax[r] - l[r] < 1e-7)) && l[r] > -HQP_INF / 2;
            int up = ((y[r] > 1e-9) || (u[r] - Ax[r] < 1e-7)) && u[r] < HQP_INF / 2;
            if (lo || up) { act[na] = r; rhs_c[na] = lo ? l[r] : u[r]; ++na; }
        }
        {                                        /* na == 0: the candidate is the unconstrained minimiser */
            const double delta = 1e-10, tol = 1e-7;
            int N = n + na;
            double* M = kkt;
            double* b = kkt + (size_t)N * N;
            memset(M, 0, sizeof(double) * (size_t)N * N);
            for (int i = 0; i < n; ++i) { M[i * N + i] = Pd[i] + delta; b[i] = -q[i]; }
            for (int a = 0; a < na; ++a) {
                for (int i = 0; i < n; ++i) { M[i * N + n + a] = A[act[a] * n + i]; M[(n + a) * N + i] = A[act[a] * n + i]; }
                M[(n + a) * N + n + a] = -delta;
                b[n + a] = rhs_c[a];
            }
            if (lu_solve(M, b, N) == 0) {
                int ok = 1;
                for (int i = 0; i < n; ++i) if (!isfinite(b[i])) ok = 0;
                for (int r = 0; r < m && ok; ++r) {
                    double s = 0.0;
                    for (int i = 0; i < n; ++i) s += A[r * n + i] * b[i];
                    if (s < l[r] - tol || s > u[r] + tol) ok = 0;
                }
                /* KKT test (qp.py _polish): feasible + stationary by construction, so it is the minimiser iff no
                 * multiplier has the wrong sign (<= 0 at a lower bound, >= 0 at an upper bound) */
                double lmax = 0.0;
                for (int a = 0; a < na && ok; ++a) { if (!isfinite(b[n + a])) ok = 0; else if (fabs(b[n + a]) > lmax) lmax = fabs(b[n + a]); }
                const double lamtol = 1e-9 * (1.0 + lmax);
                for (int a = 0; a < na && ok; ++a) {
                    const int r = act[a];
                    const int lo = ((y[r] < -1e-9) || (Ax[r] - l[r] < 1e-7)) && l[r] > -HQP_INF / 2;
                    const int up = ((y[r] > 1e-9) || (u[r] - Ax[r] < 1e-7)) && u[r] < HQP_INF / 2;
                    if (lo && !up && b[n + a] > lamtol) ok = 0;
                    if (up && !lo && b[n + a] < -lamtol) ok = 0;
                }
                if (ok) for (int i = 0; i < n; ++i) x[i] = b[i];
            }
        }
    }
    if (y_out) memcpy(y_out, y, sizeof(double) * (size_t)m);
    *pri = r_prim; *dua = r_dual; *iters_out = it;
    return status;
}

int hqp_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* The SIP step of `nprob` problems with n variables each (sip.py:294-378, default solver settings):
 *   W, xdes, xmin, xmax : [nprob*n]     (diagonal weights; xmin/xmax = the trust-region box, may be NULL together)
 *   offs [nprob+1]      : row ranges into rows [total*n] (row-major) and depths [total]
 * For each problem:  P = W + reg,  q = -W xdes,  rows: A x >= -depth + inflation,  box rows xmin <= x <= xmax.
 * Problems with no rows return xdes (sip.py:288).  If the strict solve is infeasible, or |x| reg > |negative
 * depths|, the slack problem of sip.py:332-357 is solved instead (slack_penalty = 'auto').
 * status[p]: 0 strict solution, 1 slack solution, 2 no rows, -1 failure (dx = 0).
 * iters[p] (optional): ADMM iterations spent on the problem (strict + slack). */
int hqp_sip_step_batch(int64_t nprob, int n, const double* W, const double* xdes, const double* xmin, const double* xmax,
                       const int64_t* offs, const double* rows, const double* depths, double reg, double inflation,
                       double eps_abs, int max_iter, double* dx, int32_t* status, int32_t* iters) {
    if (nprob < 0 || n <= 0 || !W || !xdes || !offs || !dx) return -1;
    int64_t maxrows = 0;
    for (int64_t p = 0; p < nprob; ++p) if (offs[p + 1] - offs[p] > maxrows) maxrows = offs[p + 1] - offs[p];
    const int box = xmin && xmax;
    const int M = (int)maxrows + (box ? n : 0);          /* rows of the largest strict problem */
    const int NS = n + (int)maxrows;                     /* variables of the largest slack problem */
    const size_t wsz = (size_t)NS * NS + 8 * (size_t)M + 4 * (size_t)NS + (size_t)(NS + M) * (NS + M + 1)
                     + (size_t)M * NS + 2 * (size_t)M + 3 * (size_t)NS + 64;
    hqp_opts o = {eps_abs, 1e-5, 0.1, 1e-6, 1.6, max_iter > 0 ? max_iter : 4000, 10, 50, 1};
    int fail = 0;
#pragma omp parallel
    {
        double* buf = (double*)malloc(sizeof(double) * wsz);
        if (!buf) {
#pragma omp atomic write
            fail = 1;
        }
#pragma omp for schedule(dynamic, 16)
        for (int64_t p = 0; p < nprob; ++p) {
            if (!buf) continue;
            const int k = (int)(offs[p + 1] - offs[p]);
            const double* Wp = W + p * n;
            const double* xd = xdes + p * n;
            double* out = dx + p * n;
            if (k == 0) {
                for (int i = 0; i < n; ++i) out[i] = xd[i];
                if (status) status[p] = 2;
                if (iters) iters[p] = 0;
                continue;
            }
            const double* Ar = rows + offs[p] * n;
            const double* b = depths + offs[p];
            const int m = k + (box ? n : 0);
            double* A = buf;                        /* m x NS max */
            double* l = A + (size_t)M * NS;
            double* u = l + M;
            double* Pd = u + M;
            double* q = Pd + NS;
            double* x = q + NS;
            double* work = x + NS;
            for (int r = 0; r < k; ++r) {
                for (int i = 0; i < n; ++i) A[r * n + i] = Ar[r * n + i];
                l[r] = -b[r] + inflation; u[r] = INFINITY;
            }
            if (box)
                for (int i = 0; i < n; ++i) {
                    for (int j = 0; j < n; ++j) A[(k + i) * n + j] = (i == j) ? 1.0 : 0.0;
                    l[k + i] = xmin[p * n + i]; u[k + i] = xmax[p * n + i];
                }
            for (int i = 0; i < n; ++i) { Pd[i] = Wp[i] + reg; q[i] = -Wp[i] * xd[i]; }
            double pri, dua;
            int its;
            int st = admm(n, m, Pd, q, A, l, u, &o, x, &pri, &dua, &its, work, NULL);
            double bnorm = 1.0, neg2 = 0.0, bmin = INFINITY;
            int nneg = 0;
            for (int r = 0; r < k; ++r) { if (b[r] < 0) { neg2 += b[r] * b[r]; ++nneg; } if (b[r] < bmin) bmin = b[r]; }
            if (nneg) bnorm = sqrt(neg2);
            double xn = 0.0;
            for (int i = 0; i < n; ++i) xn += x[i] * x[i];
            xn = sqrt(xn);
            if (st != HQP_INFEASIBLE && !(xn * reg > bnorm)) {
                for (int i = 0; i < n; ++i) out[i] = x[i];
                if (status) status[p] = 0;
                if (iters) iters[p] = its;
                continue;
            }
            const int its_strict = its;
            /* slack problem: variables [x (n) | s (k)], rows A x + s >= ..., box rows on x only */
            const double pen = (1.0 + nneg) / (fabs(bmin) + 1e-3);
            const int ns = n + k;
            for (int r = 0; r < k; ++r) {
                for (int i = 0; i < n; ++i) A[r * ns + i] = Ar[r * n + i];
                for (int j = 0; j < k; ++j) A[r * ns + n + j] = (j == r) ? 1.0 : 0.0;
            }
            if (box)
                for (int i = 0; i < n; ++i)
                    for (int j = 0; j < ns; ++j) A[(k + i) * ns + j] = (i == j) ? 1.0 : 0.0;
            for (int j = 0; j < k; ++j) { Pd[n + j] = pen; q[n + j] = 0.0; }
            st = admm(ns, m, Pd, q, A, l, u, &o, x, &pri, &dua, &its, work, NULL);
            if (st == HQP_INFEASIBLE) {
                for (int i = 0; i < n; ++i) out[i] = 0.0;
                if (status) status[p] = -1;
            } else {
                for (int i = 0; i < n; ++i) out[i] = x[i];
                if (status) status[p] = 1;
            }
            if (iters) iters[p] = its_strict + its;
        }
        free(buf);
    }
    return fail ? -2 : 0;
}

/* ---------------------------------------------------------------------------------------------------------------
 * One LARGE sparse problem with a banded structure (the trajectory QP of BASELINE config 4: 455 variables, a
 * block-tridiagonal quadratic term, rows that touch two neighbouring milestones, a box on every variable):
 *
 *     min 1/2 x'Px + q'x (+ 1/2 pen |s|^2)   s.t.  lr <= R x (+ s) <= ur,   xmin <= x <= xmax
 *
 * Same ADMM as above (and as _host/qp.py solve_qp: same iterates up to rounding, same tests every 10 iterations, same
 * rho adaptation every 50) with the x-update in its reduced form: P + sigma I + R'diag(c)R + rho_box I is banded (half
 * bandwidth bw), so it is factorised by a banded Cholesky and solved in O(n bw) per iteration -- the interpreted loop
 * around a sparse LU of the (n+m) x (n+m) KKT system spends 40 us per iteration in the solve alone.  The slack variables
 * (one per general row, each in exactly one row) are eliminated in closed form as in the device kernel (sio_qp.cu).
 * P comes as its lower band: Pb[d*n + j] = P[j+d][j], d = 0..bw.  y has mr + (has_box ? n : 0) entries, general rows
 * first.  The polish stays with the caller.  status: 0 solved, 1 iteration limit, 2 primal infeasible, -1 not positive
 * definite. */
static int band_chol(double* M, int n, int bw) {              /* M[d*n + j] = M_(j+d, j); in place: L */
    for (int j = 0; j < n; ++j) {
        double d = M[j];
        const int k0 = j - bw > 0 ? j - bw : 0;
        for (int k = k0; k < j; ++k) d -= M[(j - k) * n + k] * M[(j - k) * n + k];
        if (!(d > 0.0)) return -1;
        d = sqrt(d);
        M[j] = d;
        const int i1 = j + bw < n - 1 ? j + bw : n - 1;
        for (int i = j + 1; i <= i1; ++i) {
            double s = M[(i - j) * n + j];
            const int kk0 = i - bw > 0 ? i - bw : 0;
            for (int k = kk0 > k0 ? kk0 : k0; k < j; ++k) s -= M[(i - k) * n + k] * M[(j - k) * n + k];
            M[(i - j) * n + j] = s / d;
        }
    }
    return 0;
}
static void band_solve(const double* L, int n, int bw, double* b) {
    for (int i = 0; i < n; ++i) {
        double s = b[i];
        const int k0 = i - bw > 0 ? i - bw : 0;
        for (int k = k0; k < i; ++k) s -= L[(i - k) * n + k] * b[k];
        b[i] = s / L[i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        const int k1 = i + bw < n - 1 ? i + bw : n - 1;
        for (int k = i + 1; k <= k1; ++k) s -= L[(k - i) * n + i] * b[k];
        b[i] = s / L[i];
    }
}

int hqp_admm_banded(int n, int bw, const double* Pb, const double* q, int mr, const int* rp, const int* ci, const double* rv,
                    const double* lr, const double* ur, int has_box, const double* xmin, const double* xmax, double pen,
                    double eps_abs, double eps_rel, int max_iter, double rho, double sigma, double alpha,
                    double* x, double* s, double* y, int* status_out, int* iters_out, double* rp_out, double* rd_out) {
    const int check_every = 10, adapt_every = 50;
    const int slack = pen > 0.0;
    const int m = mr + (has_box ? n : 0);
    double* W = (double*)calloc((size_t)(bw + 1) * n + 12 * (size_t)(m + n) + 16, sizeof(double));
    if (!W) return -2;
    double* M = W;                              /* (bw+1)*n */
    double* l = M + (size_t)(bw + 1) * n;       /* m each from here */
    double* u = l + m; double* rho_v = u + m; double* z = rho_v + m; double* wv = z + m; double* zt = wv + m;
    double* dy = zt + m; double* cr = dy + m; double* ds = cr + m;
    double* xt = ds + m;                        /* n each */
    double* tmp = xt + n; double* Aty = tmp + n;
    for (int r = 0; r < mr; ++r) { l[r] = lr[r] < -HQP_INF ? -HQP_INF : lr[r]; u[r] = ur[r] > HQP_INF ? HQP_INF : ur[r]; }
    for (int i = 0; i < (has_box ? n : 0); ++i) {
        l[mr + i] = xmin[i] < -HQP_INF ? -HQP_INF : xmin[i];
        u[mr + i] = xmax[i] > HQP_INF ? HQP_INF : xmax[i];
    }
    for (int i = 0; i < n; ++i) x[i] = 0.0;
    for (int r = 0; r < mr; ++r) if (s) s[r] = 0.0;
    for (int r = 0; r < m; ++r) { y[r] = 0.0; z[r] = 0.0; dy[r] = 0.0; }
    int status = HQP_MAXIT, it = 0, rc = 0;
    double r_prim = INFINITY, r_dual = INFINITY;
#define HQP_FACTOR()                                                                                                   \
    do {                                                                                                               \
        for (int r = 0; r < m; ++r) rho_v[r] = (u[r] - l[r] < 1e-12) ? 1e3 * rho : rho;                                \
        memcpy(M, Pb, sizeof(double) * (size_t)(bw + 1) * n);                                                         \
        for (int j = 0; j < n; ++j) M[j] += sigma + (has_box ? rho_v[mr + j] : 0.0);                                   \
        for (int r = 0; r < mr; ++r) {                                                                                 \
            ds[r] = pen + sigma + rho_v[r];                                                                            \
            cr[r] = slack ? rho_v[r] - rho_v[r] * rho_v[r] / ds[r] : rho_v[r];                                         \
            for (int a = rp[r]; a < rp[r + 1]; ++a)                                                                    \
                for (int b2 = rp[r]; b2 < rp[r + 1]; ++b2) {                                                           \
                    const int i = ci[a], j = ci[b2];                                                                   \
                    if (i >= j) M[(size_t)(i - j) * n + j] += cr[r] * rv[a] * rv[b2];                                   \
                }                                                                                                      \
        }                                                                                                              \
        rc = band_chol(M, n, bw);                                                                                     \
    } while (0)
    HQP_FACTOR();
    if (rc) { free(W); *status_out = -1; *iters_out = 0; return 0; }
    for (it = 1; it <= max_iter; ++it) {
        /* rhs of the reduced system */
        for (int i = 0; i < n; ++i) xt[i] = sigma * x[i] - q[i];
        for (int r = 0; r < m; ++r) wv[r] = rho_v[r] * z[r] - y[r];
        for (int r = 0; r < mr; ++r) {
            double ws = wv[r];
            if (slack) ws -= rho_v[r] * (sigma * s[r] + wv[r]) / ds[r];
            for (int a = rp[r]; a < rp[r + 1]; ++a) xt[ci[a]] += rv[a] * ws;
        }
        if (has_box) for (int i = 0; i < n; ++i) xt[i] += wv[mr + i];
        band_solve(M, n, bw, xt);
        /* z~ = A x~ (with the eliminated slack), relaxation, projection, dual step */
        for (int r = 0; r < m; ++r) {
            double ax;
            if (r < mr) {
                ax = 0.0;
                for (int a = rp[r]; a < rp[r + 1]; ++a) ax += rv[a] * xt[ci[a]];
            } else ax = xt[r - mr];
            double zs = ax, st = 0.0;
            if (slack && r < mr) { st = (sigma * s[r] + wv[r] - rho_v[r] * ax) / ds[r]; zs = ax + st; }
            const double zr = alpha * zs + (1 - alpha) * z[r];
            const double c = zr + y[r] / rho_v[r];
            const double zn = c < l[r] ? l[r] : (c > u[r] ? u[r] : c);
            const double yn = y[r] + rho_v[r] * (zr - zn);
            dy[r] = yn - y[r];
            z[r] = zn; y[r] = yn;
            if (slack && r < mr) s[r] = alpha * st + (1 - alpha) * s[r];
        }
        for (int i = 0; i < n; ++i) x[i] = alpha * xt[i] + (1 - alpha) * x[i];
        if (it % check_every && it != max_iter) continue;
        /* residuals */
        double nAx = 0.0, nd = 0.0;
        r_prim = 0.0; r_dual = 0.0;
        for (int i = 0; i < n; ++i) Aty[i] = has_box ? y[mr + i] : 0.0;
        for (int r = 0; r < m; ++r) {
            double ax;
            if (r < mr) {
                ax = 0.0;
                for (int a = rp[r]; a < rp[r + 1]; ++a) { ax += rv[a] * x[ci[a]]; Aty[ci[a]] += rv[a] * y[r]; }
                if (slack) {
                    ax += s[r];
                    const double ps = pen * s[r];
                    r_dual = fmax(r_dual, fabs(ps + y[r]));
                    nd = fmax(nd, fmax(fabs(ps), fabs(y[r])));
                }
            } else ax = x[r - mr];
            r_prim = fmax(r_prim, fabs(ax - z[r]));
            nAx = fmax(nAx, fmax(fabs(ax), fabs(z[r])));
        }
        for (int i = 0; i < n; ++i) {                       /* P x by the band */
            double px = Pb[i] * x[i];
            for (int d = 1; d <= bw; ++d) {
                if (i + d < n) px += Pb[(size_t)d * n + i] * x[i + d];
                if (i - d >= 0) px += Pb[(size_t)d * n + (i - d)] * x[i - d];
            }
            r_dual = fmax(r_dual, fabs(px + q[i] + Aty[i]));
            nd = fmax(nd, fmax(fabs(px), fmax(fabs(Aty[i]), fabs(q[i]))));
        }
        if (r_prim <= eps_abs + eps_rel * nAx && r_dual <= eps_abs + eps_rel * nd) { status = HQP_SOLVED; break; }
        double ndy = 0.0;
        for (int r = 0; r < m; ++r) ndy = fmax(ndy, fabs(dy[r]));
        if (ndy > 1e-12) {                                  /* primal infeasibility certificate */
            double sup = 0.0, nAt = 0.0;
            for (int i = 0; i < n; ++i) tmp[i] = has_box ? dy[mr + i] / ndy : 0.0;
            for (int r = 0; r < m; ++r) {
                const double d = dy[r] / ndy;
                if (r < mr) {
                    for (int a = rp[r]; a < rp[r + 1]; ++a) tmp[ci[a]] += rv[a] * d;
                    if (slack) nAt = fmax(nAt, fabs(d));
                }
                sup += u[r] * fmax(d, 0.0) + l[r] * fmin(d, 0.0);
            }
            for (int i = 0; i < n; ++i) nAt = fmax(nAt, fabs(tmp[i]));
            if (nAt <= 1e-4 && sup < -1e-4) { status = HQP_INFEASIBLE; break; }
        }
        if (it % adapt_every == 0 && r_prim > 0 && r_dual > 0) {
            const double scale = sqrt((r_prim / fmax(nAx, 1e-12)) / (r_dual / fmax(nd, 1e-12)));
            if (scale > 5 || scale < 0.2) {
                rho = fmin(fmax(rho * scale, 1e-6), 1e6);
                HQP_FACTOR();
                if (rc) { status = -1; break; }
            }
        }
    }
#undef HQP_FACTOR
    free(W);
    *status_out = status; *iters_out = it > max_iter ? max_iter : it; *rp_out = r_prim; *rd_out = r_dual;
    return 0;
}

/* One small dense problem with a DIAGONAL quadratic term, ADMM phase only (no polish): the interpreted solver's loop
 * (_host/qp.py solve_qp) natively.  A is m x n row-major.  The caller polishes (numpy), so whenever both loops identify
 * the same active set the polished answers are the same bits.  y_out[m]; returns the status (0 solved, 1 iteration limit,
 * 2 primal infeasible) or -2 on allocation failure. */
static int solve_diag(int n, int m, const double* Pd, const double* q, const double* A, const double* l, const double* u,
                      double eps_abs, double eps_rel, int max_iter, double rho, double sigma, double alpha, int polish,
                      double* x, double* y_out, int* iters_out, double* rp_out, double* rd_out) {
    hqp_opts o;
    o.eps_abs = eps_abs; o.eps_rel = eps_rel; o.rho = rho; o.sigma = sigma; o.alpha = alpha;
    o.max_iter = max_iter; o.check_every = 10; o.adapt_every = 50; o.polish = polish;
    const size_t nw = (size_t)n * n + 8 * (size_t)m + 4 * (size_t)n + (size_t)(n + m) * (n + m + 1) + 16;
    double* work = (double*)malloc(nw * sizeof(double));
    if (!work) return -2;
    int iters = 0;
    double pri = 0.0, dua = 0.0;
    const int st = admm(n, m, Pd, q, A, l, u, &o, x, &pri, &dua, &iters, work, y_out);
    free(work);
    *iters_out = iters; *rp_out = pri; *rd_out = dua;
    return st;
}

int hqp_admm_diag(int n, int m, const double* Pd, const double* q, const double* A, const double* l, const double* u,
                  double eps_abs, double eps_rel, int max_iter, double rho, double sigma, double alpha,
                  double* x, double* y_out, int* iters_out, double* rp_out, double* rd_out) {
    return solve_diag(n, m, Pd, q, A, l, u, eps_abs, eps_rel, max_iter, rho, sigma, alpha, 0, x, y_out, iters_out, rp_out, rd_out);
}

/* The same with the active-set polish done here too (the routine hqp_sip_step_batch's problems go through): one call per
 * QP of a single-problem SIP run instead of the ADMM call plus ~0.12 ms of numpy (block matrix, LAPACK solve, tests). */
int hqp_solve_diag(int n, int m, const double* Pd, const double* q, const double* A, const double* l, const double* u,
                   double eps_abs, double eps_rel, int max_iter, double rho, double sigma, double alpha, int polish,
                   double* x, double* y_out, int* iters_out, double* rp_out, double* rd_out) {
    return solve_diag(n, m, Pd, q, A, l, u, eps_abs, eps_rel, max_iter, rho, sigma, alpha, polish, x, y_out, iters_out, rp_out, rd_out);
}
