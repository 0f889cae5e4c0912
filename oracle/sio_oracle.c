/* sio_oracle.c -- CPU ORACLE for the collision-oracle hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in Klampt 0.8.x / KrisLibrary (un-vendored,
 * unpinned C++ dependency, README.md:54), absent from /root/reference and from this image; the
 * reference holds no golden vector, known-answer test or fixture for it.  This file
 * restates the published behaviour summarised in SURVEY.md Appendix A.2-A.4 and is pinned only
 * by closed-form known answers (analytic SDFs) in tests/.  Every Klampt-side choice is a named
 * constant below.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (semiinfiniteoptimization_b200/) never does.
 *
 * What is restated (reference call site -> function here):
 *   grid.distance_point(pt)           geometryopt.py:104-108,153   -> orc_query()
 *   pc.distance_ext(grid, upperBound) geometryopt.py:117-127       -> orc_min_distance(), orc_under()
 *   pose row [(R(pt-t)) x n, n]       geometryopt.py:413-418       -> orc_pose_rows()
 *   robot.setConfig + getTransform    geometryopt.py:201-204,294   -> orc_fk()
 *   Jp(local pt)^T n                  geometryopt.py:478-482       -> orc_chain_rows()
 *
 * Data: grid samples are float32 at CELL CENTRES, z fastest (geometryopt.py:25-28); cloud points
 * are float32 xyz; transforms are row-major 3x4 float64.  Arithmetic is float64 throughout
 * (orc_*_f32 variants exist only as the single-precision CPU timing baseline).
 *
 * Build: see oracle/Makefile  (-O3 -fopenmp -ffp-contract=off)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- named Klampt-side choices ------------------------------------ */
#define ORC_CELL_CENTRED   1   /* A.1: samples at bmin + (i+1/2)h                                  */
#define ORC_OUTSIDE_ADDS_BBOX_DISTANCE 1 /* A.2 step 2/4: d = v(clamp(p)) + |p - clamp(p)|        */

typedef struct {
    const float* v;       /* nx*ny*nz samples, z fastest */
    int32_t nx, ny, nz;
    double bmin[3], bmax[3];
} orc_grid;

int orc_version(void) { return 1; }
/* torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm of bench.py asks for every host
 * core explicitly.  n <= 0: all processors. */
void orc_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : omp_get_num_procs());
#else
    (void)n;
#endif
}
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* A.2: distance and (optional) gradient w.r.t. the grid-local point p.
 * gradient = masked trilinear derivative + outward unit vector of the bbox clamp. */
static double query_local(const orc_grid* g, const double p[3], double* grad /*3 or NULL*/) {
    const int dim[3] = {g->nx, g->ny, g->nz};
    double o[3], f[3], h[3];
    int i0[3], i1[3], interior[3];
    double dout2 = 0;
    for (int a = 0; a < 3; ++a) {
        h[a] = (g->bmax[a] - g->bmin[a]) / (double)dim[a];
        double c = p[a];
        if (c < g->bmin[a]) c = g->bmin[a];
        if (c > g->bmax[a]) c = g->bmax[a];
        o[a] = p[a] - c;
        dout2 += o[a] * o[a];
        double u = (c - g->bmin[a]) / h[a] - 0.5;
        interior[a] = (u > 0.0 && u < (double)(dim[a] - 1));
        double uc = u;
        if (uc < 0.0) uc = 0.0;
        if (uc > (double)(dim[a] - 1)) uc = (double)(dim[a] - 1);
        int i = (int)floor(uc);
        if (i > dim[a] - 1) i = dim[a] - 1;
        i0[a] = i;
        i1[a] = (i + 1 < dim[a]) ? i + 1 : dim[a] - 1;
        f[a] = uc - (double)i;
    }
    const size_t sy = (size_t)g->nz, sx = (size_t)g->ny * g->nz;
#define V(I, J, K) ((double)g->v[(size_t)(I) * sx + (size_t)(J) * sy + (size_t)(K)])
    double v000 = V(i0[0], i0[1], i0[2]), v001 = V(i0[0], i0[1], i1[2]);
    double v010 = V(i0[0], i1[1], i0[2]), v011 = V(i0[0], i1[1], i1[2]);
    double v100 = V(i1[0], i0[1], i0[2]), v101 = V(i1[0], i0[1], i1[2]);
    double v110 = V(i1[0], i1[1], i0[2]), v111 = V(i1[0], i1[1], i1[2]);
#undef V
    /* z, then y, then x */
    double dz00 = v001 - v000, dz01 = v011 - v010, dz10 = v101 - v100, dz11 = v111 - v110;
    double a00 = v000 + f[2] * dz00, a01 = v010 + f[2] * dz01;
    double a10 = v100 + f[2] * dz10, a11 = v110 + f[2] * dz11;
    double dy0 = a01 - a00, dy1 = a11 - a10;
    double b0 = a00 + f[1] * dy0, b1 = a10 + f[1] * dy1;
    double dx = b1 - b0;
    double val = b0 + f[0] * dx;
    double dout = sqrt(dout2);
    if (grad) {
        double gz0 = dz00 + f[1] * (dz01 - dz00), gz1 = dz10 + f[1] * (dz11 - dz10);
        double gf[3];
        gf[0] = dx;
        gf[1] = dy0 + f[0] * (dy1 - dy0);
        gf[2] = gz0 + f[0] * (gz1 - gz0);
        for (int a = 0; a < 3; ++a) {
            double ga = interior[a] ? gf[a] / h[a] : 0.0;
            if (dout > 0.0) ga += o[a] / dout;
            grad[a] = ga;
        }
    }
#if ORC_OUTSIDE_ADDS_BBOX_DISTANCE
    return val + dout;
#else
    return val;
#endif
}

static inline void xf_apply(const double* T, double x, double y, double z, double* p) {
    p[0] = T[0] * x + T[1] * y + T[2] * z + T[3];
    p[1] = T[4] * x + T[5] * y + T[6] * z + T[7];
    p[2] = T[8] * x + T[9] * y + T[10] * z + T[11];
}

/* grid.distance_point for P points given in an input frame; T = grid-local <- input frame.
 * grad (optional) is d(distance)/d(input point) = R^T * grad_local, normalised when `normalize`
 * (A.2 step 5; zero vector stays zero). */
void orc_query(const orc_grid* g, const double* T, const double* pts, int64_t P, int normalize,
               double* d, double* grad) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < P; ++i) {
        double p[3], gl[3];
        xf_apply(T, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], p);
        d[i] = query_local(g, p, grad ? gl : NULL);
        if (grad) {
            double gx = T[0] * gl[0] + T[4] * gl[1] + T[8] * gl[2];
            double gy = T[1] * gl[0] + T[5] * gl[1] + T[9] * gl[2];
            double gz = T[2] * gl[0] + T[6] * gl[1] + T[10] * gl[2];
            if (normalize) {
                double n = sqrt(gx * gx + gy * gy + gz * gz);
                if (n > 0.0) { gx /= n; gy /= n; gz /= n; }
            }
            grad[3 * i] = gx; grad[3 * i + 1] = gy; grad[3 * i + 2] = gz;
        }
    }
}

/* the same over a float32 cloud and B transforms, per-point outputs (B*N) */
void orc_cloud_query(const orc_grid* g, const float* cloud, int64_t N, const double* T, int64_t B,
                     int normalize, double* d, double* grad) {
#pragma omp parallel for schedule(static) collapse(2)
    for (int64_t b = 0; b < B; ++b)
        for (int64_t i = 0; i < N; ++i) {
            const double* Tb = T + 12 * b;
            double p[3], gl[3];
            xf_apply(Tb, (double)cloud[3 * i], (double)cloud[3 * i + 1], (double)cloud[3 * i + 2], p);
            d[b * N + i] = query_local(g, p, grad ? gl : NULL);
            if (grad) {
                double gx = Tb[0] * gl[0] + Tb[4] * gl[1] + Tb[8] * gl[2];
                double gy = Tb[1] * gl[0] + Tb[5] * gl[1] + Tb[9] * gl[2];
                double gz = Tb[2] * gl[0] + Tb[6] * gl[1] + Tb[10] * gl[2];
                if (normalize) {
                    double n = sqrt(gx * gx + gy * gy + gz * gz);
                    if (n > 0.0) { gx /= n; gy /= n; gz /= n; }
                }
                double* o = grad + 3 * (b * N + i);
                o[0] = gx; o[1] = gy; o[2] = gz;
            }
        }
}

/* pc.distance_ext(grid): brute-force min over the cloud (A.3), lowest index wins ties.
 * grids[grid_of_b[b]] selects the grid of transform b (NULL -> grid 0).
 * argmin = -1 when bound is finite and dmin >= bound (the wrapper's "[]", geometryopt.py:411). */
void orc_min_distance(const orc_grid* grids, const int32_t* grid_of_b, const float* cloud, int64_t N,
                      const double* T, const double* bound, int64_t B, double* dmin, int64_t* argmin) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < B; ++b) {
        const orc_grid* g = grids + (grid_of_b ? grid_of_b[b] : 0);
        const double* Tb = T + 12 * b;
        double best = INFINITY;
        int64_t bi = -1;
        for (int64_t i = 0; i < N; ++i) {
            double p[3];
            xf_apply(Tb, (double)cloud[3 * i], (double)cloud[3 * i + 1], (double)cloud[3 * i + 2], p);
            double d = query_local(g, p, NULL);
            if (d < best) { best = d; bi = i; }
        }
        dmin[b] = best;
        if (bound && !(best < bound[b])) bi = -1;
        argmin[b] = bi;
    }
}

/* single-transform variant parallelised over points (B small, N large) */
void orc_min_distance_1(const orc_grid* g, const float* cloud, int64_t N, const double* T,
                        double* dmin, int64_t* argmin) {
    double best = INFINITY;
    int64_t bi = -1;
#pragma omp parallel
    {
        double lb = INFINITY;
        int64_t li = -1;
#pragma omp for schedule(static) nowait
        for (int64_t i = 0; i < N; ++i) {
            double p[3];
            xf_apply(T, (double)cloud[3 * i], (double)cloud[3 * i + 1], (double)cloud[3 * i + 2], p);
            double d = query_local(g, p, NULL);
            if (d < lb) { lb = d; li = i; }
        }
#pragma omp critical
        {
            if (lb < best || (lb == best && li >= 0 && (bi < 0 || li < bi))) { best = lb; bi = li; }
        }
    }
    *dmin = best;
    *argmin = bi;
}

/* per transform: number of points with d < bound[b], the sum of their indices and the sum of their distances
 * (size-independent fingerprints of the under-bound set: the full-size tests compare these for every transform and
 * the explicit list for a few).  Any output may be NULL.  Returns the total count. */
int64_t orc_under_counts(const orc_grid* grids, const int32_t* grid_of_b, const float* cloud, int64_t N,
                         const double* T, const double* bound, int64_t B, int64_t* count, int64_t* idx_sum, double* d_sum) {
    int64_t total = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
    for (int64_t b = 0; b < B; ++b) {
        const orc_grid* g = grids + (grid_of_b ? grid_of_b[b] : 0);
        const double* Tb = T + 12 * b;
        int64_t n = 0, si = 0;
        double sd = 0.0;
        if (bound[b] < INFINITY)
            for (int64_t i = 0; i < N; ++i) {
                double p[3];
                xf_apply(Tb, (double)cloud[3 * i], (double)cloud[3 * i + 1], (double)cloud[3 * i + 2], p);
                double d = query_local(g, p, NULL);
                if (d < bound[b]) { ++n; si += i; sd += d; }
            }
        if (count) count[b] = n;
        if (idx_sum) idx_sum[b] = si;
        if (d_sum) d_sum[b] = sd;
        total += n;
    }
    return total;
}

/* all (b, i) with d < bound[b], in (b, i) order.  Returns the total count; writes at most cap.
 * Two passes (count per transform, then fill at the prefix offsets) so that both run on all threads. */
int64_t orc_under(const orc_grid* grids, const int32_t* grid_of_b, const float* cloud, int64_t N,
                  const double* T, const double* bound, int64_t B, int64_t cap,
                  int64_t* out_b, int64_t* out_i, double* out_d) {
    int64_t* off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(B + 1));
    if (!off) return -1;
    orc_under_counts(grids, grid_of_b, cloud, N, T, bound, B, off + 1, NULL, NULL);
    off[0] = 0;
    for (int64_t b = 0; b < B; ++b) off[b + 1] += off[b];
    const int64_t total = off[B];
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < B; ++b) {
        if (off[b + 1] == off[b]) continue;
        const orc_grid* g = grids + (grid_of_b ? grid_of_b[b] : 0);
        const double* Tb = T + 12 * b;
        int64_t k = off[b];
        for (int64_t i = 0; i < N; ++i) {
            double p[3];
            xf_apply(Tb, (double)cloud[3 * i], (double)cloud[3 * i + 1], (double)cloud[3 * i + 2], p);
            double d = query_local(g, p, NULL);
            if (d < bound[b]) {
                if (k < cap) { out_b[k] = b; out_i[k] = i; out_d[k] = d; }
                ++k;
            }
        }
    }
    free(off);
    return total;
}

/* ObjectCollisionConstraint.value + df_dx (geometryopt.py:407-408, 413-418) for P world points.
 * Tpose = world <- object(grid) as row-major 3x4.  n = grad1 = -(gradient w.r.t. the world
 * point); row = [ (R (pt - t)) x n , n ]  -- the lever arm is R(pt - t) exactly as written in the
 * reference (Appendix B quirk). */
void orc_pose_rows(const orc_grid* g, const double* Tpose, const double* pts, int64_t P, int normalize,
                   double* d, double* rows) {
    /* inverse transform: grid-local <- world */
    double Ti[12];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) Ti[4 * r + c] = Tpose[4 * c + r];
        Ti[4 * r + 3] = -(Tpose[0 + r] * Tpose[3] + Tpose[4 + r] * Tpose[7] + Tpose[8 + r] * Tpose[11]);
    }
    double* grad = (double*)malloc(sizeof(double) * 3 * (size_t)(P > 0 ? P : 1));
    orc_query(g, Ti, pts, P, normalize, d, grad);
    for (int64_t i = 0; i < P; ++i) {
        double n[3] = {-grad[3 * i], -grad[3 * i + 1], -grad[3 * i + 2]};
        double q[3] = {pts[3 * i] - Tpose[3], pts[3 * i + 1] - Tpose[7], pts[3 * i + 2] - Tpose[11]};
        double r[3];
        for (int a = 0; a < 3; ++a) r[a] = Tpose[4 * a] * q[0] + Tpose[4 * a + 1] * q[1] + Tpose[4 * a + 2] * q[2];
        double* o = rows + 6 * i;
        o[0] = r[1] * n[2] - r[2] * n[1];
        o[1] = r[2] * n[0] - r[0] * n[2];
        o[2] = r[0] * n[1] - r[1] * n[0];
        o[3] = n[0]; o[4] = n[1]; o[5] = n[2];
    }
    free(grad);
}

/* ---- kinematics (Klampt RobotModel restated: serial/tree chain of revolute or prismatic links)
 * T_link = T_parent * TParent_link * Joint(axis, q);  revolute: rotation about `axis` by q,
 * prismatic: translation along `axis` by q.  All transforms row-major 3x4. */
typedef struct {
    int32_t n;
    const int32_t* parent;   /* n, -1 for roots */
    const int32_t* jtype;    /* n, 0 revolute, 1 prismatic */
    const double* tparent;   /* n*12 */
    const double* axis;      /* n*3, unit */
} orc_robot;

static void mat_mul34(const double* A, const double* B, double* C) {
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c)
            C[4 * r + c] = A[4 * r] * B[c] + A[4 * r + 1] * B[4 + c] + A[4 * r + 2] * B[8 + c];
        C[4 * r + 3] = A[4 * r] * B[3] + A[4 * r + 1] * B[7] + A[4 * r + 2] * B[11] + A[4 * r + 3];
    }
}

static void joint_xform(const double* ax, int jtype, double q, double* J) {
    if (jtype == 1) {
        J[0] = 1; J[1] = 0; J[2] = 0; J[3] = ax[0] * q;
        J[4] = 0; J[5] = 1; J[6] = 0; J[7] = ax[1] * q;
        J[8] = 0; J[9] = 0; J[10] = 1; J[11] = ax[2] * q;
        return;
    }
    double c = cos(q), s = sin(q), C = 1.0 - c;
    double x = ax[0], y = ax[1], z = ax[2];
    J[0] = c + x * x * C;     J[1] = x * y * C - z * s; J[2] = x * z * C + y * s; J[3] = 0;
    J[4] = y * x * C + z * s; J[5] = c + y * y * C;     J[6] = y * z * C - x * s; J[7] = 0;
    J[8] = z * x * C - y * s; J[9] = z * y * C + x * s; J[10] = c + z * z * C;    J[11] = 0;
}

/* q: B*n  ->  Tout: B*n*12 (world <- link) */
void orc_fk(const orc_robot* R, const double* q, int64_t B, double* Tout) {
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; ++b) {
        for (int l = 0; l < R->n; ++l) {
            double J[12], L[12];
            joint_xform(R->axis + 3 * l, R->jtype[l], q[b * R->n + l], J);
            mat_mul34(R->tparent + 12 * l, J, L);
            double* out = Tout + (b * R->n + l) * 12;
            if (R->parent[l] < 0) memcpy(out, L, sizeof(L));
            else mat_mul34(Tout + (b * R->n + R->parent[l]) * 12, L, out);
        }
    }
}

/* RobotLinkCollisionConstraint.value + df_dx (geometryopt.py:472-473, 478-482) at one config q:
 * d = SDF_link(T_link^-1 pt);  n = grad1;  Jp column j (ancestor-or-self j of `link`):
 *   revolute  w_j x (pt - o_j)   with w_j = R_j axis_j, o_j = origin of frame j (after the joint),
 *   prismatic w_j;   row = Jp^T n  (n entries). */
void orc_chain_rows(const orc_robot* R, int32_t link, const orc_grid* g, const double* q,
                    const double* pts, int64_t P, int normalize, double* d, double* rows) {
    double* Tl = (double*)malloc(sizeof(double) * 12 * (size_t)R->n);
    orc_fk(R, q, 1, Tl);
    const double* Tpose = Tl + 12 * link;
    double Ti[12];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) Ti[4 * r + c] = Tpose[4 * c + r];
        Ti[4 * r + 3] = -(Tpose[0 + r] * Tpose[3] + Tpose[4 + r] * Tpose[7] + Tpose[8 + r] * Tpose[11]);
    }
    double* grad = (double*)malloc(sizeof(double) * 3 * (size_t)(P > 0 ? P : 1));
    orc_query(g, Ti, pts, P, normalize, d, grad);
    for (int64_t i = 0; i < P; ++i) {
        double n[3] = {-grad[3 * i], -grad[3 * i + 1], -grad[3 * i + 2]};
        double* o = rows + (size_t)R->n * i;
        for (int j = 0; j < R->n; ++j) o[j] = 0.0;
        for (int j = link; j >= 0; j = R->parent[j]) {
            const double* Tj = Tl + 12 * j;
            const double* ax = R->axis + 3 * j;
            double w[3];
            for (int a = 0; a < 3; ++a) w[a] = Tj[4 * a] * ax[0] + Tj[4 * a + 1] * ax[1] + Tj[4 * a + 2] * ax[2];
            if (R->jtype[j] == 1) {
                o[j] = w[0] * n[0] + w[1] * n[1] + w[2] * n[2];
            } else {
                double r[3] = {pts[3 * i] - Tj[3], pts[3 * i + 1] - Tj[7], pts[3 * i + 2] - Tj[11]};
                double cx = w[1] * r[2] - w[2] * r[1], cy = w[2] * r[0] - w[0] * r[2], cz = w[0] * r[1] - w[1] * r[0];
                o[j] = cx * n[0] + cy * n[1] + cz * n[2];
            }
        }
    }
    free(grad);
    free(Tl);
}

/* ---- single-precision brute force, used ONLY as a CPU timing baseline (bench.py) ---------- */
static inline float query_local_f32(const orc_grid* g, const float* bminf, const float* invh, const float* hf,
                                    float px, float py, float pz) {
    const int dim[3] = {g->nx, g->ny, g->nz};
    float p[3] = {px, py, pz}, f[3];
    int i0[3], i1[3];
    float dout2 = 0.f;
    for (int a = 0; a < 3; ++a) {
        float u = (p[a] - bminf[a]) * invh[a] - 0.5f;
        float lo = -0.5f, hi = (float)dim[a] - 0.5f;
        float ub = u < lo ? lo : (u > hi ? hi : u);
        float o = (u - ub) * hf[a];
        dout2 += o * o;
        float uc = ub < 0.f ? 0.f : (ub > (float)(dim[a] - 1) ? (float)(dim[a] - 1) : ub);
        int i = (int)uc;
        i0[a] = i;
        i1[a] = (i + 1 < dim[a]) ? i + 1 : dim[a] - 1;
        f[a] = uc - (float)i;
    }
    const size_t sy = (size_t)g->nz, sx = (size_t)g->ny * g->nz;
#define V(I, J, K) (g->v[(size_t)(I) * sx + (size_t)(J) * sy + (size_t)(K)])
    float a00 = V(i0[0], i0[1], i0[2]) + f[2] * (V(i0[0], i0[1], i1[2]) - V(i0[0], i0[1], i0[2]));
    float a01 = V(i0[0], i1[1], i0[2]) + f[2] * (V(i0[0], i1[1], i1[2]) - V(i0[0], i1[1], i0[2]));
    float a10 = V(i1[0], i0[1], i0[2]) + f[2] * (V(i1[0], i0[1], i1[2]) - V(i1[0], i0[1], i0[2]));
    float a11 = V(i1[0], i1[1], i0[2]) + f[2] * (V(i1[0], i1[1], i1[2]) - V(i1[0], i1[1], i0[2]));
#undef V
    float b0 = a00 + f[1] * (a01 - a00), b1 = a10 + f[1] * (a11 - a10);
    float val = b0 + f[0] * (b1 - b0);
    return dout2 > 0.f ? val + sqrtf(dout2) : val;
}

void orc_min_distance_f32(const orc_grid* g, const float* cloud, int64_t N, const double* T, int64_t B,
                          float* dmin, int64_t* argmin) {
    float bminf[3], invh[3], hf[3];
    const int dim[3] = {g->nx, g->ny, g->nz};
    for (int a = 0; a < 3; ++a) {
        double h = (g->bmax[a] - g->bmin[a]) / dim[a];
        bminf[a] = (float)g->bmin[a]; invh[a] = (float)(1.0 / h); hf[a] = (float)h;
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < B; ++b) {
        float M[12];
        for (int k = 0; k < 12; ++k) M[k] = (float)T[12 * b + k];
        float best = INFINITY;
        int64_t bi = -1;
        for (int64_t i = 0; i < N; ++i) {
            float x = cloud[3 * i], y = cloud[3 * i + 1], z = cloud[3 * i + 2];
            float px = M[0] * x + M[1] * y + M[2] * z + M[3];
            float py = M[4] * x + M[5] * y + M[6] * z + M[7];
            float pz = M[8] * x + M[9] * y + M[10] * z + M[11];
            float d = query_local_f32(g, bminf, invh, hf, px, py, pz);
            if (d < best) { best = d; bi = i; }
        }
        dmin[b] = best;
        argmin[b] = bi;
    }
}

/* ---- hierarchy-pruned cloud-vs-grid minimum: the CPU arm with "pruning on" ---------------------------------------
 * Klampt's distance_ext does not test every point: it descends a bounding hierarchy of the cloud and discards
 * branches whose lower bound cannot beat the best distance found so far (or the caller's upperBound).  This is the
 * host restatement of that idea at the granularity the device uses (leaves of 128 points, SIO_PRUNE), written
 * independently of the product: a k-d ordering of the cloud into leaves, one bounding sphere (a cloud point + radius)
 * per leaf, a Lipschitz constant L of the trilinear field taken from the grid's own samples, and per transform a
 * BEST-FIRST visit of the leaves by lower bound d(rep) - L r, stopping at the first leaf whose bound exceeds
 * min(best so far, upperBound).  Results equal orc_min_distance's (same float64 query; lowest index on exact ties).
 * Used by tests (equality with the brute-force sweep) and by bench.py's pruned cpu_baseline leg. */
typedef struct {
    int64_t n, nleaf;
    int32_t* order;        /* [n]  leaf-ordered position -> caller index */
    float* pts;            /* [n*3] points in leaf order */
    double* rep;           /* [nleaf*4] representative point (a cloud point) + radius */
} orc_cloud_tree;

#define ORC_LEAF 128

static void kd_select(int32_t* idx, int64_t lo, int64_t hi, int64_t k, const float* xyz, int ax) {
    /* quickselect: after the call idx[lo..k) <= idx[k] <= idx(k..hi) along axis ax (ties by index) */
    while (hi - lo > 1) {
        int64_t mid = lo + (hi - lo) / 2;
        int32_t pv = idx[mid];
        float pvv = xyz[3 * (int64_t)pv + ax];
        int32_t t = idx[mid]; idx[mid] = idx[hi - 1]; idx[hi - 1] = t;
        int64_t s = lo;
        for (int64_t i = lo; i < hi - 1; ++i) {
            float v = xyz[3 * (int64_t)idx[i] + ax];
            if (v < pvv || (v == pvv && idx[i] < pv)) { t = idx[i]; idx[i] = idx[s]; idx[s] = t; ++s; }
        }
        t = idx[s]; idx[s] = idx[hi - 1]; idx[hi - 1] = t;
        if (s == k) return;
        if (k < s) hi = s; else lo = s + 1;
    }
}

static void kd_build(int32_t* idx, int64_t lo, int64_t hi, const float* xyz) {
    if (hi - lo <= ORC_LEAF) return;
    float bl[3] = {INFINITY, INFINITY, INFINITY}, bh[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = lo; i < hi; ++i)
        for (int a = 0; a < 3; ++a) {
            float v = xyz[3 * (int64_t)idx[i] + a];
            if (v < bl[a]) bl[a] = v;
            if (v > bh[a]) bh[a] = v;
        }
    int ax = 0;
    for (int a = 1; a < 3; ++a) if (bh[a] - bl[a] > bh[ax] - bl[ax]) ax = a;
    int64_t half = ((hi - lo) / 2 + ORC_LEAF - 1) / ORC_LEAF * ORC_LEAF;      /* left side in whole leaves */
    if (half >= hi - lo) half = hi - lo - 1;
    kd_select(idx, lo, hi, lo + half, xyz, ax);
    kd_build(idx, lo, lo + half, xyz);
    kd_build(idx, lo + half, hi, xyz);
}

orc_cloud_tree* orc_tree_build(const float* cloud, int64_t n) {
    orc_cloud_tree* t = (orc_cloud_tree*)calloc(1, sizeof(orc_cloud_tree));
    if (!t) return NULL;
    t->n = n;
    t->nleaf = (n + ORC_LEAF - 1) / ORC_LEAF;
    t->order = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    t->pts = (float*)malloc(sizeof(float) * 3 * (size_t)(n > 0 ? n : 1));
    t->rep = (double*)malloc(sizeof(double) * 4 * (size_t)(t->nleaf > 0 ? t->nleaf : 1));
    for (int64_t i = 0; i < n; ++i) t->order[i] = (int32_t)i;
    kd_build(t->order, 0, n, cloud);
    for (int64_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) t->pts[3 * i + a] = cloud[3 * (int64_t)t->order[i] + a];
    for (int64_t s = 0; s < t->nleaf; ++s) {
        const int64_t a0 = s * ORC_LEAF, a1 = (a0 + ORC_LEAF < n) ? a0 + ORC_LEAF : n;
        double c[3] = {0, 0, 0};
        for (int64_t i = a0; i < a1; ++i) for (int a = 0; a < 3; ++a) c[a] += t->pts[3 * i + a];
        for (int a = 0; a < 3; ++a) c[a] /= (double)(a1 - a0);
        int64_t best = a0;
        double bd = INFINITY;
        for (int64_t i = a0; i < a1; ++i) {
            double d2 = 0;
            for (int a = 0; a < 3; ++a) { double e = t->pts[3 * i + a] - c[a]; d2 += e * e; }
            if (d2 < bd) { bd = d2; best = i; }
        }
        double r2 = 0;
        for (int64_t i = a0; i < a1; ++i) {
            double d2 = 0;
            for (int a = 0; a < 3; ++a) { double e = (double)t->pts[3 * i + a] - (double)t->pts[3 * best + a]; d2 += e * e; }
            if (d2 > r2) r2 = d2;
        }
        for (int a = 0; a < 3; ++a) t->rep[4 * s + a] = t->pts[3 * best + a];
        t->rep[4 * s + 3] = sqrt(r2) * (1.0 + 1e-12);
    }
    return t;
}

void orc_tree_free(orc_cloud_tree* t) {
    if (!t) return;
    free(t->order); free(t->pts); free(t->rep); free(t);
}

/* Lipschitz constant of the trilinear interpolant w.r.t. the grid-local position: inside a cell |dv/dx| <= the largest
 * |difference| over its four x-edges / hx (likewise y, z); the maximum over cells of the norm of that vector. */
double orc_grid_lipschitz(const orc_grid* g) {
    const int nx = g->nx, ny = g->ny, nz = g->nz;
    const double h[3] = {(g->bmax[0] - g->bmin[0]) / nx, (g->bmax[1] - g->bmin[1]) / ny, (g->bmax[2] - g->bmin[2]) / nz};
    const size_t sy = (size_t)nz, sx = (size_t)ny * nz;
    double lip = 0.0;
#pragma omp parallel for reduction(max : lip) schedule(static)
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j)
            for (int k = 0; k < nz; ++k) {
                const int i1 = i + 1 < nx ? i + 1 : i, j1 = j + 1 < ny ? j + 1 : j, k1 = k + 1 < nz ? k + 1 : k;
                double e[3] = {0, 0, 0};
                const int I[2] = {i, i1}, J[2] = {j, j1}, K[2] = {k, k1};
                for (int b = 0; b < 2; ++b)
                    for (int c = 0; c < 2; ++c) {
                        double dx = fabs((double)g->v[I[1] * sx + J[b] * sy + K[c]] - (double)g->v[I[0] * sx + J[b] * sy + K[c]]);
                        double dy = fabs((double)g->v[I[b] * sx + J[1] * sy + K[c]] - (double)g->v[I[b] * sx + J[0] * sy + K[c]]);
                        double dz = fabs((double)g->v[I[b] * sx + J[c] * sy + K[1]] - (double)g->v[I[b] * sx + J[c] * sy + K[0]]);
                        if (dx > e[0]) e[0] = dx;
                        if (dy > e[1]) e[1] = dy;
                        if (dz > e[2]) e[2] = dz;
                    }
                double n2 = 0;
                for (int a = 0; a < 3; ++a) { e[a] /= h[a]; n2 += e[a] * e[a]; }
                double l = sqrt(n2);
                if (l > lip) lip = l;
            }
    return lip * (1.0 + 1e-9);
}

typedef struct { double lb; int64_t leaf; } orc_heap_item;
static int heap_cmp(const void* a, const void* b) {
    const orc_heap_item* x = (const orc_heap_item*)a; const orc_heap_item* y = (const orc_heap_item*)b;
    return (x->lb > y->lb) - (x->lb < y->lb);
}

/* dmin / argmin as orc_min_distance (caller indices, lowest index on exact ties, argmin = -1 when a finite bound is
 * not beaten).  `lip` from orc_grid_lipschitz of the same grid.  evaluated[b] (optional) = point queries spent.
 * use_bound != 0: the caller's upperBound prunes as well (what Klampt does with settings.upperBound): dmin[b] is then
 * exact only when it is below the bound -- all that PenetrationDepthGeometry.distance reads (gopt:128-130). */
void orc_min_distance_pruned(const orc_grid* g, double lip, const orc_cloud_tree* t, const double* T, const double* bound,
                             int64_t B, int use_bound, double* dmin, int64_t* argmin, int64_t* evaluated) {
#pragma omp parallel
    {
        orc_heap_item* items = (orc_heap_item*)malloc(sizeof(orc_heap_item) * (size_t)(t->nleaf > 0 ? t->nleaf : 1));
#pragma omp for schedule(dynamic, 1)
        for (int64_t b = 0; b < B; ++b) {
            const double* Tb = T + 12 * b;
            int64_t spent = 0;
            double best = INFINITY;
            int64_t bi = -1;
            for (int64_t s = 0; s < t->nleaf; ++s) {
                double p[3];
                xf_apply(Tb, t->rep[4 * s], t->rep[4 * s + 1], t->rep[4 * s + 2], p);
                double d = query_local(g, p, NULL);
                /* +1 where the sphere can leave the bbox: the overshoot distance added outside is 1-Lipschitz */
                double L = lip;
                for (int a = 0; a < 3; ++a)
                    if (p[a] - t->rep[4 * s + 3] < g->bmin[a] || p[a] + t->rep[4 * s + 3] > g->bmax[a]) { L = lip + 1.0; break; }
                items[s].lb = d - L * t->rep[4 * s + 3];
                items[s].leaf = s;
            }
            spent += t->nleaf;
            qsort(items, (size_t)t->nleaf, sizeof(orc_heap_item), heap_cmp);
            const double ub = (use_bound && bound && bound[b] < INFINITY) ? bound[b] : INFINITY;     /* Klampt's upperBound */
            for (int64_t k = 0; k < t->nleaf; ++k) {
                const double thr = best < ub ? best : ub;
                if (items[k].lb > thr) break;                /* no later leaf can hold a point at or below thr */
                const int64_t s = items[k].leaf, a0 = s * ORC_LEAF, a1 = (a0 + ORC_LEAF < t->n) ? a0 + ORC_LEAF : t->n;
                for (int64_t i = a0; i < a1; ++i) {
                    double p[3];
                    xf_apply(Tb, (double)t->pts[3 * i], (double)t->pts[3 * i + 1], (double)t->pts[3 * i + 2], p);
                    double d = query_local(g, p, NULL);
                    const int64_t ci = t->order[i];
                    if (d < best || (d == best && ci < bi)) { best = d; bi = ci; }
                }
                spent += a1 - a0;
            }
            dmin[b] = best;
            argmin[b] = (bound && !(best < bound[b])) ? -1 : bi;
            if (evaluated) evaluated[b] = spent;
        }
        free(items);
    }
}
