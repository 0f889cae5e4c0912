/* Host-side geometry preprocessing (the role Klampt's `Geometry3D.convert('VolumeGrid', res)`
 * plays for the reference at geometryopt.py:58-66): triangle mesh -> cell-centred signed distance
 * grid.  This runs ONCE per geometry, on the host, before anything is uploaded; it is NOT on the
 * oracle hot path and is not the CPU oracle (that lives in /oracle).
 *
 * Distance: exact point-triangle distance (brute force with a running-bound sphere cull).
 * Sign: generalised winding number (robust to the open / self-intersecting PSB meshes m797, m1062),
 *       evaluated only where a neighbour's distance cannot certify "same side".
 * Layout: out[(i*ny + j)*nz + k], value at cell centre bmin + (i+.5, j+.5, k+.5)*h  (z fastest,
 *       the order geometryopt.py:25-28 reads a Klampt VolumeGrid in).
 *
 * Build: gcc -O3 -march=native -fopenmp -shared -fPIC hostgeom.c -o libsio_hostgeom.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct { double x, y, z; } v3;
static inline v3 sub(v3 a, v3 b) { v3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }
static inline double dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 cross(v3 a, v3 b) { v3 r = {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; return r; }
static inline v3 madd(v3 a, v3 b, double s) { v3 r = {a.x + s * b.x, a.y + s * b.y, a.z + s * b.z}; return r; }

/* squared distance from p to triangle abc (Ericson, Real-Time Collision Detection 5.1.5) */
static double pt_tri_dist2(v3 p, v3 a, v3 b, v3 c) {
    v3 ab = sub(b, a), ac = sub(c, a), ap = sub(p, a);
    double d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0 && d2 <= 0) return dot(ap, ap);
    v3 bp = sub(p, b);
    double d3 = dot(ab, bp), d4 = dot(ac, bp);
    if (d3 >= 0 && d4 <= d3) return dot(bp, bp);
    double vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) {
        double v = d1 / (d1 - d3);
        v3 q = sub(ap, madd((v3){0, 0, 0}, ab, v));
        return dot(q, q);
    }
    v3 cp = sub(p, c);
    double d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0 && d5 <= d6) return dot(cp, cp);
    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) {
        double w = d2 / (d2 - d6);
        v3 q = sub(ap, madd((v3){0, 0, 0}, ac, w));
        return dot(q, q);
    }
    double va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        v3 q = sub(bp, madd((v3){0, 0, 0}, sub(c, b), w));
        return dot(q, q);
    }
    double den = 1.0 / (va + vb + vc);
    double v = vb * den, w = vc * den;
    v3 q = sub(ap, madd(madd((v3){0, 0, 0}, ab, v), ac, w));
    return dot(q, q);
}

/* solid angle of triangle abc seen from p (Van Oosterom & Strackee) */
static double solid_angle(v3 p, v3 a, v3 b, v3 c) {
    v3 A = sub(a, p), B = sub(b, p), C = sub(c, p);
    double la = sqrt(dot(A, A)), lb = sqrt(dot(B, B)), lc = sqrt(dot(C, C));
    double num = dot(A, cross(B, C));
    double den = la * lb * lc + dot(A, B) * lc + dot(B, C) * la + dot(C, A) * lb;
    return 2.0 * atan2(num, den);
}

static double winding(v3 p, const v3* V, const int32_t* T, int nt) {
    double s = 0;
    for (int t = 0; t < nt; ++t) s += solid_angle(p, V[T[3 * t]], V[T[3 * t + 1]], V[T[3 * t + 2]]);
    return s / (4.0 * M_PI);
}

/* returns 0 on success */
int hg_mesh_sdf(const double* verts, int nv, const int32_t* tris, int nt,
                const double* bmin, const double* h, const int32_t* dims, float* out) {
    (void)nv;
    const v3* V = (const v3*)verts;
    const int nx = dims[0], ny = dims[1], nz = dims[2];
    /* bounding spheres of the triangles for culling */
    v3* cen = (v3*)malloc(sizeof(v3) * (size_t)nt);
    double* rad = (double*)malloc(sizeof(double) * (size_t)nt);
    if (!cen || !rad) return -1;
    for (int t = 0; t < nt; ++t) {
        v3 a = V[tris[3 * t]], b = V[tris[3 * t + 1]], c = V[tris[3 * t + 2]];
        v3 m = {(a.x + b.x + c.x) / 3, (a.y + b.y + c.y) / 3, (a.z + b.z + c.z) / 3};
        double r = 0, d;
        d = dot(sub(a, m), sub(a, m)); if (d > r) r = d;
        d = dot(sub(b, m), sub(b, m)); if (d > r) r = d;
        d = dot(sub(c, m), sub(c, m)); if (d > r) r = d;
        cen[t] = m; rad[t] = sqrt(r);
    }
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j) {
            double dprev = -1; int tprev = 0; int sprev = 0;
            for (int k = 0; k < nz; ++k) {
                v3 p = {bmin[0] + (i + 0.5) * h[0], bmin[1] + (j + 0.5) * h[1], bmin[2] + (k + 0.5) * h[2]};
                /* seed the bound with the previous cell's nearest triangle */
                double best = pt_tri_dist2(p, V[tris[3 * tprev]], V[tris[3 * tprev + 1]], V[tris[3 * tprev + 2]]);
                int tb = tprev;
                for (int t = 0; t < nt; ++t) {
                    v3 q = sub(p, cen[t]);
                    double dc = sqrt(dot(q, q)) - rad[t];
                    if (dc > 0 && dc * dc >= best) continue;
                    double d2 = pt_tri_dist2(p, V[tris[3 * t]], V[tris[3 * t + 1]], V[tris[3 * t + 2]]);
                    if (d2 < best) { best = d2; tb = t; }
                }
                double d = sqrt(best);
                int sgn;
                /* the segment from the previous centre has length h[2]; if either endpoint is
                 * farther than that from the surface, the two cells are on the same side */
                if (dprev >= 0 && (dprev > 1.0001 * h[2] || d > 1.0001 * h[2])) sgn = sprev;
                else sgn = (winding(p, V, tris, nt) > 0.5) ? -1 : 1;
                out[((size_t)i * ny + j) * nz + k] = (float)(sgn * d);
                dprev = d; tprev = tb; sprev = sgn;
            }
        }
    free(cen); free(rad);
    return 0;
}

/* sign pass alone: `dist` holds the exact unsigned distances of the cell centres (from the device builder,
 * sio_mesh_distance_grid); the side is decided exactly as in hg_mesh_sdf.  returns 0 on success */
int hg_sdf_sign_pass(const double* verts, const int32_t* tris, int nt, const double* bmin, const double* h, const int32_t* dims,
                     const double* dist, float* out) {
    const v3* V = (const v3*)verts;
    const int nx = dims[0], ny = dims[1], nz = dims[2];
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j) {
            double dprev = -1; int sprev = 0;
            for (int k = 0; k < nz; ++k) {
                const size_t c = ((size_t)i * ny + j) * nz + k;
                const double d = dist[c];
                int sgn;
                if (dprev >= 0 && (dprev > 1.0001 * h[2] || d > 1.0001 * h[2])) sgn = sprev;
                else {
                    v3 p = {bmin[0] + (i + 0.5) * h[0], bmin[1] + (j + 0.5) * h[1], bmin[2] + (k + 0.5) * h[2]};
                    sgn = (winding(p, V, tris, nt) > 0.5) ? -1 : 1;
                }
                out[c] = (float)(sgn * d);
                dprev = d; sprev = sgn;
            }
        }
    return 0;
}

/* squared distance + nearest triangle for a list of points (used by tests / surface checks) */
int hg_points_mesh_dist(const double* verts, const int32_t* tris, int nt, const double* pts, int64_t np_, double* out) {
    const v3* V = (const v3*)verts;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < np_; ++i) {
        v3 p = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
        double best = INFINITY;
        for (int t = 0; t < nt; ++t) {
            double d2 = pt_tri_dist2(p, V[tris[3 * t]], V[tris[3 * t + 1]], V[tris[3 * t + 2]]);
            if (d2 < best) best = d2;
        }
        double w = winding(p, V, tris, nt);
        out[i] = (w > 0.5 ? -1.0 : 1.0) * sqrt(best);
    }
    return 0;
}
