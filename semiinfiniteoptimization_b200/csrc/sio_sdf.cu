// sio_sdf.cu -- mesh -> distance grid on the device (sm_100a): SURVEY.md row f3.
//
// What it replaces: the distance part of Geometry3D.convert('VolumeGrid', res) (geometryopt.py:58-66), which the
// host substrate does in _host/csrc/hostgeom.c.  One thread per cell, the triangles staged through shared memory in
// tiles, exact point-triangle distance (Ericson, Real-Time Collision Detection 5.1.5) in float64 without FMA
// contraction: the same arithmetic as hostgeom.c's pt_tri_dist2, so min over triangles -- and therefore the grid --
// is identical to the host builder's.  The sign (generalised winding number where a neighbour cannot certify the
// side) stays on the host: it touches only the cells near the surface.
#include "sio_common.cuh"

namespace sio {

struct V3 { double x, y, z; };
__device__ __forceinline__ V3 vsub(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ double vdot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 vmadd0(V3 b, double s) { return V3{0.0 + s * b.x, 0.0 + s * b.y, 0.0 + s * b.z}; }   // madd({0,0,0}, b, s)
__device__ __forceinline__ V3 vmadd(V3 a, V3 b, double s) { return V3{a.x + s * b.x, a.y + s * b.y, a.z + s * b.z}; }

__device__ double pt_tri_dist2(V3 p, V3 a, V3 b, V3 c) {
    V3 ab = vsub(b, a), ac = vsub(c, a), ap = vsub(p, a);
    double d1 = vdot(ab, ap), d2 = vdot(ac, ap);
    if (d1 <= 0 && d2 <= 0) return vdot(ap, ap);
    V3 bp = vsub(p, b);
    double d3 = vdot(ab, bp), d4 = vdot(ac, bp);
    if (d3 >= 0 && d4 <= d3) return vdot(bp, bp);
    double vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) {
        double v = d1 / (d1 - d3);
        V3 q = vsub(ap, vmadd0(ab, v));
        return vdot(q, q);
    }
    V3 cp = vsub(p, c);
    double d5 = vdot(ab, cp), d6 = vdot(ac, cp);
    if (d6 >= 0 && d5 <= d6) return vdot(cp, cp);
    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) {
        double w = d2 / (d2 - d6);
        V3 q = vsub(ap, vmadd0(ac, w));
        return vdot(q, q);
    }
    double va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        V3 q = vsub(bp, vmadd0(vsub(c, b), w));
        return vdot(q, q);
    }
    double den = 1.0 / (va + vb + vc);
    double v = vb * den, w = vc * den;
    V3 q = vsub(ap, vmadd(vmadd0(ab, v), ac, w));
    return vdot(q, q);
}

constexpr int kSdfThreads = 128;
constexpr int kSdfTile = 128;                 // triangles per shared-memory tile (9 KB)

// tri: [nt][9] doubles (a, b, c); out[(i*ny + j)*nz + k] = distance from the cell centre to the mesh
__global__ void __launch_bounds__(kSdfThreads)
k_mesh_distance(const double* __restrict__ tri, int32_t nt, double bx, double by, double bz, double hx, double hy, double hz,
                int32_t nx, int32_t ny, int32_t nz, double* __restrict__ out) {
    __shared__ double s_tri[kSdfTile * 9];
    const int64_t ncell = (int64_t)nx * ny * nz;
    const int64_t cell = blockIdx.x * (int64_t)kSdfThreads + threadIdx.x;
    const int64_t cc = cell < ncell ? cell : ncell - 1;
    const int k = (int)(cc % nz);
    const int j = (int)((cc / nz) % ny);
    const int i = (int)(cc / ((int64_t)nz * ny));
    const V3 p = {bx + (i + 0.5) * hx, by + (j + 0.5) * hy, bz + (k + 0.5) * hz};
    double best = INFINITY;
    for (int t0 = 0; t0 < nt; t0 += kSdfTile) {
        const int n = min(kSdfTile, nt - t0);
        __syncthreads();
        for (int e = threadIdx.x; e < n * 9; e += kSdfThreads) s_tri[e] = tri[(size_t)t0 * 9 + e];
        __syncthreads();
        for (int t = 0; t < n; ++t) {
            const double* q = s_tri + t * 9;
            const double d2 = pt_tri_dist2(p, V3{q[0], q[1], q[2]}, V3{q[3], q[4], q[5]}, V3{q[6], q[7], q[8]});
            if (d2 < best) best = d2;
        }
    }
    if (cell < ncell) out[cell] = sqrt(best);
}

void launch_mesh_distance(const double* tri, int32_t nt, const double bmin[3], const double h[3], const int32_t dims[3],
                          double* out, cudaStream_t st) {
    const int64_t ncell = (int64_t)dims[0] * dims[1] * dims[2];
    if (ncell <= 0 || nt <= 0) return;
    k_mesh_distance<<<(unsigned)((ncell + kSdfThreads - 1) / kSdfThreads), kSdfThreads, 0, st>>>(
        tri, nt, bmin[0], bmin[1], bmin[2], h[0], h[1], h[2], dims[0], dims[1], dims[2], out);
}

}  // namespace sio
