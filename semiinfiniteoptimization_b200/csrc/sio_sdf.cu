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

// =====================================================================================================================
// Row f3, second half: the inside/outside SIGN of the distance grid and the surface SAMPLER, on the device.
// Compiled with --fmad=false: every float64 operation is the plain IEEE one the host builder (hostgeom.c / geometry.py)
// and the oracle's independent restatement (oracle/ref_meshconv.py) perform, in the same order.
// =====================================================================================================================
namespace sio {

__device__ __forceinline__ V3 vcross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// solid angle of triangle abc seen from p (Van Oosterom & Strackee)
__device__ __forceinline__ double solid_angle(V3 p, V3 a, V3 b, V3 c) {
    const V3 A = vsub(a, p), B = vsub(b, p), C = vsub(c, p);
    const double la = sqrt(vdot(A, A)), lb = sqrt(vdot(B, B)), lc = sqrt(vdot(C, C));
    const double num = vdot(A, vcross(B, C));
    const double den = la * lb * lc + vdot(A, B) * lc + vdot(B, C) * la + vdot(C, A) * lb;
    return 2.0 * atan2(num, den);
}

// Side of every cell that cannot inherit it from its z-predecessor.  A cell INHERITS the side of the cell before it in
// its z-column when either of the two is farther than one cell (x 1.0001) from the surface -- the segment between their
// centres then cannot cross it; the first cell of a column and every cell near the surface get their side from the
// generalised winding number (robust to the open / self-intersecting meshes of the benchmark): > 1/2 = inside.
// side[c] = -1 / +1 decided here, 0 = inherit.  One thread per cell, triangles tiled through shared memory.
__global__ void __launch_bounds__(kSdfThreads)
k_sdf_winding(const double* __restrict__ tri, int32_t nt, double bx, double by, double bz, double hx, double hy, double hz,
              int32_t nx, int32_t ny, int32_t nz, const double* __restrict__ dist, int8_t* __restrict__ side) {
    __shared__ double s_tri[kSdfTile * 9];
    const int64_t ncell = (int64_t)nx * ny * nz;
    const int64_t cell = blockIdx.x * (int64_t)kSdfThreads + threadIdx.x;
    const int64_t cc = cell < ncell ? cell : ncell - 1;
    const int k = (int)(cc % nz);
    const int j = (int)((cc / nz) % ny);
    const int i = (int)(cc / ((int64_t)nz * ny));
    bool need = cell < ncell;
    if (need && k > 0) {
        const double lim = 1.0001 * hz;
        need = !(dist[cc - 1] > lim || dist[cc] > lim);
    }
    const bool any = __syncthreads_or(need);
    if (!any) { if (cell < ncell) side[cell] = 0; return; }
    const V3 p = {bx + (i + 0.5) * hx, by + (j + 0.5) * hy, bz + (k + 0.5) * hz};
    double s = 0.0;
    for (int t0 = 0; t0 < nt; t0 += kSdfTile) {
        const int n = min(kSdfTile, nt - t0);
        __syncthreads();
        for (int e = threadIdx.x; e < n * 9; e += kSdfThreads) s_tri[e] = tri[(size_t)t0 * 9 + e];
        __syncthreads();
        if (need)
            for (int t = 0; t < n; ++t) {
                const double* q = s_tri + t * 9;
                s += solid_angle(p, V3{q[0], q[1], q[2]}, V3{q[3], q[4], q[5]}, V3{q[6], q[7], q[8]});
            }
    }
    if (cell < ncell) side[cell] = need ? ((s / (4.0 * 3.14159265358979323846) > 0.5) ? -1 : 1) : 0;
}

// one thread per z-column: carry the side along the column and write the signed float32 sample
__global__ void k_sdf_columns(int32_t nx, int32_t ny, int32_t nz, const double* __restrict__ dist, const int8_t* __restrict__ side,
                              float* __restrict__ out) {
    const int64_t col = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (col >= (int64_t)nx * ny) return;
    int s = 1;
    for (int k = 0; k < nz; ++k) {
        const int64_t c = col * nz + k;
        const int sc = side[c];
        if (sc != 0) s = sc;
        out[c] = (float)(s * dist[c]);
    }
}

void launch_sdf_sign(const double* tri, int32_t nt, const double bmin[3], const double h[3], const int32_t dims[3],
                     const double* dist, int8_t* side, float* out, cudaStream_t st) {
    const int64_t ncell = (int64_t)dims[0] * dims[1] * dims[2];
    if (ncell <= 0 || nt <= 0) return;
    k_sdf_winding<<<(unsigned)((ncell + kSdfThreads - 1) / kSdfThreads), kSdfThreads, 0, st>>>(
        tri, nt, bmin[0], bmin[1], bmin[2], h[0], h[1], h[2], dims[0], dims[1], dims[2], dist, side);
    const int64_t ncol = (int64_t)dims[0] * dims[1];
    k_sdf_columns<<<(unsigned)((ncol + 127) / 128), 128, 0, st>>>(dims[0], dims[1], dims[2], dist, side, out);
}

// ---- surface sampler -------------------------------------------------------------------------------------------------
// Klampt's mesh -> PointCloud conversion at resolution `res` (geometryopt.py:65): the vertices plus, per triangle, the
// lattice spanned by its two shorter edges with m1 x m2 subdivisions (m = ceil(edge / res)), so that no surface point is
// farther than ~res from a sample; points shared along edges are merged.  The host lays the output out (per triangle:
// apex corner a, edge ends b, c, m1, m2 and the offset of its block) -- O(triangles) -- and the device does the rest:
//   k_lattice_fill   every lattice point a (1-u-v) + b u + c v with u + v <= 1, corners excluded, in (i, j) order
//   k_dedup_insert   hash set on the points quantised to 1e-9 m, keeping the LOWEST index of each key
//   k_dedup_flag / k_block_counts / k_scan_blocks / k_compact   order-preserving compaction of the first occurrences
struct LatticeTri { double a[3], b[3], c[3]; int32_t m1, m2; int64_t base; };

__global__ void k_lattice_fill(const LatticeTri* __restrict__ tris, int32_t nt, double* __restrict__ pts) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    const LatticeTri T = tris[t];
    int64_t o = T.base;
    for (int i = 0; i <= T.m1; ++i) {
        const double u = (double)i / (double)T.m1;
        for (int j = 0; j <= T.m2; ++j) {
            const double v = (double)j / (double)T.m2;
            if (!((u + v) <= 1.0 + 1e-12)) continue;
            if ((i == 0 && j == 0) || (i == T.m1 && j == 0) || (i == 0 && j == T.m2)) continue;
            const double w = (1.0 - u) - v;
#pragma unroll
            for (int a = 0; a < 3; ++a) pts[3 * o + a] = (T.a[a] * w + T.b[a] * u) + T.c[a] * v;
            ++o;
        }
    }
}

__device__ __forceinline__ void quant_key(const double* __restrict__ p, long long k[3]) {
    k[0] = (long long)rint(p[0] * 1e9); k[1] = (long long)rint(p[1] * 1e9); k[2] = (long long)rint(p[2] * 1e9);
}
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
__device__ __forceinline__ uint64_t key_hash(const long long k[3]) {
    return mix64((uint64_t)k[0] * 0x9E3779B97F4A7C15ULL ^ mix64((uint64_t)k[1] + 0x7F4A7C15ULL) ^ (mix64((uint64_t)k[2]) << 1));
}
constexpr uint32_t kEmpty = 0xffffffffu;

// table[slot] = lowest point index whose key lives in that slot (all indices ever stored in a slot share one key)
__global__ void k_dedup_insert(const double* __restrict__ pts, int64_t n, uint32_t* __restrict__ table, uint64_t mask) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long k[3], q[3];
    quant_key(pts + 3 * i, k);
    uint64_t h = key_hash(k) & mask;
    for (;;) {
        uint32_t cur = table[h];
        if (cur == kEmpty) {
            cur = atomicCAS(table + h, kEmpty, (uint32_t)i);
            if (cur == kEmpty) return;
        }
        quant_key(pts + 3 * (int64_t)cur, q);
        if (q[0] == k[0] && q[1] == k[1] && q[2] == k[2]) { atomicMin(table + h, (uint32_t)i); return; }
        h = (h + 1) & mask;
    }
}

__global__ void k_dedup_flag(const double* __restrict__ pts, int64_t n, const uint32_t* __restrict__ table, uint64_t mask,
                             uint8_t* __restrict__ keep) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long k[3], q[3];
    quant_key(pts + 3 * i, k);
    uint64_t h = key_hash(k) & mask;
    for (;;) {
        const uint32_t cur = table[h];
        quant_key(pts + 3 * (int64_t)cur, q);
        if (q[0] == k[0] && q[1] == k[1] && q[2] == k[2]) { keep[i] = cur == (uint32_t)i; return; }
        h = (h + 1) & mask;
    }
}

constexpr int kScanBlock = 1024;
__global__ void __launch_bounds__(kScanBlock) k_block_counts(const uint8_t* __restrict__ keep, int64_t n, uint32_t* __restrict__ counts) {
    const int64_t i = blockIdx.x * (int64_t)kScanBlock + threadIdx.x;
    const int c = __syncthreads_count(i < n && keep[i]);
    if (threadIdx.x == 0) counts[blockIdx.x] = (uint32_t)c;
}
// exclusive scan of the block counts by ONE block (a cloud of 10^8 points has 10^5 counts); total -> counts[nblocks]
__global__ void __launch_bounds__(kScanBlock) k_scan_blocks(uint32_t* __restrict__ counts, int64_t nblocks) {
    __shared__ uint32_t s[kScanBlock];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t b0 = 0; b0 < nblocks; b0 += kScanBlock) {
        const int64_t b = b0 + threadIdx.x;
        const uint32_t v = b < nblocks ? counts[b] : 0u;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < kScanBlock; o <<= 1) {
            const uint32_t t = threadIdx.x >= o ? s[threadIdx.x - o] : 0u;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        if (b < nblocks) counts[b] = carry + s[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == kScanBlock - 1) carry += s[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[nblocks] = carry;
}
__global__ void __launch_bounds__(kScanBlock) k_compact(const double* __restrict__ pts, const uint8_t* __restrict__ keep, int64_t n,
                                                        const uint32_t* __restrict__ offs, double* __restrict__ out) {
    __shared__ uint32_t wsum[kScanBlock / 32];
    const int64_t i = blockIdx.x * (int64_t)kScanBlock + threadIdx.x;
    const bool k = i < n && keep[i];
    const unsigned m = __ballot_sync(0xffffffffu, k);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) wsum[warp] = __popc(m);
    __syncthreads();
    uint32_t before = offs[blockIdx.x];
    for (int w = 0; w < warp; ++w) before += wsum[w];
    if (k) {
        const int64_t o = before + __popc(m & ((1u << lane) - 1u));
        out[3 * o] = pts[3 * i]; out[3 * o + 1] = pts[3 * i + 1]; out[3 * o + 2] = pts[3 * i + 2];
    }
}

void launch_lattice_fill(const void* tris, int32_t nt, double* pts, cudaStream_t st) {
    if (nt > 0) k_lattice_fill<<<(nt + 63) / 64, 64, 0, st>>>(reinterpret_cast<const LatticeTri*>(tris), nt, pts);
}
// pts [n*3] -> out [count*3] (first occurrence of every distinct quantised point, original order); *d_total = count.
// table: table_slots uint32 (power of two >= 2 n), keep: n bytes, counts: ceil(n / 1024) + 1 uint32
void launch_dedup(const double* pts, int64_t n, uint32_t* table, uint64_t table_slots, uint8_t* keep, uint32_t* counts, double* out,
                  cudaStream_t st) {
    if (n <= 0) return;
    cudaMemsetAsync(table, 0xff, table_slots * sizeof(uint32_t), st);
    const unsigned g = (unsigned)((n + 255) / 256);
    k_dedup_insert<<<g, 256, 0, st>>>(pts, n, table, table_slots - 1);
    k_dedup_flag<<<g, 256, 0, st>>>(pts, n, table, table_slots - 1, keep);
    const int64_t nb = (n + kScanBlock - 1) / kScanBlock;
    k_block_counts<<<(unsigned)nb, kScanBlock, 0, st>>>(keep, n, counts);
    k_scan_blocks<<<1, kScanBlock, 0, st>>>(counts, nb);
    k_compact<<<(unsigned)nb, kScanBlock, 0, st>>>(pts, keep, n, counts, out);
}

}  // namespace sio
