// sio_kernels.cu -- float32 bulk kernels of the collision oracle (sm_100a).
//
// K1/K2 of SURVEY.md 2.1: a fused rigid-transform + trilinear SDF query, replacing Klampt's
// `distance_point` / `distance_ext` inner loops (reference call sites geometryopt.py:104-108, 127).
//
// Design (DESIGN.md "Kernels"): this is gather work on CUDA cores -- no tensor cores.  Each thread
// keeps kPtsPerThread cloud points (float4, 128-bit coalesced loads of the Morton-sorted cloud) in
// registers and walks a chunk of transforms staged in shared memory, so the cloud is read from HBM
// once per chunk.  The grid is resident as BRICKS: per cell the 8 monomial coefficients of its
// trilinear interpolant (32 B = one sector), so a query is ONE 256-bit read-only load (LDG.E.256)
// at one computed address + a 7-FMA Horner evaluation, instead of eight 4-byte gathers at four
// addresses and 14 lerp operations.  History (profiles/): the 8-gather version was issue-bound at
// 101 instr/query; corner bricks brought 57; this version runs 43 instr/query (FADD.RZ
// floor/fraction on the FMA pipe, floor biases folded into the cell index, a bounding-sphere
// in-lattice test per (128-point sub-tile, transform) instead of per-point range tests, one candidate
// test per 4 queries) and is bound by the L1 data pipe: a 256-bit warp load is 8 wavefronts.
// Minima are reduced per warp with redux.sync.min.f32 and written as one float per (transform,
// 128-point sub-tile); the float64 finalize kernel (sio_fp64.cu) re-evaluates only the sub-tiles
// within tau of the minimum, which makes the final arg-min identical to the float64 oracle's.
#include "sio_common.cuh"
#include <cstdlib>

namespace sio {

// floor + fraction of a non-negative lattice coordinate with FADD.RZ (no F2I/I2F on the XU pipe):
// c + 2^23 rounded toward zero = floor(c) + 2^23, whose bit pattern is kFloorBias + floor(c).
constexpr uint32_t kFloorBias = 0x4B000000u;
__device__ __forceinline__ void split_rz(float c, uint32_t& bits, float& f) {
    const float big = 8388608.0f;
    float t = __fadd_rz(c, big);
    bits = __float_as_uint(t);
    f = c - (t - big);
}

__device__ __forceinline__ float min3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// warp-wide float minimum in one instruction (redux.sync on f32 is new with sm_100a)
__device__ __forceinline__ float warp_min_f32(float v) {
    float r;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}

// One cell of the grid as the 8 MONOMIAL coefficients of its trilinear interpolant,
//     v(fx,fy,fz) = sum_m c[m] fx^m2 fy^m1 fz^m0,   m = 4*m2 + 2*m1 + m0,
// (f = offset from the cell's low corner in cell units), 32 bytes = one sector, fetched with ONE
// 256-bit read-only load.  Horner evaluation is 7 FMAs (the corner form needs 7 FMAs + 7 subtractions)
// and the analytic gradient reuses its partial sums.  Coefficients are formed in float64 on the host.
// Storage order inside the 32 bytes: [c6 c2 | c7 c3 | c4 c0 | c5 c1] -- the four register PAIRS the packed Horner evaluation
// (fma.rn.f32x2, SASS FFMA2) consumes directly: (a,c) = fz (c7,c3) + (c6,c2); (b,d) = fz (c5,c1) + (c4,c0);
// (e,f) = fy (a,c) + (b,d); v = e fx + f.  kPos[m] = position of coefficient m.
struct Cell { float c[8]; };
__host__ __device__ constexpr int brick_pos(int m) { return m == 0 ? 5 : m == 1 ? 7 : m == 2 ? 1 : m == 3 ? 3 : m == 4 ? 4 : m == 5 ? 6 : m == 6 ? 0 : 2; }
__device__ __forceinline__ Cell load_cell(const float* __restrict__ brick, uint32_t cell) {
    Cell q;
    const float* p = brick + (size_t)cell * 8;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(q.c[0]), "=f"(q.c[1]), "=f"(q.c[2]), "=f"(q.c[3]), "=f"(q.c[4]), "=f"(q.c[5]), "=f"(q.c[6]), "=f"(q.c[7])
        : "l"(p));
    return q;
}

// ---- packed float32 pairs (sm_100a: fma/add .f32x2 -> SASS FFMA2 / FADD2, one issue slot for two lanes of work) -------
// A pair lives in a 64-bit register (two adjacent 32-bit registers).  Every operand of FFMA2 / FADD2 may also be a
// single register broadcast to both halves (SASS `R.F32`), which ptxas selects when both halves of a packed operand
// are the same register -- so a matrix entry, a fraction or a constant multiplies a pair without any duplication.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 add2_rz(f32x2 a, f32x2 b) { f32x2 r; asm("add.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// one cell as four register pairs [c6 c2 | c7 c3 | c4 c0 | c5 c1], one 256-bit read-only load
struct Cell2 { f32x2 p62, p73, p40, p51; };
__device__ __forceinline__ Cell2 load_cell2(const float* __restrict__ brick, uint32_t cell) {
    Cell2 q;
    const float* p = brick + (size_t)cell * 8;
    asm("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(q.p62), "=l"(q.p73), "=l"(q.p40), "=l"(q.p51) : "l"(p));
    return q;
}

// what the inner loop needs to know about the grid of the current transform
struct GridK {
    const float* brick;
    uint32_t cx, cz;             // cell strides: cell = ix*cx + iy*cz + iz   (cx = ny*nz, cz = nz)
    uint32_t nbias;              // -kFloorBias*(cx+cz+1) mod 2^32: folds the three floor biases into the cell index
    float hx, hy, hz;            // cell size
    float hix, hiy, hiz;         // dim-1
};

// trilinear value at lattice coordinates already inside [0, dim-1]
__device__ __forceinline__ float lattice_value(const GridK& g, float cx, float cy, float cz) {
    uint32_t bx, by, bz; float fx, fy, fz;
    split_rz(cx, bx, fx); split_rz(cy, by, fy); split_rz(cz, bz, fz);
    Cell q = load_cell(g.brick, bx * g.cx + (by * g.cz + (bz + g.nbias)));
    float a = fmaf(q.c[2], fz, q.c[0]);          // coefficient of fx fy
    float b = fmaf(q.c[6], fz, q.c[4]);          // coefficient of fx
    float c = fmaf(q.c[3], fz, q.c[1]);          // coefficient of fy
    float d = fmaf(q.c[7], fz, q.c[5]);
    return fmaf(fmaf(a, fy, b), fx, fmaf(c, fy, d));
}

// general point: clamp to the lattice, add the distance to the bbox [-1/2, dim-1/2] (Appendix A.2)
__device__ __forceinline__ float clamped_value(const GridK& g, float ux, float uy, float uz) {
    float bx = fminf(fmaxf(ux, -0.5f), g.hix + 0.5f);
    float by = fminf(fmaxf(uy, -0.5f), g.hiy + 0.5f);
    float bz = fminf(fmaxf(uz, -0.5f), g.hiz + 0.5f);
    float ox = (ux - bx) * g.hx, oy = (uy - by) * g.hy, oz = (uz - bz) * g.hz;
    float dout2 = fmaf(ox, ox, fmaf(oy, oy, oz * oz));
    float cx = fminf(fmaxf(ux, 0.0f), g.hix);
    float cy = fminf(fmaxf(uy, 0.0f), g.hiy);
    float cz = fminf(fmaxf(uz, 0.0f), g.hiz);
    float val = lattice_value(g, cx, cy, cz);
    float dout;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(dout) : "f"(dout2));
    return val + dout;
}

__device__ __forceinline__ GridK grid_of(const XfRec& r) {
    GridK g;
    g.brick = r.v; g.cx = (uint32_t)r.sx; g.cz = (uint32_t)r.sy;
    g.nbias = 0u - kFloorBias * (g.cx + g.cz + 1u);
    g.hx = r.hx; g.hy = r.hy; g.hz = r.hz; g.hix = r.hix; g.hiy = r.hiy; g.hiz = r.hiz;
    return g;
}

// Lipschitz constant of the interpolant over a sphere (centre at lattice coordinates u, radius rad in metres): the
// largest per-cell constant among the cells the sphere can touch.  Along an axis the sphere spans rad / h cells either
// way, so those cells lie within D = ceil(rad / min h) cells of the centre's cell (clamping the coordinates into the
// lattice is monotone and 1-Lipschitz, so this also holds for spheres that leave it; 1/256 cell of slack covers the
// float32 rounding of u, ~1e-5 cells); the grid carries the maxima over such neighbourhoods for D = 1 .. kLipLevels.
// On an SDF the global constant is set by a few cells on the medial axis (sqrt 3 for a box) while the field is
// 1-Lipschitz almost everywhere: on BASELINE configs[1] (the chair) the cull keeps 3.5 % of the sub-tiles with the local
// bound where the global one kept 17 %.
__device__ __forceinline__ float local_lipschitz(const XfRec& r, const GridK& g, float ux, float uy, float uz, float rad) {
    if (!r.lipd) return r.lip;
    const float rho = __fmul_ru(rad, r.ihmin) + 0.00390625f;
    if (!(rho <= (float)kLipLevels)) return r.lip;
    const int D = max(1, (int)ceilf(rho));
    const uint32_t ix = (uint32_t)fminf(fmaxf(ux, 0.0f), g.hix), iy = (uint32_t)fminf(fmaxf(uy, 0.0f), g.hiy),
                   iz = (uint32_t)fminf(fmaxf(uz, 0.0f), g.hiz);
    const uint32_t q = __ldg(r.lipd + (size_t)(D - 1) * (uint32_t)r.ncell + (ix * g.cx + iy * g.cz + iz));
    return q == 255u ? r.lip : (float)q * (1.0f / 64.0f);
}

// Is the bounding sphere (centre p, radius p.w, cloud-local) of a group of points, mapped by the
// transform of record r, entirely inside the sample lattice [0, dim-1]^3?  Then every point of the
// group takes the unclamped path.  Row a of M has norm 1/h_a, so the sphere spans r/h_a cells on axis a;
// 1/64 cell of slack covers the float32 rounding of the lattice coordinates (~1e-5 cells).
__device__ __forceinline__ bool sphere_in_lattice(const XfRec& r, float4 p) {
    float ux = fmaf(r.m[0], p.x, fmaf(r.m[1], p.y, fmaf(r.m[2], p.z, r.m[3])));
    float uy = fmaf(r.m[4], p.x, fmaf(r.m[5], p.y, fmaf(r.m[6], p.z, r.m[7])));
    float uz = fmaf(r.m[8], p.x, fmaf(r.m[9], p.y, fmaf(r.m[10], p.z, r.m[11])));
    float ex = __fdividef(p.w, r.hx) + 0.015625f, ey = __fdividef(p.w, r.hy) + 0.015625f, ez = __fdividef(p.w, r.hz) + 0.015625f;
    float m = min3(ux - ex, r.hix - ux - ex, uy - ey);
    m = min3(m, r.hiy - uy - ey, uz - ez);
    m = fminf(m, r.hiz - uz - ez);
    return m >= 0.0f;                 // false for NaN
}

// value + gradient w.r.t. the input point (cloud-local), optionally normalised
__device__ __forceinline__ float eval32_grad(const XfRec& r, float px, float py, float pz, int normalize,
                                             float& gx, float& gy, float& gz) {
    float ux = fmaf(r.m[0], px, fmaf(r.m[1], py, fmaf(r.m[2], pz, r.m[3])));
    float uy = fmaf(r.m[4], px, fmaf(r.m[5], py, fmaf(r.m[6], pz, r.m[7])));
    float uz = fmaf(r.m[8], px, fmaf(r.m[9], py, fmaf(r.m[10], pz, r.m[11])));
    float bx = fminf(fmaxf(ux, -0.5f), r.hix + 0.5f);
    float by = fminf(fmaxf(uy, -0.5f), r.hiy + 0.5f);
    float bz = fminf(fmaxf(uz, -0.5f), r.hiz + 0.5f);
    float ox = (ux - bx) * r.hx, oy = (uy - by) * r.hy, oz = (uz - bz) * r.hz;
    float dout2 = fmaf(ox, ox, fmaf(oy, oy, oz * oz));
    float cx = fminf(fmaxf(ux, 0.0f), r.hix);
    float cy = fminf(fmaxf(uy, 0.0f), r.hiy);
    float cz = fminf(fmaxf(uz, 0.0f), r.hiz);
    uint32_t bx_, by_, bz_; float fx, fy, fz;
    split_rz(cx, bx_, fx); split_rz(cy, by_, fy); split_rz(cz, bz_, fz);
    const uint32_t sx = (uint32_t)r.sx, sy = (uint32_t)r.sy;
    Cell q = load_cell(r.v, bx_ * sx + (by_ * sy + (bz_ + (0u - kFloorBias * (sx + sy + 1u)))));
    float a = fmaf(q.c[2], fz, q.c[0]);          // coefficient of fx fy
    float b = fmaf(q.c[6], fz, q.c[4]);          // coefficient of fx
    float c = fmaf(q.c[3], fz, q.c[1]);          // coefficient of fy
    float d = fmaf(q.c[7], fz, q.c[5]);
    float dx = fmaf(a, fy, b);                   // dv/dfx
    float val = fmaf(dx, fx, fmaf(c, fy, d));
    // derivative w.r.t. the cell coordinate, zero on clamped axes (Appendix A.2 step 3/5)
    float gfx = (ux > 0.0f && ux < r.hix) ? dx : 0.0f;
    float gfy = (uy > 0.0f && uy < r.hiy) ? fmaf(a, fx, c) : 0.0f;
    float gfz = (uz > 0.0f && uz < r.hiz) ? fmaf(fmaf(q.c[2], fy, q.c[6]), fx, fmaf(q.c[3], fy, q.c[7])) : 0.0f;
    float dout = sqrtf(dout2);
    if (dout2 > 0.0f) {
        float inv = 1.0f / dout;
        // outward unit vector of the bbox clamp, expressed per cell coordinate: h * o / |o|
        gfx = fmaf(r.hx * ox, inv, gfx);
        gfy = fmaf(r.hy * oy, inv, gfy);
        gfz = fmaf(r.hz * oz, inv, gfz);
    }
    // g_in = M^T gf   (M = diag(1/h) R, so this is R^T diag(1/h) gf)
    gx = fmaf(r.m[0], gfx, fmaf(r.m[4], gfy, r.m[8] * gfz));
    gy = fmaf(r.m[1], gfx, fmaf(r.m[5], gfy, r.m[9] * gfz));
    gz = fmaf(r.m[2], gfx, fmaf(r.m[6], gfy, r.m[10] * gfz));
    if (normalize) {
        float n2 = fmaf(gx, gx, fmaf(gy, gy, gz * gz));
        if (n2 > 0.0f) { float s = rsqrtf(n2); gx *= s; gy *= s; gz *= s; }
    }
    return val + dout;
}

// ---- transform records ---------------------------------------------------------------------------
__device__ __forceinline__ XfRec make_xf(const GridDev& g, const double* T, float thr) {
    XfRec r;
    for (int a = 0; a < 3; ++a) {
        double ih = g.ih[a];
        r.m[4 * a + 0] = (float)(T[4 * a + 0] * ih);
        r.m[4 * a + 1] = (float)(T[4 * a + 1] * ih);
        r.m[4 * a + 2] = (float)(T[4 * a + 2] * ih);
        r.m[4 * a + 3] = (float)((T[4 * a + 3] - g.bmin[a]) * ih - 0.5);
    }
    r.hx = (float)g.h[0]; r.hy = (float)g.h[1]; r.hz = (float)g.h[2];
    r.hix = (float)(g.nx - 1); r.hiy = (float)(g.ny - 1); r.hiz = (float)(g.nz - 1);
    r.sx = g.ny * g.nz; r.sy = g.nz; r.v = g.brick; r.lip = g.lip; r.lipd = g.lipd; r.ncell = g.nx * g.ny * g.nz; r.ihmin = __frcp_ru(fminf(r.hx, fminf(r.hy, r.hz)));
    r.thr = thr;
    return r;
}

__global__ void k_prep_xf(const GridDev* __restrict__ grids, const int32_t* __restrict__ grid_of_b,
                          const double* __restrict__ T_rel, const double* __restrict__ bound, int64_t B,
                          double tau, XfRec* __restrict__ xf) {
    int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= B) return;
    float thr = -INFINITY;                         // no under-bound list unless the bound is finite
    if (bound) {
        double bd = bound[b];
        if (bd < (double)INFINITY) thr = __double2float_ru(bd + tau);
    }
    xf[b] = make_xf(grids[grid_of_b ? grid_of_b[b] : 0], T_rel + 12 * b, thr);
}

void launch_prep_xf(const MinArgs& a, cudaStream_t st) {
    if (a.B <= 0) return;
    int threads = 128;
    int64_t blocks = (a.B + threads - 1) / threads;
    k_prep_xf<<<(unsigned)blocks, threads, 0, st>>>(a.grids, a.grid_of_b, a.T_rel, a.under_bound, a.B, a.tau, a.xf);
}

// One warp, one transform, the warp's 128 points (4 per lane): trilinear values, candidates for the
// under-bound set, and the warp's float32 minimum (returned in every lane).
// kInside: the caller has shown that all 128 points lie inside the sample lattice, so neither the
// clamps nor the bbox-overshoot distance are needed -- 9 FMA transform, 3 x (FADD.RZ, FADD, FADD)
// floor/fraction, 4 integer ops of addressing, one LDG.256, 7 FMA Horner per query.
template <bool kInside>
__device__ __forceinline__ float warp_subtile_min(const XfRec& r, const GridK& g, const float (&px)[kPtsPerThread],
                                                  const float (&py)[kPtsPerThread], const float (&pz)[kPtsPerThread],
                                                  int32_t b, int32_t base, int lane, Cand* __restrict__ cand,
                                                  unsigned long long* __restrict__ cand_count, uint32_t cand_cap) {
    float d[kPtsPerThread];
#pragma unroll
    for (int j = 0; j < kPtsPerThread; ++j) {
        float ux = fmaf(r.m[0], px[j], fmaf(r.m[1], py[j], fmaf(r.m[2], pz[j], r.m[3])));
        float uy = fmaf(r.m[4], px[j], fmaf(r.m[5], py[j], fmaf(r.m[6], pz[j], r.m[7])));
        float uz = fmaf(r.m[8], px[j], fmaf(r.m[9], py[j], fmaf(r.m[10], pz[j], r.m[11])));
        d[j] = kInside ? lattice_value(g, ux, uy, uz) : clamped_value(g, ux, uy, uz);
    }
    float best = fminf(fminf(d[0], d[1]), fminf(d[2], d[3]));
    static_assert(kPtsPerThread == 4, "min tree assumes 4 points per thread");
    if (best < r.thr) {                                    // rare: candidates for the under-bound set
#pragma unroll
        for (int j = 0; j < kPtsPerThread; ++j) {
            if (d[j] < r.thr) {
                // the counter keeps counting past the capacity: the host entry points read it, grow the buffer and re-run
                const unsigned long long slot = atomicAdd(cand_count, 1ull);
                if (slot < (unsigned long long)cand_cap) {
                    Cand c; c.b = b; c.idx = base + j * 32 + lane; c.d = (double)d[j];
                    cand[slot] = c;
                }
            }
        }
    }
    return warp_min_f32(best);
}

// The same for a sub-tile known to lie inside the sample lattice, two points per instruction: the transform of a point
// PAIR is 9 FFMA2 (matrix entries broadcast), floor / fraction 3 x (FADD2.RZ, FADD2, FADD2), and each point's Horner
// evaluation 3 FFMA2 + 1 FFMA on the brick's register pairs -- 13 floating-point instructions per query instead of 25,
// with bit-identical results (each half is the same IEEE fma / add as the scalar path).
__device__ __forceinline__ void subtile_values_packed(const XfRec& r, const GridK& g, const f32x2 (&PX)[kPtsPerThread / 2],
                                                      const f32x2 (&PY)[kPtsPerThread / 2], const f32x2 (&PZ)[kPtsPerThread / 2],
                                                      float (&d)[kPtsPerThread]) {
    const float big = 8388608.0f;
    const f32x2 BIG = pk2(big, big);
#pragma unroll
    for (int h = 0; h < kPtsPerThread / 2; ++h) {
        const f32x2 UX = fma2(pk2(r.m[0], r.m[0]), PX[h], fma2(pk2(r.m[1], r.m[1]), PY[h], fma2(pk2(r.m[2], r.m[2]), PZ[h], pk2(r.m[3], r.m[3]))));
        const f32x2 UY = fma2(pk2(r.m[4], r.m[4]), PX[h], fma2(pk2(r.m[5], r.m[5]), PY[h], fma2(pk2(r.m[6], r.m[6]), PZ[h], pk2(r.m[7], r.m[7]))));
        const f32x2 UZ = fma2(pk2(r.m[8], r.m[8]), PX[h], fma2(pk2(r.m[9], r.m[9]), PY[h], fma2(pk2(r.m[10], r.m[10]), PZ[h], pk2(r.m[11], r.m[11]))));
        const f32x2 TX = add2_rz(UX, BIG), TY = add2_rz(UY, BIG), TZ = add2_rz(UZ, BIG);
        const f32x2 FX = sub2(UX, sub2(TX, BIG)), FY = sub2(UY, sub2(TY, BIG)), FZ = sub2(UZ, sub2(TZ, BIG));
        float tx[2], ty[2], tz[2], fx[2], fy[2], fz[2];
        upk2(TX, tx[0], tx[1]); upk2(TY, ty[0], ty[1]); upk2(TZ, tz[0], tz[1]);
        upk2(FX, fx[0], fx[1]); upk2(FY, fy[0], fy[1]); upk2(FZ, fz[0], fz[1]);
        Cell2 q[2];
#pragma unroll
        for (int e = 0; e < 2; ++e)
            q[e] = load_cell2(g.brick, __float_as_uint(tx[e]) * g.cx + (__float_as_uint(ty[e]) * g.cz + (__float_as_uint(tz[e]) + g.nbias)));
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const f32x2 FZs = pk2(fz[e], fz[e]), FYs = pk2(fy[e], fy[e]);
            const f32x2 AC = fma2(FZs, q[e].p73, q[e].p62);          // (a, c): coefficients of fx fy and of fy
            const f32x2 BD = fma2(FZs, q[e].p51, q[e].p40);          // (b, d): coefficient of fx and the constant
            float ee, ff;
            upk2(fma2(FYs, AC, BD), ee, ff);                          // (a fy + b, c fy + d)
            d[2 * h + e] = fmaf(ee, fx[e], ff);
        }
    }
}

__device__ __forceinline__ float warp_subtile_min_packed(const XfRec& r, const GridK& g, const f32x2 (&PX)[kPtsPerThread / 2],
                                                         const f32x2 (&PY)[kPtsPerThread / 2], const f32x2 (&PZ)[kPtsPerThread / 2],
                                                         int32_t b, int32_t base, int lane, Cand* __restrict__ cand,
                                                         unsigned long long* __restrict__ cand_count, uint32_t cand_cap) {
    float d[kPtsPerThread];
    subtile_values_packed(r, g, PX, PY, PZ, d);
    float best = fminf(fminf(d[0], d[1]), fminf(d[2], d[3]));
    if (best < r.thr) {                                    // rare: candidates for the under-bound set
#pragma unroll
        for (int j = 0; j < kPtsPerThread; ++j) {
            if (d[j] < r.thr) {
                const unsigned long long slot = atomicAdd(cand_count, 1ull);
                if (slot < (unsigned long long)cand_cap) {
                    Cand c; c.b = b; c.idx = base + j * 32 + lane; c.d = (double)d[j];
                    cand[slot] = c;
                }
            }
        }
    }
    return warp_min_f32(best);
}

// ---- K2 bulk pass ---------------------------------------------------------------------------------
// grid  = nchunks * ntiles CTAs (chunk-major: consecutive CTAs share the transform chunk); a chunk is
//         `xfchunk` <= 32 transforms, chosen by the launcher so that the CTA count is many waves
// output: subkeys[b_local * nsub + (tile*8 + warp)] = bits of the float32 min over the warp's 128 points
// kMulti = false: every transform uses the one grid passed by value in `gk` (uniform registers);
// kMulti = true : the grid comes with each transform record (robot links).
// The cloud is padded to a whole tile with copies of its last point, so no per-point bounds test.
// Lane l of every warp tests the warp's bounding sphere (rep[sub]) against the lattice of transform l
// of the chunk; the ballot is the per-transform choice between the unclamped and the general path.
template <bool kMulti>
__global__ void __launch_bounds__(kThreads, 4)
k_min_f32(const float4* __restrict__ pts, const float4* __restrict__ rep, const GridDev* __restrict__ grids,
          const int32_t* __restrict__ grid_of_b, const double* __restrict__ T_rel, const double* __restrict__ bound, double tau,
          int64_t b_begin, int64_t Bc, int32_t ntiles, int32_t xfchunk, uint32_t* __restrict__ subkeys, Cand* __restrict__ cand,
          unsigned long long* __restrict__ cand_count, uint32_t cand_cap, GridK gk) {
    __shared__ XfRec sxf[kXfChunk];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t chunk = blockIdx.x / ntiles;
    const int32_t tile = (int32_t)(blockIdx.x - chunk * ntiles);
    const int64_t b0 = chunk * xfchunk;                  // batch-local index of the chunk's first transform
    const int nb = (int)min((int64_t)xfchunk, Bc - b0);
    if (tid < nb) {
        // fold this chunk's float64 transforms into float32 records in the prologue (no separate
        // launch): thread t builds the record of transform t
        const int64_t b = b_begin + b0 + tid;
        float thr = -INFINITY;                       // no under-bound list unless the bound is finite
        if (bound) {
            double bd = bound[b];
            if (bd < (double)INFINITY) thr = __double2float_ru(bd + tau);
        }
        sxf[tid] = make_xf(grids[grid_of_b ? grid_of_b[b] : 0], T_rel + 12 * b, thr);
    }
    // this warp's 128 contiguous points: lane l holds base + j*32 + l
    const int64_t sub = (int64_t)tile * (kThreads / 32) + warp;
    const int64_t base = sub * (32 * kPtsPerThread);
    float px[kPtsPerThread], py[kPtsPerThread], pz[kPtsPerThread];
#pragma unroll
    for (int j = 0; j < kPtsPerThread; ++j) {
        float4 p = __ldg(pts + base + j * 32 + lane);
        px[j] = p.x; py[j] = p.y; pz[j] = p.z;
    }
    f32x2 PX[kPtsPerThread / 2], PY[kPtsPerThread / 2], PZ[kPtsPerThread / 2];
#pragma unroll
    for (int h = 0; h < kPtsPerThread / 2; ++h) {
        PX[h] = pk2(px[2 * h], px[2 * h + 1]); PY[h] = pk2(py[2 * h], py[2 * h + 1]); PZ[h] = pk2(pz[2 * h], pz[2 * h + 1]);
    }
    const float4 sphere = __ldg(rep + sub);
    __syncthreads();
    const uint32_t inmask = __ballot_sync(0xffffffffu, lane < nb && sphere_in_lattice(sxf[lane < nb ? lane : 0], sphere));
    const int64_t nsub = (int64_t)ntiles * (kThreads / 32);
    float mine = INFINITY;                               // lane l keeps the minimum under transform l of the chunk
    for (int bl = 0; bl < nb; ++bl) {
        const XfRec& r = sxf[bl];
        const GridK g = kMulti ? grid_of(r) : gk;
        const int32_t b = (int32_t)(b_begin + b0 + bl);
        float m;
        if ((inmask >> bl) & 1u) m = warp_subtile_min_packed(r, g, PX, PY, PZ, b, (int32_t)base, lane, cand, cand_count, cand_cap);
        else m = warp_subtile_min<false>(r, g, px, py, pz, b, (int32_t)base, lane, cand, cand_count, cand_cap);
        mine = (lane == bl) ? m : mine;
    }
    if (lane < nb) subkeys[(b0 + lane) * nsub + sub] = __float_as_uint(mine);
}

static GridK host_gridk(const GridDev& g) {
    GridK k;
    k.brick = g.brick; k.cx = (uint32_t)(g.ny * g.nz); k.cz = (uint32_t)g.nz;
    k.nbias = 0u - kFloorBias * (k.cx + k.cz + 1u);
    k.hx = (float)g.h[0]; k.hy = (float)g.h[1]; k.hz = (float)g.h[2];
    k.hix = (float)(g.nx - 1); k.hiy = (float)(g.ny - 1); k.hiz = (float)(g.nz - 1);
    return k;
}

// transforms per CTA: as many as possible (the cloud tile is re-read once per chunk) while the launch
// still has >= 6 waves of 148 SMs x 4 resident CTAs (measured best on the C2 sweep), so the last partial wave costs little
int min_f32_chunk(int64_t Bc, int32_t ntiles) {
    static const int forced = getenv("SIO_MIN_CHUNK") ? atoi(getenv("SIO_MIN_CHUNK")) : 0;   // tuning aid
    if (forced > 0) return forced < kXfChunk ? forced : kXfChunk;
    int chunk = kXfChunk;
    while (chunk > 4 && ((Bc + chunk - 1) / chunk) * (int64_t)ntiles < 6 * 148 * 4) chunk >>= 1;
    return chunk;
}

void launch_min_f32(const MinArgs& a, cudaStream_t st) {
    if (a.Bc <= 0 || a.N <= 0) return;
    const int chunk = min_f32_chunk(a.Bc, a.ntiles);
    int64_t nchunks = (a.Bc + chunk - 1) / chunk;
    int64_t blocks = nchunks * a.ntiles;
    if (a.single_grid) {
        k_min_f32<false><<<(unsigned)blocks, kThreads, 0, st>>>(a.pts, a.rep, a.grids, a.grid_of_b, a.T_rel, a.under_bound, a.tau, a.b_begin,
                                                                a.Bc, a.ntiles, chunk, a.subkeys, a.cand, a.cand_count, a.cand_cap,
                                                                host_gridk(*a.single_grid));
    } else {
        GridK dummy{};
        k_min_f32<true><<<(unsigned)blocks, kThreads, 0, st>>>(a.pts, a.rep, a.grids, a.grid_of_b, a.T_rel, a.under_bound, a.tau, a.b_begin,
                                                               a.Bc, a.ntiles, chunk, a.subkeys, a.cand, a.cand_count, a.cand_cap, dummy);
    }
}

// ---- K2 bulk pass, grid staged in shared memory by TMA (SIO_SMEM_GRID; measured alternative, not the default) --------------
// BASELINE.json's north_star names "TMA or shared-memory staging of grid bricks".  The 32-byte bricks of even the small
// grids (C1 / C5 cube at 0.05: 27,000 cells = 864 KB) do not fit a CTA's 227 KB, but the plain padded sample array does
// (31^3 floats = 119 KB).  This variant therefore stages THAT array with one bulk asynchronous copy
// (cp.async.bulk.shared::cluster.global + mbarrier complete_tx: the TMA engine, SASS UBLKCP) per persistent CTA and
// evaluates every query from its 8 corner samples with shared-memory loads.  Same decomposition and the same sub-tile
// minima as k_min_f32 (the finalize pass is shared), general clamped path for every point.
// Result (profiles/r02_smem_vs_brick.md): slower than the brick kernel -- a query costs 8 LDS (each at least one
// data-stage wavefront, more under bank conflicts) against one 256-bit LDG, through the SAME L1 data pipe that binds.
constexpr int kSmemThreads = 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(kSmemThreads, 1)
k_min_smem(const float4* __restrict__ pts, const XfRec* __restrict__ xf, int64_t b_begin, int64_t Bc, int64_t nsub, int32_t xfchunk,
           uint32_t* __restrict__ subkeys, const float* __restrict__ gv, uint32_t n_floats, int32_t sx, int32_t sy) {
    extern __shared__ __align__(128) float sv[];
    __shared__ __align__(8) unsigned long long mbar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bulk_bytes = (n_floats * 4u) & ~15u;            // bulk copies move multiples of 16 bytes
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(bulk_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(sv)), "l"(gv), "r"(bulk_bytes), "r"(smem_u32(&mbar)) : "memory");
    }
    for (uint32_t i = bulk_bytes / 4 + tid; i < n_floats; i += kSmemThreads) sv[i] = __ldg(gv + i);     // the < 16-byte tail
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_u32(&mbar)) : "memory");
    }
    __syncthreads();
    const int64_t nchunks = (Bc + xfchunk - 1) / xfchunk;
    const int64_t nitems = nchunks * nsub;                          // one item = (transform chunk, 128-point sub-tile), one warp
    for (int64_t item = (int64_t)blockIdx.x * (kSmemThreads / 32) + warp; item < nitems; item += (int64_t)gridDim.x * (kSmemThreads / 32)) {
        const int64_t chunk = item / nsub, sub = item - chunk * nsub;
        const int64_t b0 = chunk * xfchunk;
        const int nb = (int)min((int64_t)xfchunk, Bc - b0);
        const int64_t base = sub * (32 * kPtsPerThread);
        float px[kPtsPerThread], py[kPtsPerThread], pz[kPtsPerThread];
#pragma unroll
        for (int j = 0; j < kPtsPerThread; ++j) {
            float4 p = __ldg(pts + base + j * 32 + lane);
            px[j] = p.x; py[j] = p.y; pz[j] = p.z;
        }
        float mine = INFINITY;
        for (int bl = 0; bl < nb; ++bl) {
            const XfRec* r = xf + b_begin + b0 + bl;                // warp-uniform address: one broadcast load per field
            float m[12];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(r->m) + q);
                m[4 * q] = v.x; m[4 * q + 1] = v.y; m[4 * q + 2] = v.z; m[4 * q + 3] = v.w;
            }
            const float4 h4 = __ldg(reinterpret_cast<const float4*>(&r->hx));      // hx hy hz thr
            const float4 d4 = __ldg(reinterpret_cast<const float4*>(&r->hix));     // hix hiy hiz (sx)
            float best = INFINITY;
#pragma unroll
            for (int j = 0; j < kPtsPerThread; ++j) {
                const float ux = fmaf(m[0], px[j], fmaf(m[1], py[j], fmaf(m[2], pz[j], m[3])));
                const float uy = fmaf(m[4], px[j], fmaf(m[5], py[j], fmaf(m[6], pz[j], m[7])));
                const float uz = fmaf(m[8], px[j], fmaf(m[9], py[j], fmaf(m[10], pz[j], m[11])));
                const float bx = fminf(fmaxf(ux, -0.5f), d4.x + 0.5f), by = fminf(fmaxf(uy, -0.5f), d4.y + 0.5f), bz = fminf(fmaxf(uz, -0.5f), d4.z + 0.5f);
                const float ox = (ux - bx) * h4.x, oy = (uy - by) * h4.y, oz = (uz - bz) * h4.z;
                const float dout2 = fmaf(ox, ox, fmaf(oy, oy, oz * oz));
                const float cx = fminf(fmaxf(ux, 0.0f), d4.x), cy = fminf(fmaxf(uy, 0.0f), d4.y), cz = fminf(fmaxf(uz, 0.0f), d4.z);
                uint32_t ix, iy, iz; float fx, fy, fz;
                split_rz(cx, ix, fx); split_rz(cy, iy, fy); split_rz(cz, iz, fz);
                const float* c = sv + ((ix - kFloorBias) * (uint32_t)sx + (iy - kFloorBias) * (uint32_t)sy + (iz - kFloorBias));
                // the padded copy makes corner +1 addressable on every axis (its last layer repeats the one before)
                const float v000 = c[0], v001 = c[1], v010 = c[sy], v011 = c[sy + 1];
                const float v100 = c[sx], v101 = c[sx + 1], v110 = c[sx + sy], v111 = c[sx + sy + 1];
                const float a00 = fmaf(fz, v001 - v000, v000), a01 = fmaf(fz, v011 - v010, v010);
                const float a10 = fmaf(fz, v101 - v100, v100), a11 = fmaf(fz, v111 - v110, v110);
                const float b0v = fmaf(fy, a01 - a00, a00), b1v = fmaf(fy, a11 - a10, a10);
                float dout;
                asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(dout) : "f"(dout2));
                best = fminf(best, fmaf(fx, b1v - b0v, b0v) + dout);
            }
            const float mw = warp_min_f32(best);
            mine = (lane == bl) ? mw : mine;
        }
        if (lane < nb) subkeys[(b0 + lane) * nsub + sub] = __float_as_uint(mine);
    }
}

// returns false when the variant does not apply (more than one grid, or the padded sample array exceeds the shared memory)
bool launch_min_smem(const MinArgs& a, cudaStream_t st) {
    if (!a.single_grid || a.Bc <= 0 || a.N <= 0) return false;
    const GridDev& g = *a.single_grid;
    const size_t n_floats = (size_t)(g.nx + 1) * (g.ny + 1) * (g.nz + 1);
    const size_t bytes = ((n_floats * 4 + 127) / 128) * 128;
    if (bytes > 200 * 1024) return false;
    // per device and cheap: set on every launch rather than remembered (a process may drive several devices)
    if (cudaFuncSetAttribute(k_min_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return false;
    const int64_t nsub = (int64_t)a.ntiles * (kThreads / 32);
    k_min_smem<<<a.num_sms > 0 ? a.num_sms : 148, kSmemThreads, bytes, st>>>(a.pts, a.xf, a.b_begin, a.Bc, nsub, kXfChunk, a.subkeys, g.v,
                                                                             (uint32_t)n_floats, g.sx, g.sy);
    return true;
}

// ---- K2 with sub-tile culling (SIO_PRUNE) -----------------------------------------------------------
// The GPU counterpart of the bounding-hierarchy branch-and-bound inside Klampt's distance_ext: each
// 128-point sub-tile has a representative cloud point and a bounding radius (built once with the
// cloud).  With L a Lipschitz constant of the interpolated field (from the grid's own samples; +1
// where the sphere leaves the bbox, because the bbox-overshoot term is 1-Lipschitz):
//     every point of sub-tile s has  d(b, .) >= d(b, rep_s) - L r_s      (lower bound)
//     min_i d(b, i)  <=  min_s d(b, rep_s)                               (upper bound: reps are cloud points)
// so a sub-tile whose lower bound exceeds max(upper bound, bound_b) + 2 tau can contain neither the
// minimum nor an under-bound point and is skipped; its key is set to +inf.  Results are identical to
// the un-pruned sweep.
//   k_prune_centres : lb[b][s] and ub[b]       (one query per (transform, sub-tile))
//   k_prune_select  : survivors -> work items  (compaction)
//   k_min_items     : one warp per surviving (transform, sub-tile), persistent grid
// lb is kept sub-tile-major, lb[s * Bp + b] (Bp = the chunk's transforms padded to whole warps): k_prune_select walks it
// in that order so that the survivors it appends form runs sharing their points.  The CTA's 256 sub-tiles x 32
// transforms are transposed through shared memory so that both kernels touch lb in 128-byte rows.  Every sub-tile key
// starts at +inf here (the finalize pass never opens a sub-tile k_min_items did not evaluate).
__global__ void __launch_bounds__(kThreads)
k_prune_centres(const float4* __restrict__ rep, int64_t nsub, const XfRec* __restrict__ xf, int64_t b_begin, int64_t Bc, int32_t xfchunk,
                float* __restrict__ lb, uint32_t* __restrict__ ub, uint32_t* __restrict__ subkeys) {
    __shared__ XfRec sxf[kXfChunk];
    __shared__ float tile[kThreads][kXfChunk + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t ntl = (nsub + kThreads - 1) / kThreads;
    const int64_t chunk = blockIdx.x / ntl;
    const int64_t tl = blockIdx.x - chunk * ntl;
    const int64_t b0 = chunk * xfchunk;                    // xfchunk <= kXfChunk transforms per CTA (fewer when the sweep is small)
    const int64_t Bp = (Bc + 31) & ~(int64_t)31;
    const int nb = (int)min((int64_t)xfchunk, Bc - b0);
    {
        const uint4* src = reinterpret_cast<const uint4*>(xf + b_begin + b0);
        uint4* dst = reinterpret_cast<uint4*>(sxf);
        for (int i = tid; i < nb * (int)(sizeof(XfRec) / 16); i += kThreads) dst[i] = src[i];
    }
    const int64_t s = tl * kThreads + tid;
    const bool ok = s < nsub;
    float4 p = ok ? __ldg(rep + s) : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    for (int bl = 0; bl < nb; ++bl) {
        const XfRec& r = sxf[bl];
        const GridK g = grid_of(r);
        float ux = fmaf(r.m[0], p.x, fmaf(r.m[1], p.y, fmaf(r.m[2], p.z, r.m[3])));
        float uy = fmaf(r.m[4], p.x, fmaf(r.m[5], p.y, fmaf(r.m[6], p.z, r.m[7])));
        float uz = fmaf(r.m[8], p.x, fmaf(r.m[9], p.y, fmaf(r.m[10], p.z, r.m[11])));
        float dc = clamped_value(g, ux, uy, uz);
        // distance (metres) from the centre to the nearest bbox face, negative outside
        float mx = fminf(ux + 0.5f, g.hix + 0.5f - ux) * g.hx;
        float my = fminf(uy + 0.5f, g.hiy + 0.5f - uy) * g.hy;
        float mz = fminf(uz + 0.5f, g.hiz + 0.5f - uz) * g.hz;
        float L = local_lipschitz(r, g, ux, uy, uz, p.w) + ((min3(mx, my, mz) >= p.w) ? 0.0f : 1.0f);
        tile[tid][bl] = dc - L * p.w;
        if (ok) subkeys[(b0 + bl) * nsub + s] = 0x7f800000u;
        uint32_t k = __reduce_min_sync(0xffffffffu, ok ? fkey(dc) : 0xffffffffu);
        if (lane == 0) atomicMin(ub + b_begin + b0 + bl, k);
    }
    __syncthreads();
    for (int row = warp; row < kThreads; row += kThreads / 32) {
        const int64_t sr = tl * kThreads + row;
        if (sr < nsub && lane < nb) lb[sr * Bp + b0 + lane] = tile[row][lane];
    }
}

// sub-tile-major enumeration: the threads of a warp look at ONE sub-tile under 32 consecutive transforms, so the
// survivors they append (one atomic per warp -> consecutive slots) form runs that share their points
constexpr int kSelectRows = 16;   // sub-tiles one CTA of k_prune_select walks for its 256 transforms
__global__ void __launch_bounds__(256)
k_prune_select(const float* __restrict__ lb, const uint32_t* __restrict__ ub, const double* __restrict__ bound,
               int64_t nsub, int64_t b_begin, int64_t Bc, float tau, uint32_t* __restrict__ items,
               uint32_t* __restrict__ item_count) {
    const int64_t Bp = (Bc + 31) & ~(int64_t)31;            // transforms padded to whole warps
    const int64_t bl = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool valid = bl < Bc;                             // no early exit: the CTA synchronises below
    float thr = -INFINITY;                                  // this transform's pruning threshold, once
    if (valid) {
        const int64_t b = b_begin + bl;
        thr = fkey_inv(ub[b]);
        if (bound) {
            double bd = bound[b];
            if (bd < (double)INFINITY) thr = fmaxf(thr, __double2float_ru(bd));
        }
        thr += 2.0f * tau;
    }
    // survivors of the CTA's 16 sub-tiles x 256 transforms go out sub-tile by sub-tile behind ONE atomic: per-warp
    // atomics on the single counter serialise in L2 (2 M of them per chunk were the whole cost of this kernel)
    __shared__ uint32_t cnt[kSelectRows * 8];
    __shared__ uint32_t cta_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t s0 = blockIdx.y * (int64_t)kSelectRows;
    uint32_t masks[kSelectRows];
#pragma unroll
    for (int r = 0; r < kSelectRows; ++r) {
        const int64_t s = s0 + r;
        const bool keep = valid && s < nsub && lb[s * Bp + bl] <= thr;
        masks[r] = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) cnt[r * 8 + warp] = __popc(masks[r]);
    }
    __syncthreads();
    if (warp == 0) {                                        // exclusive scan of the 128 counts, row-major
        uint32_t v[4], sum = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) { v[j] = cnt[lane * 4 + j]; sum += v[j]; }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        uint32_t run = incl - sum;
#pragma unroll
        for (int j = 0; j < 4; ++j) { cnt[lane * 4 + j] = run; run += v[j]; }
        if (lane == 31) cta_base = incl ? atomicAdd(item_count, incl) : 0u;
    }
    __syncthreads();
    const uint32_t base = cta_base;
#pragma unroll
    for (int r = 0; r < kSelectRows; ++r)
        if (masks[r] >> lane & 1u)
            items[base + cnt[r * 8 + warp] + __popc(masks[r] & ((1u << lane) - 1u))] = (uint32_t)(bl * nsub + (s0 + r));
}

constexpr int kItemRun = 16;     // consecutive work items taken by one warp (they mostly share a sub-tile)

__global__ void __launch_bounds__(kThreads, 4)
k_min_items(const float4* __restrict__ pts, const float4* __restrict__ rep, const XfRec* __restrict__ xf, int64_t b_begin,
            int64_t nsub, const uint32_t* __restrict__ items, const uint32_t* __restrict__ item_count,
            uint32_t* __restrict__ subkeys, Cand* __restrict__ cand, unsigned long long* __restrict__ cand_count, uint32_t cand_cap) {
    __shared__ XfRec sxf[kThreads / 32][kItemRun];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n = *item_count;
    const uint32_t nwarps = gridDim.x * (kThreads / 32);
    const uint32_t nsub32 = (uint32_t)nsub;
    constexpr int kRecParts = (int)(sizeof(XfRec) / 16);
    float px[kPtsPerThread], py[kPtsPerThread], pz[kPtsPerThread];
    f32x2 PX[kPtsPerThread / 2], PY[kPtsPerThread / 2], PZ[kPtsPerThread / 2];      // the same points as register pairs
    float4 sphere = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t have = 0xffffffffu;                           // sub-tile whose points sit in registers
    // runs of up to kItemRun items; shorter when there are few items per warp (a 64-pose sweep leaves 4 per warp: at
    // 16 to a run three warps in four would idle while the fourth walks its run one dependent item after the other)
    const uint32_t run = max(1u, min((uint32_t)kItemRun, (n + nwarps - 1) / nwarps));
    for (uint32_t it0 = (blockIdx.x * (kThreads / 32) + warp) * run; it0 < n; it0 += nwarps * run) {
        // the run's items and transform records arrive together: one dependent round trip per run, not per item
        const int cnt = (int)min(n - it0, run);
        const uint32_t e_l = lane < cnt ? __ldg(items + it0 + lane) : 0u;
        __syncwarp();
#pragma unroll
        for (int q0 = 0; q0 < kItemRun * kRecParts; q0 += 32) {
            if (q0 >= cnt * kRecParts) break;              // warp-uniform
            const int q = q0 + lane, rec = q / kRecParts;
            const uint32_t e = __shfl_sync(0xffffffffu, e_l, rec & (kItemRun - 1));
            if (rec < cnt)
                reinterpret_cast<uint4*>(&sxf[warp][rec])[q - rec * kRecParts] =
                    __ldg(reinterpret_cast<const uint4*>(xf + b_begin + e / nsub32) + (q - rec * kRecParts));
        }
        __syncwarp();
        for (int i = 0; i < cnt; ++i) {
            const uint32_t e = __shfl_sync(0xffffffffu, e_l, i);
            const uint32_t bl = e / nsub32;
            const uint32_t s = e - bl * nsub32;
            const XfRec& r = sxf[warp][i];
            const int64_t base = (int64_t)s * (32 * kPtsPerThread);
            if (s != have) {
#pragma unroll
                for (int j = 0; j < kPtsPerThread; ++j) {
                    float4 p = __ldg(pts + base + j * 32 + lane);
                    px[j] = p.x; py[j] = p.y; pz[j] = p.z;
                }
#pragma unroll
                for (int h = 0; h < kPtsPerThread / 2; ++h) {
                    PX[h] = pk2(px[2 * h], px[2 * h + 1]); PY[h] = pk2(py[2 * h], py[2 * h + 1]); PZ[h] = pk2(pz[2 * h], pz[2 * h + 1]);
                }
                sphere = __ldg(rep + s);
                have = s;
            }
            float m;
            if (sphere_in_lattice(r, sphere))      // warp-uniform: every lane tests the same sphere and record
                m = warp_subtile_min_packed(r, grid_of(r), PX, PY, PZ, (int32_t)(b_begin + bl), (int32_t)base, lane, cand, cand_count, cand_cap);
            else
                m = warp_subtile_min<false>(r, grid_of(r), px, py, pz, (int32_t)(b_begin + bl), (int32_t)base, lane, cand, cand_count, cand_cap);
            if (lane == 0) subkeys[e] = __float_as_uint(m);
        }
    }
}

void launch_min_pruned(const MinArgs& a, cudaStream_t st) {
    if (a.Bc <= 0 || a.N <= 0) return;
    const int64_t nsub = (int64_t)a.ntiles * (kThreads / 32);
    const int64_t ntl = (nsub + kThreads - 1) / kThreads;
    // transforms per CTA: the full chunk when the sweep is large; a small sweep (C2: 64 poses x 8,192 sub-tiles = 64 CTAs
    // at 32 per CTA, 25 us of latency for 0.8 % of the bulk pass's queries) is cut finer until it fills the machine
    int xfchunk = kXfChunk;
    const int sms = a.num_sms > 0 ? a.num_sms : 148;
    while (xfchunk > 4 && ((a.Bc + xfchunk - 1) / xfchunk) * ntl < 4 * (int64_t)sms) xfchunk >>= 1;
    const int64_t nchunks = (a.Bc + xfchunk - 1) / xfchunk;
    k_prune_centres<<<(unsigned)(nchunks * ntl), kThreads, 0, st>>>(a.rep, nsub, a.xf, a.b_begin, a.Bc, xfchunk, a.lb, a.ub, a.subkeys);
    const int64_t Bp = (a.Bc + 31) & ~(int64_t)31;
    k_prune_select<<<dim3((unsigned)((Bp + 255) / 256), (unsigned)((nsub + kSelectRows - 1) / kSelectRows)), 256, 0, st>>>(
        a.lb, a.ub, a.under_bound, nsub, a.b_begin, a.Bc, (float)a.tau, a.items, a.item_count);
    k_min_items<<<(a.num_sms > 0 ? a.num_sms : 148) * 4, kThreads, 0, st>>>(a.pts, a.rep, a.xf, a.b_begin, nsub, a.items, a.item_count, a.subkeys, a.cand,
                                              a.cand_count, a.cand_cap);
}

// ---- K1 over a resident cloud: per-point outputs ----------------------------------------------------
// out_d[b*N + i] = d, or out_g[b*N + i] = (gx, gy, gz, d); i = internal (Morton) index.
// Same decomposition as the K2 bulk pass: a warp holds 128 points (4 per lane), walks the chunk's transforms, and
// its bounding sphere decides per transform between the unclamped path (value: 7 FMA Horner; gradient: 3 more
// dot products reusing the Horner partial sums, then M' g and one rsqrt) and the general path.  Stores are
// streaming (st.global.cs): 4 B or 16 B per query, the kernel's HBM traffic.
__device__ __forceinline__ float lattice_value_grad(const XfRec& r, const GridK& g, float ux, float uy, float uz, int normalize,
                                                    float& gx, float& gy, float& gz) {
    uint32_t bx, by, bz; float fx, fy, fz;
    split_rz(ux, bx, fx); split_rz(uy, by, fy); split_rz(uz, bz, fz);
    Cell q = load_cell(g.brick, bx * g.cx + (by * g.cz + (bz + g.nbias)));
    float a = fmaf(q.c[2], fz, q.c[0]);
    float b = fmaf(q.c[6], fz, q.c[4]);
    float c = fmaf(q.c[3], fz, q.c[1]);
    float d = fmaf(q.c[7], fz, q.c[5]);
    float gfx = fmaf(a, fy, b);                                               // dv/dfx
    float val = fmaf(gfx, fx, fmaf(c, fy, d));
    float gfy = fmaf(a, fx, c);
    float gfz = fmaf(fmaf(q.c[2], fy, q.c[6]), fx, fmaf(q.c[3], fy, q.c[7]));
    // strictly inside the lattice: no clamped axis, no bbox term.  g_in = M' gf
    gx = fmaf(r.m[0], gfx, fmaf(r.m[4], gfy, r.m[8] * gfz));
    gy = fmaf(r.m[1], gfx, fmaf(r.m[5], gfy, r.m[9] * gfz));
    gz = fmaf(r.m[2], gfx, fmaf(r.m[6], gfy, r.m[10] * gfz));
    if (normalize) {
        float n2 = fmaf(gx, gx, fmaf(gy, gy, gz * gz));
        if (n2 > 0.0f) { float s = rsqrtf(n2); gx *= s; gy *= s; gz *= s; }
    }
    return val;
}

// lattice_value_grad for a sub-tile inside the sample lattice, two points per instruction where the operands allow: the
// transform and floor / fraction of a point PAIR as in subtile_values_packed; per point the brick's register pairs give
// (a, c), (b, d), (a fy + b, c fy + d) and (c2 fy + c6, c3 fy + c7) in four FFMA2, then value, dv/dfy and dv/dfz in three
// FFMA; M' gf and the normalisation again per pair.  Each half is the IEEE operation of the scalar routine, in its order.
__device__ __forceinline__ void subtile_grad_packed(const XfRec& r, const GridK& g, const f32x2 (&PX)[kPtsPerThread / 2],
                                                    const f32x2 (&PY)[kPtsPerThread / 2], const f32x2 (&PZ)[kPtsPerThread / 2],
                                                    int normalize, float4 (&out)[kPtsPerThread]) {
    const float big = 8388608.0f;
    const f32x2 BIG = pk2(big, big);
#pragma unroll
    for (int h = 0; h < kPtsPerThread / 2; ++h) {
        const f32x2 UX = fma2(pk2(r.m[0], r.m[0]), PX[h], fma2(pk2(r.m[1], r.m[1]), PY[h], fma2(pk2(r.m[2], r.m[2]), PZ[h], pk2(r.m[3], r.m[3]))));
        const f32x2 UY = fma2(pk2(r.m[4], r.m[4]), PX[h], fma2(pk2(r.m[5], r.m[5]), PY[h], fma2(pk2(r.m[6], r.m[6]), PZ[h], pk2(r.m[7], r.m[7]))));
        const f32x2 UZ = fma2(pk2(r.m[8], r.m[8]), PX[h], fma2(pk2(r.m[9], r.m[9]), PY[h], fma2(pk2(r.m[10], r.m[10]), PZ[h], pk2(r.m[11], r.m[11]))));
        const f32x2 TX = add2_rz(UX, BIG), TY = add2_rz(UY, BIG), TZ = add2_rz(UZ, BIG);
        const f32x2 FX = sub2(UX, sub2(TX, BIG)), FY = sub2(UY, sub2(TY, BIG)), FZ = sub2(UZ, sub2(TZ, BIG));
        float tx[2], ty[2], tz[2], fx[2], fy[2], fz[2];
        upk2(TX, tx[0], tx[1]); upk2(TY, ty[0], ty[1]); upk2(TZ, tz[0], tz[1]);
        upk2(FX, fx[0], fx[1]); upk2(FY, fy[0], fy[1]); upk2(FZ, fz[0], fz[1]);
        Cell2 q[2];
#pragma unroll
        for (int e = 0; e < 2; ++e)
            q[e] = load_cell2(g.brick, __float_as_uint(tx[e]) * g.cx + (__float_as_uint(ty[e]) * g.cz + (__float_as_uint(tz[e]) + g.nbias)));
        float val[2], gfx[2], gfy[2], gfz[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const f32x2 FZs = pk2(fz[e], fz[e]), FYs = pk2(fy[e], fy[e]);
            const f32x2 AC = fma2(FZs, q[e].p73, q[e].p62);          // (a, c)
            const f32x2 BD = fma2(FZs, q[e].p51, q[e].p40);          // (b, d)
            float a, c, ff, u, v;
            upk2(AC, a, c);
            upk2(fma2(FYs, AC, BD), gfx[e], ff);                      // (dv/dfx = a fy + b, c fy + d)
            upk2(fma2(FYs, q[e].p73, q[e].p51), u, v);                // (c2 fy + c6, c3 fy + c7) in the scalar routine's naming
            val[e] = fmaf(gfx[e], fx[e], ff);
            gfy[e] = fmaf(a, fx[e], c);
            gfz[e] = fmaf(u, fx[e], v);
        }
        const f32x2 GFX = pk2(gfx[0], gfx[1]), GFY = pk2(gfy[0], gfy[1]), GFZ = pk2(gfz[0], gfz[1]);
        f32x2 GX = fma2(pk2(r.m[0], r.m[0]), GFX, fma2(pk2(r.m[4], r.m[4]), GFY, mul2(pk2(r.m[8], r.m[8]), GFZ)));
        f32x2 GY = fma2(pk2(r.m[1], r.m[1]), GFX, fma2(pk2(r.m[5], r.m[5]), GFY, mul2(pk2(r.m[9], r.m[9]), GFZ)));
        f32x2 GZ = fma2(pk2(r.m[2], r.m[2]), GFX, fma2(pk2(r.m[6], r.m[6]), GFY, mul2(pk2(r.m[10], r.m[10]), GFZ)));
        float gx[2], gy[2], gz[2];
        if (normalize) {
            const f32x2 N2 = fma2(GX, GX, fma2(GY, GY, mul2(GZ, GZ)));
            float n2[2], sc[2];
            upk2(N2, n2[0], n2[1]);
#pragma unroll
            for (int e = 0; e < 2; ++e) sc[e] = n2[e] > 0.0f ? rsqrtf(n2[e]) : 1.0f;     // x * 1.0f == x: the scalar routine leaves a zero gradient alone
            const f32x2 S = pk2(sc[0], sc[1]);
            GX = mul2(GX, S); GY = mul2(GY, S); GZ = mul2(GZ, S);
        }
        upk2(GX, gx[0], gx[1]); upk2(GY, gy[0], gy[1]); upk2(GZ, gz[0], gz[1]);
#pragma unroll
        for (int e = 0; e < 2; ++e) out[2 * h + e] = make_float4(gx[e], gy[e], gz[e], val[e]);
    }
}

template <bool kGrad>
__global__ void __launch_bounds__(kThreads, 4)
k_cloud_query_f32(const float4* __restrict__ pts, const float4* __restrict__ rep, int64_t N, int32_t ntiles,
                  const XfRec* __restrict__ xf, int64_t B, int32_t xfchunk, int normalize, float* __restrict__ out_d,
                  float4* __restrict__ out_g) {
    __shared__ XfRec sxf[kXfChunk];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t chunk = blockIdx.x / ntiles;
    const int32_t tile = (int32_t)(blockIdx.x - chunk * ntiles);
    const int64_t b0 = chunk * xfchunk;
    const int nb = (int)min((int64_t)xfchunk, B - b0);
    {
        const uint4* src = reinterpret_cast<const uint4*>(xf + b0);
        uint4* dst = reinterpret_cast<uint4*>(sxf);
        for (int i = tid; i < nb * (int)(sizeof(XfRec) / 16); i += kThreads) dst[i] = src[i];
    }
    const int64_t sub = (int64_t)tile * (kThreads / 32) + warp;
    const int64_t base = sub * (32 * kPtsPerThread);       // the cloud is padded to a multiple of kTile
    float px[kPtsPerThread], py[kPtsPerThread], pz[kPtsPerThread];
#pragma unroll
    for (int j = 0; j < kPtsPerThread; ++j) {
        float4 p = __ldg(pts + base + j * 32 + lane);
        px[j] = p.x; py[j] = p.y; pz[j] = p.z;
    }
    f32x2 PX[kPtsPerThread / 2], PY[kPtsPerThread / 2], PZ[kPtsPerThread / 2];      // the same points as register pairs (sub-tiles inside the lattice)
#pragma unroll
    for (int h = 0; h < kPtsPerThread / 2; ++h) {
        PX[h] = pk2(px[2 * h], px[2 * h + 1]); PY[h] = pk2(py[2 * h], py[2 * h + 1]); PZ[h] = pk2(pz[2 * h], pz[2 * h + 1]);
    }
    const float4 sphere = __ldg(rep + sub);
    __syncthreads();
    const uint32_t inmask = __ballot_sync(0xffffffffu, lane < nb && sphere_in_lattice(sxf[lane < nb ? lane : 0], sphere));
    for (int bl = 0; bl < nb; ++bl) {
        const XfRec& r = sxf[bl];
        const GridK g = grid_of(r);
        const bool inside = (inmask >> bl) & 1u;
        const int64_t row = (b0 + bl) * N;
        if (kGrad && inside) {                             // packed value + gradient
            float4 o4[kPtsPerThread];
            subtile_grad_packed(r, g, PX, PY, PZ, normalize, o4);
#pragma unroll
            for (int j = 0; j < kPtsPerThread; ++j) {
                const int64_t i = base + j * 32 + lane;
                if (i < N) __stcs(out_g + row + i, o4[j]);
            }
            continue;
        }
        if (!kGrad && inside) {                            // the K2 bulk kernel's packed evaluation (two points per FFMA2), same values
            float d4[kPtsPerThread];
            subtile_values_packed(r, g, PX, PY, PZ, d4);
#pragma unroll
            for (int j = 0; j < kPtsPerThread; ++j) {
                const int64_t i = base + j * 32 + lane;
                if (i < N) __stcs(out_d + row + i, d4[j]);
            }
            continue;
        }
#pragma unroll
        for (int j = 0; j < kPtsPerThread; ++j) {
            const int64_t i = base + j * 32 + lane;
            float gx = 0.f, gy = 0.f, gz = 0.f, d;
            if (kGrad) {
                if (inside) {
                    float ux = fmaf(r.m[0], px[j], fmaf(r.m[1], py[j], fmaf(r.m[2], pz[j], r.m[3])));
                    float uy = fmaf(r.m[4], px[j], fmaf(r.m[5], py[j], fmaf(r.m[6], pz[j], r.m[7])));
                    float uz = fmaf(r.m[8], px[j], fmaf(r.m[9], py[j], fmaf(r.m[10], pz[j], r.m[11])));
                    d = lattice_value_grad(r, g, ux, uy, uz, normalize, gx, gy, gz);
                } else {
                    d = eval32_grad(r, px[j], py[j], pz[j], normalize, gx, gy, gz);
                }
                if (i < N) __stcs(out_g + row + i, make_float4(gx, gy, gz, d));
            } else {
                float ux = fmaf(r.m[0], px[j], fmaf(r.m[1], py[j], fmaf(r.m[2], pz[j], r.m[3])));
                float uy = fmaf(r.m[4], px[j], fmaf(r.m[5], py[j], fmaf(r.m[6], pz[j], r.m[7])));
                float uz = fmaf(r.m[8], px[j], fmaf(r.m[9], py[j], fmaf(r.m[10], pz[j], r.m[11])));
                d = inside ? lattice_value(g, ux, uy, uz) : clamped_value(g, ux, uy, uz);
                if (i < N) __stcs(out_d + row + i, d);
            }
        }
    }
}

void launch_cloud_query_f32(const float4* pts, const float4* rep, int64_t N, const XfRec* xf, int64_t B, int normalize,
                            float* out_d, float4* out_g, cudaStream_t st) {
    if (B <= 0 || N <= 0) return;
    const int32_t ntiles = (int32_t)((N + kTile - 1) / kTile);
    const int chunk = min_f32_chunk(B, ntiles);
    const int64_t nchunks = (B + chunk - 1) / chunk;
    if (out_g)
        k_cloud_query_f32<true><<<(unsigned)(ntiles * nchunks), kThreads, 0, st>>>(pts, rep, N, ntiles, xf, B, chunk, normalize, nullptr, out_g);
    else
        k_cloud_query_f32<false><<<(unsigned)(ntiles * nchunks), kThreads, 0, st>>>(pts, rep, N, ntiles, xf, B, chunk, normalize, out_d, nullptr);
}

// ---- K1 for caller-supplied float64 points, float32 arithmetic --------------------------------------
__global__ void k_query_f32(const GridDev* __restrict__ grid, const double* __restrict__ T,
                            const double* __restrict__ pts, int64_t P, int normalize,
                            double* __restrict__ d, double* __restrict__ grad) {
    __shared__ XfRec r;
    if (threadIdx.x == 0) r = make_xf(*grid, T, -INFINITY);
    __syncthreads();
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= P) return;
    float x = (float)pts[3 * i], y = (float)pts[3 * i + 1], z = (float)pts[3 * i + 2];
    float gx, gy, gz;
    d[i] = (double)eval32_grad(r, x, y, z, normalize, gx, gy, gz);
    if (grad) { grad[3 * i] = gx; grad[3 * i + 1] = gy; grad[3 * i + 2] = gz; }
}

void launch_query_f32(const GridDev* grid, const double* T, const double* pts, int64_t P, int normalize,
                      double* d, double* grad, cudaStream_t st) {
    if (P <= 0) return;
    int threads = 128;
    k_query_f32<<<(unsigned)((P + threads - 1) / threads), threads, 0, st>>>(grid, T, pts, P, normalize, d, grad);
}

// ---- gather microbenchmark (bench.py: the L2 gather roofline denominator) ----------------------------
__global__ void k_gather_bench(const float* __restrict__ arr, const uint32_t* __restrict__ idx,
                               int64_t n_gathers, float* __restrict__ sink) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    float acc = 0.f;
    for (; i + 3 * stride < n_gathers; i += 4 * stride) {
        uint32_t a = __ldg(idx + i), b = __ldg(idx + i + stride), c = __ldg(idx + i + 2 * stride), e = __ldg(idx + i + 3 * stride);
        acc += __ldg(arr + a) + __ldg(arr + b) + __ldg(arr + c) + __ldg(arr + e);
    }
    for (; i < n_gathers; i += stride) acc += __ldg(arr + __ldg(idx + i));
    if (acc == 123.456f) sink[0] = acc;      // keep the loads alive
}

void launch_gather_bench(const float* arr, int64_t n_elems, const uint32_t* idx, int64_t n_gathers,
                         float* sink, cudaStream_t st) {
    (void)n_elems;
    k_gather_bench<<<148 * 8, 256, 0, st>>>(arr, idx, n_gathers, sink);
}

}  // namespace sio
