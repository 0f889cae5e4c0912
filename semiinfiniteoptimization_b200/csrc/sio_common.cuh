// sio_common.cuh -- device-side records shared by the fp32 bulk kernels (sio_kernels.cu), the
// fp64 kernels (sio_fp64.cu) and the C-ABI host code (sio_abi.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sio {

constexpr int kThreads = 256;        // CTA size of the bulk kernels
constexpr int kPtsPerThread = 4;     // points held in registers per thread
constexpr int kTile = kThreads * kPtsPerThread;   // points per tile (1024)
constexpr int kXfChunk = 32;         // transforms processed per CTA of the min kernel

// One resident SDF grid.  `v` is a PADDED copy: (nx+1)*(ny+1)*(nz+1) samples whose last layer in
// every axis replicates the one before it, so corner i0+1 is always addressable and equals the
// clamped corner of Appendix A.2 step 3.
struct GridDev {
    const float* v;
    // the same samples as 32-byte BRICKS: cell (i,j,k) -> the 8 monomial coefficients of its trilinear
    // interpolant, c[4a+2b+c] multiplying fx^a fy^b fz^c (f = offset from the low corner in cell units,
    // formed in float64 from the 8 corner samples), at ((i*ny + j)*nz + k)*8, 32-byte aligned: one 256-bit
    // load fetches everything a query needs (the fp32 kernels read this; the fp64 kernels read `v`)
    const float* brick;
    int32_t nx, ny, nz;
    int32_t sx, sy;              // element strides of the padded copy: sy = nz+1, sx = (ny+1)*(nz+1)
    double bmin[3], bmax[3], h[3], ih[3];   // ih = 1/h (host-computed: no fp64 divides on the device)
    float vmin, vmax;            // range of the samples
    float lip;                   // Lipschitz constant of the trilinear interpolant w.r.t. local position
    // LOCAL Lipschitz constants for the sub-tile culling: lipd[(D-1) * ncell + cell] = ceil(64 * the largest per-cell
    // constant among the cells within D cells (Chebyshev, D = 1 .. kLipLevels) of `cell`), 255 = use `lip`
    const uint8_t* lipd;
};
constexpr int kLipLevels = 8;

// fp32 transform record of one query b: cloud-local point -> dual-lattice cell coordinates
// u = M p + t  (grid affine (p - bmin)/h - 1/2 folded in), plus what the inner loop needs.
struct __align__(16) XfRec {
    float m[12];                 // rows of [M | t]
    float hx, hy, hz;            // cell size (bbox-overshoot distance is measured in metres)
    float thr;                   // candidate threshold: bound + tau (+inf: no under-bound list)
    float hix, hiy, hiz;         // dim-1 (upper clamp of the interpolation coordinate)
    int32_t sx;                  // cell stride in x: ny*nz
    const float* v;              // the grid's bricks (monomial coefficients)
    int32_t sy;                  // cell stride in y: nz
    float lip;                   // Lipschitz constant of the grid's interpolant (sub-tile culling)
    const uint8_t* lipd;         // the grid's local constants (GridDev::lipd), or null
    int32_t ncell;               // cells per level of lipd
    float ihmin;                 // 1 / (smallest cell size), rounded up
};
static_assert(sizeof(XfRec) == 112, "XfRec layout");

// candidate / under-bound record.  The fp32 pass writes (b, internal index, float d); the fp64
// re-check rewrites it in place to (b or -1, caller index, exact d).
struct __align__(16) Cand {
    int32_t b;
    int32_t idx;
    double d;
};

// order-preserving map float -> uint32 (min on the key == min on the float, -0 < +0)
__host__ __device__ inline uint32_t fkey(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
// the same key from the float's bit pattern (the bulk pass stores raw float bits per sub-tile)
__host__ __device__ inline uint32_t fkey_bits(uint32_t u) { return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__host__ __device__ inline float fkey_inv(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}

// ---- launchers implemented in the .cu files (all enqueue on `st`, none synchronises) --------
struct MinArgs {
    const float4* pts; const int32_t* perm; int64_t N;
    const GridDev* grids; const int32_t* grid_of_b;   // device
    const GridDev* single_grid;                         // HOST copy when every transform uses one grid, else null
    const double* T_rel; const double* bound;          // device; bound may be null
    const double* under_bound;                          // = bound when the under-bound set is wanted, else null (minimum only)
    int64_t B;                                          // total transforms (xf, T_rel, bound, outputs are [B])
    int64_t b_begin, Bc;                                // the batch [b_begin, b_begin+Bc) this launch covers
    double tau;
    uint32_t* subkeys; int32_t ntiles;                  // scratch [Bc * ntiles*8]: float32 min (raw bits) per 128-point sub-tile
    XfRec* xf;                                          // scratch [B]
    Cand* cand; unsigned long long* cand_count; uint32_t cand_cap;   // 64-bit counter: a C5-scale sweep can offer > 2^32 candidates
    int32_t* n_under;                                   // device [B] counters of the exact under-bound set, or null
    const float4* rep;                                  // bounding sphere (a cloud point + radius) per 128-point sub-tile
    // sub-tile culling (SIO_PRUNE): lower bounds, upper bounds, work list
    float* lb; uint32_t* ub; uint32_t* items; uint32_t* item_count;
    double* dmin; int64_t* argmin;                      // device outputs
    int normalize;
    int num_sms;                                        // persistent grids are sized from it
};
void launch_prep_xf(const MinArgs& a, cudaStream_t st);
void launch_min_f32(const MinArgs& a, cudaStream_t st);
void launch_min_pruned(const MinArgs& a, cudaStream_t st);
bool launch_min_smem(const MinArgs& a, cudaStream_t st);      // TMA-staged shared-memory variant (needs a.xf); false = not applicable
void launch_finalize(const MinArgs& a, bool all_tiles, bool emit_under, int n_recheck, cudaStream_t st);
void launch_push_results(void* const* d_boards, int nranks, int rank, int64_t slot_bytes, const double* dmin, const int64_t* argmin,
                         const int32_t* n_under, int64_t B, cudaStream_t st);

void launch_cloud_query_f32(const float4* pts, const float4* rep, int64_t N, const XfRec* xf, int64_t B, int normalize,
                            float* out_d, float4* out_g, cudaStream_t st);
void launch_query_f64(const GridDev* grid, const double* T /*12, device*/, const double* pts, int64_t P,
                      int normalize, double* d, double* grad, cudaStream_t st);
void launch_query_f32(const GridDev* grid, const double* T, const double* pts, int64_t P,
                      int normalize, double* d, double* grad, cudaStream_t st);
void launch_pose_rows(const GridDev* grid, const double* T_pose, const int64_t* offsets, int64_t nprob,
                      const double* pts, int64_t total, int normalize, double* d, double* rows, cudaStream_t st);

struct RobotDev {
    int32_t n;
    const int32_t* parent; const int32_t* jtype;
    const double* tparent; const double* axis;
};
void launch_fk(RobotDev r, const double* q, int64_t B, double* T_out, cudaStream_t st);
// T_rel[(s*n_sel + l)] = T_link(s, links[l])^-1 * T_cloud ; grid_of_b[(s*n_sel+l)] = l
void launch_link_rel(const double* T_links, int32_t n_links, const int32_t* links, int32_t n_sel, int64_t S,
                     const double* T_cloud, double* T_rel, int32_t* grid_of_b, cudaStream_t st);
void launch_pair_rel(const double* T_links, int32_t n_links, const int32_t* link_of_pair, const int32_t* slot_of_link, int64_t P,
                     const double* T_cloud, double* T_rel, int32_t* grid_of_b, cudaStream_t st);
void launch_chain_rows(RobotDev r, const GridDev* grids, const int32_t* link_grid, const double* T_links,
                       const int32_t* cfg_of_pt, const int32_t* link_of_pt, const double* pts, int64_t P,
                       int normalize, double* d, double* rows, cudaStream_t st);
// the strict QPs of one lock-step SIP iteration, one warp per problem (sio_qp.cu); returns -1 for an unsupported n
size_t sip_qp_scratch_bytes(int num_sms);
// rows of problem p: offs[p] .. offs[p+1] (ragged), or -- offs == NULL -- cnt[p] rows at p * stride (cnt[p] < 0: skipped)
// deferral of slow problems (see k_sip_qp): list/count/done set = capped launch that lists the problems it gave up on;
// only/only_n/done set = follow-up launch over exactly those problems
struct QpDefer {
    int32_t* list = nullptr; uint32_t* count = nullptr; int32_t* done = nullptr;
    uint8_t* mark = nullptr;            // capped launch: mark[p] = 1 for every problem it lists (nobody else writes it)
    const uint8_t* busy = nullptr;      // capped launch: problems with busy[p] == 1 are out with an earlier follow-up launch
    const int32_t* only = nullptr; const uint32_t* only_n = nullptr;
    int32_t* wide = nullptr; uint32_t* wide_n = nullptr;   // capped launch: room for the problems with two rows per lane (sio_qp.cu)
    // host side only: with these set the two-row launch runs BESIDE the one-row launch on `side` (fork / join by the two
    // events) instead of after it
    cudaStream_t side = nullptr; cudaEvent_t fork = nullptr, join = nullptr;
};
int launch_sip_qp(int n, int64_t nprob, const double* W, const double* xdes, const double* xmin, const double* xmax,
                  const int64_t* offs, const int32_t* cnt, int32_t stride, const double* rows, const double* depths, double reg,
                  double inflation, double eps_abs, int max_iter, double* dx, int32_t* status, void* scratch, int num_sms,
                  cudaStream_t st, const QpDefer* defer = nullptr);
// mesh -> unsigned distance at every cell centre (sio_sdf.cu); tri = [nt][9] doubles
void launch_mesh_distance(const double* tri, int32_t nt, const double bmin[3], const double h[3], const int32_t dims[3],
                          double* out, cudaStream_t st);
// inside/outside sign of a distance grid (winding number near the surface, inherited along z elsewhere) -> float32 SDF
void launch_sdf_sign(const double* tri, int32_t nt, const double bmin[3], const double h[3], const int32_t dims[3],
                     const double* dist, int8_t* side, float* out, cudaStream_t st);
// surface sampler: lattice points of every triangle (host-built layout records), then order-preserving de-duplication
struct LatticeTriHost { double a[3], b[3], c[3]; int32_t m1, m2; int64_t base; };
void launch_lattice_fill(const void* tris, int32_t nt, double* pts, cudaStream_t st);
void launch_dedup(const double* pts, int64_t n, uint32_t* table, uint64_t table_slots, uint8_t* keep, uint32_t* counts, double* out,
                  cudaStream_t st);
void launch_gather_bench(const float* arr, int64_t n_elems, const uint32_t* idx, int64_t n_gathers,
                         float* sink, cudaStream_t st);

}  // namespace sio
