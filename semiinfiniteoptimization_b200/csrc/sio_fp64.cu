// sio_fp64.cu -- float64 kernels of the collision oracle (sm_100a), compiled with --fmad=false so
// the arithmetic is the plain IEEE sequence of SURVEY.md Appendix A.2 (no contraction).
//
//   k_finalize   exact (d, index) minimum per transform from the fp32 pass's sub-tile minima
//                (extra CTAs of the same launch re-evaluate the fp32 pass's under-bound candidates exactly)
//   k_query_f64  K1 in float64 (PenetrationDepthGeometry.distance / normal, gopt:104-108, 153)
//   k_pose_rows  K3 pose rows  [ (R(pt-t)) x n , n ]            (gopt:413-418)
//   k_fk / k_link_rel / k_chain_rows  K4 kinematics + link rows   (gopt:201-204, 479-482)
#include "sio_common.cuh"
#include "sio_eval64.cuh"

namespace sio {

// ---- finalize: one CTA per transform (+ extra CTAs re-checking the under-bound candidates) -------------------
constexpr int kFinThreads = 256;
constexpr int kSubPts = 32 * kPtsPerThread;      // points per sub-tile (128)
constexpr int kFinSubs = kFinThreads / kSubPts;  // sub-tiles evaluated per pass (4)
constexpr int kMaxList = 512;

__device__ __forceinline__ bool lex_less(double d1, int64_t i1, double d2, int64_t i2) {
    return d1 < d2 || (d1 == d2 && i1 >= 0 && (i2 < 0 || i1 < i2));
}

// exact re-evaluation of the fp32 pass's under-bound candidates, strided over the `nthreads` threads of the
// re-check CTAs; kept entries are counted per transform in n_under.
// More candidates than the buffer holds: the list is incomplete, so no count can be exact -- every n_under[b] of the
// call is set to -1 (never a silently low count).  The host entry points see the counter, grow the buffer and run the
// sweep again; a caller of the _dev entry point sees the -1s (and SIO_ERR_OVERFLOW from sio_under_fetch).
__device__ void recheck_candidates(const float4* __restrict__ pts, const int32_t* __restrict__ perm, int64_t N,
                                   const GridDev* __restrict__ grids, const int32_t* __restrict__ grid_of_b,
                                   const double* __restrict__ T_rel, const double* __restrict__ bound,
                                   Cand* __restrict__ cand, const unsigned long long* __restrict__ cand_count, uint32_t cand_cap,
                                   int32_t* __restrict__ n_under, int64_t B, uint32_t first, uint32_t nthreads) {
    const unsigned long long offered = *cand_count;
    if (offered > (unsigned long long)cand_cap) {
        if (n_under)
            for (int64_t b = first; b < B; b += nthreads) n_under[b] = -1;
        return;
    }
    const uint32_t n = (uint32_t)offered;
    for (uint32_t c = first; c < n; c += nthreads) {
        Cand r = cand[c];
        if (r.idx >= N) { r.b = -1; cand[c] = r; continue; }        // a padding copy of the last point
        const GridDev& g = grids[grid_of_b ? grid_of_b[r.b] : 0];
        float4 q = __ldg(pts + r.idx);
        double p[3];
        xf64(T_rel + 12 * (int64_t)r.b, (double)q.x, (double)q.y, (double)q.z, p);
        double d = eval64(g, p, nullptr);
        Cand o;
        o.b = (d < bound[r.b]) ? r.b : -1;
        o.idx = perm[r.idx];
        o.d = d;
        cand[c] = o;
        if (o.b >= 0 && n_under) atomicAdd(n_under + o.b, 1);
    }
}

// blocks [0, Bc): exact (d, index) minimum of transform b_begin + blockIdx.x from the fp32 sub-tile minima;
// blocks [Bc, Bc + n_recheck): candidate re-check (only launched with the last batch of a float32 call)
__global__ void __launch_bounds__(kFinThreads)
k_finalize(const float4* __restrict__ pts, const int32_t* __restrict__ perm, int64_t N,
           const GridDev* __restrict__ grids, const int32_t* __restrict__ grid_of_b,
           const double* __restrict__ T_rel, const double* __restrict__ bound, double tau, int64_t b_begin, int64_t Bc,
           const uint32_t* __restrict__ subkeys, int64_t nsub, int all_tiles, int emit_under,
           Cand* __restrict__ cand, unsigned long long* __restrict__ cand_count, uint32_t cand_cap,
           int32_t* __restrict__ n_under, double* __restrict__ dmin, int64_t* __restrict__ argmin) {
    const int tid = threadIdx.x;
    if ((int64_t)blockIdx.x >= Bc) {
        recheck_candidates(pts, perm, N, grids, grid_of_b, T_rel, bound, cand, cand_count, cand_cap, n_under, b_begin + Bc,
                           (uint32_t)((blockIdx.x - Bc) * kFinThreads + tid), (uint32_t)((gridDim.x - Bc) * kFinThreads));
        return;
    }
    __shared__ uint32_t s_wkey[kFinThreads / 32];
    __shared__ int32_t s_list[kMaxList];
    __shared__ int32_t s_nlist;
    __shared__ double s_d[kFinThreads / 32];
    __shared__ int64_t s_i[kFinThreads / 32];
    __shared__ GridDev s_g;
    __shared__ double s_T[12];
    const int lane = tid & 31, warp = tid >> 5;
    const int64_t bl = blockIdx.x;                 // batch-local (indexes subkeys)
    const int64_t b = b_begin + bl;                // global (indexes T_rel, bound, outputs)
    if (tid == 0) { s_g = grids[grid_of_b ? grid_of_b[b] : 0]; s_nlist = 0; }
    if (tid < 12) s_T[tid] = T_rel[12 * b + tid];
    __syncthreads();
    bool scan_all = all_tiles != 0;
    if (!scan_all) {
        // 1) fp32 minimum over the sub-tile minima (kept in registers for pass 2)
        const uint32_t* keys = subkeys + bl * nsub;
        uint32_t k = 0xffffffffu;
        for (int64_t s = tid; s < nsub; s += kFinThreads) k = min(k, fkey_bits(__ldg(keys + s)));
        k = __reduce_min_sync(0xffffffffu, k);
        if (lane == 0) s_wkey[warp] = k;
        __syncthreads();
        k = s_wkey[lane < kFinThreads / 32 ? lane : 0];
        k = __reduce_min_sync(0xffffffffu, k);
        const float m32 = fkey_inv(k);
        // 2) sub-tiles that may hold the exact minimum: min <= m32 + tau  (fp32 error << tau)
        const uint32_t kthr = fkey(__double2float_ru((double)m32 + tau));
        for (int64_t s = tid; s < nsub; s += kFinThreads) {
            if (fkey_bits(__ldg(keys + s)) <= kthr) {
                int slot = atomicAdd(&s_nlist, 1);
                if (slot < kMaxList) s_list[slot] = (int32_t)s;
            }
        }
        __syncthreads();
        if (s_nlist > kMaxList) scan_all = true;     // degenerate (flat field): fall back to every sub-tile
    }
    // 3) exact evaluation, kFinSubs sub-tiles per pass
    double best = INFINITY;
    int64_t besti = -1;
    const double bnd = (emit_under && bound) ? bound[b] : -INFINITY;
    const int64_t nloop = scan_all ? nsub : (int64_t)s_nlist;
    for (int64_t l0 = 0; l0 < nloop; l0 += kFinSubs) {
        const int64_t l = l0 + tid / kSubPts;
        if (l >= nloop) continue;
        const int64_t s = scan_all ? l : (int64_t)s_list[l];
        const int64_t i = s * kSubPts + (tid % kSubPts);
        if (i < N) {
            float4 q = __ldg(pts + i);
            double p[3];
            xf64(s_T, (double)q.x, (double)q.y, (double)q.z, p);
            double d = eval64(s_g, p, nullptr);
            int64_t oi = perm[i];
            if (lex_less(d, oi, best, besti)) { best = d; besti = oi; }
            if (d < bnd && bnd < (double)INFINITY) {        // fp64 mode: final under-bound entries
                const unsigned long long slot = atomicAdd(cand_count, 1ull);      // n_under stays exact past the capacity
                if (slot < (unsigned long long)cand_cap) { Cand c; c.b = (int32_t)b; c.idx = (int32_t)oi; c.d = d; cand[slot] = c; }
                if (n_under) atomicAdd(n_under + b, 1);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, best, o);
        const int64_t oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (lex_less(od, oi, best, besti)) { best = od; besti = oi; }
    }
    if (lane == 0) { s_d[warp] = best; s_i[warp] = besti; }
    __syncthreads();
    if (tid == 0) {
        double dm = s_d[0];
        int64_t am = s_i[0];
        for (int w = 1; w < kFinThreads / 32; ++w)
            if (lex_less(s_d[w], s_i[w], dm, am)) { dm = s_d[w]; am = s_i[w]; }
        if (bound && !(dm < bound[b])) am = -1;          // "[]": nothing under Klampt's upperBound
        dmin[b] = dm;
        argmin[b] = am;
    }
}

// n_recheck > 0 appends that many candidate re-check CTAs to the launch (float32 calls that want the under-bound set)
void launch_finalize(const MinArgs& a, bool all_tiles, bool emit_under, int n_recheck, cudaStream_t st) {
    if (a.Bc <= 0) return;
    int64_t nsub = (int64_t)a.ntiles * (kThreads / 32);
    k_finalize<<<(unsigned)(a.Bc + n_recheck), kFinThreads, 0, st>>>(a.pts, a.perm, a.N, a.grids, a.grid_of_b, a.T_rel, a.bound,
                                                                      a.tau, a.b_begin, a.Bc, a.subkeys, nsub, all_tiles ? 1 : 0,
                                                                      emit_under ? 1 : 0, a.cand, a.cand_count, a.cand_cap, a.n_under,
                                                                      a.dmin, a.argmin);
}

// ---- result boards: the compact per-transform results stored straight into every rank's memory over NVLink ------------
// One CTA per destination rank.  A board is nranks slots of slot_bytes; slot r of EVERY board receives rank r's records
// [dmin f64 x B | argmin i64 x B | n_under i32 x B].  boards[] are device pointers valid on THIS device: the local board
// and the peers' boards mapped through CUDA IPC (other processes) or peer access (same process).  The stores are plain
// 8-byte writes over NVLink / NVSwitch; __threadfence_system() orders them before the kernel's completion, so that a
// host-side barrier after the ranks' streams have drained (what bench.py and batch.solve_sharded do anyway) is all the
// synchronisation the exchange needs -- no collective launch, no staging, no host copies.
__global__ void k_push_results(void* const* __restrict__ boards, int rank, int64_t slot_bytes, const double* __restrict__ dmin,
                               const int64_t* __restrict__ argmin, const int32_t* __restrict__ n_under, int64_t B) {
    char* slot = reinterpret_cast<char*>(boards[blockIdx.x]) + (int64_t)rank * slot_bytes;
    double* o_d = reinterpret_cast<double*>(slot);
    int64_t* o_a = reinterpret_cast<int64_t*>(slot + 8 * B);
    int32_t* o_n = reinterpret_cast<int32_t*>(slot + 16 * B);
    for (int64_t b = threadIdx.x; b < B; b += blockDim.x) {
        o_d[b] = dmin[b];
        o_a[b] = argmin[b];
        o_n[b] = n_under ? n_under[b] : 0;
    }
    __threadfence_system();
}

void launch_push_results(void* const* d_boards, int nranks, int rank, int64_t slot_bytes, const double* dmin, const int64_t* argmin,
                         const int32_t* n_under, int64_t B, cudaStream_t st) {
    if (nranks <= 0 || B <= 0) return;
    k_push_results<<<nranks, 128, 0, st>>>(d_boards, rank, slot_bytes, dmin, argmin, n_under, B);
}

// ---- K1 float64 ---------------------------------------------------------------------------------------
__global__ void k_query_f64(const GridDev* __restrict__ grid, const double* __restrict__ T,
                            const double* __restrict__ pts, int64_t P, int normalize,
                            double* __restrict__ d, double* __restrict__ grad) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= P) return;
    const GridDev g = *grid;
    double p[3], gl[3];
    xf64(T, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], p);
    d[i] = eval64(g, p, grad ? gl : nullptr);
    if (grad) grad_to_input(T, gl, normalize, grad + 3 * i);
}

void launch_query_f64(const GridDev* grid, const double* T, const double* pts, int64_t P, int normalize,
                      double* d, double* grad, cudaStream_t st) {
    if (P <= 0) return;
    int threads = 128;
    k_query_f64<<<(unsigned)((P + threads - 1) / threads), threads, 0, st>>>(grid, T, pts, P, normalize, d, grad);
}

// ---- K3 pose rows ----------------------------------------------------------------------------------------
__global__ void k_pose_rows(const GridDev* __restrict__ grid, const double* __restrict__ T_pose,
                            const int64_t* __restrict__ offsets, int64_t nprob, const double* __restrict__ pts,
                            int64_t total, int normalize, double* __restrict__ d, double* __restrict__ rows) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= total) return;
    // owning problem: largest p with offsets[p] <= i
    int64_t lo = 0, hi = nprob - 1;
    while (lo < hi) {
        int64_t mid = (lo + hi + 1) >> 1;
        if (offsets[mid] <= i) lo = mid; else hi = mid - 1;
    }
    const double* T = T_pose + 12 * lo;
    double Ti[12];
    invert34(T, Ti);
    const GridDev g = *grid;
    double p[3], gl[3], gw[3];
    const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    xf64(Ti, x, y, z, p);
    d[i] = eval64(g, p, gl);
    if (!rows) return;
    grad_to_input(Ti, gl, normalize, gw);
    const double n0 = -gw[0], n1 = -gw[1], n2 = -gw[2];            // grad1
    const double q0 = x - T[3], q1 = y - T[7], q2 = z - T[11];
    // lever arm R (pt - t), exactly as the reference writes it (gopt:415)
    const double r0 = T[0] * q0 + T[1] * q1 + T[2] * q2;
    const double r1 = T[4] * q0 + T[5] * q1 + T[6] * q2;
    const double r2 = T[8] * q0 + T[9] * q1 + T[10] * q2;
    double* o = rows + 6 * i;
    o[0] = r1 * n2 - r2 * n1;
    o[1] = r2 * n0 - r0 * n2;
    o[2] = r0 * n1 - r1 * n0;
    o[3] = n0; o[4] = n1; o[5] = n2;
}

void launch_pose_rows(const GridDev* grid, const double* T_pose, const int64_t* offsets, int64_t nprob,
                      const double* pts, int64_t total, int normalize, double* d, double* rows, cudaStream_t st) {
    if (total <= 0) return;
    int threads = 128;
    k_pose_rows<<<(unsigned)((total + threads - 1) / threads), threads, 0, st>>>(grid, T_pose, offsets, nprob, pts,
                                                                                 total, normalize, d, rows);
}

// ---- K4 kinematics -----------------------------------------------------------------------------------------
__device__ __forceinline__ void joint34(const double* ax, int jt, double q, double* J) {
    if (jt == 1) {
        J[0] = 1; J[1] = 0; J[2] = 0; J[3] = ax[0] * q;
        J[4] = 0; J[5] = 1; J[6] = 0; J[7] = ax[1] * q;
        J[8] = 0; J[9] = 0; J[10] = 1; J[11] = ax[2] * q;
        return;
    }
    double s, c;
    sincos(q, &s, &c);
    const double C = 1.0 - c, x = ax[0], y = ax[1], z = ax[2];
    J[0] = c + x * x * C;     J[1] = x * y * C - z * s; J[2] = x * z * C + y * s; J[3] = 0;
    J[4] = y * x * C + z * s; J[5] = c + y * y * C;     J[6] = y * z * C - x * s; J[7] = 0;
    J[8] = z * x * C - y * s; J[9] = z * y * C + x * s; J[10] = c + z * z * C;    J[11] = 0;
}

// one thread per configuration; links are visited in index order (parents precede children)
__global__ void k_fk(RobotDev r, const double* __restrict__ q, int64_t B, double* __restrict__ T_out) {
    int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= B) return;
    for (int l = 0; l < r.n; ++l) {
        double J[12], L[12], W[12];
        joint34(r.axis + 3 * l, r.jtype[l], q[b * r.n + l], J);
        mul34(r.tparent + 12 * l, J, L);
        double* out = T_out + (b * r.n + l) * 12;
        if (r.parent[l] < 0) {
#pragma unroll
            for (int k = 0; k < 12; ++k) out[k] = L[k];
        } else {
            const double* Pm = T_out + (b * r.n + r.parent[l]) * 12;
            double Pl[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) Pl[k] = Pm[k];
            mul34(Pl, L, W);
#pragma unroll
            for (int k = 0; k < 12; ++k) out[k] = W[k];
        }
    }
}

void launch_fk(RobotDev r, const double* q, int64_t B, double* T_out, cudaStream_t st) {
    if (B <= 0) return;
    int threads = 64;
    k_fk<<<(unsigned)((B + threads - 1) / threads), threads, 0, st>>>(r, q, B, T_out);
}

__global__ void k_link_rel(const double* __restrict__ T_links, int32_t n_links, const int32_t* __restrict__ links,
                           int32_t n_sel, int64_t S, const double* __restrict__ T_cloud,
                           double* __restrict__ T_rel, int32_t* __restrict__ grid_of_b) {
    int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= S * n_sel) return;
    int64_t s = b / n_sel;
    int l = (int)(b - s * n_sel);
    const double* T = T_links + (s * n_links + links[l]) * 12;
    double Tl[12], Ti[12], Tc[12], R[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) { Tl[k] = T[k]; Tc[k] = T_cloud[k]; }
    invert34(Tl, Ti);
    mul34(Ti, Tc, R);
#pragma unroll
    for (int k = 0; k < 12; ++k) T_rel[12 * b + k] = R[k];
    grid_of_b[b] = l;
}

void launch_link_rel(const double* T_links, int32_t n_links, const int32_t* links, int32_t n_sel, int64_t S,
                     const double* T_cloud, double* T_rel, int32_t* grid_of_b, cudaStream_t st) {
    int64_t n = S * n_sel;
    if (n <= 0) return;
    int threads = 128;
    k_link_rel<<<(unsigned)((n + threads - 1) / threads), threads, 0, st>>>(T_links, n_links, links, n_sel, S, T_cloud,
                                                                            T_rel, grid_of_b);
}

// pair p: T_rel[p] = T_link(p, link_of_pair[p])^-1 * T_cloud, grid_of_b[p] = slot_of_link[link_of_pair[p]]
__global__ void k_pair_rel(const double* __restrict__ T_links, int32_t n_links, const int32_t* __restrict__ link_of_pair,
                           const int32_t* __restrict__ slot_of_link, int64_t P, const double* __restrict__ T_cloud,
                           double* __restrict__ T_rel, int32_t* __restrict__ grid_of_b) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int l = link_of_pair[p];
    const double* T = T_links + (p * n_links + l) * 12;
    double Tl[12], Ti[12], Tc[12], R[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) { Tl[k] = T[k]; Tc[k] = T_cloud[k]; }
    invert34(Tl, Ti);
    mul34(Ti, Tc, R);
#pragma unroll
    for (int k = 0; k < 12; ++k) T_rel[12 * p + k] = R[k];
    grid_of_b[p] = slot_of_link[l];
}

void launch_pair_rel(const double* T_links, int32_t n_links, const int32_t* link_of_pair, const int32_t* slot_of_link, int64_t P,
                     const double* T_cloud, double* T_rel, int32_t* grid_of_b, cudaStream_t st) {
    if (P <= 0) return;
    int threads = 128;
    k_pair_rel<<<(unsigned)((P + threads - 1) / threads), threads, 0, st>>>(T_links, n_links, link_of_pair, slot_of_link, P, T_cloud,
                                                                            T_rel, grid_of_b);
}

// value + Jp^T n for points attached to links (gopt:478-482).  One thread per point.
__global__ void k_chain_rows(RobotDev r, const GridDev* __restrict__ grids, const int32_t* __restrict__ link_grid,
                             const double* __restrict__ T_links, const int32_t* __restrict__ cfg_of_pt,
                             const int32_t* __restrict__ link_of_pt, const double* __restrict__ pts, int64_t P,
                             int normalize, double* __restrict__ d, double* __restrict__ rows) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int link = link_of_pt[i];
    const int64_t cfg = cfg_of_pt ? cfg_of_pt[i] : 0;
    const double* Tl = T_links + cfg * r.n * 12;
    double T[12], Ti[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = Tl[12 * link + k];
    invert34(T, Ti);
    const GridDev g = grids[link_grid[link]];
    const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    double p[3], gl[3], gw[3];
    xf64(Ti, x, y, z, p);
    d[i] = eval64(g, p, gl);
    if (!rows) return;
    grad_to_input(Ti, gl, normalize, gw);
    const double n0 = -gw[0], n1 = -gw[1], n2 = -gw[2];
    double* o = rows + (size_t)r.n * i;
    for (int j = 0; j < r.n; ++j) o[j] = 0.0;
    for (int j = link; j >= 0; j = r.parent[j]) {
        const double* Tj = Tl + 12 * j;
        const double* ax = r.axis + 3 * j;
        const double w0 = Tj[0] * ax[0] + Tj[1] * ax[1] + Tj[2] * ax[2];
        const double w1 = Tj[4] * ax[0] + Tj[5] * ax[1] + Tj[6] * ax[2];
        const double w2 = Tj[8] * ax[0] + Tj[9] * ax[1] + Tj[10] * ax[2];
        if (r.jtype[j] == 1) {
            o[j] = w0 * n0 + w1 * n1 + w2 * n2;
        } else {
            const double a0 = x - Tj[3], a1 = y - Tj[7], a2 = z - Tj[11];
            const double c0 = w1 * a2 - w2 * a1, c1 = w2 * a0 - w0 * a2, c2 = w0 * a1 - w1 * a0;
            o[j] = c0 * n0 + c1 * n1 + c2 * n2;
        }
    }
}

void launch_chain_rows(RobotDev r, const GridDev* grids, const int32_t* link_grid, const double* T_links,
                       const int32_t* cfg_of_pt, const int32_t* link_of_pt, const double* pts, int64_t P,
                       int normalize, double* d, double* rows, cudaStream_t st) {
    if (P <= 0) return;
    int threads = 64;
    k_chain_rows<<<(unsigned)((P + threads - 1) / threads), threads, 0, st>>>(r, grids, link_grid, T_links, cfg_of_pt,
                                                                              link_of_pt, pts, P, normalize, d, rows);
}

}  // namespace sio
