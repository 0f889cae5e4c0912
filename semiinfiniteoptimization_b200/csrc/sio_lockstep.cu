// sio_lockstep.cu -- B independent object-pose SIP problems advanced in lock step ENTIRELY on the device (sm_100a).
//
// What it replaces: B sequential runs of optimizeCollFree (geometryopt.py:1039-1081) on optimizeSemiInfinite
// (sip.py:795-1241, default flags) with an ObjectCollisionConstraint (gopt:394-420) -- BASELINE.json configs[4].
// batch.optimizeCollFreeBatchFast already issues one device call per oracle request of an outer iteration, but keeps the
// per-problem bookkeeping in numpy on the host; at 65,536 problems that host work is most of the wall time.  Here the
// bookkeeping is kernels too (one thread per problem): objective / step (so3 log and exp maps), the oracle's drop rule
// and exclusion hash (sip.py:203-213, 280-285), merit score, trust region and weight schedule (sip.py:1159-1215).  The
// host only sequences launches and reads a few counters per round (live problems, problems entering a sweep).
// Rounds are not outer iterations: a problem whose QP went to the background launch (see k_sip_qp) skips rounds and keeps
// its own iteration index, so the slowest QP of a round delays one problem, not all of them.
// Heavy lifting is the existing kernels: K2 sweeps (sio_min_distance_dev), the QPs (k_sip_qp on the padded layout),
// and value / Jacobian-row evaluation of the instantiated parameters (same arithmetic as k_pose_rows).
//
// Layout: everything per problem is padded to `cap` parameter slots: P / key [B][cap][3], D / Dn [B][cap], rows [B][cap][6].
// The arithmetic and its order follow batch.optimizeCollFreeBatchFast statement for statement (float64, no FMA
// contraction), so both drivers walk every problem through the same iterations.
#include "sio_eval64.cuh"
#include "sio_lockstep.cuh"

namespace sio {

constexpr int kLsThreads = 128;

// ---- so3 maps, statement for statement the host's (_host/so3.py, batch._log_so3_vec / _exp_so3_vec) ---------------
__device__ void so3_log(const double* M, double* out) {            // M row-major 3x3
    double c = (M[0] + M[4] + M[8] - 1.0) * 0.5;
    c = fmin(1.0, fmax(-1.0, c));
    const double th = acos(c);
    const double v[3] = {M[7] - M[5], M[2] - M[6], M[3] - M[1]};
    if (th < 1e-8) { out[0] = 0.5 * v[0]; out[1] = 0.5 * v[1]; out[2] = 0.5 * v[2]; return; }
    if (fabs(th - 3.141592653589793) < 1e-5) {
        // near pi the antisymmetric part vanishes; recover the axis from R + I
        double A[9];
        for (int i = 0; i < 9; ++i) A[i] = (M[i] + ((i % 4 == 0) ? 1.0 : 0.0)) * 0.5;
        int k = 0;
        if (A[4] > A[0]) k = 1;
        if (A[8] > A[4 * k]) k = 2;
        const double s = sqrt(fmax(A[4 * k], 1e-300));
        double ax[3] = {A[k] / s, A[3 + k] / s, A[6 + k] / s};
        if (ax[0] * v[0] + ax[1] * v[1] + ax[2] * v[2] < 0) { ax[0] = -ax[0]; ax[1] = -ax[1]; ax[2] = -ax[2]; }
        out[0] = ax[0] * th; out[1] = ax[1] * th; out[2] = ax[2] * th;
        return;
    }
    const double f = th / (2.0 * sin(th));
    out[0] = v[0] * f; out[1] = v[1] * f; out[2] = v[2] * f;
}

__device__ void so3_exp(const double* w, double* M) {
    const double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    const double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double K2[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) K2[3 * i + j] = K[3 * i] * K[j] + K[3 * i + 1] * K[3 + j] + K[3 * i + 2] * K[6 + j];
    double a, b;
    if (th < 1e-12) { a = 1.0; b = 0.5; }
    else { a = sin(th) / th; b = (1 - cos(th)) / (th * th); }
    for (int i = 0; i < 9; ++i) M[i] = ((i % 4 == 0) ? 1.0 : 0.0) + a * K[i] + b * K2[i];
}

// f = |tdes - t|^2 + |log(Rdes R')|^2 and the step towards the target [log(Rdes R'), tdes - t] (gopt:383-389)
__device__ double objective(const double* Rdes, const double* tdes, const double* R, const double* t, double* dxdes) {
    double M[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[3 * i + j] = Rdes[3 * i] * R[3 * j] + Rdes[3 * i + 1] * R[3 * j + 1] + Rdes[3 * i + 2] * R[3 * j + 2];
    double rv[3];
    so3_log(M, rv);
    const double dt[3] = {tdes[0] - t[0], tdes[1] - t[1], tdes[2] - t[2]};
    if (dxdes) { dxdes[0] = rv[0]; dxdes[1] = rv[1]; dxdes[2] = rv[2]; dxdes[3] = dt[0]; dxdes[4] = dt[1]; dxdes[5] = dt[2]; }
    return (dt[0] * dt[0] + dt[1] * dt[1] + dt[2] * dt[2]) + (rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
}

__device__ __forceinline__ double row_min(const double* D, int n) {
    double m = INFINITY;
    for (int k = 0; k < n; ++k) m = fmin(m, D[k]);
    return m;
}

// T_rel = T_obj^-1 T_env for the pose (R, t), in the host driver's order of operations
__device__ void rel_transform(const double* R, const double* t, const double* Tenv, double* out) {
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) out[4 * i + j] = R[i] * Tenv[j] + R[3 + i] * Tenv[4 + j] + R[6 + i] * Tenv[8 + j];
        const double a = R[i] * Tenv[3] + R[3 + i] * Tenv[7] + R[6 + i] * Tenv[11];
        const double b = R[i] * t[0] + R[3 + i] * t[1] + R[6 + i] * t[2];
        out[4 * i + 3] = a - b;
    }
}

// appends the world point y to problem p unless its exclusion key is present (sip.py:278-285); returns the slot or -1
__device__ int instantiate(const LsState& S, int64_t p, const double* y, double excl) {
    long long k[3];
    for (int a = 0; a < 3; ++a) k[a] = (long long)trunc(y[a] / excl);          // int() truncates toward zero
    const int n = S.cnt[p];
    long long* keys = S.key + p * S.cap * 3;
    for (int s = 0; s < n; ++s)
        if (keys[3 * s] == k[0] && keys[3 * s + 1] == k[1] && keys[3 * s + 2] == k[2]) return -1;
    if (n >= S.cap) { atomicAdd(S.counters + 3, 1ull); return -1; }
    double* P = S.P + (p * S.cap + n) * 3;
    for (int a = 0; a < 3; ++a) { P[a] = y[a]; keys[3 * n + a] = k[a]; }
    S.cnt[p] = n + 1;
    return n;
}

// ---- kernels --------------------------------------------------------------------------------------------------
// Round bookkeeping.  Problems advance through their OWN outer iterations (pit): one whose QP was handed to the
// background launch (ph = 1) sits rounds out until the answer flag is up, then joins the next round at the step
// phase (ph = 2); everything a problem does, and the order it does it in, is what the plain lock-step loop did.
__global__ void k_ls_begin(LsState S, LsSettings cfg) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= S.B || !S.live[p]) return;
    if (S.ph[p] == 1) {
        if (*reinterpret_cast<volatile int32_t*>(S.qdone + p)) { S.ph[p] = 2; atomicAdd(S.counters + 4, 1ull); }
        atomicAdd(S.counters + 1, 1ull);
        return;
    }
    const int it = S.pit[p];
    if (it >= cfg.max_iters) { S.live[p] = 0; return; }    // out of iterations: stays 'not converged'
    S.niter[p] = it;
    S.mu[p] = (it == 0) || ((it - 1) % 100 == 0);          // sip.py:855-870
    objective(S.Rdes + 9 * p, S.tdes + 3 * p, S.R + 9 * p, S.t + 3 * p, S.dxdes + 6 * p);
    atomicAdd(S.counters + 1, 1ull);
    atomicAdd(S.counters + 4, 1ull);
}

// value (and Jacobian row) of every instantiated parameter at the current (next = 0) or proposed (next = 1) pose;
// one thread per (problem, slot); the arithmetic is k_pose_rows'.  write_d = 0: rows only (after the oracle has
// appended a parameter, whose depth is the sweep's exact minimum and must stay).
__global__ void k_ls_values(LsState S, int next, int want_rows, int write_d) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= S.B * S.cap) return;
    const int64_t p = e / S.cap;
    const int slot = (int)(e - p * S.cap);
    double* out = next ? S.Dn : S.D;
    const bool active = next ? S.go[p] : (S.live[p] && S.ph[p] == 0 && (write_d || S.mu[p]));
    if (!active) return;
    if (slot >= S.cnt[p]) { if (write_d) out[e] = INFINITY; return; }
    double T[12], Ti[12];
    const double* R = (next ? S.Rn : S.R) + 9 * p;
    const double* t = (next ? S.tn : S.t) + 3 * p;
    for (int i = 0; i < 3; ++i) { T[4 * i] = R[3 * i]; T[4 * i + 1] = R[3 * i + 1]; T[4 * i + 2] = R[3 * i + 2]; T[4 * i + 3] = t[i]; }
    invert34(T, Ti);
    const double* y = S.P + e * 3;
    double q[3], gl[3], gw[3];
    xf64(Ti, y[0], y[1], y[2], q);
    const double dval = eval64(*S.grid, q, want_rows ? gl : nullptr);
    if (write_d) out[e] = dval;
    if (!want_rows) return;
    grad_to_input(Ti, gl, 1, gw);
    const double n0 = -gw[0], n1 = -gw[1], n2 = -gw[2];
    const double q0 = y[0] - T[3], q1 = y[1] - T[7], q2 = y[2] - T[11];
    const double r0 = T[0] * q0 + T[1] * q1 + T[2] * q2;             // lever arm R (pt - t), as the reference writes it (gopt:415)
    const double r1 = T[4] * q0 + T[5] * q1 + T[6] * q2;
    const double r2 = T[8] * q0 + T[9] * q1 + T[10] * q2;
    double* o = S.rows + e * 6;
    o[0] = r1 * n2 - r2 * n1; o[1] = r2 * n0 - r0 * n2; o[2] = r0 * n1 - r1 * n0;
    o[3] = n0; o[4] = n1; o[5] = n2;
}

// oracle at x, first half (sip.py:201-219): drop parameters at or above the drop value, queue the problem for a sweep
__global__ void k_ls_oracle_pre(LsState S, LsSettings cfg) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= S.B || !S.live[p] || S.ph[p] != 0 || !S.mu[p] || !S.upd[p]) return;
    double* D = S.D + p * S.cap;
    double* P = S.P + p * S.cap * 3;
    long long* K = S.key + p * S.cap * 3;
    const int n = S.cnt[p];
    int m = 0;
    for (int s = 0; s < n; ++s) {
        if (D[s] >= cfg.drop_value) continue;
        if (m != s) {
            D[m] = D[s];
            for (int a = 0; a < 3; ++a) { P[3 * m + a] = P[3 * s + a]; K[3 * m + a] = K[3 * s + a]; }
        }
        ++m;
    }
    for (int s = m; s < n; ++s) D[s] = INFINITY;
    S.cnt[p] = m;
    const double olddmin = fmin(0.0, row_min(D, m));
    const unsigned long long k = atomicAdd(S.counters + 0, 1ull);
    S.list[k] = (int32_t)p;
    rel_transform(S.R + 9 * p, S.t + 3 * p, S.Tenv, S.T_rel + 12 * k);
    S.bound[k] = olddmin;
}

__global__ void k_ls_oracle_post(LsState S, LsSettings cfg, int64_t count) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int64_t p = S.list[k];
    if (S.arg[k] >= 0 && S.dmin[k] < S.bound[k]) {
        const int slot = instantiate(S, p, S.cloud_world + 3 * S.arg[k], cfg.excl);
        if (slot >= 0) S.D[p * S.cap + slot] = S.dmin[k];
    }
}

// depths -> g(x), first-iteration bookkeeping, the oracle schedule, and the QP inputs of the problem
__global__ void k_ls_gx(LsState S, LsSettings cfg) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= S.B) return;
    if (!S.live[p]) { S.qk[p] = -1; return; }
    if (S.ph[p] != 0) { if (S.ph[p] == 2) S.qk[p] = -1; return; }     // ph = 1: the background launch still reads qk
    S.qdef[p] = 0;
    const int it = S.pit[p];
    const double g = row_min(S.D + p * S.cap, S.cnt[p]);
    S.gxc[p] = g;
    if (it == 0) {
        S.gx0[p] = g;
        if (g < S.min_cv[p]) S.min_cv[p] = g;
    }
    S.upd[p] = (it % 100 == 0);
    S.qk[p] = S.cnt[p];
    for (int i = 0; i < 6; ++i) { S.qW[6 * p + i] = S.w[p]; S.qlo[6 * p + i] = -S.trs[p]; S.qhi[6 * p + i] = S.trs[p]; }
}

// after the QP: convergence test, proposed step, merit of the current point (sip.py:980-981, 1146-1151)
__global__ void k_ls_step(LsState S, LsSettings cfg) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= S.B) return;
    S.go[p] = 0;
    if (!S.live[p] || S.ph[p] == 1) return;
    if (S.ph[p] == 0 && S.qdef[p]) { S.ph[p] = 1; return; }    // deferred: the answer comes from the background launch
    if ((S.qstatus[p] & 0xff) == 3) {
        // the QP kernel produced no step (its slack re-solve was infeasible too): the reference raises here
        // (results.x is None, sip.py:358-378).  The problem stops with its own status instead of stepping along
        // whatever S.dx holds.
        S.status[p] = 2;
        S.live[p] = 0;
        S.ph[p] = 0;
        S.gx[p] = S.gxc[p];
        atomicAdd(S.counters + 5, 1ull);
        return;
    }
    const double* dx = S.dx + 6 * p;
    double n2 = 0.0;
    for (int i = 0; i < 6; ++i) n2 += dx[i] * dx[i];
    if (n2 < cfg.xeps2) {
        S.status[p] = 1;
        S.live[p] = 0;
        S.ph[p] = 0;
        S.gx[p] = S.gxc[p];
        return;
    }
    S.go[p] = 1;
    double E[9];
    so3_exp(dx, E);
    const double* R = S.R + 9 * p;
    double* Rn = S.Rn + 9 * p;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Rn[3 * i + j] = E[3 * i] * R[j] + E[3 * i + 1] * R[3 + j] + E[3 * i + 2] * R[6 + j];
    for (int a = 0; a < 3; ++a) S.tn[3 * p + a] = S.t[3 * p + a] + dx[3 + a];
    S.sorig[p] = fmax(0.0, -S.gxc[p]) + S.w[p] * S.fx[p];
    S.ninst0[p] = S.cnt[p];
    S.fxn[p] = objective(S.Rdes + 9 * p, S.tdes + 3 * p, Rn, S.tn + 3 * p, nullptr);
}

__global__ void k_ls_score_pre(LsState S) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= S.B || !S.go[p]) return;
    const double g = row_min(S.Dn + p * S.cap, S.cnt[p]);
    S.gxn[p] = g;
    S.metric[p] = fmax(0.0, -g);
    const unsigned long long k = atomicAdd(S.counters + 0, 1ull);
    S.list[k] = (int32_t)p;
    rel_transform(S.Rn + 9 * p, S.tn + 3 * p, S.Tenv, S.T_rel + 12 * k);
    S.bound[k] = g;
}

// discovery at the proposed point, accept / reject, trust region, weight schedule (sip.py:156-165, 1159-1215)
__global__ void k_ls_score_post(LsState S, LsSettings cfg, int64_t count) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int64_t p = S.list[k];
    double gxn = S.gxn[p], metric = S.metric[p];
    if (S.arg[k] >= 0 && S.dmin[k] < gxn) {
        gxn = S.dmin[k];
        const int slot = instantiate(S, p, S.cloud_world + 3 * S.arg[k], cfg.excl);
        if (slot >= 0 && S.dmin[k] < 0) metric = fmax(metric, -S.dmin[k]);
    }
    const double snew = metric + S.w[p] * S.fxn[p];
    bool accept = !(snew >= S.sorig[p]);
    if (S.cnt[p] > S.ninst0[p] && gxn < 0) { S.upd[p] = 0; accept = false; }
    const double* dx = S.dx + 6 * p;
    if (accept) {
        for (int i = 0; i < 9; ++i) S.R[9 * p + i] = S.Rn[9 * p + i];
        for (int a = 0; a < 3; ++a) S.t[3 * p + a] = S.tn[3 * p + a];
        S.fx[p] = S.fxn[p];
        S.gx[p] = gxn;
        S.trs[p] *= 2.5;
    } else {
        S.gx[p] = S.gxc[p];
        double n2 = 0.0, mx = 0.0;
        for (int i = 0; i < 6; ++i) { n2 += dx[i] * dx[i]; mx = fmax(mx, fabs(dx[i])); }
        S.trs[p] *= 0.5;
        if (S.trs[p] > sqrt(n2)) S.trs[p] = mx * 0.5;
    }
    const double scale = 0.75 * ((atan(S.gx[p] * 3) * 2 / 3.141592653589793 + 0.5) * 1.75 + 0.25);
    S.w[p] = fmax(S.w[p] * scale, cfg.w_min);
    S.ph[p] = 0;
    S.pit[p] += 1;
    atomicAdd(S.counters + 2, 1ull);
}

__global__ void k_ls_init(LsState S, LsSettings cfg) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= S.B) return;
    S.cnt[p] = 0; S.w[p] = cfg.w0; S.trs[p] = 0.5; S.upd[p] = 1; S.live[p] = 1; S.go[p] = 0;
    S.ph[p] = 0; S.mu[p] = 1; S.qdef[p] = 0; S.pit[p] = 0; S.qdone[p] = 0; S.qstatus[p] = 0;
    S.min_cv[p] = cfg.min_cv; S.status[p] = 0; S.niter[p] = 0;
    S.gx[p] = INFINITY; S.gx0[p] = INFINITY; S.gxc[p] = INFINITY;
    S.fx[p] = objective(S.Rdes + 9 * p, S.tdes + 3 * p, S.R + 9 * p, S.t + 3 * p, nullptr);
    for (int i = 0; i < 6; ++i) S.dx[6 * p + i] = 0.0;     // the arena is reused across calls
}

// poses as the caller hands them over (row-major 3x4, initial then desired) -> R (3x3), t, Rdes, tdes
__global__ void __launch_bounds__(kLsThreads)
k_ls_unpack_poses(LsState S, const double* __restrict__ raw) {
    const int64_t p = (int64_t)blockIdx.x * kLsThreads + threadIdx.x;
    if (p >= S.B) return;
    const double* Ti = raw + 12 * p;
    const double* Td = raw + 12 * (S.B + p);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) { S.R[9 * p + 3 * i + j] = Ti[4 * i + j]; S.Rdes[9 * p + 3 * i + j] = Td[4 * i + j]; }
        S.t[3 * p + i] = Ti[4 * i + 3]; S.tdes[3 * p + i] = Td[4 * i + 3];
    }
}

// the instantiated parameters, packed: problem p's cnt[p] points go to packed[3 * offs[p] ...] (one warp per problem)
__global__ void __launch_bounds__(kLsThreads)
k_ls_pack_params(LsState S, const int64_t* __restrict__ offs, double* __restrict__ packed) {
    const int64_t p = ((int64_t)blockIdx.x * kLsThreads + threadIdx.x) >> 5;
    if (p >= S.B) return;
    const int n3 = 3 * S.cnt[p];
    const double* src = S.P + p * (int64_t)S.cap * 3;
    double* dst = packed + 3 * offs[p];
    for (int i = threadIdx.x & 31; i < n3; i += 32) dst[i] = src[i];
}

static unsigned grid_for(int64_t n) { return (unsigned)((n + kLsThreads - 1) / kLsThreads); }

void ls_launch_init(const LsState& S, const LsSettings& c, cudaStream_t st) { k_ls_init<<<grid_for(S.B), kLsThreads, 0, st>>>(S, c); }
void ls_launch_begin(const LsState& S, const LsSettings& c, cudaStream_t st) { k_ls_begin<<<grid_for(S.B), kLsThreads, 0, st>>>(S, c); }
void ls_launch_values(const LsState& S, int next, int want_rows, int write_d, cudaStream_t st) {
    k_ls_values<<<grid_for(S.B * S.cap), kLsThreads, 0, st>>>(S, next, want_rows, write_d);
}
void ls_launch_oracle_pre(const LsState& S, const LsSettings& c, cudaStream_t st) { k_ls_oracle_pre<<<grid_for(S.B), kLsThreads, 0, st>>>(S, c); }
void ls_launch_oracle_post(const LsState& S, const LsSettings& c, int64_t n, cudaStream_t st) {
    if (n > 0) k_ls_oracle_post<<<grid_for(n), kLsThreads, 0, st>>>(S, c, n);
}
void ls_launch_gx(const LsState& S, const LsSettings& c, cudaStream_t st) { k_ls_gx<<<grid_for(S.B), kLsThreads, 0, st>>>(S, c); }
void ls_launch_step(const LsState& S, const LsSettings& c, cudaStream_t st) { k_ls_step<<<grid_for(S.B), kLsThreads, 0, st>>>(S, c); }
void ls_launch_score_pre(const LsState& S, cudaStream_t st) { k_ls_score_pre<<<grid_for(S.B), kLsThreads, 0, st>>>(S); }
void ls_launch_score_post(const LsState& S, const LsSettings& c, int64_t n, cudaStream_t st) {
    if (n > 0) k_ls_score_post<<<grid_for(n), kLsThreads, 0, st>>>(S, c, n);
}

void ls_launch_unpack_poses(const LsState& S, const double* raw, cudaStream_t st) { k_ls_unpack_poses<<<grid_for(S.B), kLsThreads, 0, st>>>(S, raw); }
void ls_launch_pack_params(const LsState& S, const int64_t* offs, double* packed, cudaStream_t st) {
    k_ls_pack_params<<<grid_for(S.B * 32), kLsThreads, 0, st>>>(S, offs, packed);
}

}  // namespace sio
