// sio_qp.cu -- the QPs of one lock-step SIP iteration on the device (sm_100a), one warp per problem.
//
// What it replaces: the OSQP solves of ConstraintGenerationData.solve_qp_osqp (sip.py:294-378), one call per
// problem per outer iteration,
//     min (x - xdes)' W (x - xdes) + reg |x|^2   s.t.  A x + b >= inflation,  xmin <= x <= xmax,
// and, when that strict problem is infeasible or its solution out of scale, the slack problem of
// sip.py:332-357 (one slack variable per row, penalty 'auto') -- for the 65,536 problems of BASELINE.json
// configs[4] advancing together.  The algorithm is the one of _host/qp.py / _host/csrc/hostqp.c (OSQP's ADMM:
// over-relaxation, adaptive rho, residual test every 10 iterations, primal-infeasibility certificate,
// active-set polish), in float64.
//
// Mapping: lane r owns constraint rows r and r+32 (the k instantiated-parameter rows, then the n box rows) and, in
// the slack problem, the slack variables of its rows; x lives replicated in every lane.  The quadratic term is
// diagonal and every slack variable sits in exactly one row, so the ADMM linear system reduces (Schur complement
// over the slack block) to an n x n system whose explicit inverse K is folded into the rows once per rho (lane r keeps
// K a_r), so that one n-value warp sum per iteration yields the whole x-update; everything else is lane-local.  The polish solves its KKT system (slack unknowns eliminated the same
// way) with a warp-parallel Gaussian elimination in a per-warp scratch (L2-resident global memory: the ADMM loop
// itself touches no memory).  Warps are persistent and draw problems from an atomic counter, because the ADMM
// iteration count varies from tens to thousands between problems.
// Problems with more than 64 rows are flagged for the host solver.
#include "sio_common.cuh"
#include <cstdio>
#include <cstdlib>

namespace sio {

constexpr int kQpWarps = 4;                 // warps per CTA
constexpr int kQpCtasPerSm = 2;             // resident CTAs per SM the register budget is set for
constexpr int kQpCtasNarrow = 4;            // the same for the one-row-per-lane build (128 registers: its loop-invariant operands sit in shared memory)
constexpr int kQpMaxRows = 64;              // 2 rows per lane
constexpr int kQpMaxVars = 8;
constexpr int kQpMaxKkt = kQpMaxVars + kQpMaxRows;     // largest polish system (x + active rows)
constexpr int kQpScratch = kQpMaxKkt * (kQpMaxKkt + 1);   // doubles per warp
// Polish systems of up to kQpSmemKkt unknowns (x + active rows: nearly all of them) are factorised in shared memory:
// the elimination is a chain of dependent loads and stores with two __syncwarp per column, 15-20 us per problem at
// L2 latency (a quarter of a typical solve), a few at shared-memory latency.  Larger systems use the global scratch.
constexpr int kQpSmemKkt = 24;
constexpr int kQpSmemScratch = kQpSmemKkt * (kQpSmemKkt + 1);

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Maxima of NON-NEGATIVE doubles through their bit patterns, which order like the values: three or four integer
// instructions where fmax's NaN handling costs seven, and a warp maximum in two integer warp reductions (high words, then
// low words among the lanes that hold the largest high word) where a shuffle butterfly takes five dependent levels.  A
// NaN orders above everything here (fmax drops it): the residual tests then fail and the solve runs to its iteration limit.
__device__ __forceinline__ double max_nn(double a, double b) {
    return __longlong_as_double(max(__double_as_longlong(a), __double_as_longlong(b)));
}
__device__ __forceinline__ double warp_max_nn(double v) {
    const int hi = __double2hiint(v);
    const int mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, (hi == mhi) ? (unsigned)__double2loint(v) : 0u);
    return __hiloint2double(mhi, (int)mlo);
}

// Sum (or maximum) over the warp of M (<= 8) doubles per lane, delivered to every lane.  A butterfly per value costs
// 5 M 64-bit shuffles; here three halving steps (xor 16, 8, 4: a lane keeps one half of its values and ships the
// other) leave one value per lane, two butterfly levels finish it inside each group of four lanes, and M broadcasts
// hand every total to every lane: 6 values = 8 + 6 shuffles instead of 30.  The maximum pads with 0: for
// non-negative values only.
template <int M, bool MAX = false>
struct AllRed {
    static constexpr int H0 = (M + 1) / 2, H1 = (H0 + 1) / 2, H2 = (H1 + 1) / 2;
    static_assert(M >= 1 && M <= 8 && H2 == 1, "three halving steps must reach one value");
    __device__ __forceinline__ static double comb(double a, double b) { return MAX ? fmax(a, b) : a + b; }
    __host__ __device__ static constexpr int src(int i) {   // lane (bits 1, 0 clear) that ends up with total i
        int l = 0;
        if (i >= H0) { l |= 16; i -= H0; }
        if (i >= H1) { l |= 8; i -= H1; }
        if (i >= H2) { l |= 4; }
        return l;
    }
    template <int LEN, int HALF>
    __device__ __forceinline__ static void halve(double (&v)[M], bool up, int mask) {
#pragma unroll
        for (int k = 0; k < HALF; ++k) {
            const double lo = v[k], hi = (k + HALF < LEN) ? v[k + HALF] : 0.0;
            const double send = up ? lo : hi, keep = up ? hi : lo;
            v[k] = comb(keep, __shfl_xor_sync(0xffffffffu, send, mask));
        }
    }
    __device__ __forceinline__ static void all(double (&v)[M], int lane) {
        halve<M, H0>(v, lane & 16, 16);
        halve<H0, H1>(v, lane & 8, 8);
        halve<H1, H2>(v, lane & 4, 4);
        double r = v[0];
        r = comb(r, __shfl_xor_sync(0xffffffffu, r, 2));
        r = comb(r, __shfl_xor_sync(0xffffffffu, r, 1));
#pragma unroll
        for (int i = 0; i < M; ++i) v[i] = __shfl_sync(0xffffffffu, r, src(i));
    }
};
template <int M> using AllSum = AllRed<M, false>;
template <int M> using AllMax = AllRed<M, true>;

// The same sum for the ADMM loop's one reduction per iteration, without the selects: lane L keeps its 8 slots (M
// values and zero padding) in the order slot j <-> value (L >> 2 & 7) ^ j, so that in every halving step "keep the first
// half, ship the second" is the right thing for every lane; the lane ends with total (L >> 2 & 7).
struct Sum8 {
    __device__ __forceinline__ static int slot_value(int lane, int j) { return ((lane >> 2) & 7) ^ j; }
    __device__ __forceinline__ static double scatter(double (&v)[8]) {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[4 + k], 16);
#pragma unroll
        for (int k = 0; k < 2; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[2 + k], 8);
        double r = v[0] + __shfl_xor_sync(0xffffffffu, v[1], 4);
        r += __shfl_xor_sync(0xffffffffu, r, 2);
        r += __shfl_xor_sync(0xffffffffu, r, 1);
        return r;
    }
    template <int M>
    __device__ __forceinline__ static void gather(double r, double (&out)[M]) {
#pragma unroll
        for (int i = 0; i < M; ++i) out[i] = __shfl_sync(0xffffffffu, r, i << 2);
    }
};

// Symmetric positive definite N x N system matrix -> its explicit inverse (packed lower triangle, registers).
// The ADMM loop applies the inverse as N independent dot products: a Cholesky forward/back substitution is a
// serial dependency chain several times longer, and this kernel is bound by exactly that latency.
template <int N>
struct SpdInverse {
    double A[N * (N + 1) / 2];              // in: lower triangle of the matrix; out: lower triangle of its inverse
    __device__ __forceinline__ double& at(int i, int j) { return A[i * (i + 1) / 2 + j]; }
    __device__ __forceinline__ double sym(int i, int j) const { return i >= j ? A[i * (i + 1) / 2 + j] : A[j * (j + 1) / 2 + i]; }
    __device__ __forceinline__ bool invert() {
        double inv[N];
#pragma unroll
        for (int j = 0; j < N; ++j) {                      // Cholesky, in place
            double d = at(j, j);
#pragma unroll
            for (int k = 0; k < j; ++k) d -= at(j, k) * at(j, k);
            if (!(d > 0.0)) return false;
            d = sqrt(d);
            at(j, j) = d;
            inv[j] = 1.0 / d;
#pragma unroll
            for (int i = j + 1; i < N; ++i) {
                double s = at(i, j);
#pragma unroll
                for (int k = 0; k < j; ++k) s -= at(i, k) * at(j, k);
                at(i, j) = s * inv[j];
            }
        }
        double X[N][N];                                    // columns of the inverse: solve L L' X = I
#pragma unroll
        for (int c = 0; c < N; ++c) {
            double b[N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                double s = (i == c) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 0; k < i; ++k) s -= at(i, k) * b[k];
                b[i] = s * inv[i];
            }
#pragma unroll
            for (int i = N - 1; i >= 0; --i) {
                double s = b[i];
#pragma unroll
                for (int k = i + 1; k < N; ++k) s -= at(k, i) * b[k];
                b[i] = s * inv[i];
            }
#pragma unroll
            for (int i = 0; i < N; ++i) X[i][c] = b[i];
        }
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) at(i, j) = 0.5 * (X[i][j] + X[j][i]);
        return true;
    }
    __device__ __forceinline__ void apply(double (&b)[N]) const {
        double r[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < N; ++j) s += sym(i, j) * b[j];
            r[i] = s;
        }
#pragma unroll
        for (int i = 0; i < N; ++i) b[i] = r[i];
    }
};

// Gaussian elimination with partial pivoting on the NN x NN row-major system M x = b in the warp's scratch.  The pivot of
// a column is its first largest |entry|, as in the host's serial routine (hostqp.c lu_solve), and each entry sees the same
// operations in the same order, so the answer is the same.  Returns false on a zero pivot.  Result in b.
// Every step is spread over the whole warp: the pivot search is two integer warp maxima over the bit patterns of the
// (non-negative) magnitudes and an integer minimum over the rows that hold it; the elimination of column c runs over (row, column) pairs, lane = 8 * row
// group + column group, with the multiplier of a row kept in the entry it annihilates; only the back substitution is
// serial.  (One row per lane with a serial column loop left three lanes in four idle on a typical 12 x 12 system and
// cost a sixth of the kernel's time: profiles/r02_qp_phases.md.)
__device__ bool warp_lu_solve(double* M, double* b, int NN, int lane) {
    const int lr = lane >> 3, lk = lane & 7;
    for (int c = 0; c < NN; ++c) {
        double v = -1.0;                                   // a NaN never wins, as in the host's comparison
        int pr = NN;
        for (int r = c + lane; r < NN; r += 32) {
            const double a = fabs(M[r * NN + c]);
            if (a > v) { v = a; pr = r; }
        }
        const int hi = __double2hiint(v);
        const int mhi = __reduce_max_sync(0xffffffffu, hi);
        if (mhi < 0) return false;
        const unsigned lo = (hi == mhi) ? (unsigned)__double2loint(v) : 0u;
        const unsigned mlo = __reduce_max_sync(0xffffffffu, lo);
        if (mhi == 0 && mlo == 0u) return false;           // zero pivot
        const int pv = __reduce_min_sync(0xffffffffu, (hi == mhi && lo == mlo) ? pr : 0x7fffffff);
        if (pv != c) {
            for (int kk = lane; kk <= NN; kk += 32) {
                double* pc = kk < NN ? M + c * NN + kk : b + c;
                double* pp = kk < NN ? M + pv * NN + kk : b + pv;
                const double t = *pc; *pc = *pp; *pp = t;
            }
        }
        __syncwarp();
        const double piv = M[c * NN + c];
        for (int r = c + 1 + lane; r < NN; r += 32) M[r * NN + c] = M[r * NN + c] / piv;     // the row's multiplier, in place
        __syncwarp();
        for (int r = c + 1 + lr; r < NN; r += 4) {
            const double f = M[r * NN + c];
            if (f == 0.0) continue;
#pragma unroll 2
            for (int kk = c + 1 + lk; kk <= NN; kk += 8) {
                double* dst = kk < NN ? M + r * NN + kk : b + r;
                const double src = kk < NN ? M[c * NN + kk] : b[c];
                *dst -= f * src;
            }
        }
        __syncwarp();
    }
    if (lane == 0)
        for (int r = NN - 1; r >= 0; --r) {
            double s = b[r];
#pragma unroll 4
            for (int kk = r + 1; kk < NN; ++kk) s -= M[r * NN + kk] * b[kk];
            b[r] = s / M[r * NN + r];
        }
    __syncwarp();
    return true;
}

// one lane's share of a problem: NR constraint rows (and, in the slack problem, their slack variables).  NR = 1 serves
// problems with at most 32 rows (nearly all of them) at half the per-iteration work of NR = 2.
// SM (the one-row build of the lock-step solver's capped launch): everything the iteration loop only READS -- the lane's
// row a_r, its K a_r (kb), its row of K (krow), and the problem's P, q -- lives in a per-warp block of shared memory instead
// of registers, column `lane` of a [value][32] array each (P and q once, broadcast).  That is 64 registers less across the
// loop, which lets the build fit 128 registers = 4 CTAs x 4 warps per SM with no spill inside the loop: the kernel is
// bound by the latency of one dependent chain per warp, so resident warps are throughput.  The loads do not sit on that
// chain (their addresses never change), and the arithmetic is the same operations on the same values.
template <int N> __host__ __device__ constexpr int qp_ops_doubles() { return (N + 8 + N) * 32 + 2 * N; }
template <int N, int NR, bool SM>
struct Rows {
    double areg[SM ? 1 : NR][SM ? 1 : N], l[NR], u[NR];
    bool have[NR], slack[NR];               // row present; row has a slack variable (the k parameter rows)
    double* op;                             // SM: the warp's operand block + lane
    __device__ __forceinline__ double a(int s, int i) const { return SM ? op[i * 32] : areg[s][i]; }
    __device__ __forceinline__ void set_a(int s, int i, double v) { if (SM) op[i * 32] = v; else areg[s][i] = v; }
    __device__ __forceinline__ double kb(int j) const { return op[(N + j) * 32]; }
    __device__ __forceinline__ double krow(int j) const { return op[(N + 8 + j) * 32]; }
};

enum { kAdmmSolved = 0, kAdmmMaxIter = 1, kAdmmInfeasible = 2 };

// ADMM + polish.  SLACK = false: variables x (N).  SLACK = true: variables [x | one slack per parameter row],
// the slack block carrying the quadratic weight `pen`.  x (and sv, the lane's slack values) hold the answer.
// One instantiation serves both (a runtime flag): the unrolled body is large and two copies thrash the
// instruction cache (ncu r01_qp: 64 % of the stall samples were instruction fetch).
template <int N, int NR, bool SM>
__device__ __forceinline__ int admm(const bool SLACK, const Rows<N, NR, SM>& R, const double (&Pd_)[SM ? 1 : N], const double (&q_)[SM ? 1 : N], double pen, int k, int m, double eps_abs,
                    double eps_rel, int max_iter, double* kkt, double* skkt, int lane, double (&x)[N], double (&sv)[NR], int& iters) {
    static_assert(!SM || NR == 1, "shared-memory operands: one row per lane");
    constexpr double kInf = 1e20, sigma = 1e-6, alpha = 1.6;
    // P and q: registers, or (SM) the broadcast tail of the warp's operand block
    const double* const pq = SM ? R.op - lane + (N + 8 + N) * 32 : nullptr;
    auto Pd = [&](int i) -> double { return SM ? pq[i] : Pd_[SM ? 0 : i]; };
    auto q = [&](int i) -> double { return SM ? pq[N + i] : q_[SM ? 0 : i]; };
    constexpr int check_every = 10, adapt_every = 50;
    double z[NR], y[NR], rv[NR], dy[NR], Ax[NR], ds[NR];
#pragma unroll
    for (int s = 0; s < NR; ++s) { z[s] = 0.0; y[s] = 0.0; dy[s] = 0.0; Ax[s] = 0.0; ds[s] = 1.0; }
    double rho = 0.1;
    double irv[NR], ids[NR];                                  // reciprocals: no float64 divisions inside the loop
    // The x-update is x~ = K (A'w + sigma x - q), K the inverse of the reduced system matrix.  By linearity it is
    // sum over rows of (K a_r) w_r + K (sigma x - q): each lane keeps kb = K a_r for its rows (in Sum8's slot order), so
    // that ONE warp sum delivers K A'w, and the lane that the sum leaves total i on keeps row i of K (krow; kq = krow . q)
    // to add (K (sigma x - q))_i before the totals are broadcast.  K itself lives only inside set_rho.
    double kb_[SM ? 1 : NR][SM ? 1 : 8], krow_[SM ? 1 : N], kq = 0.0;
    auto KB = [&](int s, int j) -> double { return SM ? R.kb(j) : kb_[SM ? 0 : s][SM ? 0 : j]; };
    auto KROW = [&](int j) -> double { return SM ? R.krow(j) : krow_[SM ? 0 : j]; };
    auto set_rho = [&]() {
        SpdInverse<N> K;
#pragma unroll
        for (int s = 0; s < NR; ++s) {
            rv[s] = (R.u[s] - R.l[s] < 1e-12) ? 1e3 * rho : rho;
            ds[s] = pen + sigma + rv[s];                  // diagonal of the slack block of the system matrix
            irv[s] = 1.0 / rv[s];
            ids[s] = 1.0 / ds[s];
        }
        constexpr int kTri = N * (N + 1) / 2, kChunk = 7;
        static_assert(kTri % kChunk == 0, "the packed triangle is summed seven entries at a time");
#pragma unroll
        for (int c0 = 0; c0 < kTri; c0 += kChunk) {
            double part[kChunk];
#pragma unroll
            for (int e = 0; e < kChunk; ++e) {
                int i = 0;                                 // (i, j) of packed entry c0 + e
                while ((i + 1) * (i + 2) / 2 <= c0 + e) ++i;
                const int j = c0 + e - i * (i + 1) / 2;
                double v = 0.0;
#pragma unroll
                for (int s = 0; s < NR; ++s)
                    if (R.have[s]) {
                        const double wgt = (SLACK && R.slack[s]) ? rv[s] - rv[s] * rv[s] * ids[s] : rv[s];   // Schur complement
                        v += wgt * R.a(s, i) * R.a(s, j);
                    }
                part[e] = v;
            }
            AllSum<kChunk>::all(part, lane);
#pragma unroll
            for (int e = 0; e < kChunk; ++e) {
                int i = 0;
                while ((i + 1) * (i + 2) / 2 <= c0 + e) ++i;
                const int j = c0 + e - i * (i + 1) / 2;
                K.A[c0 + e] = part[e] + ((i == j) ? Pd(i) + sigma : 0.0);
            }
        }
        if (!K.invert()) return false;
#pragma unroll
        for (int s = 0; s < NR; ++s) {
            double ka[N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                double v = 0.0;
#pragma unroll
                for (int j = 0; j < N; ++j) v += K.sym(i, j) * R.a(s, j);
                ka[i] = v;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int want = Sum8::slot_value(lane, j);
                double v = 0.0;
#pragma unroll
                for (int i = 0; i < N; ++i) v = (want == i) ? ka[i] : v;
                if (SM) R.op[(N + j) * 32] = v; else kb_[SM ? 0 : s][SM ? 0 : j] = v;
            }
        }
        kq = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double v = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) v = (((lane >> 2) & 7) == i) ? K.sym(i, j) : v;
            if (SM) R.op[(N + 8 + j) * 32] = v; else krow_[SM ? 0 : j] = v;
            kq += v * q(j);
        }
        return true;
    };
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = 0.0;
#pragma unroll
    for (int s = 0; s < NR; ++s) sv[s] = 0.0;
    // set_rho's unrolled body is a quarter of the solver's instructions and the kernel is bound by instruction fetch
    // outside its iteration loop (profiles/r02_qp_phases.md), so it exists ONCE: an outer loop per value of rho, the
    // iteration loop inside.  The iterate is parked in the warp's (otherwise idle) polish scratch across the call --
    // left to the compiler it stays live in registers and every problem's first refactorisation pays the spills.
    static_assert((5 * NR + N) * 32 <= kQpSmemScratch, "the parked iterate fits the polish scratch");
    double* const nook = skkt + lane;
    auto park = [&]() {
#pragma unroll
        for (int s = 0; s < NR; ++s) {
            nook[(5 * s + 0) * 32] = z[s]; nook[(5 * s + 1) * 32] = y[s]; nook[(5 * s + 2) * 32] = sv[s];
            nook[(5 * s + 3) * 32] = dy[s]; nook[(5 * s + 4) * 32] = Ax[s];
        }
#pragma unroll
        for (int i = 0; i < N; ++i) nook[(5 * NR + i) * 32] = x[i];
    };
    auto unpark = [&]() {
#pragma unroll
        for (int s = 0; s < NR; ++s) {
            z[s] = nook[(5 * s + 0) * 32]; y[s] = nook[(5 * s + 1) * 32]; sv[s] = nook[(5 * s + 2) * 32];
            dy[s] = nook[(5 * s + 3) * 32]; Ax[s] = nook[(5 * s + 4) * 32];
        }
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = nook[(5 * NR + i) * 32];
    };
    park();
    int st = kAdmmMaxIter;
    int to_check = check_every;                              // iterations to the next multiple of check_every
    for (int it = 1;;) {
    const bool refactored = !__any_sync(0xffffffffu, !set_rho());
    unpark();
    if (!refactored) break;
    bool new_rho = false;
    for (; it <= max_iter; ++it) {
        ++iters;
        double xt[N], w[NR], stl[NR], yold[NR];
        const bool chk = --to_check == 0 || it == max_iter;             // residual test this iteration (qp.py: every 10)
        if (to_check == 0) to_check = check_every;
        {
            double part[8];
#pragma unroll
            for (int s = 0; s < NR; ++s) {
                w[s] = rv[s] * z[s] - y[s];
                double ws = w[s];
                if (SLACK && R.slack[s]) ws -= rv[s] * (sigma * sv[s] + w[s]) * ids[s];     // eliminate the slack unknown
#pragma unroll
                for (int j = 0; j < 8; ++j) part[j] = (s == 0) ? KB(0, j) * ws : fma(KB(s, j), ws, part[j]);   // kb = 0 on absent rows
            }
            double kx = KROW(0) * x[0];
#pragma unroll
            for (int j = 1; j < N; ++j) kx = fma(KROW(j), x[j], kx);
            Sum8::gather(Sum8::scatter(part) + fma(sigma, kx, -kq), xt);
        }
        double ndy = 0.0;
#pragma unroll
        for (int s = 0; s < NR; ++s) {
            double ax = R.a(s, 0) * xt[0];
#pragma unroll
            for (int i = 1; i < N; ++i) ax = fma(R.a(s, i), xt[i], ax);
            double zs = ax;
            if (SLACK && R.slack[s]) {
                stl[s] = (sigma * sv[s] + w[s] - rv[s] * ax) * ids[s];
                zs = ax + stl[s];
            }
            const double zr = alpha * zs + (1.0 - alpha) * z[s];
            const double cc = zr + y[s] * irv[s];
            const double zn = cc < R.l[s] ? R.l[s] : (cc > R.u[s] ? R.u[s] : cc);     // l <= u
            const double yn = y[s] + rv[s] * (zr - zn);
            yold[s] = y[s];
            if (R.have[s]) { z[s] = zn; y[s] = yn; }
            if (SLACK && R.slack[s]) sv[s] = alpha * stl[s] + (1.0 - alpha) * sv[s];
        }
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = alpha * xt[i] + (1.0 - alpha) * x[i];
        if (!chk) continue;
#pragma unroll
        for (int s = 0; s < NR; ++s) {                       // only the infeasibility certificate reads dy
            dy[s] = y[s] - yold[s];                          // 0 on absent rows: their y never moves
            ndy = max_nn(ndy, fabs(dy[s]));
        }
        // residuals (qp.py solve_qp: every 10 iterations)
        double mx[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};       // r_prim, |Ax| / |z|, slack dual residual, its scale, |dy|, |dy| over slack rows
        double Aty[N];
#pragma unroll
        for (int i = 0; i < N; ++i) Aty[i] = 0.0;
#pragma unroll
        for (int s = 0; s < NR; ++s) {
            double v = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) { const double ai = R.a(s, i); v += ai * x[i]; Aty[i] += ai * y[s]; }
            if (SLACK && R.slack[s]) {
                v += sv[s];
                const double ps = pen * sv[s];                          // dual residual of the slack variable: pen s + y
                mx[2] = max_nn(mx[2], fabs(ps + y[s]));
                mx[3] = max_nn(mx[3], max_nn(fabs(ps), fabs(y[s])));
                mx[5] = max_nn(mx[5], fabs(dy[s]));
            }
            Ax[s] = v;
            if (R.have[s]) { mx[0] = max_nn(mx[0], fabs(v - z[s])); mx[1] = max_nn(mx[1], max_nn(fabs(v), fabs(z[s]))); }
        }
        mx[4] = ndy;
#pragma unroll
        for (int i = 0; i < 6; ++i) mx[i] = warp_max_nn(mx[i]);
        AllSum<N>::all(Aty, lane);
        const double r_prim = mx[0], nAx = mx[1];
        double r_dual = mx[2], nd = mx[3];
        ndy = mx[4];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double t = Aty[i];
            const double px = Pd(i) * x[i];
            r_dual = max_nn(r_dual, fabs(px + q(i) + t));
            nd = max_nn(nd, max_nn(fabs(px), max_nn(fabs(t), fabs(q(i)))));
        }
        if (__all_sync(0xffffffffu, r_prim <= eps_abs + eps_rel * nAx && r_dual <= eps_abs + eps_rel * nd)) { st = kAdmmSolved; break; }
        if (__all_sync(0xffffffffu, ndy > 1e-12)) {       // primal infeasibility certificate
            double t[N + 1];                                // A' dy / |dy| and, last, the support term
#pragma unroll
            for (int i = 0; i <= N; ++i) t[i] = 0.0;
            const double indy = 1.0 / ndy;
#pragma unroll
            for (int s = 0; s < NR; ++s) {
                const double d = dy[s] * indy;
#pragma unroll
                for (int i = 0; i < N; ++i) t[i] += R.a(s, i) * d;
                if (R.have[s]) t[N] += R.u[s] * fmax(d, 0.0) + R.l[s] * fmin(d, 0.0);
            }
            AllSum<N + 1>::all(t, lane);
            double nAt = mx[5] * indy;                      // the slack columns of A' dy
#pragma unroll
            for (int i = 0; i < N; ++i) nAt = max_nn(nAt, fabs(t[i]));
            if (__all_sync(0xffffffffu, nAt <= 1e-4 && t[N] < -1e-4)) { st = kAdmmInfeasible; break; }
        }
        if (it % adapt_every == 0 && __all_sync(0xffffffffu, r_prim > 0 && r_dual > 0)) {
            const double scale = sqrt((r_prim / fmax(nAx, 1e-12)) / (r_dual / fmax(nd, 1e-12)));
            if (__all_sync(0xffffffffu, scale > 5 || scale < 0.2)) {
                rho = fmin(fmax(rho * scale, 1e-6), 1e6);
                park();
                new_rho = true;
                ++it;
                break;
            }
        }
    }
    if (!new_rho) break;
    }
    if (st != kAdmmSolved) return st;
    // ---- polish on the active set (qp.py _polish): KKT system over [variables | active rows] ----------------
    int flag[NR], side[NR];                                // side: -1 held at the lower bound only, +1 upper only, 0 both / none
    double rhs[NR];
#pragma unroll
    for (int s = 0; s < NR; ++s) {
        const bool lo = R.have[s] && ((y[s] < -1e-9) || (Ax[s] - R.l[s] < 1e-7)) && R.l[s] > -kInf / 2;
        const bool up = R.have[s] && ((y[s] > 1e-9) || (R.u[s] - Ax[s] < 1e-7)) && R.u[s] < kInf / 2;
        flag[s] = lo || up;
        side[s] = (lo && !up) ? -1 : ((up && !lo) ? 1 : 0);
        rhs[s] = lo ? R.l[s] : R.u[s];
    }
    const unsigned m0 = __ballot_sync(0xffffffffu, flag[0]), m1 = NR > 1 ? __ballot_sync(0xffffffffu, flag[NR - 1]) : 0u;
    const int na = __popc(m0) + __popc(m1);
    // na == 0: the candidate is the unconstrained minimiser (the system below is then diagonal)
    // Unknowns [x | multipliers of the active rows].  A slack variable s_r (weight pen + delta, one entry in its own
    // row) is eliminated in closed form: s_r = -lambda_r / (pen + delta), which adds 1 / (pen + delta) to that row's
    // -delta diagonal; slack variables of inactive rows are zero.
    const int NN = N + na;
    double* M = (NN <= kQpSmemKkt) ? skkt : kkt;
    double* bb = M + NN * NN;
    for (int e = lane; e < NN * (NN + 1); e += 32) M[e] = 0.0;
    __syncwarp();
    const double delta = 1e-10;
    #pragma unroll
    for (int i = 0; i < N; ++i)
        if (lane == i) { M[i * NN + i] = Pd(i) + delta; bb[i] = -q(i); }
    int idx[NR];
#pragma unroll
    for (int s = 0; s < NR; ++s) idx[s] = 0;
#pragma unroll
    for (int s = 0; s < NR; ++s) {
        if (flag[s]) {                                     // active rows keep their row order
            idx[s] = N + ((s == 0) ? __popc(m0 & ((1u << lane) - 1u)) : __popc(m0) + __popc(m1 & ((1u << lane) - 1u)));
#pragma unroll
            for (int i = 0; i < N; ++i) { const double ai = R.a(s, i); M[i * NN + idx[s]] = ai; M[idx[s] * NN + i] = ai; }
            M[idx[s] * NN + idx[s]] = -delta - ((SLACK && R.slack[s]) ? 1.0 / (pen + delta) : 0.0);
            bb[idx[s]] = rhs[s];
        }
    }
    __syncwarp();
    if (!warp_lu_solve(M, bb, NN, lane)) return kAdmmSolved;
    double xp[N], sp[NR];
#pragma unroll
    for (int s = 0; s < NR; ++s) sp[s] = 0.0;
    int bad = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        xp[i] = bb[i];
        if (!isfinite(xp[i])) bad = 1;
    }
    // KKT test (qp.py _polish): xp is stationary by construction; feasible + multipliers of the right sign
    // (<= 0 on a row held at its lower bound, >= 0 at its upper bound) = the minimiser
    double lam[NR], lmax = 0.0;
#pragma unroll
    for (int s = 0; s < NR; ++s) {
        lam[s] = flag[s] ? bb[idx[s]] : 0.0;
        if (!isfinite(lam[s])) bad = 1;
        lmax = fmax(lmax, fabs(lam[s]));
    }
    lmax = warp_max(lmax);
    const double lamtol = 1e-9 * (1.0 + lmax);
#pragma unroll
    for (int s = 0; s < NR; ++s) {
        double v = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) v += R.a(s, i) * xp[i];
        if (SLACK && R.slack[s]) {
            sp[s] = flag[s] ? -lam[s] / (pen + delta) : 0.0;
            if (!isfinite(sp[s])) bad = 1;
            v += sp[s];
        }
        if (R.have[s] && (v < R.l[s] - 1e-7 || v > R.u[s] + 1e-7)) bad = 1;
        if ((side[s] < 0 && lam[s] > lamtol) || (side[s] > 0 && lam[s] < -lamtol)) bad = 1;
    }
    bad = __any_sync(0xffffffffu, bad);
    __syncwarp();
    if (!bad) {
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = xp[i];
#pragma unroll
        for (int s = 0; s < NR; ++s) sv[s] = sp[s];
    }
    return kAdmmSolved;
}

// status codes written per problem (low byte; the ADMM iteration count sits above it)
enum { kQpStrict = 0, kQpSlack = 1, kQpNoRows = 2, kQpToHost = 3, kQpDeferred = 4 };

template <int N, int NR, bool SM = false>
__device__ __forceinline__ void solve_rows(int64_t p, int k, int m, int64_t base, bool box, const double* __restrict__ W,
                                           const double* __restrict__ xdes, const double* __restrict__ xmin, const double* __restrict__ xmax,
                                           const double* __restrict__ rows, const double* __restrict__ depths, double reg, double inflation,
                                           double eps_abs, double eps_rel, int max_iter, double* kkt, double* skkt, double* sop, int lane, double* __restrict__ dx,
                                           int32_t* __restrict__ status, const QpDefer& df) {
    constexpr double kInf = 1e20;
    double Pd[SM ? 1 : N], q[SM ? 1 : N];
    Rows<N, NR, SM> R;
    R.op = SM ? sop + lane : nullptr;
    if (SM) {
        __syncwarp();                                          // the previous problem's readers are done with the block
        if (lane < N) {
            const double w = W[p * N + lane];
            sop[(N + 8 + N) * 32 + lane] = w + reg;
            sop[(N + 8 + N) * 32 + N + lane] = -w * xdes[p * N + lane];
        }
        __syncwarp();
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double w = W[p * N + i];
            Pd[SM ? 0 : i] = w + reg;
            q[SM ? 0 : i] = -w * xdes[p * N + i];
        }
    }
    double bneg2 = 0.0, bmin = INFINITY;
    int nneg = 0;
#pragma unroll
    for (int s = 0; s < NR; ++s) {
        const int r = lane + 32 * s;
        R.have[s] = r < m;
        R.slack[s] = r < k;
        R.l[s] = -kInf; R.u[s] = kInf;
#pragma unroll
        for (int i = 0; i < N; ++i) R.set_a(s, i, 0.0);
        if (r < k) {
            const double* src = rows + (base + r) * N;
#pragma unroll
            for (int i = 0; i < N; ++i) R.set_a(s, i, src[i]);
            const double b = depths[base + r];
            R.l[s] = -b + inflation;
            if (b < 0) { bneg2 += b * b; ++nneg; }
            bmin = fmin(bmin, b);
        } else if (r < m) {
#pragma unroll
            for (int i = 0; i < N; ++i) R.set_a(s, i, (i == r - k) ? 1.0 : 0.0);
            R.l[s] = fmax(xmin[p * N + (r - k)], -kInf);
            R.u[s] = fmin(xmax[p * N + (r - k)], kInf);
        }
    }
    // |negative depths| is summed in row order on the host; a fixed butterfly order here differs only in rounding
    bneg2 = warp_sum(bneg2);
    nneg = __reduce_add_sync(0xffffffffu, nneg);
    bmin = -warp_max(-bmin);
    const double bnorm = nneg ? sqrt(bneg2) : 1.0;
    double x[N], sv[NR];
    // strict problem first; if it is infeasible or its solution out of scale, the reference's slack re-solve
    // (sip.py:332-357, slack_penalty = 'auto').  One call site in a loop: the unrolled solver body exists once.
    int iters = 0, code = kQpStrict;
    double pen = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
        const int st = admm<N, NR, SM>(pass == 1, R, Pd, q, pen, k, m, eps_abs, eps_rel, max_iter, kkt, skkt, lane, x, sv, iters);
        if (df.list && st == kAdmmMaxIter) {              // capped launch: out of budget -> the follow-up launch solves it in full
            if (lane == 0) {
                status[p] = kQpDeferred;
                df.mark[p] = 1;
                df.done[p] = 0;
                df.list[atomicAdd(df.count, 1u)] = (int32_t)p;
            }
            return;
        }
        if (pass == 1) {
            if (st == kAdmmInfeasible) { if (lane == 0) status[p] = kQpToHost; return; }
            code = kQpSlack;
            break;
        }
        double xn = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) xn += x[i] * x[i];
        if (__all_sync(0xffffffffu, !(st == kAdmmInfeasible || sqrt(xn) * reg > bnorm))) break;
        pen = (1.0 + nneg) / (fabs(bmin) + 1e-3);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) if (lane == i) dx[p * N + i] = x[i];
    if (lane == 0) status[p] = code | (iters << 8);        // ADMM iterations spent, for diagnostics
}

// Three builds of the kernel.  kQpFull serves any problem (plain calls, and the follow-up launches over the stragglers).
// The capped launch of the lock-step solver is split in two, back to back on one stream: kQpNarrow compiles in only the
// one-row-per-lane solver and lists the problems with more than 32 rows in df.wide; kQpWide then solves exactly those with
// the two-row solver.  In one kernel the two-row instantiation (70 % of the instructions) sets the register allocation
// and the one-row loop -- where the launch spends its time -- runs with spills; apart it has 222 registers and none
// (profiles/r02_qp_phases.md: in-line QP time of BASELINE configs[4] 0.217 -> 0.16 s).
enum { kQpFull = 0, kQpNarrow = 1, kQpWide = 2, kQpNarrowReg = 3 };   // kQpNarrowReg: the one-row build with every operand in registers (168, 3 CTAs per SM)

template <int N, int MODE>
__device__ void solve_problem(int64_t p, const double* __restrict__ W, const double* __restrict__ xdes, const double* __restrict__ xmin,
                              const double* __restrict__ xmax, const int64_t* __restrict__ offs, const int32_t* __restrict__ cnt,
                              int32_t stride, const double* __restrict__ rows,
                              const double* __restrict__ depths, double reg, double inflation, double eps_abs, double eps_rel,
                              int max_iter, double* kkt, double* skkt, double* sop, int lane, double* __restrict__ dx, int32_t* __restrict__ status,
                              const QpDefer& df) {
    constexpr double kInf = 1e20;
    // ragged layout: rows of problem p at offs[p] .. offs[p+1];  padded layout (lock-step solver): cnt[p] rows at
    // p * stride, cnt[p] < 0 = problem not taking part
    // every branch below depends on warp-uniform VALUES; routing them through votes / shuffles makes them uniform for the
    // compiler too, which otherwise wraps each of the loop's shuffles in a reconvergence sequence
    const int k = __shfl_sync(0xffffffffu, offs ? (int)(offs[p + 1] - offs[p]) : cnt[p], 0);
    const int64_t base = offs ? offs[p] : p * (int64_t)stride;
    if (k < 0) return;
    const bool box = xmin != nullptr;
    if (k == 0) {                                          // no rows: xdes, box ignored (sip.py:288)
        if (lane < N) dx[p * N + lane] = xdes[p * N + lane];
        if (lane == 0) status[p] = kQpNoRows;
        return;
    }
    const int m = k + (box ? N : 0);
    if (m > kQpMaxRows) { if (lane == 0) status[p] = kQpToHost; return; }
    if constexpr (MODE == kQpWide) {                       // the narrow launch listed it: two rows per lane
        solve_rows<N, 2>(p, k, m, base, box, W, xdes, xmin, xmax, rows, depths, reg, inflation, eps_abs, eps_rel, max_iter, kkt, skkt, sop, lane, dx, status, df);
    } else if (m <= 32) {
        solve_rows<N, 1, MODE == kQpNarrow>(p, k, m, base, box, W, xdes, xmin, xmax, rows, depths, reg, inflation, eps_abs, eps_rel, max_iter, kkt, skkt, sop, lane, dx, status, df);
    } else if constexpr (MODE == kQpNarrow || MODE == kQpNarrowReg) {
        // k_qp_list_wide listed it for the two-row launch
    } else {
        solve_rows<N, 2>(p, k, m, base, box, W, xdes, xmin, xmax, rows, depths, reg, inflation, eps_abs, eps_rel, max_iter, kkt, skkt, sop, lane, dx, status, df);
    }
}

// persistent warps: each draws the next problem from *counter (reset to 0 by the launcher).
// Deferral (the lock-step solver, QpDefer): a launch with df.list set runs every problem under a small iteration budget
// and lists those that exhaust it; a follow-up launch with df.only set solves exactly the listed problems from scratch
// under the real limit -- the same arithmetic as an uncapped solve -- on another stream, and raises df.done[p] when
// problem p's answer is in memory.  A launch lasts as long as its slowest problem; this keeps the 4000-iteration
// stragglers (0.5 % of the problems) off the critical path of the other 99.5 %.
template <int N, int MODE>
__global__ void __launch_bounds__(kQpWarps * 32, MODE == kQpNarrow ? kQpCtasNarrow : (MODE == kQpNarrowReg ? 3 : kQpCtasPerSm))
k_sip_qp(int64_t nprob, const double* __restrict__ W, const double* __restrict__ xdes, const double* __restrict__ xmin,
         const double* __restrict__ xmax, const int64_t* __restrict__ offs, const int32_t* __restrict__ cnt, int32_t stride,
         const double* __restrict__ rows,
         const double* __restrict__ depths, double reg, double inflation, double eps_abs, double eps_rel, int max_iter,
         double* __restrict__ scratch, unsigned long long* __restrict__ counter, double* __restrict__ dx, int32_t* __restrict__ status,
         QpDefer df) {
    const int lane = threadIdx.x & 31;
    double* kkt = scratch + (blockDim.x == 32 ? (size_t)blockIdx.x : (size_t)blockIdx.x * kQpWarps + (threadIdx.x >> 5)) * kQpScratch;
    __shared__ double s_kkt[kQpWarps][kQpSmemScratch];
    double* skkt = s_kkt[threadIdx.x >> 5];
    __shared__ double s_op[MODE == kQpNarrow ? kQpWarps : 1][MODE == kQpNarrow ? qp_ops_doubles<N>() : 1];   // Rows<.., SM>
    double* sop = s_op[MODE == kQpNarrow ? (threadIdx.x >> 5) : 0];
    const unsigned long long limit = MODE == kQpWide ? (unsigned long long)*df.wide_n
                                                     : (df.only ? (unsigned long long)*df.only_n : (unsigned long long)nprob);
    for (;;) {
        unsigned long long i = 0;
        if (lane == 0) i = atomicAdd(counter, 1ull);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= limit) return;
        const int64_t p = MODE == kQpWide ? (int64_t)df.wide[i] : (df.only ? (int64_t)df.only[i] : (int64_t)i);
        if (MODE != kQpWide && df.busy && df.busy[p] == 1) continue;
        solve_problem<N, MODE>(p, W, xdes, xmin, xmax, offs, cnt, stride, rows, depths, reg, inflation, eps_abs, eps_rel, max_iter, kkt, skkt, sop, lane, dx, status, df);
        if (df.only) {                                     // publish: answer first, flag second
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicExch(df.done + p, 1);
        }
    }
}

// The problems of a capped launch that need two rows per lane (32 < rows <= 64), listed BEFORE either solver kernel runs
// so that the one-row and the two-row launch can share the device: late in a BASELINE configs[4] solve half of the live
// problems are of this kind, and run one after the other each launch lasts as long as its own slowest problem.
__global__ void k_qp_list_wide(int64_t nprob, const int32_t* __restrict__ cnt, int nbox, QpDefer df) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= nprob) return;
    if (df.busy && df.busy[p] == 1) return;
    const int k = cnt[p];
    if (k <= 0) return;
    const int m = k + nbox;
    if (m > 32 && m <= kQpMaxRows) df.wide[atomicAdd(df.wide_n, 1u)] = (int32_t)p;
}

int sip_qp_grid(int num_sms) { return num_sms * kQpCtasPerSm; }           // 8 resident warps per SM at 255 registers: measured best (3 x 168 and 4 x 128 spill and run slower alone and loaded)
// polish scratch of the one-row launch's warps, then of the two-row launch's (they run side by side), then the two work counters
static size_t qp_narrow_doubles(int num_sms) { return (size_t)num_sms * kQpCtasNarrow * kQpWarps * kQpScratch; }
size_t sip_qp_scratch_bytes(int num_sms) { return (qp_narrow_doubles(num_sms) + (size_t)num_sms * kQpCtasPerSm * kQpWarps * kQpScratch) * sizeof(double) + 128; }

// scratch: sip_qp_scratch_bytes(num_sms) bytes; its last 128 bytes hold the work counters
int launch_sip_qp(int n, int64_t nprob, const double* W, const double* xdes, const double* xmin, const double* xmax,
                  const int64_t* offs, const int32_t* cnt, int32_t stride, const double* rows, const double* depths, double reg,
                  double inflation, double eps_abs, int max_iter, double* dx, int32_t* status, void* scratch, int num_sms,
                  cudaStream_t st, const QpDefer* defer) {
    if (nprob <= 0) return 0;
    const QpDefer df = defer ? *defer : QpDefer{};
    // a follow-up launch over a few hundred stragglers runs beside other kernels for milliseconds: one warp per CTA, so
    // that each straggler holds one warp's registers on its SM and not four
    const int threads = df.only ? 32 : kQpWarps * 32;
    const int blocks = df.only ? sip_qp_grid(num_sms) * kQpWarps : sip_qp_grid(num_sms);
    double* kkt = reinterpret_cast<double*>(scratch);
    unsigned long long* counter = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(scratch) + sip_qp_scratch_bytes(num_sms) - 128);
    unsigned long long* counter2 = counter + 8;
    cudaMemsetAsync(counter, 0, 128, st);
#define SIO_QP_LAUNCH(N, MODE)                                                                                                   \
    k_sip_qp<N, MODE><<<blocks, threads, 0, st>>>(nprob, W, xdes, xmin, xmax, offs, cnt, stride, rows, depths, reg, inflation, eps_abs, \
                                                  1e-5, max_iter, kkt, counter, dx, status, df)
    if (n != 6 && n != 7) return -1;
    if (df.wide) {                                         // capped launch: the two-row problems listed first, then both solvers
        if (offs) return -1;                               // padded layout only (the lock-step solver's)
        static const bool once = [] {                      // SIO_LS_PROFILE: resident CTAs per SM the driver grants each build
            if (getenv("SIO_LS_PROFILE")) {
                int a = 0, b = 0;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_sip_qp<6, kQpNarrow>, kQpWarps * 32, 0);
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_sip_qp<6, kQpWide>, kQpWarps * 32, 0);
                fprintf(stderr, "[sio qp] resident CTAs per SM: one-row build %d, two-row build %d\n", a, b);
            }
            return true;
        }();
        (void)once;
        cudaMemsetAsync(df.wide_n, 0, sizeof(uint32_t), st);
        k_qp_list_wide<<<(unsigned)((nprob + 255) / 256), 256, 0, st>>>(nprob, cnt, xmin ? n : 0, df);
        const bool beside = df.side && df.fork && df.join;
        cudaStream_t wst = beside ? df.side : st;
        if (beside) { cudaEventRecord(df.fork, st); cudaStreamWaitEvent(wst, df.fork, 0); }
        static const bool reg_build = getenv("SIO_QP_REGS") && atoi(getenv("SIO_QP_REGS")) != 0;   // A/B aid: operands in registers
        auto launch_narrow = [&]() {
            if (reg_build) {
                const int blocks = num_sms * 3;
                if (n == 6) SIO_QP_LAUNCH(6, kQpNarrowReg); else SIO_QP_LAUNCH(7, kQpNarrowReg);
            } else {
                const int blocks = num_sms * kQpCtasNarrow;
                if (n == 6) SIO_QP_LAUNCH(6, kQpNarrow); else SIO_QP_LAUNCH(7, kQpNarrow);
            }
        };
        if (!beside) launch_narrow();
        {
            double* kkt_n = kkt;
            double* kkt = kkt_n + qp_narrow_doubles(num_sms);
            unsigned long long* counter = counter2;
            cudaStream_t st = wst;
            // beside the one-row launch: ONE two-row CTA per SM (255 registers x 128 threads = half the register file), which
            // leaves room for a one-row CTA (168 x 128) next to it; alone, two per SM
            const int blocks = beside ? num_sms : sip_qp_grid(num_sms);
            if (n == 6) SIO_QP_LAUNCH(6, kQpWide); else SIO_QP_LAUNCH(7, kQpWide);
        }
        if (beside) {
            // the two-row launch goes first: it has the fewer, longer problems, and the one-row CTAs fill the SMs around it
            cudaEventRecord(df.join, wst);
            launch_narrow();
            cudaStreamWaitEvent(st, df.join, 0);
        }
    } else {
        if (n == 6) SIO_QP_LAUNCH(6, kQpFull); else SIO_QP_LAUNCH(7, kQpFull);
    }
#undef SIO_QP_LAUNCH
    return 0;
}

}  // namespace sio
