// sio_abi.cu -- host side of the C-ABI declared in include/sio_abi.h: context, resident handles,
// scratch arenas and the launch sequences of the oracle calls.  No CPU implementation of any query
// exists here: without a CUDA device sio_ctx_create fails with SIO_ERR_NODEVICE.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <chrono>
#include <thread>
#include <vector>

#include "../../include/sio_abi.h"
#include "sio_common.cuh"
#include "sio_lockstep.cuh"

using namespace sio;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(e_ == cudaErrorMemoryAllocation ? SIO_ERR_NOMEM : SIO_ERR_CUDA, "%s: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                               \
    } while (0)

// grow-only device / pinned-host buffers; growth is geometric (a lock-step SIP run asks for a little more
// every iteration, and cudaFree / cudaMalloc / cudaMallocHost each cost a device synchronisation and milliseconds)
inline size_t grown(size_t bytes, size_t cap) { return std::max(std::max(bytes, (size_t)256), cap + cap / 2); }
// bumped by every re-allocation of a scratch buffer: a captured CUDA graph holds the buffers' addresses AND the
// driver's registration of the pinned ones, so a graph captured before any re-allocation must never be replayed
// (freeing a pinned buffer and getting the same address back is not "the same buffer" to the driver: replaying a
// graph with such a copy node crashed inside cudaGraphLaunch)
std::atomic<uint64_t> g_realloc_gen{0};
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        size_t want = grown(bytes, cap);
        g_realloc_gen.fetch_add(1);
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) { g_realloc_gen.fetch_add(1); cudaFree(p); } p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};
struct HostBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        size_t want = grown(bytes, cap);
        g_realloc_gen.fetch_add(1);
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) { g_realloc_gen.fetch_add(1); cudaFreeHost(p); } p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct GridH {
    bool live = false;
    float* d_v = nullptr;
    float* d_brick = nullptr;
    uint8_t* d_lipd = nullptr;
    GridDev meta{};
    GridDev* d_meta = nullptr;     // single-element device copy (K1 kernels take a pointer)
};
struct CloudH {
    bool live = false;
    int64_t n = 0;
    float4* d_pts = nullptr;       // Morton-sorted
    float4* d_rep = nullptr;       // per 128-point sub-tile: representative cloud point (xyz) + bounding radius (w)
    int32_t* d_perm = nullptr;     // internal -> caller index
    std::vector<int32_t> perm;
    double max_abs = 0.0;          // largest |coordinate|: the float32 transform error grows with it
};
struct RobotH {
    bool live = false;
    int32_t n = 0;
    int32_t *d_parent = nullptr, *d_jtype = nullptr;
    double *d_tparent = nullptr, *d_axis = nullptr;
    RobotDev dev() const { return RobotDev{n, d_parent, d_jtype, d_tparent, d_axis}; }
};

struct UnderEntry { int64_t b, idx; double d; };

}  // namespace

struct sio_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_bulk0 = nullptr, ev_bulk1 = nullptr;   // bracket the fp32 bulk launches of the last K2 call
    bool bulk_timed = false;
    double tau = 0.0;                 // <= 0: automatic
    int64_t launches = 0;
    std::vector<GridH> grids;
    std::vector<CloudH> clouds;
    std::vector<RobotH> robots;
    // scratch
    DevBuf s_lb, s_ub, s_items, s_icnt;
    int64_t scratch_limit = (int64_t)(256u << 20);   // bytes of sub-tile minima kept per K2 batch (sio_ctx_set_scratch_limit)
    int64_t qp_iters_total = 0, qp_iters_max = 0, qp_at_limit = 0;      // ADMM iterations of the last sio_sip_step_qp call, summed over problems
    DevBuf s_qp;                     // per-warp polish scratch + work counter of the SIP-step QP kernel
    DevBuf s_ls;                     // per-problem state of the lock-step solver (sio_lockstep_pose_solve)
    DevBuf s_ls_cloud;               // the environment cloud in world coordinates (float64), kept between calls:
    sio_handle ls_cloud_h = -1;      //   by contract it is a function of (cloud, T_env), so the 6 MB upload is skipped when
    double ls_cloud_T[12] = {0};     //   both are those of the previous call
    // background QP launches of the lock-step solver (the deferred stragglers): their own streams and polish scratch
    static constexpr int kLsBack = 3, kLsRing = 8;
    cudaStream_t ls_stream[kLsBack] = {nullptr, nullptr, nullptr};
    DevBuf s_qpb[kLsBack];
    cudaEvent_t ls_evA[kLsRing] = {}, ls_evB[kLsRing] = {};
    cudaStream_t qp_side = nullptr;   // the two-row QP launch of a round runs here, beside the one-row launch on the main stream
    cudaEvent_t qp_fork = nullptr, qp_join = nullptr;
    HostBuf h_icnt;
    int n_icnt = 0;                   // batches of the last pruned call (their survivor counts sit in h_icnt)
    int64_t prune_total = 0;          // (transform, sub-tile) pairs of the last pruned call
    DevBuf s_xf, s_keys, s_gridtab, s_cand, s_cnt, s_in0, s_in1, s_in2, s_in3, s_out0, s_out1, s_out2, s_tlinks, s_trel, s_gob;
    HostBuf h_a, h_b, h_c, h_cnt, h_ls, h_ls_out, h_ls_pack;
    DevBuf s_ls_pack;
    uint32_t cand_cap = 1u << 20;     // under-bound candidate records the buffer holds (sio_ctx_set_under_capacity; host entry points grow it)
    GridDev single_grid_host{};       // host copy handed to the single-grid bulk kernel by value
    std::vector<sio_handle> staged_grids;   // handles whose table currently sits in s_gridtab (grids are immutable)
    double staged_tau = 0.0;
    // host-pointer K2 calls repeated with the same signature are replayed as ONE CUDA graph (H2D, memsets, kernels, D2H)
    struct HostGraph {
        std::vector<int64_t> key;        // everything the captured launch sequence depends on (see host_graph_key)
        cudaGraphExec_t exec = nullptr;
        int n_icnt = 0; int64_t prune_total = 0; int64_t launches = 0; bool f64 = false;
    } hgraph;
    std::vector<int64_t> last_host_key;  // signature of the previous host-pointer K2 call (a repeat is what gets captured)
    bool host_graphs = true;
    // result boards (multi-GPU exchange by peer stores): this rank, the boards of all ranks as device pointers
    int board_rank = -1, board_n = 0;
    int64_t board_slot = 0;
    void* board_local = nullptr;
    DevBuf s_boards;
    // state of the last min-distance call (for sio_under_fetch)
    bool under_valid = false, under_pending = false, under_final = false;
    std::vector<UnderEntry> under;
};

namespace {

constexpr unsigned long long kMaxCand = (1ull << 31) - 1;      // record slots are 32-bit in the kernels (16 B each: 34 GB)

bool grid_ok(sio_ctx* c, sio_handle h) { return h >= 0 && (size_t)h < c->grids.size() && c->grids[h].live; }
bool cloud_ok(sio_ctx* c, sio_handle h) { return h >= 0 && (size_t)h < c->clouds.size() && c->clouds[h].live; }
bool robot_ok(sio_ctx* c, sio_handle h) { return h >= 0 && (size_t)h < c->robots.size() && c->robots[h].live; }

// Wait for a stream the way a latency-bound caller wants it: the single-problem paths (PenetrationDepthGeometry.distance,
// value / df_dx batches of a few points) wait for ~10-100 us of device work per call, and a blocking
// cudaStreamSynchronize costs 10-20 us of wake-up latency on top.  Poll for up to ~0.5 ms, then block.
// the same without giving up: for loops that wait for milliseconds of device work many times over (the lock-step solver)
inline cudaError_t spin_stream(cudaStream_t st) {
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        const cudaError_t e = cudaStreamQuery(st);
        if (e != cudaErrorNotReady) return e;
        if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(200)) break;    // something is wrong or huge: block
    }
    return cudaStreamSynchronize(st);
}
inline cudaError_t wait_stream(cudaStream_t st) {
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        const cudaError_t e = cudaStreamQuery(st);
        if (e != cudaErrorNotReady) return e;
        if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(500)) break;
    }
    return cudaStreamSynchronize(st);
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// the automatic near-tie window: generous multiple of the fp32 evaluation error, which scales with the
// magnitude of the samples and of the coordinates (DESIGN.md "Exactness")
double auto_tau(const GridDev* g, int n) {
    double scale = 1.0;
    for (int i = 0; i < n; ++i) {
        double diag = 0;
        for (int a = 0; a < 3; ++a) diag += (g[i].bmax[a] - g[i].bmin[a]) * (g[i].bmax[a] - g[i].bmin[a]);
        scale = std::max(scale, std::sqrt(diag) + std::max(std::fabs((double)g[i].vmin), std::fabs((double)g[i].vmax)));
    }
    return 1e-5 * scale;
}

uint64_t morton3(uint32_t x, uint32_t y, uint32_t z) {   // 21 bits per axis
    auto spread = [](uint64_t v) {
        v &= 0x1fffffULL;
        v = (v | v << 32) & 0x1f00000000ffffULL;
        v = (v | v << 16) & 0x1f0000ff0000ffULL;
        v = (v | v << 8) & 0x100f00f00f00f00fULL;
        v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
        v = (v | v << 2) & 0x1249249249249249ULL;
        return v;
    };
    return spread(x) | (spread(y) << 1) | (spread(z) << 2);
}

}  // namespace

extern "C" {

int sio_abi_version(void) { return SIO_ABI_VERSION; }
const char* sio_last_error(void) { return g_err.c_str(); }

int sio_ctx_create(int device, sio_ctx** out) {
    if (!out) return fail(SIO_ERR_INVALID, "sio_ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(SIO_ERR_NODEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(SIO_ERR_INVALID, "device %d out of range [0,%d)", device, ndev);
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(SIO_ERR_NODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    sio_ctx* c = new sio_ctx();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return fail(SIO_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    cudaEventCreate(&c->ev_bulk0);
    cudaEventCreate(&c->ev_bulk1);
    if (getenv("SIO_HOST_GRAPHS")) c->host_graphs = atoi(getenv("SIO_HOST_GRAPHS")) != 0;
    *out = c;
    return SIO_OK;
}

int sio_ctx_destroy(sio_ctx* c) {
    if (!c) return SIO_OK;
    DeviceGuard guard(c->device);
    cudaStreamSynchronize(c->stream);
    for (int k = 0; k < sio_ctx::kLsBack; ++k) if (c->ls_stream[k]) cudaStreamSynchronize(c->ls_stream[k]);
    if (c->qp_side) { cudaStreamSynchronize(c->qp_side); cudaStreamDestroy(c->qp_side); cudaEventDestroy(c->qp_fork); cudaEventDestroy(c->qp_join); }
    for (auto& g : c->grids) if (g.live) { cudaFree(g.d_v); cudaFree(g.d_brick); cudaFree(g.d_lipd); cudaFree(g.d_meta); }
    for (auto& p : c->clouds) if (p.live) { cudaFree(p.d_pts); cudaFree(p.d_rep); cudaFree(p.d_perm); }
    for (auto& r : c->robots) if (r.live) { cudaFree(r.d_parent); cudaFree(r.d_jtype); cudaFree(r.d_tparent); cudaFree(r.d_axis); }
    c->s_qp.release();
    if (c->hgraph.exec) cudaGraphExecDestroy(c->hgraph.exec);
    c->s_boards.release();
    c->s_lb.release(); c->s_ub.release(); c->s_items.release(); c->s_icnt.release(); c->h_icnt.release(); c->s_ls.release(); c->s_ls_cloud.release();
    for (int k = 0; k < sio_ctx::kLsBack; ++k) { c->s_qpb[k].release(); if (c->ls_stream[k]) cudaStreamDestroy(c->ls_stream[k]); }
    for (int k = 0; k < sio_ctx::kLsRing; ++k) { if (c->ls_evA[k]) cudaEventDestroy(c->ls_evA[k]); if (c->ls_evB[k]) cudaEventDestroy(c->ls_evB[k]); }
    DevBuf* bufs[] = {&c->s_xf, &c->s_keys, &c->s_gridtab, &c->s_cand, &c->s_cnt, &c->s_in0, &c->s_in1, &c->s_in2, &c->s_in3,
                      &c->s_out0, &c->s_out1, &c->s_out2, &c->s_tlinks, &c->s_trel, &c->s_gob};
    for (auto* b : bufs) b->release();
    c->h_a.release(); c->h_b.release(); c->h_c.release(); c->h_cnt.release(); c->h_ls.release(); c->h_ls_out.release(); c->h_ls_pack.release(); c->s_ls_pack.release();
    cudaEventDestroy(c->ev_bulk0); cudaEventDestroy(c->ev_bulk1);
    cudaStreamDestroy(c->stream);
    delete c;
    return SIO_OK;
}

int sio_ctx_sync(sio_ctx* c) {
    if (!c) return fail(SIO_ERR_INVALID, "null context");
    DeviceGuard guard(c->device);
    CU(cudaStreamSynchronize(c->stream));
    return SIO_OK;
}
void* sio_ctx_stream(sio_ctx* c) { return c ? (void*)c->stream : nullptr; }
int sio_ctx_set_tau(sio_ctx* c, double tau) { if (!c) return fail(SIO_ERR_INVALID, "null context"); c->tau = tau; return SIO_OK; }
int sio_ctx_set_scratch_limit(sio_ctx* c, int64_t bytes) {
    if (!c) return fail(SIO_ERR_INVALID, "null context");
    c->scratch_limit = bytes > 0 ? bytes : (int64_t)(256u << 20);
    return SIO_OK;
}
int sio_ctx_set_under_capacity(sio_ctx* c, int64_t entries) {
    if (!c) return fail(SIO_ERR_INVALID, "null context");
    if (entries > (int64_t)kMaxCand) return fail(SIO_ERR_INVALID, "sio_ctx_set_under_capacity: at most %lld entries", (long long)kMaxCand);
    c->cand_cap = entries > 0 ? (uint32_t)entries : (1u << 20);
    return SIO_OK;
}
int sio_ctx_num_sms(sio_ctx* c) { return c ? c->num_sms : 0; }
int sio_ctx_set_host_graphs(sio_ctx* c, int on) {
    if (!c) return fail(SIO_ERR_INVALID, "null context");
    c->host_graphs = on != 0;
    if (!on && c->hgraph.exec) { cudaGraphExecDestroy(c->hgraph.exec); c->hgraph.exec = nullptr; c->hgraph.key.clear(); }
    return SIO_OK;
}

// ---- result boards ---------------------------------------------------------------------------------------------------
int sio_board_create(sio_ctx* c, int64_t bytes, void** dptr, unsigned char ipc_handle[64]) {
    if (!c || bytes <= 0 || !dptr) return fail(SIO_ERR_INVALID, "sio_board_create: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard guard(c->device);
    void* p = nullptr;
    CU(cudaMalloc(&p, (size_t)bytes));
    CU(cudaMemset(p, 0, (size_t)bytes));
    if (ipc_handle) {
        cudaIpcMemHandle_t h;
        cudaError_t e = cudaIpcGetMemHandle(&h, p);
        if (e != cudaSuccess) { cudaFree(p); return fail(SIO_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
        std::memcpy(ipc_handle, &h, 64);
    }
    *dptr = p;
    return SIO_OK;
}
int sio_board_open(sio_ctx* c, const unsigned char ipc_handle[64], void** dptr) {
    if (!c || !ipc_handle || !dptr) return fail(SIO_ERR_INVALID, "sio_board_open: bad argument");
    DeviceGuard guard(c->device);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, ipc_handle, 64);
    CU(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SIO_OK;
}
int sio_board_close(sio_ctx* c, void* dptr, int is_local) {
    if (!c || !dptr) return fail(SIO_ERR_INVALID, "sio_board_close: bad argument");
    DeviceGuard guard(c->device);
    cudaStreamSynchronize(c->stream);
    if (is_local) { if (c->board_local == dptr) { c->board_n = 0; c->board_local = nullptr; } CU(cudaFree(dptr)); }
    else CU(cudaIpcCloseMemHandle(dptr));
    return SIO_OK;
}
int sio_ctx_set_boards(sio_ctx* c, int32_t rank, int32_t nranks, void* const* boards, int64_t slot_bytes) {
    if (!c) return fail(SIO_ERR_INVALID, "null context");
    if (nranks <= 0) { c->board_n = 0; c->board_rank = -1; c->board_local = nullptr; return SIO_OK; }
    if (!boards || rank < 0 || rank >= nranks || slot_bytes <= 0 || slot_bytes % 8) return fail(SIO_ERR_INVALID, "sio_ctx_set_boards: bad argument");
    DeviceGuard guard(c->device);
    CU(cudaStreamSynchronize(c->stream));
    CU(c->s_boards.ensure((size_t)nranks * sizeof(void*)));
    CU(cudaMemcpy(c->s_boards.p, boards, (size_t)nranks * sizeof(void*), cudaMemcpyHostToDevice));
    c->board_rank = rank; c->board_n = nranks; c->board_slot = slot_bytes; c->board_local = boards[rank];
    return SIO_OK;
}
int sio_board_read(sio_ctx* c, void* host_out, int64_t bytes) {
    if (!c || !host_out || bytes <= 0 || !c->board_local) return fail(SIO_ERR_INVALID, "sio_board_read: no board set");
    if (bytes > c->board_slot * c->board_n) return fail(SIO_ERR_INVALID, "sio_board_read: %lld bytes asked, the board has %lld", (long long)bytes, (long long)(c->board_slot * c->board_n));
    DeviceGuard guard(c->device);
    CU(cudaMemcpyAsync(host_out, c->board_local, (size_t)bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return SIO_OK;
}
int64_t sio_ctx_kernel_launches(sio_ctx* c) { return c ? c->launches : 0; }

// ---- geometry --------------------------------------------------------------------------------------
int sio_grid_create(sio_ctx* c, const float* values, const int32_t dims[3], const double bmin[3],
                    const double bmax[3], sio_handle* out) {
    if (!c || !values || !dims || !bmin || !bmax || !out) return fail(SIO_ERR_INVALID, "sio_grid_create: null argument");
    for (int a = 0; a < 3; ++a) {
        if (dims[a] < 1) return fail(SIO_ERR_INVALID, "sio_grid_create: dims[%d]=%d", a, dims[a]);
        if (!(bmax[a] > bmin[a])) return fail(SIO_ERR_INVALID, "sio_grid_create: empty bbox on axis %d", a);
    }
    if ((int64_t)(dims[0] + 1) * (dims[1] + 1) * (dims[2] + 1) >= (1LL << 31))
        return fail(SIO_ERR_INVALID, "sio_grid_create: grid too large for 32-bit cell indexing");
    DeviceGuard guard(c->device);
    const int nx = dims[0], ny = dims[1], nz = dims[2];
    const int px = nx + 1, py = ny + 1, pz = nz + 1;
    // padded copy: the extra far layer in each axis replicates the last sample
    std::vector<float> pad((size_t)px * py * pz);
    float vmin = std::numeric_limits<float>::infinity(), vmax = -vmin;
    for (int i = 0; i < px; ++i) {
        int si = std::min(i, nx - 1);
        for (int j = 0; j < py; ++j) {
            int sj = std::min(j, ny - 1);
            const float* src = values + ((size_t)si * ny + sj) * nz;
            float* dst = pad.data() + ((size_t)i * py + j) * pz;
            std::memcpy(dst, src, sizeof(float) * nz);
            dst[nz] = src[nz - 1];
        }
    }
    for (size_t k = 0, n = (size_t)nx * ny * nz; k < n; ++k) {
        float v = values[k];
        if (!(v == v)) return fail(SIO_ERR_INVALID, "sio_grid_create: NaN sample at flat index %zu", k);
        vmin = std::min(vmin, v); vmax = std::max(vmax, v);
    }
    // bricks: one 32-byte record per cell (i0 in [0, dim-1]; corner +1 comes from the padded copy), first as the
    // 8 corner samples [v000 v001 v010 v011 v100 v101 v110 v111], converted to monomial coefficients below
    std::vector<float> brick((size_t)nx * ny * nz * 8);
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j)
            for (int k = 0; k < nz; ++k) {
                float* dst = brick.data() + (((size_t)i * ny + j) * nz + k) * 8;
                for (int di = 0; di < 2; ++di)
                    for (int dj = 0; dj < 2; ++dj)
                        for (int dk = 0; dk < 2; ++dk)
                            dst[di * 4 + dj * 2 + dk] = pad[((size_t)(i + di) * py + (j + dj)) * pz + (k + dk)];
            }
    // Lipschitz constant of the trilinear interpolant: inside a cell |dv/dx| <= max over its four x-edges of
    // |difference| / hx (likewise y, z), so sqrt of the sum of squares bounds the gradient norm there; take the
    // maximum over cells.  Used only by the sub-tile culling (a valid bound keeps pruned results exact).
    double lip = 0.0;
    std::vector<uint8_t> lipq((size_t)kLipLevels * nx * ny * nz);       // local constants, see GridDev::lipd
    {
        const double hx = (bmax[0] - bmin[0]) / nx, hy = (bmax[1] - bmin[1]) / ny, hz = (bmax[2] - bmin[2]) / nz;
        for (size_t cidx = 0, nc = (size_t)nx * ny * nz; cidx < nc; ++cidx) {
            const float* q = brick.data() + cidx * 8;
            double ex = 0, ey = 0, ez = 0;
            for (int m = 0; m < 4; ++m) {
                ex = std::max(ex, (double)std::fabs(q[4 + m] - q[m]));                       // x-edges: (1,j,k)-(0,j,k)
                ey = std::max(ey, (double)std::fabs(q[(m & 2) * 2 + 2 + (m & 1)] - q[(m & 2) * 2 + (m & 1)]));  // y-edges
                ez = std::max(ez, (double)std::fabs(q[2 * m + 1] - q[2 * m]));               // z-edges
            }
            ex /= hx; ey /= hy; ez /= hz;
            const double lc = std::sqrt(ex * ex + ey * ey + ez * ez);
            lip = std::max(lip, lc);
            const double q64 = std::ceil(lc * (1.0 + 1e-6) * 64.0 + 1e-4);        // rounded up: stays a bound
            lipq[cidx] = (uint8_t)(q64 < 255.0 ? q64 : 255.0);                    // 255 = fall back to the global constant
        }
        // level D = the maximum over the cells within D cells (Chebyshev): D dilations by one cell, each separable
        const size_t nc = (size_t)nx * ny * nz;
        std::vector<uint8_t> cur(lipq.begin(), lipq.begin() + nc), tmp(nc);
        const int dims[3] = {nx, ny, nz};
        const size_t strides[3] = {(size_t)ny * nz, (size_t)nz, 1};
        for (int D = 1; D <= kLipLevels; ++D) {
            for (int ax = 0; ax < 3; ++ax) {
                for (size_t i = 0; i < nc; ++i) {
                    const int pos = (int)((i / strides[ax]) % dims[ax]);
                    uint8_t m = cur[i];
                    if (pos > 0) m = std::max(m, cur[i - strides[ax]]);
                    if (pos + 1 < dims[ax]) m = std::max(m, cur[i + strides[ax]]);
                    tmp[i] = m;
                }
                cur.swap(tmp);
            }
            std::copy(cur.begin(), cur.end(), lipq.begin() + (size_t)(D - 1) * nc);
        }
    }
    // corner samples -> monomial coefficients (Moebius transform per axis bit, in float64)
    for (size_t cidx = 0, nc = (size_t)nx * ny * nz; cidx < nc; ++cidx) {
        float* q = brick.data() + cidx * 8;
        double m[8];
        for (int k = 0; k < 8; ++k) m[k] = q[k];
        for (int bit = 1; bit < 8; bit <<= 1)
            for (int k = 0; k < 8; ++k)
                if (k & bit) m[k] -= m[k ^ bit];
        // stored as the register pairs of the packed Horner evaluation: [c6 c2 | c7 c3 | c4 c0 | c5 c1] (sio_kernels.cu)
        static const int pos[8] = {5, 7, 1, 3, 4, 6, 0, 2};
        for (int k = 0; k < 8; ++k) q[pos[k]] = (float)m[k];
    }
    GridH g;
    g.meta.lip = (float)(lip * (1.0 + 1e-6)) + 1e-6f;
    CU(cudaMalloc(&g.d_v, pad.size() * sizeof(float)));
    CU(cudaMemcpyAsync(g.d_v, pad.data(), pad.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMalloc(&g.d_lipd, lipq.size()));
    CU(cudaMemcpyAsync(g.d_lipd, lipq.data(), lipq.size(), cudaMemcpyHostToDevice, c->stream));
    g.meta.lipd = g.d_lipd;
    CU(cudaMalloc(&g.d_brick, brick.size() * sizeof(float)));
    CU(cudaMemcpyAsync(g.d_brick, brick.data(), brick.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    g.meta.v = g.d_v;
    g.meta.brick = g.d_brick;
    g.meta.nx = nx; g.meta.ny = ny; g.meta.nz = nz;
    g.meta.sy = pz; g.meta.sx = py * pz;
    for (int a = 0; a < 3; ++a) {
        g.meta.bmin[a] = bmin[a]; g.meta.bmax[a] = bmax[a];
        g.meta.h[a] = (bmax[a] - bmin[a]) / (double)dims[a];
        g.meta.ih[a] = 1.0 / g.meta.h[a];
    }
    g.meta.vmin = vmin; g.meta.vmax = vmax;
    CU(cudaMalloc(&g.d_meta, sizeof(GridDev)));
    CU(cudaMemcpyAsync(g.d_meta, &g.meta, sizeof(GridDev), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    g.live = true;
    c->grids.push_back(g);
    *out = (sio_handle)(c->grids.size() - 1);
    return SIO_OK;
}

int sio_grid_destroy(sio_ctx* c, sio_handle h) {
    if (!c || !grid_ok(c, h)) return fail(SIO_ERR_INVALID, "sio_grid_destroy: bad handle %d", h);
    DeviceGuard guard(c->device);
    cudaStreamSynchronize(c->stream);
    g_realloc_gen.fetch_add(1);
    cudaFree(c->grids[h].d_v); cudaFree(c->grids[h].d_brick); cudaFree(c->grids[h].d_lipd); cudaFree(c->grids[h].d_meta);
    c->grids[h] = GridH();
    c->staged_grids.clear();
    return SIO_OK;
}

// the resident order of a cloud (pure host work, no device needed): order_out[internal] = caller index
int sio_cloud_order(const float* xyz, int64_t n, int32_t* order_out) {
    if (n < 0 || (n > 0 && (!xyz || !order_out))) return fail(SIO_ERR_INVALID, "sio_cloud_order: bad argument");
    // Point order (one-off, host).  Sub-tiles of 128 consecutive points are the unit of the bulk kernels (one warp) and
    // of the culling (one bounding sphere), so they must be spatially tight: a k-d ordering -- median splits along the
    // longest axis, left halves sized in whole sub-tiles -- down to leaves of 128, Morton order inside a leaf so that the
    // lanes of a warp touch neighbouring cells.  (Plain Morton chunks straddle the curve's jumps: on the C5 scene 15 %
    // of them had radii above 15 cm, up to 1.9 m, and survived every cull; k-d leaves: 3 %, at most 0.35 m.)
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            float v = xyz[3 * i + a];
            if (!std::isfinite(v)) return fail(SIO_ERR_INVALID, "sio_cloud_create: non-finite coordinate at point %lld", (long long)i);
            lo[a] = std::min(lo[a], v); hi[a] = std::max(hi[a], v);
        }
    std::vector<std::pair<uint64_t, int32_t>> keys((size_t)n);
    float span = 0.f;
    for (int a = 0; a < 3; ++a) span = std::max(span, hi[a] - lo[a]);
    const double sc = span > 0 ? 2097151.0 / span : 0.0;
    for (int64_t i = 0; i < n; ++i) {
        uint32_t q[3];
        for (int a = 0; a < 3; ++a) q[a] = (uint32_t)std::min(2097151.0, std::max(0.0, (xyz[3 * i + a] - lo[a]) * sc));
        keys[i] = {morton3(q[0], q[1], q[2]), (int32_t)i};
    }
    {
        const int64_t leaf = 32 * kPtsPerThread;
        std::vector<std::pair<int64_t, int64_t>> todo;
        if (n > 0) todo.push_back({0, n});
        while (!todo.empty()) {
            const int64_t a0 = todo.back().first, a1 = todo.back().second;
            todo.pop_back();
            if (a1 - a0 <= leaf) { std::sort(keys.begin() + a0, keys.begin() + a1); continue; }
            float blo[3] = {INFINITY, INFINITY, INFINITY}, bhi[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (int64_t i = a0; i < a1; ++i)
                for (int a = 0; a < 3; ++a) {
                    const float v = xyz[3 * (int64_t)keys[i].second + a];
                    blo[a] = std::min(blo[a], v); bhi[a] = std::max(bhi[a], v);
                }
            int ax = 0;
            for (int a = 1; a < 3; ++a) if (bhi[a] - blo[a] > bhi[ax] - blo[ax]) ax = a;
            const int64_t left = std::min(a1 - a0 - 1, ((a1 - a0) / 2 + leaf - 1) / leaf * leaf);
            std::nth_element(keys.begin() + a0, keys.begin() + a0 + left, keys.begin() + a1,
                             [&](const std::pair<uint64_t, int32_t>& u, const std::pair<uint64_t, int32_t>& v) {
                                 const float fu = xyz[3 * (int64_t)u.second + ax], fv = xyz[3 * (int64_t)v.second + ax];
                                 return fu < fv || (fu == fv && u.second < v.second);
                             });
            todo.push_back({a0 + left, a1});
            todo.push_back({a0, a0 + left});
        }
    }
    for (int64_t i = 0; i < n; ++i) order_out[i] = keys[i].second;
    return SIO_OK;
}

int sio_cloud_create(sio_ctx* c, const float* xyz, int64_t n, sio_handle* out) {
    if (!c || !out || n < 0 || (n > 0 && !xyz)) return fail(SIO_ERR_INVALID, "sio_cloud_create: bad argument");
    if (n >= (1LL << 31) - kTile) return fail(SIO_ERR_INVALID, "sio_cloud_create: more than 2^31 points");
    DeviceGuard guard(c->device);
    CloudH p;
    p.n = n;
    std::vector<int32_t> order((size_t)n);
    {
        const int rc = sio_cloud_order(xyz, n, order.data());
        if (rc != SIO_OK) return rc;
    }
    const int64_t npad = ((n + kTile - 1) / kTile) * kTile;
    std::vector<float4> pts((size_t)std::max<int64_t>(npad, 1));
    p.perm.resize((size_t)std::max<int64_t>(npad, 1), -1);
    for (int64_t i = 0; i < n; ++i) {
        int32_t s = order[i];
        pts[i] = make_float4(xyz[3 * s], xyz[3 * s + 1], xyz[3 * s + 2], 0.f);
        p.perm[i] = s;
        for (int a = 0; a < 3; ++a) p.max_abs = std::max(p.max_abs, (double)std::fabs(xyz[3 * s + a]));
    }
    // pad to a whole tile with copies of the last point: the bulk kernel needs no bounds test, a duplicate
    // cannot lower a minimum, and the fp64 stages only ever look at indices < n
    for (int64_t i = n; i < npad; ++i) pts[i] = n > 0 ? pts[n - 1] : make_float4(0.f, 0.f, 0.f, 0.f);
    // sub-tile culling data: for every 128 consecutive (Morton-ordered) points, the point nearest to their
    // centroid and the radius of the sphere about it that contains all of them (rounded up)
    const int64_t sub = 32 * kPtsPerThread;
    const int64_t nsubt = std::max<int64_t>(npad / sub, 1);
    std::vector<float4> reps((size_t)nsubt, make_float4(0.f, 0.f, 0.f, 0.f));
    for (int64_t s = 0; s * sub < npad; ++s) {
        const float4* q = pts.data() + s * sub;
        double cx = 0, cy = 0, cz = 0;
        for (int64_t i = 0; i < sub; ++i) { cx += q[i].x; cy += q[i].y; cz += q[i].z; }
        cx /= sub; cy /= sub; cz /= sub;
        int64_t best = 0; double bd = 1e300;
        for (int64_t i = 0; i < sub; ++i) {
            double d2 = (q[i].x - cx) * (q[i].x - cx) + (q[i].y - cy) * (q[i].y - cy) + (q[i].z - cz) * (q[i].z - cz);
            if (d2 < bd) { bd = d2; best = i; }
        }
        double r2 = 0;
        for (int64_t i = 0; i < sub; ++i) {
            double dx = (double)q[i].x - q[best].x, dy = (double)q[i].y - q[best].y, dz = (double)q[i].z - q[best].z;
            r2 = std::max(r2, dx * dx + dy * dy + dz * dz);
        }
        reps[s] = make_float4(q[best].x, q[best].y, q[best].z, (float)(std::sqrt(r2) * (1.0 + 1e-6)) + 1e-7f);
    }
    CU(cudaMalloc(&p.d_pts, pts.size() * sizeof(float4)));
    CU(cudaMalloc(&p.d_rep, reps.size() * sizeof(float4)));
    CU(cudaMemcpyAsync(p.d_rep, reps.data(), reps.size() * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMalloc(&p.d_perm, p.perm.size() * sizeof(int32_t)));
    CU(cudaMemcpyAsync(p.d_pts, pts.data(), pts.size() * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(p.d_perm, p.perm.data(), p.perm.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    p.live = true;
    c->clouds.push_back(std::move(p));
    *out = (sio_handle)(c->clouds.size() - 1);
    return SIO_OK;
}

int sio_cloud_destroy(sio_ctx* c, sio_handle h) {
    if (!c || !cloud_ok(c, h)) return fail(SIO_ERR_INVALID, "sio_cloud_destroy: bad handle %d", h);
    DeviceGuard guard(c->device);
    cudaStreamSynchronize(c->stream);
    g_realloc_gen.fetch_add(1);
    cudaFree(c->clouds[h].d_pts); cudaFree(c->clouds[h].d_rep); cudaFree(c->clouds[h].d_perm);
    c->clouds[h] = CloudH();
    return SIO_OK;
}

int64_t sio_cloud_size(sio_ctx* c, sio_handle h) { return (c && cloud_ok(c, h)) ? c->clouds[h].n : -1; }

int sio_cloud_perm(sio_ctx* c, sio_handle h, int32_t* perm_out) {
    if (!c || !cloud_ok(c, h) || !perm_out) return fail(SIO_ERR_INVALID, "sio_cloud_perm: bad argument");
    std::memcpy(perm_out, c->clouds[h].perm.data(), sizeof(int32_t) * (size_t)c->clouds[h].n);
    return SIO_OK;
}

// ---- K1 ----------------------------------------------------------------------------------------------
int sio_query(sio_ctx* c, sio_handle grid, const double T[12], const double* pts, int64_t P, int flags,
              double* d, double* grad) {
    if (!c || !grid_ok(c, grid) || !T || P < 0 || (P > 0 && (!pts || !d)))
        return fail(SIO_ERR_INVALID, "sio_query: bad argument");
    if (P == 0) return SIO_OK;
    DeviceGuard guard(c->device);
    cudaStream_t st = c->stream;
    // one pinned block each way: [T | pts] -> device, [d | grad] <- device (a point query is latency, not bandwidth)
    const size_t in_bytes = (12 + (size_t)P * 3) * sizeof(double);
    const size_t out_bytes = ((size_t)P + (grad ? (size_t)P * 3 : 0)) * sizeof(double);
    CU(c->h_a.ensure(in_bytes));
    CU(c->h_b.ensure(out_bytes));
    CU(c->s_in0.ensure(in_bytes));
    CU(c->s_out0.ensure(out_bytes));
    double* hin = c->h_a.as<double>();
    std::memcpy(hin, T, 12 * sizeof(double));
    std::memcpy(hin + 12, pts, (size_t)P * 3 * sizeof(double));
    CU(cudaMemcpyAsync(c->s_in0.p, hin, in_bytes, cudaMemcpyHostToDevice, st));
    const double* dT = c->s_in0.as<double>();
    const double* dpts = dT + 12;
    double* dd = c->s_out0.as<double>();
    double* dg = grad ? dd + P : nullptr;
    const int normalize = (flags & SIO_NO_NORMALIZE) ? 0 : 1;
    if (flags & SIO_F64)
        launch_query_f64(c->grids[grid].d_meta, dT, dpts, P, normalize, dd, dg, st);
    else
        launch_query_f32(c->grids[grid].d_meta, dT, dpts, P, normalize, dd, dg, st);
    c->launches += 1;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->h_b.p, dd, out_bytes, cudaMemcpyDeviceToHost, st));
    CU(wait_stream(st));
    std::memcpy(d, c->h_b.p, (size_t)P * sizeof(double));
    if (grad) std::memcpy(grad, c->h_b.as<double>() + P, (size_t)P * 3 * sizeof(double));
    return SIO_OK;
}

// builds the device grid table for a call; returns the automatic tau through *tau
static int stage_grid_table(sio_ctx* c, const sio_handle* grids, int32_t n_grids, double* tau) {
    if (n_grids <= 0 || !grids) return fail(SIO_ERR_INVALID, "no grids given");
    if ((int)c->staged_grids.size() == n_grids && std::equal(grids, grids + n_grids, c->staged_grids.begin())) {
        *tau = c->tau > 0 ? c->tau : c->staged_tau;
        return SIO_OK;                      // same table as the previous call: nothing to copy
    }
    c->staged_grids.clear();
    std::vector<GridDev> tab((size_t)n_grids);
    for (int i = 0; i < n_grids; ++i) {
        if (!grid_ok(c, grids[i])) return fail(SIO_ERR_INVALID, "bad grid handle %d at position %d", grids[i], i);
        tab[i] = c->grids[grids[i]].meta;
    }
    CU(c->s_gridtab.ensure(tab.size() * sizeof(GridDev)));
    // pageable -> device copy of a stack/vector object: synchronous w.r.t. the host, so `tab` may die
    CU(cudaMemcpyAsync(c->s_gridtab.p, tab.data(), tab.size() * sizeof(GridDev), cudaMemcpyHostToDevice, c->stream));
    c->staged_tau = auto_tau(tab.data(), n_grids);
    c->staged_grids.assign(grids, grids + n_grids);
    *tau = c->tau > 0 ? c->tau : c->staged_tau;
    return SIO_OK;
}

int sio_cloud_query_dev(sio_ctx* c, sio_handle grid, sio_handle cloud, const double* dT_rel, int64_t B, int flags,
                        float* out_d, float* out_g4) {
    if (!c || !grid_ok(c, grid) || !cloud_ok(c, cloud) || B < 0 || (B > 0 && !dT_rel) || (!out_d && !out_g4))
        return fail(SIO_ERR_INVALID, "sio_cloud_query_dev: bad argument");
    if (flags & SIO_F64) return fail(SIO_ERR_INVALID, "sio_cloud_query_dev: SIO_F32 only");
    if (B == 0) return SIO_OK;
    DeviceGuard guard(c->device);
    double tau;
    int rc = stage_grid_table(c, &grid, 1, &tau);
    if (rc) return rc;
    CU(c->s_xf.ensure((size_t)B * sizeof(XfRec)));
    MinArgs a{};
    a.grids = c->s_gridtab.as<GridDev>(); a.grid_of_b = nullptr; a.T_rel = dT_rel; a.bound = nullptr;
    a.B = B; a.tau = tau; a.xf = c->s_xf.as<XfRec>();
    launch_prep_xf(a, c->stream);
    const CloudH& p = c->clouds[cloud];
    launch_cloud_query_f32(p.d_pts, p.d_rep, p.n, a.xf, B, (flags & SIO_NO_NORMALIZE) ? 0 : 1, out_g4 ? nullptr : out_d,
                           reinterpret_cast<float4*>(out_g4), c->stream);
    c->launches += 2;
    CU(cudaGetLastError());
    return SIO_OK;
}

// ---- K2 ----------------------------------------------------------------------------------------------
int sio_min_distance_dev(sio_ctx* c, const sio_handle* grids, int32_t n_grids, const int32_t* d_grid_of_b,
                         sio_handle cloud, const double* dT_rel, const double* d_bound, int64_t B, int flags,
                         double* d_dmin, int64_t* d_argmin, int32_t* d_n_under) {
    if (!c || !cloud_ok(c, cloud) || B < 0 || (B > 0 && (!dT_rel || !d_dmin || !d_argmin)))
        return fail(SIO_ERR_INVALID, "sio_min_distance_dev: bad argument");
    if (B >= (1LL << 31)) return fail(SIO_ERR_INVALID, "sio_min_distance_dev: B too large");
    DeviceGuard guard(c->device);
    c->under.clear(); c->under_valid = false; c->under_pending = false;
    if (B == 0) { c->under_valid = true; return SIO_OK; }
    double tau;
    int rc = stage_grid_table(c, grids, n_grids, &tau);
    if (rc) return rc;
    cudaStream_t st = c->stream;
    const CloudH& p = c->clouds[cloud];
    // the automatic near-tie window also has to cover the float32 error of the transform itself, ~2e-7 x the magnitude of
    // the cloud-local coordinates (cancellation in M p + t): a cloud stored hundreds of metres from its origin widens it.
    // (A T_rel whose translation is far larger than both geometries is the caller's to cover with sio_ctx_set_tau.)
    if (!(c->tau > 0)) tau = std::max(tau, 1e-5 * p.max_abs);
    const bool f64 = (flags & SIO_F64) != 0;
    const int32_t ntiles = (int32_t)((p.n + kTile - 1) / kTile);
    const int64_t nsub = (int64_t)ntiles * (kThreads / 32);
    // bound the sub-tile key scratch to ~256 MB by sweeping the transforms in batches
    int64_t Bc = B;
    if (!f64 && nsub > 0) {
        int64_t maxb = c->scratch_limit / (nsub * 4);
        maxb = std::max<int64_t>(kXfChunk, (maxb / kXfChunk) * kXfChunk);
        Bc = std::min(B, maxb);
    }
    CU(c->s_cnt.ensure(sizeof(unsigned long long)));
    {
        cudaError_t e = c->s_cand.ensure((size_t)c->cand_cap * sizeof(Cand));
        if (e != cudaSuccess) {
            const unsigned long long asked = c->cand_cap;
            c->cand_cap = 1u << 20;                           // the next call starts from the default again
            cudaGetLastError();
            return fail(SIO_ERR_NOMEM, "sio_min_distance_dev: no memory for %llu under-bound candidate records (%s)", asked, cudaGetErrorString(e));
        }
    }
    CU(cudaMemsetAsync(c->s_cnt.p, 0, sizeof(unsigned long long), st));
    if (d_n_under) CU(cudaMemsetAsync(d_n_under, 0, (size_t)B * sizeof(int32_t), st));
    MinArgs a{};
    a.pts = p.d_pts; a.perm = p.d_perm; a.N = p.n; a.rep = p.d_rep;
    a.grids = c->s_gridtab.as<GridDev>(); a.grid_of_b = d_grid_of_b;
    a.single_grid = nullptr;
    if (n_grids == 1) { c->single_grid_host = c->grids[grids[0]].meta; a.single_grid = &c->single_grid_host; }
    // no counter array = the caller wants the minimum only: no candidates are produced, no re-check runs
    const bool want_under = d_bound && d_n_under;
    a.T_rel = dT_rel; a.bound = d_bound; a.under_bound = want_under ? d_bound : nullptr; a.B = B; a.tau = tau; a.ntiles = ntiles;
    a.cand = c->s_cand.as<Cand>(); a.cand_count = c->s_cnt.as<unsigned long long>(); a.cand_cap = c->cand_cap; a.num_sms = c->num_sms;
    a.n_under = want_under ? d_n_under : nullptr;
    a.dmin = d_dmin; a.argmin = d_argmin;
    const bool prune = !f64 && (flags & SIO_PRUNE) && p.n > 0;
    const int max_batches = 4096;
    if (!f64) {
        CU(c->s_keys.ensure((size_t)std::max<int64_t>(1, Bc * nsub) * sizeof(uint32_t)));
        a.subkeys = c->s_keys.as<uint32_t>();
    }
    c->n_icnt = 0; c->prune_total = 0;
    if (prune) {
        if ((B + Bc - 1) / Bc > max_batches) return fail(SIO_ERR_INVALID, "sio_min_distance_dev: too many batches for SIO_PRUNE");
        CU(c->s_xf.ensure((size_t)B * sizeof(XfRec)));
        CU(c->s_lb.ensure((size_t)(((Bc + 31) & ~(int64_t)31) * nsub) * sizeof(float)));      // sub-tile-major, transforms padded to whole warps
        CU(c->s_items.ensure((size_t)(Bc * nsub) * sizeof(uint32_t)));
        CU(c->s_ub.ensure((size_t)B * sizeof(uint32_t)));
        CU(c->s_icnt.ensure(sizeof(uint32_t) * max_batches));
        CU(c->h_icnt.ensure(sizeof(uint32_t) * max_batches));
        CU(cudaMemsetAsync(c->s_ub.p, 0xff, (size_t)B * sizeof(uint32_t), st));
        CU(cudaMemsetAsync(c->s_icnt.p, 0, sizeof(uint32_t) * ((B + Bc - 1) / Bc), st));
        a.xf = c->s_xf.as<XfRec>(); a.lb = c->s_lb.as<float>(); a.ub = c->s_ub.as<uint32_t>();
        a.items = c->s_items.as<uint32_t>();
        launch_prep_xf(a, st);
        c->launches += 1;
        c->prune_total = B * nsub;
    }
    const bool smem_variant = !f64 && !prune && (flags & SIO_SMEM_GRID) != 0;
    if (smem_variant) {
        if (want_under) return fail(SIO_ERR_INVALID, "SIO_SMEM_GRID computes minima only: pass n_under = NULL");
        CU(c->s_xf.ensure((size_t)B * sizeof(XfRec)));
        a.xf = c->s_xf.as<XfRec>();
        launch_prep_xf(a, st);
        c->launches += 1;
    }
    c->bulk_timed = false;
    // the bulk launch is bracketed by timing events unless the caller is capturing the call into a CUDA graph
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    const bool one_batch = Bc >= B && cap == cudaStreamCaptureStatusNone;
    for (int64_t b0 = 0; b0 < B; b0 += Bc) {
        a.b_begin = b0; a.Bc = std::min(Bc, B - b0);
        if (prune) {
            a.item_count = c->s_icnt.as<uint32_t>() + c->n_icnt;
            if (one_batch) cudaEventRecord(c->ev_bulk0, st);
            launch_min_pruned(a, st);
            if (one_batch) { cudaEventRecord(c->ev_bulk1, st); c->bulk_timed = true; }
            c->n_icnt += 1;
            c->launches += 3;
        } else if (!f64 && smem_variant) {
            if (one_batch) cudaEventRecord(c->ev_bulk0, st);
            if (!launch_min_smem(a, st)) return fail(SIO_ERR_INVALID, "SIO_SMEM_GRID: needs one grid whose padded samples fit 200 KB");
            if (one_batch) { cudaEventRecord(c->ev_bulk1, st); c->bulk_timed = true; }
            c->launches += 1;
        } else if (!f64) {
            if (one_batch) cudaEventRecord(c->ev_bulk0, st);
            launch_min_f32(a, st);
            if (one_batch) { cudaEventRecord(c->ev_bulk1, st); c->bulk_timed = true; }
            c->launches += (p.n > 0);
        }
        // the last batch's launch also carries the CTAs that re-check the float32 pass's under-bound candidates
        const bool last = b0 + Bc >= B;
        launch_finalize(a, /*all_tiles=*/f64, /*emit_under=*/f64 && want_under, (!f64 && want_under && last) ? c->num_sms : 0, st);
        c->launches += 1;
    }
    if (c->board_n > 0 && (flags & SIO_PUSH_BOARDS)) {      // the results also go to every rank's board, over NVLink
        if (B * 20 > c->board_slot) return fail(SIO_ERR_INVALID, "sio_min_distance_dev: %lld results do not fit a board slot of %lld bytes", (long long)B, (long long)c->board_slot);
        launch_push_results(c->s_boards.as<void*>(), c->board_n, c->board_rank, c->board_slot, d_dmin, d_argmin, want_under ? d_n_under : nullptr, B, st);
        c->launches += 1;
    }
    CU(cudaGetLastError());
    c->under_pending = want_under;
    c->under_final = f64;
    c->under_valid = !c->under_pending;
    return SIO_OK;
}

// pulls the candidate records of the last call, filters and sorts them into ctx->under
static int collect_under(sio_ctx* c) {
    if (c->under_valid) return SIO_OK;
    cudaStream_t st = c->stream;
    CU(c->h_c.ensure(sizeof(unsigned long long)));
    CU(cudaMemcpyAsync(c->h_c.p, c->s_cnt.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const unsigned long long n = *c->h_c.as<unsigned long long>();
    if (n > c->cand_cap)
        return fail(SIO_ERR_OVERFLOW, "under-bound candidate buffer overflow: %llu candidates, capacity %u (sio_ctx_set_under_capacity)", n, c->cand_cap);
    c->under.clear();
    if (n > 0) {
        std::vector<Cand> recs(n);
        CU(cudaMemcpy(recs.data(), c->s_cand.p, (size_t)n * sizeof(Cand), cudaMemcpyDeviceToHost));
        // (b, idx) order: bucket by transform (a C5-scale set has tens of millions of entries), then sort each bucket
        int32_t bmax = -1;
        for (const Cand& r : recs) bmax = std::max(bmax, r.b);
        std::vector<size_t> start((size_t)bmax + 2, 0);
        for (const Cand& r : recs) if (r.b >= 0) ++start[(size_t)r.b + 1];
        for (size_t b = 1; b < start.size(); ++b) start[b] += start[b - 1];
        c->under.resize(start.back());
        {
            std::vector<size_t> fill(start.begin(), start.end() - 1);
            for (const Cand& r : recs) if (r.b >= 0) c->under[fill[(size_t)r.b]++] = {r.b, r.idx, r.d};
        }
        for (size_t b = 0; b + 1 < start.size(); ++b)
            std::sort(c->under.begin() + start[b], c->under.begin() + start[b + 1],
                      [](const UnderEntry& x, const UnderEntry& y) { return x.idx < y.idx; });
    }
    c->under_valid = true; c->under_pending = false;
    return SIO_OK;
}

int sio_under_fetch(sio_ctx* c, int64_t capacity, int64_t* b_out, int64_t* idx_out, double* d_out, int64_t* total) {
    if (!c || !total) return fail(SIO_ERR_INVALID, "sio_under_fetch: bad argument");
    DeviceGuard guard(c->device);
    int rc = collect_under(c);
    if (rc) return rc;
    *total = (int64_t)c->under.size();
    int64_t m = std::min<int64_t>(capacity, *total);
    for (int64_t i = 0; i < m; ++i) {
        if (b_out) b_out[i] = c->under[i].b;
        if (idx_out) idx_out[i] = c->under[i].idx;
        if (d_out) d_out[i] = c->under[i].d;
    }
    return (*total > capacity && capacity > 0) ? fail(SIO_ERR_OVERFLOW, "under list has %lld entries, capacity %lld", (long long)*total, (long long)capacity) : SIO_OK;
}

// A host entry point has enqueued its sweep and its output copies: wait for them and see whether the under-bound
// candidates fitted.  Returns 0 = done, 1 = the buffer was too small and has been re-sized: run the sweep again
// (the float32 pass is deterministic, so the second run offers the same candidates), < 0 = error.
static int under_settle(sio_ctx* c, bool want_under) {
    cudaStream_t st = c->stream;
    if (!want_under) { CU(wait_stream(st)); return 0; }
    CU(c->h_cnt.ensure(sizeof(unsigned long long)));
    CU(cudaMemcpyAsync(c->h_cnt.p, c->s_cnt.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CU(wait_stream(st));
    const unsigned long long n = *c->h_cnt.as<unsigned long long>();
    if (n <= c->cand_cap) return 0;
    if (n > kMaxCand)
        return fail(SIO_ERR_OVERFLOW, "%llu under-bound candidates in one call, more than the %llu the list can hold: split the batch", n, kMaxCand);
    c->cand_cap = (uint32_t)std::min<unsigned long long>(kMaxCand, n + n / 8 + 4096);
    return 1;
}

// host-pointer K2: inputs are packed into one pinned block and sent with ONE H2D copy; outputs come back
// with ONE D2H copy:   in  = [T_rel B*12 f64 | bound B f64 | grid_of_b B i32]
//                      out = [dmin B f64 | argmin B i64 | n_under B i32]
// everything the launch sequence of a host-pointer K2 call depends on: arguments' shapes, settings, and the addresses
// of every buffer it touches (they are grow-only, so equal addresses = nothing was re-allocated)
static std::vector<int64_t> host_graph_key(sio_ctx* c, const sio_handle* grids, int32_t n_grids, bool has_gob, sio_handle cloud,
                                           bool has_bound, int64_t B, int flags, bool want_n) {
    std::vector<int64_t> k;
    k.reserve(32 + n_grids);
    k.push_back(n_grids);
    for (int i = 0; i < n_grids; ++i) k.push_back(grids[i]);
    const int64_t scalars[] = {cloud, B, flags, has_gob, has_bound, want_n, (int64_t)c->cand_cap, c->scratch_limit, c->board_n, c->board_rank};
    k.insert(k.end(), std::begin(scalars), std::end(scalars));
    int64_t taubits;
    std::memcpy(&taubits, &c->tau, 8);
    k.push_back(taubits);
    k.push_back((int64_t)g_realloc_gen.load());
    const void* bufs[] = {c->h_a.p, c->h_b.p, c->h_cnt.p, c->s_in0.p, c->s_out0.p, c->s_cand.p, c->s_cnt.p, c->s_keys.p, c->s_xf.p, c->s_lb.p,
                          c->s_ub.p, c->s_items.p, c->s_icnt.p, c->s_gridtab.p, c->s_boards.p};
    for (const void* p : bufs) k.push_back((int64_t)(intptr_t)p);
    return k;
}

// host-pointer K2: inputs are packed into one pinned block and sent with ONE H2D copy; outputs come back
// with ONE D2H copy:   in  = [T_rel B*12 f64 | bound B f64 | grid_of_b B i32]
//                      out = [dmin B f64 | argmin B i64 | n_under B i32]
// A call that repeats the previous call's signature is captured into a CUDA graph (copies, memsets, kernels) and every
// further repeat replays it with one launch: the SIP loop issues the same call thousands of times, and at 64 poses the
// seven driver calls of the plain sequence cost as much as a fifth of the sweep itself.
static int min_distance_host(sio_ctx* c, const sio_handle* grids, int32_t n_grids, const int32_t* grid_of_b, sio_handle cloud,
                             const double* T_rel, const double* bound, int64_t B, int flags, double* dmin, int64_t* argmin,
                             int32_t* n_under) {
    cudaStream_t st = c->stream;
    const size_t oT = 0, oB = (size_t)B * 96, oG = oB + (size_t)B * 8, in_bytes = oG + (size_t)B * 4;
    const size_t oD = 0, oA = (size_t)B * 8, oN = oA + (size_t)B * 8, out_bytes = oN + (size_t)B * 4;
    CU(c->h_a.ensure(in_bytes));
    CU(c->h_b.ensure(out_bytes));
    CU(c->h_cnt.ensure(sizeof(unsigned long long)));
    CU(c->s_in0.ensure(in_bytes));
    CU(c->s_out0.ensure(out_bytes));
    char* hin = c->h_a.as<char>();
    std::memcpy(hin + oT, T_rel, (size_t)B * 96);
    if (bound) std::memcpy(hin + oB, bound, (size_t)B * 8);
    if (grid_of_b) {
        for (int64_t b = 0; b < B; ++b)
            if (grid_of_b[b] < 0 || grid_of_b[b] >= n_grids) return fail(SIO_ERR_INVALID, "grid_of_b[%lld]=%d out of range", (long long)b, grid_of_b[b]);
        std::memcpy(hin + oG, grid_of_b, (size_t)B * 4);
    }
    const bool want_under = bound && n_under;
    char* din = c->s_in0.as<char>();
    char* dout = c->s_out0.as<char>();
    auto enqueue = [&]() -> int {                           // the call's whole device-side sequence, on the context's stream
        CU(cudaMemcpyAsync(c->s_in0.p, hin, in_bytes, cudaMemcpyHostToDevice, st));
        int rc = sio_min_distance_dev(c, grids, n_grids, grid_of_b ? reinterpret_cast<const int32_t*>(din + oG) : nullptr, cloud,
                                      reinterpret_cast<const double*>(din + oT), bound ? reinterpret_cast<const double*>(din + oB) : nullptr,
                                      B, flags, reinterpret_cast<double*>(dout + oD), reinterpret_cast<int64_t*>(dout + oA),
                                      n_under ? reinterpret_cast<int32_t*>(dout + oN) : nullptr);
        if (rc) return rc;
        CU(cudaMemcpyAsync(c->h_b.p, dout, out_bytes, cudaMemcpyDeviceToHost, st));
        if (want_under) CU(cudaMemcpyAsync(c->h_cnt.p, c->s_cnt.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        return SIO_OK;
    };
    auto overflowed = [&]() { return want_under && *c->h_cnt.as<unsigned long long>() > c->cand_cap; };
    bool done = false;
    const std::vector<int64_t> key = c->host_graphs ? host_graph_key(c, grids, n_grids, grid_of_b != nullptr, cloud, bound != nullptr, B, flags, n_under != nullptr)
                                                    : std::vector<int64_t>();
    if (c->host_graphs && c->hgraph.exec && c->hgraph.key == key) {            // replay
        CU(cudaGraphLaunch(c->hgraph.exec, st));
        CU(wait_stream(st));
        c->under.clear();
        c->under_pending = want_under; c->under_valid = !want_under; c->under_final = c->hgraph.f64;
        c->bulk_timed = false; c->n_icnt = c->hgraph.n_icnt; c->prune_total = c->hgraph.prune_total;
        c->launches += c->hgraph.launches;
        done = !overflowed();                               // an overflow falls through to the plain path, which re-sizes
    } else if (c->host_graphs && !key.empty() && key == c->last_host_key) {    // second identical call in a row: capture it
        if (c->hgraph.exec) { cudaGraphExecDestroy(c->hgraph.exec); c->hgraph.exec = nullptr; c->hgraph.key.clear(); }
        const int64_t launches0 = c->launches;
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            const int rc = enqueue();
            const cudaError_t e = cudaStreamEndCapture(st, &graph);
            if (rc == SIO_OK && e == cudaSuccess && graph && cudaGraphInstantiate(&c->hgraph.exec, graph, 0) == cudaSuccess) {
                c->hgraph.key = key;
                c->hgraph.n_icnt = c->n_icnt; c->hgraph.prune_total = c->prune_total; c->hgraph.f64 = (flags & SIO_F64) != 0;
                c->hgraph.launches = c->launches - launches0;
                if (graph) cudaGraphDestroy(graph);
                CU(cudaGraphLaunch(c->hgraph.exec, st));
                CU(wait_stream(st));
                done = !overflowed();
            } else {                                        // capture refused something: stay on the plain path for good
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                c->hgraph.exec = nullptr; c->hgraph.key.clear();
                c->host_graphs = false;
            }
        }
    }
    while (!done) {
        int rc = enqueue();
        if (rc) return rc;
        rc = under_settle(c, want_under);
        if (rc < 0) return rc;
        if (rc == 0) done = true;
    }
    c->last_host_key = c->host_graphs ? host_graph_key(c, grids, n_grids, grid_of_b != nullptr, cloud, bound != nullptr, B, flags, n_under != nullptr)
                                      : std::vector<int64_t>();
    const char* hout = c->h_b.as<char>();
    std::memcpy(dmin, hout + oD, (size_t)B * 8);
    std::memcpy(argmin, hout + oA, (size_t)B * 8);
    if (n_under) {
        if (bound) std::memcpy(n_under, hout + oN, (size_t)B * 4);
        else std::memset(n_under, 0, (size_t)B * 4);
    }
    return SIO_OK;
}

int sio_min_distance(sio_ctx* c, const sio_handle* grids, int32_t n_grids, const int32_t* grid_of_b, sio_handle cloud,
                     const double* T_rel, const double* bound, int64_t B, int flags, double* dmin, int64_t* argmin,
                     int32_t* n_under) {
    if (!c || B < 0 || (B > 0 && (!T_rel || !dmin || !argmin))) return fail(SIO_ERR_INVALID, "sio_min_distance: bad argument");
    DeviceGuard guard(c->device);
    if (B == 0) { c->under.clear(); c->under_valid = true; return SIO_OK; }
    return min_distance_host(c, grids, n_grids, grid_of_b, cloud, T_rel, bound, B, flags, dmin, argmin, n_under);
}

// ---- K3 ----------------------------------------------------------------------------------------------
int sio_pose_rows(sio_ctx* c, sio_handle grid, const double* T_pose, const int64_t* offsets, int64_t nprob,
                  const double* pts, int flags, double* d, double* rows) {
    if (!c || !grid_ok(c, grid) || nprob < 0 || (nprob > 0 && (!T_pose || !offsets)))
        return fail(SIO_ERR_INVALID, "sio_pose_rows: bad argument");
    if (nprob == 0) return SIO_OK;
    const int64_t total = offsets[nprob];
    if (offsets[0] != 0 || total < 0) return fail(SIO_ERR_INVALID, "sio_pose_rows: offsets must start at 0 and be non-decreasing");
    for (int64_t p = 0; p < nprob; ++p)
        if (offsets[p + 1] < offsets[p]) return fail(SIO_ERR_INVALID, "sio_pose_rows: offsets decrease at %lld", (long long)p);
    if (total == 0) return SIO_OK;
    if (!pts || !d) return fail(SIO_ERR_INVALID, "sio_pose_rows: null pts/d");
    DeviceGuard guard(c->device);
    cudaStream_t st = c->stream;
    // one pinned staging block each way: [T_pose | pts | offsets] -> device, [d | rows] <- device
    const size_t oT = 0, oP = (size_t)nprob * 12 * sizeof(double), oO = oP + (size_t)total * 3 * sizeof(double);
    const size_t in_bytes = oO + (size_t)(nprob + 1) * sizeof(int64_t);
    const size_t oD = 0, oR = (size_t)total * sizeof(double), out_bytes = oR + (rows ? (size_t)total * 6 * sizeof(double) : 0);
    CU(c->h_a.ensure(in_bytes));
    CU(c->h_b.ensure(out_bytes));
    CU(c->s_in0.ensure(in_bytes));
    CU(c->s_out0.ensure(out_bytes));
    char* h = c->h_a.as<char>();
    std::memcpy(h + oT, T_pose, oP);
    std::memcpy(h + oP, pts, (size_t)total * 3 * sizeof(double));
    std::memcpy(h + oO, offsets, (size_t)(nprob + 1) * sizeof(int64_t));
    CU(cudaMemcpyAsync(c->s_in0.p, h, in_bytes, cudaMemcpyHostToDevice, st));
    char* din = c->s_in0.as<char>();
    char* dout = c->s_out0.as<char>();
    launch_pose_rows(c->grids[grid].d_meta, reinterpret_cast<double*>(din + oT), reinterpret_cast<int64_t*>(din + oO), nprob,
                     reinterpret_cast<double*>(din + oP), total, (flags & SIO_NO_NORMALIZE) ? 0 : 1,
                     reinterpret_cast<double*>(dout + oD), rows ? reinterpret_cast<double*>(dout + oR) : nullptr, st);
    c->launches += 1;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->h_b.p, dout, out_bytes, cudaMemcpyDeviceToHost, st));
    CU(wait_stream(st));
    std::memcpy(d, c->h_b.as<char>() + oD, (size_t)total * sizeof(double));
    if (rows) std::memcpy(rows, c->h_b.as<char>() + oR, (size_t)total * 6 * sizeof(double));
    return SIO_OK;
}

// ---- lock-step pose solver ----------------------------------------------------------------------------------
}  // extern "C"
namespace {
// one device allocation carved into the solver's arrays; freed when it goes out of scope
struct Arena {                                              // carves a context-owned buffer; frees nothing
    char* base = nullptr; size_t used = 0, cap = 0;
    template <class T> T* take(size_t n) {
        used = (used + 255) / 256 * 256;
        T* p = reinterpret_cast<T*>(base + used);
        used += n * sizeof(T);
        return p;
    }
};
template <class F> size_t carve(Arena& A, LsState& S, int64_t B, int cap, F&& extra) {
    S.R = A.take<double>(9 * B); S.t = A.take<double>(3 * B); S.Rdes = A.take<double>(9 * B); S.tdes = A.take<double>(3 * B);
    S.P = A.take<double>((size_t)B * cap * 3); S.key = A.take<long long>((size_t)B * cap * 3); S.cnt = A.take<int32_t>(B);
    S.D = A.take<double>((size_t)B * cap); S.Dn = A.take<double>((size_t)B * cap); S.rows = A.take<double>((size_t)B * cap * 6);
    S.w = A.take<double>(B); S.trs = A.take<double>(B); S.fx = A.take<double>(B); S.gx = A.take<double>(B);
    S.gx0 = A.take<double>(B); S.min_cv = A.take<double>(B); S.gxc = A.take<double>(B);
    S.upd = A.take<uint8_t>(B); S.live = A.take<uint8_t>(B); S.go = A.take<uint8_t>(B);
    S.status = A.take<int32_t>(B); S.niter = A.take<int32_t>(B); S.ninst0 = A.take<int32_t>(B);
    S.dxdes = A.take<double>(6 * B); S.dx = A.take<double>(6 * B); S.Rn = A.take<double>(9 * B); S.tn = A.take<double>(3 * B);
    S.fxn = A.take<double>(B); S.gxn = A.take<double>(B); S.sorig = A.take<double>(B); S.metric = A.take<double>(B);
    S.qW = A.take<double>(6 * B); S.qlo = A.take<double>(6 * B); S.qhi = A.take<double>(6 * B);
    S.qk = A.take<int32_t>(B); S.qstatus = A.take<int32_t>(B);
    S.list = A.take<int32_t>(B); S.counters = A.take<unsigned long long>(8);
    S.ph = A.take<uint8_t>(B); S.mu = A.take<uint8_t>(B); S.qdef = A.take<uint8_t>(B); S.pit = A.take<int32_t>(B); S.qdone = A.take<int32_t>(B);
    S.T_rel = A.take<double>(12 * B); S.bound = A.take<double>(B); S.dmin = A.take<double>(B); S.arg = A.take<int64_t>(B);
    extra();
    return A.used;
}
}  // namespace
extern "C" {

int sio_lockstep_pose_solve(sio_ctx* c, sio_handle grid, sio_handle cloud, const double T_env[12], const double* cloud_world,
                            int64_t B, const double* T_init, const double* T_des, const sio_lockstep_settings* cfg,
                            double* T_out, int32_t* status, int32_t* iterations, double* fx, double* gx, double* gx0,
                            int32_t* n_params, double* params, int64_t stats[6]) {
    if (!c || !grid_ok(c, grid) || !cloud_ok(c, cloud) || !T_env || !cloud_world || B < 0 || !cfg ||
        (B > 0 && (!T_init || !T_des || !T_out)))
        return fail(SIO_ERR_INVALID, "sio_lockstep_pose_solve: bad argument");
    if (cfg->cap < 1 || cfg->cap > 58) return fail(SIO_ERR_INVALID, "sio_lockstep_pose_solve: cap must be in [1, 58], got %d", cfg->cap);
    if (B >= (1LL << 31) / 64) return fail(SIO_ERR_INVALID, "sio_lockstep_pose_solve: too many problems");
    if (stats) stats[0] = stats[1] = stats[2] = stats[3] = stats[4] = stats[5] = 0;
    if (B == 0) return SIO_OK;
    DeviceGuard guard(c->device);
    cudaStream_t st = c->stream;
    // every way out of this call -- errors included -- first waits for the background QP launches: they read and write
    // the context-owned arena that the next call re-carves
    struct BackgroundDrain {
        sio_ctx* c;
        ~BackgroundDrain() {
            for (int k = 0; k < sio_ctx::kLsBack; ++k) if (c->ls_stream[k]) cudaStreamSynchronize(c->ls_stream[k]);
            if (c->qp_side) cudaStreamSynchronize(c->qp_side);
        }
    } drain{c};
    const auto entry0 = std::chrono::steady_clock::now();
    const int cap = cfg->cap;
    const int64_t N = c->clouds[cloud].n;
    LsState S{};
    S.B = B; S.N = N; S.cap = cap;
    for (int k = 0; k < 12; ++k) S.Tenv[k] = T_env[k];
    double* d_cloud = nullptr;
    Arena A;
    {   // size pass on a null arena, then the real one
        Arena probe; LsState tmp{};
        size_t need = carve(probe, tmp, B, cap, [&] {
            for (int k = 0; k < sio_ctx::kLsRing; ++k) probe.take<int32_t>(B);
            probe.take<uint32_t>(sio_ctx::kLsRing);
            probe.take<int32_t>(B);
            probe.take<uint32_t>(1);
            probe.take<double>(24 * B);
        });
        CU(c->s_ls.ensure(need + 4096));                  // kept by the context: releasing ~350 MB per call cost tens of ms
        A.base = c->s_ls.as<char>();
        A.cap = need + 4096;
    }
    int32_t* dlist[sio_ctx::kLsRing];                      // problems each of the last rounds' capped QP launches gave up on
    uint32_t* dcount = nullptr;
    int32_t* wlist = nullptr;                              // problems of the current round with more than 32 QP rows
    uint32_t* wcount = nullptr;
    double* raw_poses = nullptr;                           // T_init, T_des as handed over; unpacked on the device
    carve(A, S, B, cap, [&] {
        for (int k = 0; k < sio_ctx::kLsRing; ++k) dlist[k] = A.take<int32_t>(B);
        dcount = A.take<uint32_t>(sio_ctx::kLsRing);
        wlist = A.take<int32_t>(B);
        wcount = A.take<uint32_t>(1);
        raw_poses = A.take<double>(24 * B);
    });
    {
        const bool same = c->ls_cloud_h == cloud && c->s_ls_cloud.p && std::memcmp(c->ls_cloud_T, T_env, sizeof(c->ls_cloud_T)) == 0;
        if (!same) {
            c->ls_cloud_h = -1;
            CU(c->s_ls_cloud.ensure((size_t)std::max<int64_t>(N, 1) * 3 * sizeof(double)));
            CU(cudaMemcpyAsync(c->s_ls_cloud.p, cloud_world, (size_t)N * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
            c->ls_cloud_h = cloud;
            std::memcpy(c->ls_cloud_T, T_env, sizeof(c->ls_cloud_T));
        }
        d_cloud = c->s_ls_cloud.as<double>();
    }
    S.cloud_world = d_cloud;
    S.grid = c->grids[grid].d_meta;
    // poses: two copies as they are (row-major 3x4); R (3x3) and t are split out on the device
    CU(cudaMemcpyAsync(raw_poses, T_init, (size_t)B * 12 * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(raw_poses + 12 * B, T_des, (size_t)B * 12 * 8, cudaMemcpyHostToDevice, st));
    LsSettings L{};
    L.max_iters = cfg->max_iters; L.cap = cap;
    L.xeps2 = cfg->xepsilon * cfg->xepsilon * 2;           // len((R,t)) == 2 in the reference (sip.py:836-837)     quirk
    L.inflation = cfg->constraint_inflation; L.drop_value = cfg->constraint_drop_value; L.min_cv = cfg->minimum_constraint_value;
    L.w0 = cfg->initial_objective_score_weight; L.w_min = cfg->minimum_objective_score_weight;
    L.excl = cfg->parameter_exclusion_distance; L.reg = cfg->qp_regularization;
    CU(cudaMemsetAsync(S.counters, 0, 8 * sizeof(unsigned long long), st));
    CU(cudaMemsetAsync(S.key, 0, (size_t)B * cap * 3 * sizeof(long long), st));
    ls_launch_unpack_poses(S, raw_poses, st);
    ls_launch_init(S, L, st);
    CU(c->s_qp.ensure(sip_qp_scratch_bytes(c->num_sms)));
    // the loop's counters come back twice per round into pinned memory (a pageable destination makes the "async" copy a
    // staged, synchronous one), and the host SPINS on the stream for them: a blocking synchronise adds tens of
    // microseconds of wake-up latency per read-back, 128 of them per solve, with the device idle meanwhile
    CU(c->h_ls.ensure(8 * sizeof(unsigned long long)));
    unsigned long long* const hc = c->h_ls.as<unsigned long long>();
    int64_t queries = 0;
    // device time of the QP launches and of the sweeps, read back at the end (events per iteration)
    struct EventBag {                                      // destroyed on every way out of the call
        std::vector<cudaEvent_t> v;
        ~EventBag() { for (cudaEvent_t e : v) cudaEventDestroy(e); }
    } ev_bag, span_bag;
    std::vector<cudaEvent_t>& ev = ev_bag.v;
    auto mark = [&]() { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); ev.push_back(e); };
    // SIO_LS_PROFILE=1: device time of every phase of the loop, printed to stderr at the end (a development aid)
    const bool prof = getenv("SIO_LS_PROFILE") != nullptr && atoi(getenv("SIO_LS_PROFILE")) == 1;
    const bool prof_total = getenv("SIO_LS_PROFILE") != nullptr && !prof;      // 2: only the call's wall time, no extra syncs
    enum { kPhPre = 0, kPhPost, kPhGx, kPhStep, kPhScore, kPhCount };
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> spans;
    cudaEvent_t span_open = nullptr;
    auto span_begin = [&]() { if (prof) { cudaEventCreate(&span_open); span_bag.v.push_back(span_open); cudaEventRecord(span_open, st); } };
    auto span_end = [&](int label) {
        if (prof) { cudaEvent_t e; cudaEventCreate(&e); span_bag.v.push_back(e); cudaEventRecord(e, st); spans.push_back({label, {span_open, e}}); }
    };
    const auto wall0 = std::chrono::steady_clock::now();
    if (prof) {
        cudaStreamSynchronize(st);
        fprintf(stderr, "[sio lockstep] setup (arena, staging, init) %.1f ms\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - entry0).count());
    }
    auto read_counters = [&]() -> cudaError_t {
        cudaError_t e = cudaMemcpyAsync(hc, S.counters, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return e;
        return spin_stream(st);
    };
    auto sweep = [&](int64_t n) -> int {                   // K2 over the queued problems: minimum / arg-minimum only
        queries += n * N;
        mark();
        int rc = sio_min_distance_dev(c, &grid, 1, nullptr, cloud, S.T_rel, S.bound, n, cfg->sweep_flags, S.dmin, S.arg, nullptr);
        mark();
        if (prof && rc == 0) {
            int64_t evald = 0, total = 0;
            sio_ctx_last_prune_stats(c, &evald, &total);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[ev.size() - 2], ev[ev.size() - 1]);
            fprintf(stderr, "[sio lockstep] sweep: %lld problems, %.2f ms, %lld of %lld sub-tiles evaluated (%.2f%%), %.0f G point-queries/s resolved\n",
                    (long long)n, ms, (long long)evald, (long long)total, total ? 100.0 * evald / total : 0.0, n * (double)N / ms * 1e-6);
        }
        return rc;
    };
    std::vector<size_t> qp_marks;
    const int qp_max_iter = getenv("SIO_QP_MAXIT") ? atoi(getenv("SIO_QP_MAXIT")) : 4000;   // diagnostic only: OSQP's limit is 4000
    // Deferral of QP stragglers (SIO_LS_DEFER = ADMM iteration budget of the in-line launch, 0 = off): see k_sip_qp
    const int defer_cap = getenv("SIO_LS_DEFER") ? atoi(getenv("SIO_LS_DEFER")) : 800;   // measured best of 150 ... 1600 at 8,192 and 65,536 problems
    const bool defer = defer_cap > 0 && defer_cap < qp_max_iter;
    const bool qp_beside = !(getenv("SIO_QP_BESIDE") && atoi(getenv("SIO_QP_BESIDE")) == 0);   // 0: the two-row launch after the one-row launch
    if (defer) {
        for (int k = 0; k < sio_ctx::kLsBack; ++k) {
            if (!c->ls_stream[k]) {                        // lowest priority: the main stream's CTAs go first when an SM frees up
                int lo_prio = 0, hi_prio = 0;
                CU(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
                CU(cudaStreamCreateWithPriority(&c->ls_stream[k], cudaStreamNonBlocking, lo_prio));
            }
            CU(c->s_qpb[k].ensure(sip_qp_scratch_bytes(c->num_sms)));
        }
        if (!c->qp_side) {
            CU(cudaStreamCreateWithFlags(&c->qp_side, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&c->qp_fork, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&c->qp_join, cudaEventDisableTiming));
        }
        for (int k = 0; k < sio_ctx::kLsRing; ++k) {
            if (!c->ls_evA[k]) CU(cudaEventCreateWithFlags(&c->ls_evA[k], cudaEventDisableTiming));
            if (!c->ls_evB[k]) CU(cudaEventCreateWithFlags(&c->ls_evB[k], cudaEventDisableTiming));
        }
    }
    // SIO_LS_INJECT_QPFAIL=<problem>: fault injection for the tests of the "QP without a solution" path (status 2)
    const int64_t inject_fail = getenv("SIO_LS_INJECT_QPFAIL") ? atoll(getenv("SIO_LS_INJECT_QPFAIL")) : -1;
    const int64_t max_rounds = (int64_t)cfg->max_iters * 64 + 64;    // a problem sits a few rounds out per deferred QP
    int64_t round = 0, rounds_with_qp = 0;
    for (;; ++round) {
        if (round >= max_rounds) return fail(SIO_ERR_INVALID, "sio_lockstep_pose_solve: no progress after %lld rounds", (long long)round);
        span_begin();
        CU(cudaMemsetAsync(S.counters, 0, 2 * sizeof(unsigned long long), st));
        CU(cudaMemsetAsync(S.counters + 4, 0, sizeof(unsigned long long), st));
        ls_launch_begin(S, L, st);
        ls_launch_values(S, 0, 1, 1, st);
        ls_launch_oracle_pre(S, L, st);                    // only problems whose iteration refreshes the oracle at x queue up
        span_end(kPhPre);
        CU(read_counters());
        if (hc[1] == 0) break;                             // no live problem left
        if (hc[4] == 0) {                                  // every live problem is waiting for a background QP: wait with them
            if (!defer) return fail(SIO_ERR_INVALID, "sio_lockstep_pose_solve: live problems but none taking part");
            for (int k = 0; k < sio_ctx::kLsBack; ++k) CU(cudaStreamSynchronize(c->ls_stream[k]));
            continue;
        }
        {
            const int64_t n = (int64_t)hc[0];
            if (n > 0) {
                int rc = sweep(n); if (rc) return rc;
                span_begin();
                ls_launch_oracle_post(S, L, n, st);
                ls_launch_values(S, 0, 1, 0, st);          // rows only: the appended depths are the sweeps' exact minima
                span_end(kPhPost);
            }
        }
        span_begin();
        ls_launch_gx(S, L, st);
        span_end(kPhGx);
        qp_marks.push_back(ev.size());
        if (prof) CU(cudaMemsetAsync(S.qstatus, 0, (size_t)B * 4, st));   // the kernel skips problems that are not taking part
        const int slot = (int)(rounds_with_qp % sio_ctx::kLsRing);
        QpDefer dfA;
        if (defer) {
            if (rounds_with_qp >= sio_ctx::kLsRing) CU(cudaStreamWaitEvent(st, c->ls_evB[slot], 0));   // the slot's previous list is consumed
            CU(cudaMemsetAsync(dcount + slot, 0, sizeof(uint32_t), st));
            dfA.list = dlist[slot]; dfA.count = dcount + slot; dfA.done = S.qdone; dfA.mark = S.qdef; dfA.busy = S.ph;
            dfA.wide = wlist; dfA.wide_n = wcount;
            // side by side only while the round is latency-bound (fewer problems than ~4 per resident one-row warp): there
            // each launch lasts as long as its slowest problem and the two overlap; a throughput-bound round (65,536
            // problems: 17 k still live at the end) is served best by each build's full grid, one after the other
            // (measured: 8,192 problems 58.9 -> 56.0 ms, 65,536 problems 237 -> 247 ms when forced everywhere)
            if (qp_beside && (int64_t)hc[4] <= 4 * 16 * (int64_t)c->num_sms) { dfA.side = c->qp_side; dfA.fork = c->qp_fork; dfA.join = c->qp_join; }
        }
        mark();
        if (launch_sip_qp(6, B, S.qW, S.dxdes, S.qlo, S.qhi, nullptr, S.qk, cap, S.rows, S.D, L.reg, L.inflation, 1e-5,
                          defer ? defer_cap : qp_max_iter, S.dx, S.qstatus, c->s_qp.p, c->num_sms, st, defer ? &dfA : nullptr))
            return fail(SIO_ERR_INVALID, "sio_lockstep_pose_solve: QP launch failed");
        mark();
        if (defer) {                                       // the listed problems, in full, behind the main stream's back
            const int k = (int)(rounds_with_qp % sio_ctx::kLsBack);
            CU(cudaEventRecord(c->ls_evA[slot], st));
            CU(cudaStreamWaitEvent(c->ls_stream[k], c->ls_evA[slot], 0));
            QpDefer dfB;
            dfB.only = dlist[slot]; dfB.only_n = dcount + slot; dfB.done = S.qdone;
            if (launch_sip_qp(6, B, S.qW, S.dxdes, S.qlo, S.qhi, nullptr, S.qk, cap, S.rows, S.D, L.reg, L.inflation, 1e-5, qp_max_iter, S.dx,
                              S.qstatus, c->s_qpb[k].p, c->num_sms, c->ls_stream[k], &dfB))
                return fail(SIO_ERR_INVALID, "sio_lockstep_pose_solve: QP launch failed");
            CU(cudaEventRecord(c->ls_evB[slot], c->ls_stream[k]));
            c->launches += 1;
        }
        if (inject_fail >= 0 && inject_fail < B && round == 0) {   // test aid: pretend problem p's QP had no solution
            const int32_t code = 3;
            CU(cudaMemcpyAsync(S.qstatus + inject_fail, &code, sizeof(code), cudaMemcpyHostToDevice, st));
            CU(cudaStreamSynchronize(st));
        }
        ++rounds_with_qp;
        if (prof) {                                        // per-launch ADMM iteration statistics
            std::vector<int32_t> hs((size_t)B), hk((size_t)B);
            CU(cudaMemcpyAsync(hs.data(), S.qstatus, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(hk.data(), S.qk, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            long long tot = 0, mx = 0, nz = 0, lim = 0, deferred = 0, big = 0;
            for (int64_t p = 0; p < B; ++p) {
                if (hk[p] < 0) continue;
                const long long n = hs[p] >> 8;
                tot += n; mx = std::max(mx, n); ++nz; lim += n >= 4000; deferred += (hs[p] & 0xff) == 4;
                big += hk[p] + 6 > 32;
            }
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[ev.size() - 2], ev[ev.size() - 1]);
            fprintf(stderr, "[sio lockstep] round %lld: qp %.2f ms, %lld problems (%lld with 2 rows/lane), %lld ADMM its (max %lld, %lld at limit), %lld deferred\n",
                    (long long)round, ms, nz, big, tot, mx, lim, deferred);
        }
        span_begin();
        ls_launch_step(S, L, st);
        ls_launch_values(S, 1, 0, 1, st);
        CU(cudaMemsetAsync(S.counters, 0, sizeof(unsigned long long), st));
        ls_launch_score_pre(S, st);
        span_end(kPhStep);
        CU(read_counters());
        const int64_t n = (int64_t)hc[0];
        if (n > 0) { int rc = sweep(n); if (rc) return rc; }
        span_begin();
        ls_launch_score_post(S, L, n, st);
        span_end(kPhScore);
        c->launches += 8;
        CU(cudaGetLastError());
    }
    if (defer)
        for (int k = 0; k < sio_ctx::kLsBack; ++k) CU(cudaStreamSynchronize(c->ls_stream[k]));
    CU(read_counters());
    const auto loop1 = std::chrono::steady_clock::now();
    // results.  Everything comes back through ONE pinned block kept by the context, and the instantiated parameters
    // PACKED (a device gather of the cnt[p] points per problem: a quarter of the B * cap * 3 dense array on BASELINE
    // configs[4]); worker threads then spread the rows over the caller's arrays.  Copying the dense array straight into
    // the caller's freshly allocated memory took 24 ms of a 260 ms solve -- page faults, mostly.
    {
        const size_t oR = 0, ot = oR + (size_t)B * 9 * 8, ofx = ot + (size_t)B * 3 * 8, ogx = ofx + (size_t)B * 8, og0 = ogx + (size_t)B * 8;
        const size_t ost = og0 + (size_t)B * 8, oit = ost + (size_t)B * 4, ocn = oit + (size_t)B * 4, small = ocn + (size_t)B * 4;
        CU(c->h_ls_out.ensure(small));
        char* hb = c->h_ls_out.as<char>();
        CU(cudaMemcpyAsync(hb + oR, S.R, (size_t)B * 9 * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(hb + ot, S.t, (size_t)B * 3 * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(hb + ofx, S.fx, (size_t)B * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(hb + ogx, S.gx, (size_t)B * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(hb + og0, S.gx0, (size_t)B * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(hb + ost, S.status, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(hb + oit, S.niter, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(hb + ocn, S.cnt, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
        CU(spin_stream(st));
        const double* hR = reinterpret_cast<const double*>(hb + oR);
        const double* ht = reinterpret_cast<const double*>(hb + ot);
        const int32_t* hcnt = reinterpret_cast<const int32_t*>(hb + ocn);
        std::vector<int64_t> offs;
        int64_t total = 0;
        const double* hpk = nullptr;
        if (params) {
            offs.resize((size_t)B + 1);
            for (int64_t p = 0; p < B; ++p) { offs[p] = total; total += std::max(0, std::min<int32_t>(hcnt[p], cap)); }
            offs[B] = total;
            // sized for full rows: a second solve of the same batch never re-allocates (pinned allocations cost milliseconds)
            CU(c->s_ls_pack.ensure((size_t)B * 8 + (size_t)B * cap * 24));
            CU(c->h_ls_pack.ensure((size_t)B * cap * 24));
            int64_t* d_offs = c->s_ls_pack.as<int64_t>();
            double* d_pk = reinterpret_cast<double*>(c->s_ls_pack.as<char>() + (size_t)B * 8);
            CU(cudaMemcpyAsync(d_offs, offs.data(), (size_t)B * 8, cudaMemcpyHostToDevice, st));
            ls_launch_pack_params(S, d_offs, d_pk, st);
            CU(cudaMemcpyAsync(c->h_ls_pack.p, d_pk, (size_t)total * 24, cudaMemcpyDeviceToHost, st));
            hpk = c->h_ls_pack.as<double>();
        }
        auto spread = [&](int64_t p0, int64_t p1, bool with_params) {
            for (int64_t p = p0; p < p1; ++p) {
                for (int i = 0; i < 3; ++i) {
                    for (int j = 0; j < 3; ++j) T_out[12 * p + 4 * i + j] = hR[9 * p + 3 * i + j];
                    T_out[12 * p + 4 * i + 3] = ht[3 * p + i];
                }
                if (with_params) std::memcpy(params + p * (int64_t)cap * 3, hpk + 3 * offs[p], (size_t)(offs[p + 1] - offs[p]) * 24);
            }
        };
        // the small arrays while the packed parameters are still in flight
        if (fx) std::memcpy(fx, hb + ofx, (size_t)B * 8);
        if (gx) std::memcpy(gx, hb + ogx, (size_t)B * 8);
        if (gx0) std::memcpy(gx0, hb + og0, (size_t)B * 8);
        if (status) std::memcpy(status, hb + ost, (size_t)B * 4);
        if (iterations) std::memcpy(iterations, hb + oit, (size_t)B * 4);
        if (n_params) std::memcpy(n_params, hb + ocn, (size_t)B * 4);
        if (params) CU(spin_stream(st));
        const int nthr = (int)std::max<int64_t>(1, std::min<int64_t>({8, (int64_t)std::thread::hardware_concurrency(), B / 2048}));
        if (nthr <= 1) {
            spread(0, B, params != nullptr);
        } else {
            std::vector<std::thread> pool;
            for (int k = 0; k < nthr; ++k) pool.emplace_back(spread, B * k / nthr, B * (k + 1) / nthr, params != nullptr);
            for (auto& th : pool) th.join();
        }
    }
    double qp_ms = 0.0, sweep_ms = 0.0;
    {
        std::vector<char> is_qp(ev.size(), 0);
        for (size_t k : qp_marks) is_qp[k] = 1;
        for (size_t k = 0; k + 1 < ev.size(); k += 2) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[k], ev[k + 1]);
            (is_qp[k] ? qp_ms : sweep_ms) += ms;
        }
    }
    if (prof) {
        double ph[kPhCount] = {0, 0, 0, 0, 0};
        for (auto& sp : spans) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, sp.second.first, sp.second.second);
            ph[sp.first] += ms;
        }
        const double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
        fprintf(stderr, "[sio lockstep] loop+readback wall %.1f ms: qp %.1f, sweeps %.1f, begin/values/oracle_pre %.1f, oracle_post/rows %.1f, gx %.1f, "
                        "step/values/score_pre %.1f, score_post %.1f\n", wall, qp_ms, sweep_ms, ph[kPhPre], ph[kPhPost], ph[kPhGx], ph[kPhStep], ph[kPhScore]);
    }
    if (prof_total)
        fprintf(stderr, "[sio lockstep] call %.1f ms (B = %lld): set-up %.1f, loop %.1f, read-back %.1f\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - entry0).count(), (long long)B,
                std::chrono::duration<double, std::milli>(wall0 - entry0).count(),
                std::chrono::duration<double, std::milli>(loop1 - wall0).count(),
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - loop1).count());
    if (stats) {
        stats[0] = (int64_t)hc[2]; stats[1] = queries; stats[2] = (int64_t)(qp_ms * 1e3); stats[3] = (int64_t)hc[3];
        stats[4] = (int64_t)(sweep_ms * 1e3); stats[5] = (int64_t)hc[5];
    }
    return SIO_OK;
}

// ---- mesh -> distance grid ----------------------------------------------------------------------------------
int sio_mesh_distance_grid(sio_ctx* c, const double* verts, int32_t nv, const int32_t* tris, int32_t nt,
                           const double bmin[3], const double h[3], const int32_t dims[3], double* out_dist) {
    if (!c || !verts || !tris || !bmin || !h || !dims || !out_dist || nv <= 0 || nt <= 0)
        return fail(SIO_ERR_INVALID, "sio_mesh_distance_grid: bad argument");
    const int64_t ncell = (int64_t)dims[0] * dims[1] * dims[2];
    if (dims[0] < 1 || dims[1] < 1 || dims[2] < 1 || ncell >= (1LL << 31)) return fail(SIO_ERR_INVALID, "sio_mesh_distance_grid: bad dims");
    std::vector<double> tri((size_t)nt * 9);
    for (int t = 0; t < nt; ++t)
        for (int v = 0; v < 3; ++v) {
            const int32_t idx = tris[3 * t + v];
            if (idx < 0 || idx >= nv) return fail(SIO_ERR_INVALID, "sio_mesh_distance_grid: triangle %d refers to vertex %d", t, idx);
            for (int a = 0; a < 3; ++a) tri[(size_t)t * 9 + v * 3 + a] = verts[3 * (size_t)idx + a];
        }
    DeviceGuard guard(c->device);
    cudaStream_t st = c->stream;
    CU(c->s_in0.ensure(tri.size() * sizeof(double)));
    CU(c->s_out0.ensure((size_t)ncell * sizeof(double)));
    CU(cudaMemcpyAsync(c->s_in0.p, tri.data(), tri.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    launch_mesh_distance(c->s_in0.as<double>(), nt, bmin, h, dims, c->s_out0.as<double>(), st);
    c->launches += 1;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out_dist, c->s_out0.p, (size_t)ncell * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SIO_OK;
}

// mesh -> SIGNED distance grid, distances and sides on the device
int sio_mesh_sdf_grid(sio_ctx* c, const double* verts, int32_t nv, const int32_t* tris, int32_t nt,
                      const double bmin[3], const double h[3], const int32_t dims[3], float* out_sdf) {
    if (!c || !verts || !tris || !bmin || !h || !dims || !out_sdf || nv <= 0 || nt <= 0)
        return fail(SIO_ERR_INVALID, "sio_mesh_sdf_grid: bad argument");
    const int64_t ncell = (int64_t)dims[0] * dims[1] * dims[2];
    if (dims[0] < 1 || dims[1] < 1 || dims[2] < 1 || ncell >= (1LL << 31)) return fail(SIO_ERR_INVALID, "sio_mesh_sdf_grid: bad dims");
    std::vector<double> tri((size_t)nt * 9);
    for (int t = 0; t < nt; ++t)
        for (int v = 0; v < 3; ++v) {
            const int32_t idx = tris[3 * t + v];
            if (idx < 0 || idx >= nv) return fail(SIO_ERR_INVALID, "sio_mesh_sdf_grid: triangle %d refers to vertex %d", t, idx);
            for (int a = 0; a < 3; ++a) tri[(size_t)t * 9 + v * 3 + a] = verts[3 * (size_t)idx + a];
        }
    DeviceGuard guard(c->device);
    cudaStream_t st = c->stream;
    CU(c->s_in0.ensure(tri.size() * sizeof(double)));
    CU(c->s_out0.ensure((size_t)ncell * sizeof(double)));
    CU(c->s_out1.ensure((size_t)ncell * sizeof(float)));
    CU(c->s_out2.ensure((size_t)ncell));
    CU(cudaMemcpyAsync(c->s_in0.p, tri.data(), tri.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    launch_mesh_distance(c->s_in0.as<double>(), nt, bmin, h, dims, c->s_out0.as<double>(), st);
    launch_sdf_sign(c->s_in0.as<double>(), nt, bmin, h, dims, c->s_out0.as<double>(), c->s_out2.as<int8_t>(), c->s_out1.as<float>(), st);
    c->launches += 3;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out_sdf, c->s_out1.p, (size_t)ncell * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SIO_OK;
}

// layout of the surface sampler's output (host, O(triangles)): the vertices first, then one block per triangle,
// triangles grouped by their (m1, m2) in lexicographic order, ascending triangle index inside a group -- the order the
// host builder emits.  Returns the number of points BEFORE duplicates are merged, or -1 (message set).
static int64_t surface_layout(const double* verts, int32_t nv, const int32_t* tris, int32_t nt, double res, std::vector<LatticeTriHost>* recs) {
    int64_t total = nv;
    if (!(res > 0 && std::isfinite(res) && nt > 0)) return total;
    struct Item { int32_t m1, m2, t; };
    std::vector<Item> items;
    std::vector<LatticeTriHost> tmp((size_t)nt);
    for (int t = 0; t < nt; ++t) {
        const double* P[3];
        for (int v = 0; v < 3; ++v) {
            const int32_t idx = tris[3 * t + v];
            if (idx < 0 || idx >= nv) { fail(SIO_ERR_INVALID, "sio_mesh_surface_sample: triangle %d refers to vertex %d", t, idx); return -1; }
            P[v] = verts + 3 * (size_t)idx;
        }
        auto dist = [](const double* x, const double* y) {
            const double d0 = x[0] - y[0], d1 = x[1] - y[1], d2 = x[2] - y[2];
            return std::sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        };
        const double e[3] = {dist(P[2], P[1]), dist(P[0], P[2]), dist(P[1], P[0])};       // edge opposite vertex v
        int apex = 0;
        if (e[1] > e[apex]) apex = 1;
        if (e[2] > e[apex]) apex = 2;
        const double *A = P[apex], *Bp = P[(apex + 1) % 3], *Cp = P[(apex + 2) % 3];
        LatticeTriHost r;
        for (int a = 0; a < 3; ++a) { r.a[a] = A[a]; r.b[a] = Bp[a]; r.c[a] = Cp[a]; }
        r.m1 = (int32_t)std::max(1.0, std::ceil(dist(Bp, A) / res - 1e-9));
        r.m2 = (int32_t)std::max(1.0, std::ceil(dist(Cp, A) / res - 1e-9));
        r.base = 0;
        tmp[t] = r;
        if (!(r.m1 == 1 && r.m2 == 1)) items.push_back({r.m1, r.m2, t});
    }
    std::stable_sort(items.begin(), items.end(), [](const Item& x, const Item& y) { return x.m1 != y.m1 ? x.m1 < y.m1 : x.m2 < y.m2; });
    if (recs) recs->reserve(items.size());
    for (const Item& it : items) {
        LatticeTriHost r = tmp[it.t];
        int64_t cnt = 0;
        for (int i = 0; i <= r.m1; ++i) {
            const double u = (double)i / (double)r.m1;
            for (int j = 0; j <= r.m2; ++j) {
                const double v = (double)j / (double)r.m2;
                if ((u + v) <= 1.0 + 1e-12 && !((i == 0 && j == 0) || (i == r.m1 && j == 0) || (i == 0 && j == r.m2))) ++cnt;
            }
        }
        r.base = total;
        total += cnt;
        if (recs) recs->push_back(r);
    }
    return total;
}

// upper bound of the sampler's output (host work only): what a caller sizes out_pts with
int64_t sio_mesh_surface_sample_bound(const double* verts, int32_t nv, const int32_t* tris, int32_t nt, double res) {
    if (!verts || nv <= 0 || nt < 0 || (nt > 0 && !tris)) { fail(SIO_ERR_INVALID, "sio_mesh_surface_sample_bound: bad argument"); return -1; }
    return surface_layout(verts, nv, tris, nt, res, nullptr);
}

// mesh -> surface point cloud at resolution `res`
int sio_mesh_surface_sample(sio_ctx* c, const double* verts, int32_t nv, const int32_t* tris, int32_t nt, double res,
                            double* out_pts, int64_t capacity, int64_t* n_out) {
    if (!c || !verts || !n_out || nv <= 0 || nt < 0 || (nt > 0 && !tris)) return fail(SIO_ERR_INVALID, "sio_mesh_surface_sample: bad argument");
    std::vector<LatticeTriHost> recs;
    const int64_t total = surface_layout(verts, nv, tris, nt, res, &recs);
    if (total < 0) return SIO_ERR_INVALID;
    if (total >= (1LL << 31)) return fail(SIO_ERR_INVALID, "sio_mesh_surface_sample: %lld lattice points: too many", (long long)total);
    DeviceGuard guard(c->device);
    cudaStream_t st = c->stream;
    uint64_t slots = 1024;
    while (slots < 2 * (uint64_t)total) slots <<= 1;
    const int64_t nb = (total + 1023) / 1024;
    CU(c->s_in0.ensure(std::max<size_t>(recs.size(), 1) * sizeof(LatticeTriHost)));
    CU(c->s_in1.ensure((size_t)total * 3 * sizeof(double)));          // all lattice points
    CU(c->s_out0.ensure((size_t)total * 3 * sizeof(double)));         // compacted
    CU(c->s_in2.ensure((size_t)slots * sizeof(uint32_t)));
    CU(c->s_out2.ensure((size_t)total));
    CU(c->s_in3.ensure((size_t)(nb + 1) * sizeof(uint32_t)));
    CU(cudaMemcpyAsync(c->s_in1.p, verts, (size_t)nv * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    if (!recs.empty()) {
        CU(cudaMemcpyAsync(c->s_in0.p, recs.data(), recs.size() * sizeof(LatticeTriHost), cudaMemcpyHostToDevice, st));
        launch_lattice_fill(c->s_in0.p, (int32_t)recs.size(), c->s_in1.as<double>(), st);
    }
    launch_dedup(c->s_in1.as<double>(), total, c->s_in2.as<uint32_t>(), slots, c->s_out2.as<uint8_t>(), c->s_in3.as<uint32_t>(),
                 c->s_out0.as<double>(), st);
    c->launches += 6 + (recs.empty() ? 0 : 1);
    CU(cudaGetLastError());
    uint32_t kept = 0;
    CU(cudaMemcpyAsync(&kept, c->s_in3.as<uint32_t>() + nb, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *n_out = (int64_t)kept;
    if (!out_pts) return SIO_OK;                                      // count only
    if ((int64_t)kept > capacity) return fail(SIO_ERR_OVERFLOW, "sio_mesh_surface_sample: %u points, capacity %lld", kept, (long long)capacity);
    CU(cudaMemcpy(out_pts, c->s_out0.p, (size_t)kept * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    return SIO_OK;
}

// ---- SIP-step QPs ----------------------------------------------------------------------------------------
int sio_sip_step_qp(sio_ctx* c, int32_t n, int64_t nprob, const double* W, const double* xdes, const double* xmin,
                    const double* xmax, const int64_t* offsets, const double* rows, const double* depths, double reg,
                    double inflation, double eps_abs, int32_t max_iter, double* dx, int32_t* status) {
    if (!c || nprob < 0 || (nprob > 0 && (!W || !xdes || !offsets || !dx || !status)) || ((xmin == nullptr) != (xmax == nullptr)))
        return fail(SIO_ERR_INVALID, "sio_sip_step_qp: bad argument");
    if (n != 6 && n != 7) return fail(SIO_ERR_INVALID, "sio_sip_step_qp: n must be 6 or 7, got %d", n);
    if (nprob == 0) return SIO_OK;
    const int64_t total = offsets[nprob];
    if (offsets[0] != 0 || total < 0 || (total > 0 && (!rows || !depths))) return fail(SIO_ERR_INVALID, "sio_sip_step_qp: bad offsets/rows");
    DeviceGuard guard(c->device);
    cudaStream_t st = c->stream;
    // one staging block: [W | xdes | xmin | xmax | depths | rows | offsets]
    const size_t v = (size_t)nprob * n * sizeof(double);
    const size_t oW = 0, oX = v, oLo = 2 * v, oHi = 3 * v, oD = 4 * v, oR = oD + (size_t)total * sizeof(double);
    const size_t oO = oR + (size_t)total * n * sizeof(double), in_bytes = oO + (size_t)(nprob + 1) * sizeof(int64_t);
    CU(c->h_a.ensure(in_bytes));
    CU(c->s_in0.ensure(in_bytes));
    char* h = c->h_a.as<char>();
    std::memcpy(h + oW, W, v);
    std::memcpy(h + oX, xdes, v);
    if (xmin) { std::memcpy(h + oLo, xmin, v); std::memcpy(h + oHi, xmax, v); }
    if (total > 0) { std::memcpy(h + oD, depths, (size_t)total * sizeof(double)); std::memcpy(h + oR, rows, (size_t)total * n * sizeof(double)); }
    std::memcpy(h + oO, offsets, (size_t)(nprob + 1) * sizeof(int64_t));
    CU(cudaMemcpyAsync(c->s_in0.p, h, in_bytes, cudaMemcpyHostToDevice, st));
    const size_t out_bytes = v + (size_t)nprob * sizeof(int32_t);
    CU(c->s_out0.ensure(out_bytes));
    CU(c->h_b.ensure(out_bytes));
    CU(c->s_qp.ensure(sip_qp_scratch_bytes(c->num_sms)));
    char* d = c->s_in0.as<char>();
    char* o = c->s_out0.as<char>();
    if (launch_sip_qp(n, nprob, reinterpret_cast<double*>(d + oW), reinterpret_cast<double*>(d + oX),
                      xmin ? reinterpret_cast<double*>(d + oLo) : nullptr, xmin ? reinterpret_cast<double*>(d + oHi) : nullptr,
                      reinterpret_cast<int64_t*>(d + oO), nullptr, 0, reinterpret_cast<double*>(d + oR), reinterpret_cast<double*>(d + oD), reg,
                      inflation, eps_abs, max_iter > 0 ? max_iter : 4000, reinterpret_cast<double*>(o),
                      reinterpret_cast<int32_t*>(o + v), c->s_qp.p, c->num_sms, st))
        return fail(SIO_ERR_INVALID, "sio_sip_step_qp: unsupported n");
    c->launches += 1;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->h_b.p, o, out_bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    std::memcpy(dx, c->h_b.p, v);
    std::memcpy(status, c->h_b.as<char>() + v, (size_t)nprob * sizeof(int32_t));
    c->qp_iters_total = 0; c->qp_iters_max = 0; c->qp_at_limit = 0;
    const int lim = max_iter > 0 ? max_iter : 4000;
    for (int64_t i = 0; i < nprob; ++i) {
        const int64_t its = status[i] >> 8;
        c->qp_iters_total += its;
        if (its > c->qp_iters_max) c->qp_iters_max = its;
        if (its >= lim) c->qp_at_limit += 1;
        status[i] &= 0xff;
    }
    return SIO_OK;
}

// ---- K4 ----------------------------------------------------------------------------------------------
int sio_robot_create(sio_ctx* c, int32_t n, const int32_t* parent, const int32_t* jtype, const double* tparent,
                     const double* axis, sio_handle* out) {
    if (!c || n <= 0 || !parent || !tparent || !axis || !out) return fail(SIO_ERR_INVALID, "sio_robot_create: bad argument");
    for (int i = 0; i < n; ++i)
        if (parent[i] >= i) return fail(SIO_ERR_INVALID, "sio_robot_create: parent[%d]=%d must precede its child", i, parent[i]);
    DeviceGuard guard(c->device);
    RobotH r;
    r.n = n;
    std::vector<int32_t> jt(n, 0);
    if (jtype) jt.assign(jtype, jtype + n);
    CU(cudaMalloc(&r.d_parent, n * sizeof(int32_t)));
    CU(cudaMalloc(&r.d_jtype, n * sizeof(int32_t)));
    CU(cudaMalloc(&r.d_tparent, (size_t)n * 12 * sizeof(double)));
    CU(cudaMalloc(&r.d_axis, (size_t)n * 3 * sizeof(double)));
    CU(cudaMemcpy(r.d_parent, parent, n * sizeof(int32_t), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(r.d_jtype, jt.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(r.d_tparent, tparent, (size_t)n * 12 * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(r.d_axis, axis, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice));
    r.live = true;
    c->robots.push_back(r);
    *out = (sio_handle)(c->robots.size() - 1);
    return SIO_OK;
}

int sio_robot_destroy(sio_ctx* c, sio_handle h) {
    if (!c || !robot_ok(c, h)) return fail(SIO_ERR_INVALID, "sio_robot_destroy: bad handle %d", h);
    DeviceGuard guard(c->device);
    cudaStreamSynchronize(c->stream);
    RobotH& r = c->robots[h];
    cudaFree(r.d_parent); cudaFree(r.d_jtype); cudaFree(r.d_tparent); cudaFree(r.d_axis);
    r = RobotH();
    return SIO_OK;
}

int sio_fk(sio_ctx* c, sio_handle robot, const double* q, int64_t B, double* T_out) {
    if (!c || !robot_ok(c, robot) || B < 0 || (B > 0 && (!q || !T_out))) return fail(SIO_ERR_INVALID, "sio_fk: bad argument");
    if (B == 0) return SIO_OK;
    DeviceGuard guard(c->device);
    cudaStream_t st = c->stream;
    const RobotH& r = c->robots[robot];
    CU(c->s_in0.ensure((size_t)B * r.n * sizeof(double)));
    CU(c->s_tlinks.ensure((size_t)B * r.n * 12 * sizeof(double)));
    CU(cudaMemcpyAsync(c->s_in0.p, q, (size_t)B * r.n * sizeof(double), cudaMemcpyHostToDevice, st));
    launch_fk(r.dev(), c->s_in0.as<double>(), B, c->s_tlinks.as<double>(), st);
    c->launches += 1;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(T_out, c->s_tlinks.p, (size_t)B * r.n * 12 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(wait_stream(st));
    return SIO_OK;
}

int sio_chain_rows(sio_ctx* c, sio_handle robot, const sio_handle* link_grids, const double* q, int64_t n_cfg,
                   const int32_t* cfg_of_pt, const int32_t* link_of_pt, const double* pts, int64_t P, int flags,
                   double* d, double* rows) {
    if (!c || !robot_ok(c, robot) || !link_grids || n_cfg <= 0 || !q || P < 0 || (P > 0 && (!link_of_pt || !pts || !d)))
        return fail(SIO_ERR_INVALID, "sio_chain_rows: bad argument");
    if (P == 0) return SIO_OK;
    DeviceGuard guard(c->device);
    cudaStream_t st = c->stream;
    const RobotH& r = c->robots[robot];
    // grid table over the links that have one; link_grid[l] = slot or -1
    std::vector<GridDev> tab;
    std::vector<int32_t> slot(r.n, -1);
    for (int l = 0; l < r.n; ++l)
        if (link_grids[l] >= 0) {
            if (!grid_ok(c, link_grids[l])) return fail(SIO_ERR_INVALID, "sio_chain_rows: bad grid handle for link %d", l);
            slot[l] = (int32_t)tab.size();
            tab.push_back(c->grids[link_grids[l]].meta);
        }
    for (int64_t i = 0; i < P; ++i) {
        int l = link_of_pt[i];
        if (l < 0 || l >= r.n || slot[l] < 0) return fail(SIO_ERR_INVALID, "sio_chain_rows: point %lld on link %d which has no grid", (long long)i, l);
        if (cfg_of_pt && (cfg_of_pt[i] < 0 || cfg_of_pt[i] >= n_cfg)) return fail(SIO_ERR_INVALID, "sio_chain_rows: cfg_of_pt[%lld] out of range", (long long)i);
    }
    c->staged_grids.clear();
    CU(c->s_gridtab.ensure(tab.size() * sizeof(GridDev)));
    CU(cudaMemcpyAsync(c->s_gridtab.p, tab.data(), tab.size() * sizeof(GridDev), cudaMemcpyHostToDevice, st));
    CU(c->s_gob.ensure((size_t)r.n * sizeof(int32_t)));
    CU(cudaMemcpyAsync(c->s_gob.p, slot.data(), (size_t)r.n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CU(c->s_in0.ensure((size_t)n_cfg * r.n * sizeof(double)));
    CU(cudaMemcpyAsync(c->s_in0.p, q, (size_t)n_cfg * r.n * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(c->s_tlinks.ensure((size_t)n_cfg * r.n * 12 * sizeof(double)));
    launch_fk(r.dev(), c->s_in0.as<double>(), n_cfg, c->s_tlinks.as<double>(), st);
    CU(c->s_in1.ensure((size_t)P * 3 * sizeof(double)));
    CU(cudaMemcpyAsync(c->s_in1.p, pts, (size_t)P * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(c->s_in2.ensure((size_t)P * sizeof(int32_t)));
    CU(cudaMemcpyAsync(c->s_in2.p, link_of_pt, (size_t)P * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    const int32_t* d_cfg = nullptr;
    if (cfg_of_pt) {
        CU(c->s_in3.ensure((size_t)P * sizeof(int32_t)));
        CU(cudaMemcpyAsync(c->s_in3.p, cfg_of_pt, (size_t)P * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        d_cfg = c->s_in3.as<int32_t>();
    }
    CU(c->s_out0.ensure((size_t)P * sizeof(double)));
    CU(c->s_out1.ensure((size_t)P * r.n * sizeof(double)));
    launch_chain_rows(r.dev(), c->s_gridtab.as<GridDev>(), c->s_gob.as<int32_t>(), c->s_tlinks.as<double>(), d_cfg,
                      c->s_in2.as<int32_t>(), c->s_in1.as<double>(), P, (flags & SIO_NO_NORMALIZE) ? 0 : 1,
                      c->s_out0.as<double>(), rows ? c->s_out1.as<double>() : nullptr, st);
    c->launches += 2;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(d, c->s_out0.p, (size_t)P * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (rows) CU(cudaMemcpyAsync(rows, c->s_out1.p, (size_t)P * r.n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(wait_stream(st));
    return SIO_OK;
}

int sio_robot_min_distance(sio_ctx* c, sio_handle robot, const sio_handle* link_grids, const int32_t* links, int32_t n_sel,
                           sio_handle cloud, const double T_cloud[12], const double* q, int64_t S, const double* bound,
                           int flags, double* dmin, int64_t* argmin, int32_t* n_under) {
    if (!c || !robot_ok(c, robot) || !link_grids || !links || n_sel <= 0 || !cloud_ok(c, cloud) || !T_cloud || S < 0 ||
        (S > 0 && (!q || !dmin || !argmin)))
        return fail(SIO_ERR_INVALID, "sio_robot_min_distance: bad argument");
    if (S == 0) { c->under.clear(); c->under_valid = true; return SIO_OK; }
    DeviceGuard guard(c->device);
    cudaStream_t st = c->stream;
    const RobotH& r = c->robots[robot];
    std::vector<sio_handle> sel((size_t)n_sel);
    for (int l = 0; l < n_sel; ++l) {
        if (links[l] < 0 || links[l] >= r.n) return fail(SIO_ERR_INVALID, "sio_robot_min_distance: link %d out of range", links[l]);
        sel[l] = link_grids[links[l]];
        if (!grid_ok(c, sel[l])) return fail(SIO_ERR_INVALID, "sio_robot_min_distance: link %d has no grid", links[l]);
    }
    const int64_t B = S * n_sel;
    CU(c->s_in0.ensure((size_t)S * r.n * sizeof(double)));
    CU(cudaMemcpyAsync(c->s_in0.p, q, (size_t)S * r.n * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(c->s_tlinks.ensure((size_t)S * r.n * 12 * sizeof(double)));
    launch_fk(r.dev(), c->s_in0.as<double>(), S, c->s_tlinks.as<double>(), st);
    CU(c->s_in2.ensure((size_t)n_sel * sizeof(int32_t) + 12 * sizeof(double) + 16));
    // [links | pad | T_cloud] in one staging buffer
    int32_t* d_links = c->s_in2.as<int32_t>();
    double* d_Tc = reinterpret_cast<double*>(reinterpret_cast<char*>(c->s_in2.p) + (((size_t)n_sel * sizeof(int32_t) + 15) / 16) * 16);
    CU(cudaMemcpyAsync(d_links, links, (size_t)n_sel * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_Tc, T_cloud, 12 * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(c->s_trel.ensure((size_t)B * 12 * sizeof(double)));
    CU(c->s_gob.ensure((size_t)B * sizeof(int32_t)));
    launch_link_rel(c->s_tlinks.as<double>(), r.n, d_links, n_sel, S, d_Tc, c->s_trel.as<double>(), c->s_gob.as<int32_t>(), st);
    c->launches += 2;
    const double* d_bound = nullptr;
    if (bound) {
        CU(c->s_in1.ensure((size_t)B * sizeof(double)));
        CU(cudaMemcpyAsync(c->s_in1.p, bound, (size_t)B * sizeof(double), cudaMemcpyHostToDevice, st));
        d_bound = c->s_in1.as<double>();
    }
    CU(c->s_out0.ensure((size_t)B * sizeof(double)));
    CU(c->s_out1.ensure((size_t)B * sizeof(int64_t)));
    CU(c->s_out2.ensure((size_t)B * sizeof(int32_t)));
    for (;;) {
        int rc = sio_min_distance_dev(c, sel.data(), n_sel, c->s_gob.as<int32_t>(), cloud, c->s_trel.as<double>(), d_bound, B, flags,
                                      c->s_out0.as<double>(), c->s_out1.as<int64_t>(), n_under ? c->s_out2.as<int32_t>() : nullptr);
        if (rc) return rc;
        CU(cudaMemcpyAsync(dmin, c->s_out0.p, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(argmin, c->s_out1.p, (size_t)B * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        if (n_under && bound) CU(cudaMemcpyAsync(n_under, c->s_out2.p, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        rc = under_settle(c, bound && n_under);
        if (rc < 0) return rc;
        if (rc == 0) break;
    }
    if (n_under && !bound) std::memset(n_under, 0, sizeof(int32_t) * (size_t)B);
    return SIO_OK;
}

// ---- measurement helpers --------------------------------------------------------------------------------
int sio_ctx_last_qp_iters(sio_ctx* c, int64_t* iters, int64_t* max_single, int64_t* n_at_limit) {
    if (!c || !iters) return fail(SIO_ERR_INVALID, "sio_ctx_last_qp_iters: bad argument");
    *iters = c->qp_iters_total;
    if (max_single) *max_single = c->qp_iters_max;
    if (n_at_limit) *n_at_limit = c->qp_at_limit;
    return SIO_OK;
}

int sio_ctx_last_prune_stats(sio_ctx* c, int64_t* evaluated, int64_t* total) {
    if (!c || !evaluated || !total) return fail(SIO_ERR_INVALID, "sio_ctx_last_prune_stats: bad argument");
    DeviceGuard guard(c->device);
    *evaluated = 0; *total = c->prune_total;
    if (c->n_icnt > 0) {
        CU(cudaMemcpyAsync(c->h_icnt.p, c->s_icnt.p, sizeof(uint32_t) * c->n_icnt, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        for (int i = 0; i < c->n_icnt; ++i) *evaluated += c->h_icnt.as<uint32_t>()[i];
    }
    return SIO_OK;
}

int sio_robot_pairs_min_distance(sio_ctx* c, sio_handle robot, const sio_handle* link_grids, sio_handle cloud,
                                 const double T_cloud[12], const double* q, const int32_t* link_of_pair, int64_t P,
                                 const double* bound, int flags, double* dmin, int64_t* argmin, int32_t* n_under) {
    if (!c || !robot_ok(c, robot) || !link_grids || !cloud_ok(c, cloud) || !T_cloud || P < 0 ||
        (P > 0 && (!q || !link_of_pair || !dmin || !argmin)))
        return fail(SIO_ERR_INVALID, "sio_robot_pairs_min_distance: bad argument");
    if (P == 0) { c->under.clear(); c->under_valid = true; return SIO_OK; }
    DeviceGuard guard(c->device);
    cudaStream_t st = c->stream;
    const RobotH& r = c->robots[robot];
    // the grids of the links that occur, in link order; slot_of_link maps a link to its position in that list
    std::vector<int32_t> slot((size_t)r.n, -1);
    for (int64_t p = 0; p < P; ++p) {
        const int l = link_of_pair[p];
        if (l < 0 || l >= r.n) return fail(SIO_ERR_INVALID, "sio_robot_pairs_min_distance: link %d out of range", l);
        if (!grid_ok(c, link_grids[l])) return fail(SIO_ERR_INVALID, "sio_robot_pairs_min_distance: link %d has no grid", l);
        slot[l] = 0;
    }
    std::vector<sio_handle> sel;
    for (int l = 0; l < r.n; ++l) if (slot[l] == 0) { slot[l] = (int32_t)sel.size(); sel.push_back(link_grids[l]); }
    // staging: [q | link_of_pair | slot_of_link | T_cloud | bound]
    const size_t oQ = 0, oL = (size_t)P * r.n * sizeof(double), oS = oL + (((size_t)P * 4 + 15) / 16) * 16;
    const size_t oT = oS + (((size_t)r.n * 4 + 15) / 16) * 16, oB = oT + 12 * sizeof(double), in_bytes = oB + (bound ? (size_t)P * 8 : 0);
    CU(c->h_a.ensure(in_bytes));
    CU(c->s_in0.ensure(in_bytes));
    char* h = c->h_a.as<char>();
    std::memcpy(h + oQ, q, oL);
    std::memcpy(h + oL, link_of_pair, (size_t)P * 4);
    std::memcpy(h + oS, slot.data(), (size_t)r.n * 4);
    std::memcpy(h + oT, T_cloud, 12 * sizeof(double));
    if (bound) std::memcpy(h + oB, bound, (size_t)P * 8);
    CU(cudaMemcpyAsync(c->s_in0.p, h, in_bytes, cudaMemcpyHostToDevice, st));
    char* d = c->s_in0.as<char>();
    CU(c->s_tlinks.ensure((size_t)P * r.n * 12 * sizeof(double)));
    launch_fk(r.dev(), reinterpret_cast<double*>(d + oQ), P, c->s_tlinks.as<double>(), st);
    CU(c->s_trel.ensure((size_t)P * 12 * sizeof(double)));
    CU(c->s_gob.ensure((size_t)P * sizeof(int32_t)));
    launch_pair_rel(c->s_tlinks.as<double>(), r.n, reinterpret_cast<int32_t*>(d + oL), reinterpret_cast<int32_t*>(d + oS), P,
                    reinterpret_cast<double*>(d + oT), c->s_trel.as<double>(), c->s_gob.as<int32_t>(), st);
    c->launches += 2;
    const size_t oD = 0, oA = (size_t)P * 8, oN = oA + (size_t)P * 8, out_bytes = oN + (size_t)P * 4;
    CU(c->s_out0.ensure(out_bytes));
    CU(c->h_b.ensure(out_bytes));
    char* o = c->s_out0.as<char>();
    for (;;) {
        int rc = sio_min_distance_dev(c, sel.data(), (int32_t)sel.size(), c->s_gob.as<int32_t>(), cloud, c->s_trel.as<double>(),
                                      bound ? reinterpret_cast<double*>(d + oB) : nullptr, P, flags, reinterpret_cast<double*>(o + oD),
                                      reinterpret_cast<int64_t*>(o + oA), n_under ? reinterpret_cast<int32_t*>(o + oN) : nullptr);
        if (rc) return rc;
        CU(cudaMemcpyAsync(c->h_b.p, o, out_bytes, cudaMemcpyDeviceToHost, st));
        rc = under_settle(c, bound && n_under);
        if (rc < 0) return rc;
        if (rc == 0) break;
    }
    const char* ho = c->h_b.as<char>();
    std::memcpy(dmin, ho + oD, (size_t)P * 8);
    std::memcpy(argmin, ho + oA, (size_t)P * 8);
    if (n_under) { if (bound) std::memcpy(n_under, ho + oN, (size_t)P * 4); else std::memset(n_under, 0, (size_t)P * 4); }
    return SIO_OK;
}

int sio_ctx_last_bulk_ms(sio_ctx* c, double* ms) {
    if (!c || !ms) return fail(SIO_ERR_INVALID, "sio_ctx_last_bulk_ms: bad argument");
    if (!c->bulk_timed) return fail(SIO_ERR_INVALID, "sio_ctx_last_bulk_ms: the last call had no single-batch fp32 bulk launch");
    DeviceGuard guard(c->device);
    CU(cudaEventSynchronize(c->ev_bulk1));
    float f = 0;
    CU(cudaEventElapsedTime(&f, c->ev_bulk0, c->ev_bulk1));
    *ms = (double)f;
    return SIO_OK;
}

int sio_bench_gather(sio_ctx* c, int64_t n_elems, int64_t n_gathers, int iters, double* gbps) {
    if (!c || n_elems <= 0 || n_gathers <= 0 || iters <= 0 || !gbps) return fail(SIO_ERR_INVALID, "sio_bench_gather: bad argument");
    DeviceGuard guard(c->device);
    cudaStream_t st = c->stream;
    float* arr = nullptr; uint32_t* idx = nullptr; float* sink = nullptr;
    CU(cudaMalloc(&arr, (size_t)n_elems * sizeof(float)));
    CU(cudaMalloc(&idx, (size_t)n_gathers * sizeof(uint32_t)));
    CU(cudaMalloc(&sink, sizeof(float)));
    CU(cudaMemsetAsync(arr, 0, (size_t)n_elems * sizeof(float), st));
    std::vector<uint32_t> h((size_t)n_gathers);
    uint64_t s = 0x9E3779B97F4A7C15ULL;
    for (auto& v : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = (uint32_t)(s % (uint64_t)n_elems); }
    CU(cudaMemcpy(idx, h.data(), h.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) launch_gather_bench(arr, n_elems, idx, n_gathers, sink, st);
    CU(cudaEventRecord(e0, st));
    for (int it = 0; it < iters; ++it) launch_gather_bench(arr, n_elems, idx, n_gathers, sink, st);
    CU(cudaEventRecord(e1, st));
    CU(cudaEventSynchronize(e1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    c->launches += iters + 3;
    *gbps = (double)n_gathers * 4.0 * iters / (ms * 1e-3) / 1e9;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(arr); cudaFree(idx); cudaFree(sink);
    return SIO_OK;
}

}  // extern "C"
