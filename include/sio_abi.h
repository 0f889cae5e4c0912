/* sio_abi.h -- C-ABI of the B200 collision oracle (libsio_b200.so).
 *
 * This is the drop-in boundary for the collision-oracle hot path of
 * krishauser/SemiInfiniteOptimization.  The reference reaches that path through Klampt's SWIG
 * bindings; each entry point below names the reference call site(s) (file:line under
 * /root/reference) whose work it replaces.  The Python classes that keep the reference's names
 * (`PenetrationDepthGeometry`, `ObjectCollisionConstraint`, ... in
 * semiinfiniteoptimization_b200/geometryopt.py) bind these symbols with ctypes; INTEGRATION.md
 * shows the stub.
 *
 * Conventions
 *   - plain pointers and sizes only; no exceptions cross the boundary.  Every function returns
 *     SIO_OK (0) or a negative status; sio_last_error() gives the message (thread-local).
 *   - rigid transforms are row-major 3x4 double[12] = [R | t]  (Klampt's se3 (R,t) has R
 *     column-major; the Python shim converts).
 *   - grid samples: float32, values at CELL CENTRES, flat order (i*ny + j)*nz + k (z fastest) --
 *     the order geometryopt.py:25-28 reads a Klampt VolumeGrid in.
 *   - cloud points: float32 xyz[N*3].  Indices returned are indices into THAT array (the
 *     library's internal Morton reordering is invisible).
 *   - precision: SIO_F32 = float32 arithmetic for the bulk pass + float64 re-evaluation of every
 *     near-tie, so minima / arg-minima / under-bound sets equal the float64 oracle's;
 *     SIO_F64 = float64 arithmetic everywhere (validation mode).
 *   - "host" entry points take host pointers and copy in/out on the context's stream; "_dev"
 *     entry points take device pointers, enqueue on the context's stream and do not synchronise.
 */
#ifndef SIO_ABI_H
#define SIO_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIO_ABI_VERSION 2

/* status codes */
#define SIO_OK               0
#define SIO_ERR_INVALID     -1   /* bad argument / handle                       */
#define SIO_ERR_CUDA        -2   /* CUDA runtime error (message has the detail) */
#define SIO_ERR_NOMEM       -3
#define SIO_ERR_NODEVICE    -4   /* no usable CUDA device: the product has no CPU fallback */
#define SIO_ERR_OVERFLOW    -5   /* an output capacity was too small (see sio_under_fetch, sio_min_distance_dev) */

/* flags */
#define SIO_F32             0x0
#define SIO_F64             0x1
#define SIO_NO_NORMALIZE    0x2   /* return the raw analytic gradient instead of the unit one */
#define SIO_WANT_GRAD       0x4
#define SIO_PRUNE           0x8   /* K2: skip 128-point sub-tiles whose Lipschitz lower bound rules them out (results identical) */
#define SIO_PUSH_BOARDS     0x20  /* K2 (_dev): after the call's kernels, store its results into every rank's result board (sio_ctx_set_boards) */
#define SIO_SMEM_GRID       0x10  /* K2, one grid of <= 200 KB of samples: bulk pass from a TMA-staged shared-memory copy of the grid
                                     (a measured alternative, slower than the default brick kernel; minimum / arg-minimum only) */

typedef struct sio_ctx sio_ctx;
typedef int32_t sio_handle;

int         sio_abi_version(void);
const char* sio_last_error(void);

/* ---- context: one per device, owns a stream, scratch arenas and all handles --------------- */
int   sio_ctx_create(int device, sio_ctx** out);
int   sio_ctx_destroy(sio_ctx* ctx);
int   sio_ctx_sync(sio_ctx* ctx);
void* sio_ctx_stream(sio_ctx* ctx);                 /* the cudaStream_t kernels are launched on */
int   sio_ctx_set_tau(sio_ctx* ctx, double tau);    /* near-tie window (metres) of the f64 re-check; <=0 -> automatic */
/* bytes of per-(transform, sub-tile) scratch a K2 call may hold at once; larger sweeps run in batches of transforms
 * (default 256 MiB; <= 0 restores it).  Results do not depend on it. */
int   sio_ctx_set_scratch_limit(sio_ctx* ctx, int64_t bytes);
/* under-bound candidate records a K2 call can hold (default 2^20, 16 B each; <= 0 restores it).  The host entry points
 * (sio_min_distance, sio_robot_*_min_distance) do not need it: they size the buffer from the sweep's own count and run
 * the sweep again, so their n_under and under-bound lists are always exact (or the call fails with SIO_ERR_NOMEM /
 * SIO_ERR_OVERFLOW above 2^31-1 records).  Callers of sio_min_distance_dev size it themselves. */
int   sio_ctx_set_under_capacity(sio_ctx* ctx, int64_t entries);
/* host-pointer K2 calls that repeat the previous call's signature are replayed as one CUDA graph (default on; results
 * do not depend on it; sio_ctx_last_bulk_ms is not available for a replayed call) */
int   sio_ctx_set_host_graphs(sio_ctx* ctx, int on);
int   sio_ctx_num_sms(sio_ctx* ctx);
int64_t sio_ctx_kernel_launches(sio_ctx* ctx);      /* kernels launched by this context so far */

/* ---- result boards: the multi-GPU exchange of the compact K2 results as peer stores over NVLink --------
 * The path has no data-path collective; what crosses the links are the per-transform records (20 bytes each).  Instead
 * of a collective launch per reporting interval, every rank owns a BOARD of nranks slots; after sio_ctx_set_boards,
 * the tail of every sio_min_distance_dev call made with SIO_PUSH_BOARDS stores this rank's records [dmin f64 x B | argmin i64 x B | n_under i32 x B]
 * into slot `rank` of EVERY rank's board (the peers' boards are mapped here through CUDA IPC), with a system-scope
 * fence.  Once the ranks' streams have drained and the ranks have met at a host barrier, every board holds every
 * rank's latest results; sio_board_read copies the local one out.  One process per GPU:
 *     sio_board_create (local board + its 64-byte IPC handle)  ->  exchange the handles (torch.distributed)
 *     ->  sio_board_open per peer  ->  sio_ctx_set_boards(rank, nranks, pointers, slot_bytes). */
int sio_board_create(sio_ctx* ctx, int64_t bytes, void** dptr, unsigned char ipc_handle[64]);
int sio_board_open(sio_ctx* ctx, const unsigned char ipc_handle[64], void** dptr);
int sio_board_close(sio_ctx* ctx, void* dptr, int is_local);
int sio_ctx_set_boards(sio_ctx* ctx, int32_t rank, int32_t nranks, void* const* boards, int64_t slot_bytes);   /* nranks <= 0: off */
int sio_board_read(sio_ctx* ctx, void* host_out, int64_t bytes);

/* ---- resident geometry ----------------------------------------------------------------------
 * replaces: Geometry3D.convert(...) results being handed to Klampt's collision data structures,
 * geometryopt.py:58-66 (PenetrationDepthGeometry.__init__).  Built once, immutable. */
int sio_grid_create(sio_ctx* ctx, const float* values, const int32_t dims[3],
                    const double bmin[3], const double bmax[3], sio_handle* out);
int sio_grid_destroy(sio_ctx* ctx, sio_handle grid);
int sio_cloud_create(sio_ctx* ctx, const float* xyz, int64_t n, sio_handle* out);
/* the order sio_cloud_create keeps the points in (host work only, no device or context involved): a k-d ordering into
 * leaves of 128 consecutive points -- the unit one warp evaluates and one bounding sphere culls --, Morton order inside
 * a leaf.  order_out[internal] = caller index; sio_cloud_perm returns the same for a resident cloud. */
int sio_cloud_order(const float* xyz, int64_t n, int32_t* order_out);
int sio_cloud_destroy(sio_ctx* ctx, sio_handle cloud);
int64_t sio_cloud_size(sio_ctx* ctx, sio_handle cloud);

/* ---- K1: point queries ----------------------------------------------------------------------
 * replaces: grid.distance_point(pt[,settings]) -- geometryopt.py:104-108 (PDG.distance(point)),
 * :139, :153 (PDG.normal).  T_grid_from_in maps the frame the points are given in (normally
 * world) to grid-local, i.e. it is the INVERSE of the pose set by setTransform (gopt:85-88).
 * d[P]; grad[P*3] = d(distance)/d(input point) (Klampt's grad1 is its negative), may be NULL. */
int sio_query(sio_ctx* ctx, sio_handle grid, const double T_grid_from_in[12],
              const double* pts, int64_t P, int flags, double* d, double* grad);

/* the same over a resident cloud and B transforms, outputs stay on the device:
 * out_d[B*N] float (and out_g[B*N] float4 = (gx,gy,gz,d) when non-NULL); SIO_F32 only.
 * Point order is the cloud's INTERNAL order (throughput mode; use sio_cloud_perm to map). */
int sio_cloud_query_dev(sio_ctx* ctx, sio_handle grid, sio_handle cloud, const double* dT_rel,
                        int64_t B, int flags, float* out_d, float* out_g4);
int sio_cloud_perm(sio_ctx* ctx, sio_handle cloud, int32_t* perm_out /* internal -> caller index */);

/* ---- K2: cloud-vs-grid minimum + active set -------------------------------------------------
 * replaces: pc.distance_ext(grid, settings{upperBound}) -- geometryopt.py:117-127, called from
 * ObjectCollisionConstraint.minvalue (gopt:409-412), RobotLinkCollisionConstraint.minvalue
 * (gopt:474-477) and the milestone / bisection sweeps of the trajectory constraints
 * (gopt:610, 640, 748, 822, 896).  Batched over B transforms:
 *   T_rel[b]   grid-local <- cloud-local of query b  (= T_grid(b)^-1 * T_cloud)
 *   grid_of_b  index into grids[] per query (NULL: all use grids[0])
 *   bound[b]   Klampt's upperBound; +inf or bound==NULL means none
 * outputs per b:  dmin[b] (exact float64 value of the winner), argmin[b] (cloud index, lowest
 * index on exact ties; -1 when bound is finite and dmin >= bound, the wrapper's "[]" case,
 * gopt:411/476), n_under[b] = #{i : d(b,i) < bound[b]} (0 for an infinite bound).
 * The compact under-bound list (b, i, d) sorted by (b, i) is fetched with sio_under_fetch.
 * n_under == NULL asks for the minimum / arg-minimum only (what the reference's minvalue consumes): no
 * under-bound set is produced, sio_under_fetch then returns an empty list.
 * Capacity: sio_min_distance never returns a truncated set -- when the sweep offers more candidates than the
 * context's buffer holds it grows the buffer and sweeps again.  sio_min_distance_dev cannot wait for the count: if the
 * buffer (sio_ctx_set_under_capacity) is too small, EVERY d_n_under[b] of the call is written as -1 (dmin / argmin
 * stay exact) and sio_under_fetch returns SIO_ERR_OVERFLOW with the number of records needed in its message. */
int sio_min_distance(sio_ctx* ctx, const sio_handle* grids, int32_t n_grids, const int32_t* grid_of_b,
                     sio_handle cloud, const double* T_rel, const double* bound, int64_t B, int flags,
                     double* dmin, int64_t* argmin, int32_t* n_under);
int sio_min_distance_dev(sio_ctx* ctx, const sio_handle* grids, int32_t n_grids, const int32_t* d_grid_of_b,
                         sio_handle cloud, const double* dT_rel, const double* d_bound, int64_t B, int flags,
                         double* d_dmin, int64_t* d_argmin, int32_t* d_n_under /* [B] or NULL */);
/* under-bound list of the LAST sio_min_distance[_dev] call on this context (synchronises).
 * *total = exact number of entries; at most `capacity` are written. */
int sio_under_fetch(sio_ctx* ctx, int64_t capacity, int64_t* b_out, int64_t* idx_out, double* d_out,
                    int64_t* total);

/* ---- K3: Jacobian rows ----------------------------------------------------------------------
 * replaces: ObjectCollisionConstraint.value + df_dx -- geometryopt.py:407-408, 413-418.
 * Batched over `nprob` independent poses; problem p owns points [offsets[p], offsets[p+1]).
 * T_pose[p] = world <- object(grid).  d[i]; rows[i*6] = [ (R(pt - t)) x n , n ], n = grad1. */
int sio_pose_rows(sio_ctx* ctx, sio_handle grid, const double* T_pose, const int64_t* offsets,
                  int64_t nprob, const double* pts, int flags, double* d, double* rows);

/* ---- K4: kinematics -------------------------------------------------------------------------
 * replaces: RobotModel.setConfig + RobotModelLink.getTransform -- geometryopt.py:201-204, 294-295,
 * 636-637, 658-659, 893-895; link.getLocalPosition / getPositionJacobian -- gopt:479-482, 670-672,
 * 923-925.  jtype: 0 revolute, 1 prismatic.  tparent: n*12 row-major, axis: n*3. */
int sio_robot_create(sio_ctx* ctx, int32_t n_links, const int32_t* parent, const int32_t* jtype,
                     const double* tparent, const double* axis, sio_handle* out);
int sio_robot_destroy(sio_ctx* ctx, sio_handle robot);
int sio_fk(sio_ctx* ctx, sio_handle robot, const double* q, int64_t B, double* T_out /* B*n*12 */);
/* RobotLinkCollisionConstraint / RobotTrajectoryCollisionConstraint value + df_dx
 * (gopt:472-473, 478-482, 704-722, 906-925): point i is attached to link[i] with grid
 * grids[link_grid[link[i]]] at configuration q[cfg[i]]; rows[i*n] = Jp^T n. */
int sio_chain_rows(sio_ctx* ctx, sio_handle robot, const sio_handle* link_grids /* n_links, -1 = none */,
                   const double* q, int64_t n_cfg, const int32_t* cfg_of_pt, const int32_t* link_of_pt,
                   const double* pts, int64_t P, int flags, double* d, double* rows);

/* fused FK -> relative transforms -> K2 for robot links against one static cloud (pose sweeps and
 * trajectory time-sample sweeps, gopt:736-759, 802-858 batched).  Query (s, l) uses configuration
 * q[s], link links[l], grid link_grids[links[l]]; T_cloud = world <- cloud.  Outputs are
 * [S*n_sel] in (s, l) order; the under-bound list uses b = s*n_sel + l. */
int sio_robot_min_distance(sio_ctx* ctx, sio_handle robot, const sio_handle* link_grids,
                           const int32_t* links, int32_t n_sel, sio_handle cloud, const double T_cloud[12],
                           const double* q, int64_t S, const double* bound /* S*n_sel or NULL */, int flags,
                           double* dmin, int64_t* argmin, int32_t* n_under);
/* the same for an explicit list of (configuration, link) PAIRS: query p uses configuration q[p] (n values) and link
 * link_of_pair[p].  This is what the second phase of a Lipschitz-scheduled time sweep issues: only the interior time
 * samples of the (edge, link) pairs whose lower bound (intervalLipschitzMinimum, gopt:554-563, 761-797) can still
 * beat the best milestone value are evaluated. */
int sio_robot_pairs_min_distance(sio_ctx* ctx, sio_handle robot, const sio_handle* link_grids, sio_handle cloud,
                                 const double T_cloud[12], const double* q, const int32_t* link_of_pair, int64_t P,
                                 const double* bound /* P or NULL */, int flags, double* dmin, int64_t* argmin,
                                 int32_t* n_under);

/* ---- measurement helpers (bench.py only) ---------------------------------------------------- */
/* ---- B independent object-pose problems in lock step, entirely on the device --------------------------
 * replaces: B runs of optimizeCollFree (geometryopt.py:1039-1081) = optimizeSemiInfinite (sip.py:795-1241, default
 * flags) with one ObjectCollisionConstraint (gopt:394-420) each -- BASELINE.json configs[4].  The moving body owns
 * `grid`, the environment owns `cloud` (pose T_env = world <- cloud; cloud_world [N*3] = its points in world
 * coordinates, caller order -- a function of (cloud, T_env): the library keeps its device copy and re-uploads it only
 * when either differs from the previous call's).  T_init / T_des / T_out are row-major 3x4 poses of the moving body.
 * Per problem: status (1 = 'local optimum', 0 = iteration limit), iterations, f(x), g(x) = min over instantiated
 * parameters, g at the start, number of instantiated parameters and (optional) the parameters themselves
 * [B*cap*3]: the first n_params[p] points of row p are written, the rest of the row is left as it was.  status 2 = the problem's QP had no solution, not even with slack variables (the reference fails there,
 * sip.py:358-378); the problem stops at its current pose.  stats[0..5] = outer iterations, point queries resolved by
 * the sweeps, device microseconds in the in-line QP launches, parameters that did not fit into `cap` slots, device
 * microseconds in the sweeps, problems stopped with status 2.
 * Problems advance independently: a QP that exhausts a small ADMM budget is finished by a background launch while
 * its problem sits rounds out (environment SIO_LS_DEFER = budget, default 800, 0 = off); outputs do not depend on it. */
typedef struct {
    int32_t max_iters;                     /* sip.py:65   (50) */
    int32_t cap;                           /* parameter slots per problem (<= 58: the device QP takes 64 rows) */
    double xepsilon;                       /* sip.py:66   (1e-4) */
    double constraint_inflation;           /* sip.py:68   (1e-3) */
    double constraint_drop_value;          /* sip.py:69   (0.1) */
    double minimum_constraint_value;       /* sip.py:70   (-inf) */
    double initial_objective_score_weight; /* sip.py:71   (0.1) */
    double minimum_objective_score_weight; /* sip.py:73   (1e-5) */
    double parameter_exclusion_distance;   /* sip.py:74   (1e-3) */
    double qp_regularization;              /* sip.py:75   (1e-3) */
    int32_t sweep_flags;                   /* SIO_F32 [| SIO_PRUNE] or SIO_F64 */
} sio_lockstep_settings;
int sio_lockstep_pose_solve(sio_ctx* ctx, sio_handle grid, sio_handle cloud, const double T_env[12], const double* cloud_world,
                            int64_t B, const double* T_init, const double* T_des, const sio_lockstep_settings* settings,
                            double* T_out, int32_t* status, int32_t* iterations, double* fx, double* gx, double* gx0,
                            int32_t* n_params, double* params /* B*cap*3 or NULL */, int64_t stats[6]);

/* ---- geometry preprocessing ------------------------------------------------
 * replaces: the distance computation inside geom.convert('VolumeGrid', res) -- geometryopt.py:58-66.
 * out_dist[(i*ny + j)*nz + k] = exact distance from the cell centre bmin + (i+.5, j+.5, k+.5) h to the triangle mesh
 * (verts [nv*3] float64, tris [nt*3] int32).  The caller applies the inside/outside sign (host, winding number near
 * the surface only) and narrows to float32; the values equal the host builder's bit for bit. */
int sio_mesh_distance_grid(sio_ctx* ctx, const double* verts, int32_t nv, const int32_t* tris, int32_t nt,
                           const double bmin[3], const double h[3], const int32_t dims[3], double* out_dist);

/* The whole conversion on the device: distances as above, then the side of every cell -- the generalised winding number
 * (> 1/2 = inside; robust to the open / self-intersecting benchmark meshes) for the first cell of each z-column and for
 * every cell within one cell size of the surface, inherited from the z-predecessor elsewhere -- and the float32 samples
 * out_sdf[(i*ny + j)*nz + k] that sio_grid_create takes.  Equal to the host builder's grid (tests). */
int sio_mesh_sdf_grid(sio_ctx* ctx, const double* verts, int32_t nv, const int32_t* tris, int32_t nt,
                      const double bmin[3], const double h[3], const int32_t dims[3], float* out_sdf);
/* replaces: geom.convert('PointCloud', pcres) -- geometryopt.py:65.  The mesh vertices plus, per triangle, the lattice
 * spanned by its two shorter edges with ceil(edge / res) subdivisions (no surface point farther than ~res from a
 * sample), points shared between triangles merged (quantised to 1e-9 m, first occurrence kept, order preserved).
 * out_pts [capacity*3] float64 (NULL: count only); *n_out = number of points.  sio_mesh_surface_sample_bound (host work
 * only, no device) returns the number of points before merging: an upper bound to size out_pts with. */
int64_t sio_mesh_surface_sample_bound(const double* verts, int32_t nv, const int32_t* tris, int32_t nt, double res);
int sio_mesh_surface_sample(sio_ctx* ctx, const double* verts, int32_t nv, const int32_t* tris, int32_t nt, double res,
                            double* out_pts, int64_t capacity, int64_t* n_out);

/* ---- the QPs of one lock-step SIP iteration ------------------------------------
 * replaces: ConstraintGenerationData.solve_qp_osqp -> osqp.OSQP().setup/solve -- sip.py:294-331, one
 * call per problem per outer iteration (sip.py:964), here for `nprob` problems at once, one warp each:
 *     min (x - xdes)' diag(W) (x - xdes) + reg |x|^2   s.t.  rows x + depths >= inflation,  xmin <= x <= xmax
 *   W, xdes, xmin, xmax [nprob*n] (xmin/xmax may both be NULL); offsets [nprob+1] into rows [total*n], depths [total]
 *   n = 6 (pose problems) or 7 (TX90 configurations).
 * status[p]: 0 = strict problem solved, 1 = solved by the reference's slack re-solve (strict problem infeasible or
 * out of scale, sip.py:332-357), 2 = no rows (dx = xdes, sip.py:288), 3 = not handled on the device (more than 64
 * rows, or a polish system larger than the kernel's scratch): the caller solves those on the host (hostqp.c),
 * dx[p] is left untouched. */
int sio_sip_step_qp(sio_ctx* ctx, int32_t n, int64_t nprob, const double* W, const double* xdes, const double* xmin,
                    const double* xmax, const int64_t* offsets, const double* rows, const double* depths, double reg,
                    double inflation, double eps_abs, int32_t max_iter, double* dx, int32_t* status);

/* device time (ms, CUDA events on the context's stream) of the fp32 bulk kernel launches
 * (k_min_f32) of the LAST sio_min_distance[_dev] call; synchronises on that call's end event. */
int sio_ctx_last_bulk_ms(sio_ctx* ctx, double* ms);
/* SIO_PRUNE bookkeeping of the LAST K2 call: (transform, sub-tile) pairs evaluated vs total. */
int sio_ctx_last_prune_stats(sio_ctx* ctx, int64_t* evaluated, int64_t* total);
/* ADMM iterations of the LAST sio_sip_step_qp call: summed over its problems, the largest single count, and the number
 * of solves that ran into the iteration limit (diagnostics; the last two may be NULL). */
int sio_ctx_last_qp_iters(sio_ctx* ctx, int64_t* iters, int64_t* max_single, int64_t* n_at_limit);
/* random 4-byte gathers over an L2-resident float array of `n_elems`: returns achieved GB/s of
 * gathered bytes (32 B sectors are NOT counted, only the 4 B used) in *gbps. */
int sio_bench_gather(sio_ctx* ctx, int64_t n_elems, int64_t n_gathers, int iters, double* gbps);

#ifdef __cplusplus
}
#endif
#endif /* SIO_ABI_H */
