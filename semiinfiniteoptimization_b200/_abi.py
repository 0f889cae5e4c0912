"""ctypes binding of the C-ABI in include/sio_abi.h (libsio_b200.so).

This is the only door from the Python mirror of the reference's API to the device.  There is no
CPU implementation behind it: if the library is missing, or no sm_100 device is usable,
`Context()` raises -- the oracle path never silently falls back (SURVEY.md 8b "Errors").
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libsio_b200.so")

SIO_OK = 0
SIO_ERR_OVERFLOW = -5
SIO_F32 = 0x0
SIO_F64 = 0x1
SIO_NO_NORMALIZE = 0x2
SIO_PRUNE = 0x8
SIO_SMEM_GRID = 0x10
SIO_PUSH_BOARDS = 0x20

# every symbol include/sio_abi.h declares (tests check the built library exports each of them)
ABI_SYMBOLS = (
    "sio_abi_version", "sio_last_error",
    "sio_ctx_create", "sio_ctx_destroy", "sio_ctx_sync", "sio_ctx_stream", "sio_ctx_set_tau", "sio_ctx_set_scratch_limit",
    "sio_ctx_set_under_capacity", "sio_ctx_set_host_graphs", "sio_ctx_num_sms",
    "sio_board_create", "sio_board_open", "sio_board_close", "sio_ctx_set_boards", "sio_board_read",
    "sio_ctx_kernel_launches",
    "sio_grid_create", "sio_grid_destroy", "sio_cloud_create", "sio_cloud_order", "sio_cloud_destroy", "sio_cloud_size",
    "sio_query", "sio_cloud_query_dev", "sio_cloud_perm",
    "sio_min_distance", "sio_min_distance_dev", "sio_under_fetch",
    "sio_pose_rows", "sio_sip_step_qp", "sio_mesh_distance_grid", "sio_mesh_sdf_grid", "sio_mesh_surface_sample", "sio_mesh_surface_sample_bound", "sio_lockstep_pose_solve",
    "sio_robot_create", "sio_robot_destroy", "sio_fk", "sio_chain_rows", "sio_robot_min_distance", "sio_robot_pairs_min_distance",
    "sio_bench_gather", "sio_ctx_last_bulk_ms", "sio_ctx_last_prune_stats", "sio_ctx_last_qp_iters",
)

_lib = None


class SioError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "sio error %d: %s" % (code, msg))
        self.code = code


def load_library():
    """dlopen libsio_b200.so and declare prototypes.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the collision oracle has no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    P, I32, I64, D = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.sio_abi_version.restype = C.c_int
    L.sio_last_error.restype = C.c_char_p
    L.sio_ctx_create.argtypes = [C.c_int, C.POINTER(P)]
    L.sio_ctx_destroy.argtypes = [P]
    L.sio_ctx_sync.argtypes = [P]
    L.sio_ctx_stream.argtypes = [P]
    L.sio_ctx_stream.restype = P
    L.sio_ctx_set_tau.argtypes = [P, D]
    L.sio_ctx_set_scratch_limit.argtypes = [P, I64]
    L.sio_ctx_set_under_capacity.argtypes = [P, I64]
    L.sio_ctx_set_host_graphs.argtypes = [P, C.c_int]
    L.sio_board_create.argtypes = [P, I64, C.POINTER(P), P]
    L.sio_board_open.argtypes = [P, P, C.POINTER(P)]
    L.sio_board_close.argtypes = [P, P, C.c_int]
    L.sio_ctx_set_boards.argtypes = [P, I32, I32, P, I64]
    L.sio_board_read.argtypes = [P, P, I64]
    L.sio_ctx_num_sms.argtypes = [P]
    L.sio_ctx_kernel_launches.argtypes = [P]
    L.sio_ctx_kernel_launches.restype = I64
    L.sio_grid_create.argtypes = [P, P, P, P, P, C.POINTER(I32)]
    L.sio_grid_destroy.argtypes = [P, I32]
    L.sio_cloud_create.argtypes = [P, P, I64, C.POINTER(I32)]
    L.sio_cloud_order.argtypes = [P, I64, P]
    L.sio_cloud_destroy.argtypes = [P, I32]
    L.sio_cloud_size.argtypes = [P, I32]
    L.sio_cloud_size.restype = I64
    L.sio_query.argtypes = [P, I32, P, P, I64, C.c_int, P, P]
    L.sio_cloud_query_dev.argtypes = [P, I32, I32, P, I64, C.c_int, P, P]
    L.sio_cloud_perm.argtypes = [P, I32, P]
    L.sio_min_distance.argtypes = [P, P, I32, P, I32, P, P, I64, C.c_int, P, P, P]
    L.sio_min_distance_dev.argtypes = [P, P, I32, P, I32, P, P, I64, C.c_int, P, P, P]
    L.sio_under_fetch.argtypes = [P, I64, P, P, P, C.POINTER(I64)]
    L.sio_pose_rows.argtypes = [P, I32, P, P, I64, P, C.c_int, P, P]
    L.sio_sip_step_qp.argtypes = [P, I32, I64, P, P, P, P, P, P, P, D, D, D, I32, P, P]
    L.sio_mesh_distance_grid.argtypes = [P, P, I32, P, I32, P, P, P, P]
    L.sio_mesh_sdf_grid.argtypes = [P, P, I32, P, I32, P, P, P, P]
    L.sio_mesh_surface_sample.argtypes = [P, P, I32, P, I32, D, P, I64, C.POINTER(I64)]
    L.sio_mesh_surface_sample_bound.argtypes = [P, I32, P, I32, D]
    L.sio_mesh_surface_sample_bound.restype = I64
    L.sio_lockstep_pose_solve.argtypes = [P, I32, I32, P, P, I64, P, P, P, P, P, P, P, P, P, P, P, P]
    L.sio_robot_create.argtypes = [P, I32, P, P, P, P, C.POINTER(I32)]
    L.sio_robot_destroy.argtypes = [P, I32]
    L.sio_fk.argtypes = [P, I32, P, I64, P]
    L.sio_chain_rows.argtypes = [P, I32, P, P, I64, P, P, P, I64, C.c_int, P, P]
    L.sio_robot_min_distance.argtypes = [P, I32, P, P, I32, I32, P, P, I64, P, C.c_int, P, P, P]
    L.sio_robot_pairs_min_distance.argtypes = [P, I32, P, I32, P, P, P, I64, P, C.c_int, P, P, P]
    L.sio_bench_gather.argtypes = [P, I64, I64, C.c_int, C.POINTER(D)]
    L.sio_ctx_last_bulk_ms.argtypes = [P, C.POINTER(D)]
    L.sio_ctx_last_prune_stats.argtypes = [P, C.POINTER(I64), C.POINTER(I64)]
    L.sio_ctx_last_qp_iters.argtypes = [P, C.POINTER(I64), C.POINTER(I64), C.POINTER(I64)]
    _lib = L
    return L


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def cloud_order(xyz):
    """The order sio_cloud_create keeps a cloud in (host work only: usable without a GPU): internal -> caller index."""
    L = load_library()
    p = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    out = np.empty(len(p), dtype=np.int32)
    rc = L.sio_cloud_order(p.ctypes.data, len(p), out.ctypes.data)
    if rc != SIO_OK:
        raise SioError(rc, L.sio_last_error().decode())
    return out


def _ptr(a):
    return None if a is None else a.ctypes.data


def lockstep_outputs(B, cap=48, want_params=True):
    """The per-problem result arrays of Context.lockstep_pose_solve for B problems."""
    return {"T": np.empty((B, 12)), "status": np.empty(B, dtype=np.int32), "iterations": np.empty(B, dtype=np.int32),
            "fx": np.empty(B), "gx": np.empty(B), "gx0": np.empty(B), "n_params": np.empty(B, dtype=np.int32),
            "params": np.empty((B, cap, 3)) if want_params else None}


class Context:
    """One device context (stream + scratch + resident handles).  Calls are serialised on its stream."""

    def __init__(self, device=0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.sio_ctx_create(int(device), C.byref(h))
        if rc != SIO_OK:
            raise SioError(rc, self.L.sio_last_error().decode())
        self.h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "h", None):
            self.L.sio_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != SIO_OK:
            raise SioError(rc, self.L.sio_last_error().decode())

    # -- context --
    def sync(self):
        self._ck(self.L.sio_ctx_sync(self.h))

    @property
    def stream(self):
        return self.L.sio_ctx_stream(self.h)

    def set_tau(self, tau):
        self._ck(self.L.sio_ctx_set_tau(self.h, float(tau)))

    def set_scratch_limit(self, nbytes):
        self._ck(self.L.sio_ctx_set_scratch_limit(self.h, int(nbytes)))

    def set_under_capacity(self, entries):
        """Under-bound candidate records a K2 call holds (the host entry points grow it on their own)."""
        self._ck(self.L.sio_ctx_set_under_capacity(self.h, int(entries)))

    def set_host_graphs(self, on):
        """Replay repeated host-pointer K2 calls as one CUDA graph (default on)."""
        self._ck(self.L.sio_ctx_set_host_graphs(self.h, 1 if on else 0))

    # -- result boards (multi-GPU exchange by peer stores) --
    def board_create(self, nbytes):
        """(device pointer, 64-byte IPC handle) of a zeroed local board."""
        p = C.c_void_p()
        h = (C.c_ubyte * 64)()
        self._ck(self.L.sio_board_create(self.h, int(nbytes), C.byref(p), h))
        return p.value, bytes(h)

    def board_open(self, handle):
        p = C.c_void_p()
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        self._ck(self.L.sio_board_open(self.h, buf, C.byref(p)))
        return p.value

    def board_close(self, ptr, is_local):
        self._ck(self.L.sio_board_close(self.h, C.c_void_p(ptr), 1 if is_local else 0))

    def set_boards(self, rank, ptrs, slot_bytes):
        if not ptrs:
            self._ck(self.L.sio_ctx_set_boards(self.h, 0, 0, None, 0))
            return
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        self._ck(self.L.sio_ctx_set_boards(self.h, int(rank), len(ptrs), arr, int(slot_bytes)))

    def board_read(self, nranks, slot_bytes, B):
        """The local board as (dmin[nranks,B], argmin[nranks,B], n_under[nranks,B])."""
        raw = np.empty(nranks * slot_bytes, dtype=np.uint8)
        self._ck(self.L.sio_board_read(self.h, raw.ctypes.data, raw.nbytes))
        raw = raw.reshape(nranks, slot_bytes)
        d = np.ascontiguousarray(raw[:, :8 * B]).view(np.float64)
        a = np.ascontiguousarray(raw[:, 8 * B:16 * B]).view(np.int64)
        n = np.ascontiguousarray(raw[:, 16 * B:20 * B]).view(np.int32)
        return d, a, n

    @property
    def num_sms(self):
        return self.L.sio_ctx_num_sms(self.h)

    @property
    def kernel_launches(self):
        return int(self.L.sio_ctx_kernel_launches(self.h))

    # -- geometry --
    def grid_create(self, values, bmin, bmax):
        v = np.ascontiguousarray(values, dtype=np.float32)
        if v.ndim != 3:
            raise ValueError("grid values must be a 3-D array (nx,ny,nz), z fastest")
        dims = np.array(v.shape, dtype=np.int32)
        bmin, bmax = _f64(bmin), _f64(bmax)
        out = C.c_int32(-1)
        self._ck(self.L.sio_grid_create(self.h, v.ctypes.data, dims.ctypes.data, bmin.ctypes.data, bmax.ctypes.data, C.byref(out)))
        return out.value

    def grid_destroy(self, g):
        self._ck(self.L.sio_grid_destroy(self.h, g))

    def cloud_create(self, xyz):
        p = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        out = C.c_int32(-1)
        self._ck(self.L.sio_cloud_create(self.h, p.ctypes.data, len(p), C.byref(out)))
        return out.value

    def cloud_destroy(self, c):
        self._ck(self.L.sio_cloud_destroy(self.h, c))

    def cloud_size(self, c):
        return int(self.L.sio_cloud_size(self.h, c))

    def cloud_perm(self, c):
        out = np.empty(self.cloud_size(c), dtype=np.int32)
        self._ck(self.L.sio_cloud_perm(self.h, c, out.ctypes.data))
        return out

    # -- K1 --
    def query(self, grid, T_grid_from_in, pts, flags=SIO_F64, want_grad=True):
        T = _f64(T_grid_from_in).reshape(12)
        p = _f64(pts).reshape(-1, 3)
        d = np.empty(len(p))
        g = np.empty((len(p), 3)) if want_grad else None
        self._ck(self.L.sio_query(self.h, grid, T.ctypes.data, p.ctypes.data, len(p), int(flags), d.ctypes.data, _ptr(g)))
        return (d, g) if want_grad else d

    def cloud_query_dev(self, grid, cloud, dT_rel_ptr, B, flags, out_d_ptr, out_g4_ptr):
        self._ck(self.L.sio_cloud_query_dev(self.h, grid, cloud, dT_rel_ptr, int(B), int(flags), out_d_ptr, out_g4_ptr))

    # -- K2 --
    def min_distance(self, grids, cloud, T_rel, bound=None, grid_of_b=None, flags=SIO_F32, want_under=True):
        """Returns dmin[B], argmin[B], n_under[B]; the under-bound list comes from under_fetch()."""
        if np.isscalar(grids):
            grids = [grids]
        ga = np.ascontiguousarray(grids, dtype=np.int32)
        T = _f64(T_rel).reshape(-1, 12)
        B = len(T)
        bnd = None if bound is None else _f64(np.broadcast_to(bound, (B,)))
        gob = None if grid_of_b is None else np.ascontiguousarray(grid_of_b, dtype=np.int32)
        dmin = np.empty(B)
        arg = np.empty(B, dtype=np.int64)
        nu = np.zeros(B, dtype=np.int32) if want_under else None
        self._ck(self.L.sio_min_distance(self.h, ga.ctypes.data, len(ga), _ptr(gob), cloud, T.ctypes.data, _ptr(bnd), B,
                                         int(flags), dmin.ctypes.data, arg.ctypes.data, _ptr(nu)))
        return dmin, arg, nu

    def min_distance_plan(self, grids, cloud, B, flags=SIO_F32, with_bound=True, want_under=True):
        """A prebound `min_distance` for a loop that repeats one call shape: see MinDistancePlan."""
        return MinDistancePlan(self, grids, cloud, B, flags, with_bound, want_under)

    def min_distance_dev(self, grids, cloud, dT_rel_ptr, d_bound_ptr, B, flags, d_dmin_ptr, d_argmin_ptr, d_gob_ptr=None,
                         d_n_under_ptr=None):
        ga = np.ascontiguousarray(grids, dtype=np.int32)
        self._ck(self.L.sio_min_distance_dev(self.h, ga.ctypes.data, len(ga), d_gob_ptr, cloud, dT_rel_ptr, d_bound_ptr, int(B),
                                             int(flags), d_dmin_ptr, d_argmin_ptr, d_n_under_ptr))

    def under_fetch(self, capacity=1 << 20):
        b = np.empty(capacity, dtype=np.int64)
        i = np.empty(capacity, dtype=np.int64)
        d = np.empty(capacity)
        tot = C.c_int64(0)
        self._ck(self.L.sio_under_fetch(self.h, capacity, b.ctypes.data, i.ctypes.data, d.ctypes.data, C.byref(tot)))
        n = tot.value
        return b[:n], i[:n], d[:n]

    # -- K3 --
    _POSE_ROWS_CAP = 256

    def pose_rows(self, grid, T_pose, offsets, pts, flags=SIO_F64, want_rows=True):
        if len(offsets) == 2 and len(pts) <= self._POSE_ROWS_CAP:
            # ONE pose and a handful of points -- the call a single-problem SIP run makes three times per iteration: inputs
            # go into arrays this context keeps (their addresses fetched once), outputs come back as copies
            sc = getattr(self, "_pr", None)
            if sc is None:
                cap = self._POSE_ROWS_CAP
                arrs = (np.empty(12), np.zeros(2, dtype=np.int64), np.empty((cap, 3)), np.empty(cap), np.empty((cap, 6)))
                sc = self._pr = arrs + tuple(a.ctypes.data for a in arrs)
            T, off, p, d, rows, pT, poff, pp, pd, prows = sc
            P = len(pts)
            T[...] = np.reshape(T_pose, 12)
            off[1] = P
            if P:
                p[:P] = pts
            self._ck(self.L.sio_pose_rows(self.h, grid, pT, poff, 1, pp, int(flags), pd, prows if want_rows else None))
            return d[:P].copy(), (rows[:P].copy() if want_rows else None)
        T = _f64(T_pose).reshape(-1, 12)
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        p = _f64(pts).reshape(-1, 3)
        assert len(off) == len(T) + 1 and off[-1] == len(p)
        d = np.empty(len(p))
        rows = np.empty((len(p), 6)) if want_rows else None
        self._ck(self.L.sio_pose_rows(self.h, grid, T.ctypes.data, off.ctypes.data, len(T), p.ctypes.data, int(flags),
                                      d.ctypes.data, _ptr(rows)))
        return d, rows

    # -- lock-step pose solver (device resident) --
    def lockstep_pose_solve(self, grid, cloud, T_env, cloud_world, T_init, T_des, settings, cap=48, sweep_flags=SIO_F32 | SIO_PRUNE,
                            want_params=True, out=None):
        """B object-pose SIP problems in lock step on the device (sio_lockstep_pose_solve).  `settings`: a
        SemiInfiniteOptimizationSettings.  Returns a dict of per-problem arrays and the stats.  `out`: the same dict
        with caller-owned C-contiguous arrays of B rows (slices of a larger batch's arrays) to be filled in place."""
        class _S(C.Structure):
            _fields_ = [("max_iters", C.c_int32), ("cap", C.c_int32), ("xepsilon", C.c_double), ("constraint_inflation", C.c_double),
                        ("constraint_drop_value", C.c_double), ("minimum_constraint_value", C.c_double),
                        ("initial_objective_score_weight", C.c_double), ("minimum_objective_score_weight", C.c_double),
                        ("parameter_exclusion_distance", C.c_double), ("qp_regularization", C.c_double), ("sweep_flags", C.c_int32)]
        Ti = _f64(T_init).reshape(-1, 12)
        Td = _f64(T_des).reshape(-1, 12)
        B = len(Ti)
        assert len(Td) == B
        cw = _f64(cloud_world).reshape(-1, 3)
        Te = _f64(T_env).reshape(12)
        s = _S(int(settings.max_iters), int(cap), float(settings.xepsilon), float(settings.constraint_inflation),
               float(settings.constraint_drop_value), float(settings.minimum_constraint_value),
               float(settings.initial_objective_score_weight), float(settings.minimum_objective_score_weight),
               float(settings.parameter_exclusion_distance), float(settings.qp_solver_params.get('regularizationFactor', 1e-2)),
               int(sweep_flags))
        if out is None:
            out = lockstep_outputs(B, cap, want_params)
        else:
            shapes = lockstep_outputs(0, cap, want_params)
            for k, ref in shapes.items():
                a = out[k]
                if ref is None:
                    assert a is None
                    continue
                assert a.dtype == ref.dtype and a.shape == (B,) + ref.shape[1:] and a.flags.c_contiguous, k
        stats = np.zeros(6, dtype=np.int64)
        self._ck(self.L.sio_lockstep_pose_solve(self.h, grid, cloud, Te.ctypes.data, cw.ctypes.data, B, Ti.ctypes.data, Td.ctypes.data,
                                                C.byref(s), out["T"].ctypes.data, out["status"].ctypes.data,
                                                out["iterations"].ctypes.data, out["fx"].ctypes.data, out["gx"].ctypes.data,
                                                out["gx0"].ctypes.data, out["n_params"].ctypes.data, _ptr(out["params"]),
                                                stats.ctypes.data))
        out["stats"] = {"outer_iterations": int(stats[0]), "point_queries": int(stats[1]), "qp_device_s": stats[2] * 1e-6,
                        "slot_overflows": int(stats[3]), "sweep_device_s": stats[4] * 1e-6, "qp_failures": int(stats[5])}
        return out

    # -- geometry preprocessing --
    def mesh_distance_grid(self, verts, tris, bmin, h, dims):
        """Unsigned distance from every cell centre to the mesh, float64 (nx, ny, nz)."""
        V = _f64(verts).reshape(-1, 3)
        T = np.ascontiguousarray(tris, dtype=np.int32).reshape(-1, 3)
        b, hh = _f64(bmin).reshape(3), _f64(h).reshape(3)
        d3 = np.ascontiguousarray(dims, dtype=np.int32).reshape(3)
        out = np.empty(tuple(int(v) for v in d3))
        self._ck(self.L.sio_mesh_distance_grid(self.h, V.ctypes.data, len(V), T.ctypes.data, len(T), b.ctypes.data, hh.ctypes.data,
                                               d3.ctypes.data, out.ctypes.data))
        return out

    def mesh_sdf_grid(self, verts, tris, bmin, h, dims):
        """Signed distance samples float32 (nx, ny, nz) at the cell centres: distances AND sides on the device."""
        V = _f64(verts).reshape(-1, 3)
        T = np.ascontiguousarray(tris, dtype=np.int32).reshape(-1, 3)
        b, hh = _f64(bmin).reshape(3), _f64(h).reshape(3)
        d3 = np.ascontiguousarray(dims, dtype=np.int32).reshape(3)
        out = np.empty(tuple(int(v) for v in d3), dtype=np.float32)
        self._ck(self.L.sio_mesh_sdf_grid(self.h, V.ctypes.data, len(V), T.ctypes.data, len(T), b.ctypes.data, hh.ctypes.data,
                                          d3.ctypes.data, out.ctypes.data))
        return out

    def mesh_surface_sample(self, verts, tris, res):
        """Surface point cloud (n, 3) float64 at resolution `res` (vertices + per-triangle lattices, duplicates merged)."""
        V = _f64(verts).reshape(-1, 3)
        T = np.ascontiguousarray(tris, dtype=np.int32).reshape(-1, 3)
        tp = T.ctypes.data if len(T) else None
        bound = self.L.sio_mesh_surface_sample_bound(V.ctypes.data, len(V), tp, len(T), float(res))     # host only: points before merging
        if bound < 0:
            raise SioError(-1, self.L.sio_last_error().decode())
        out = np.empty((bound, 3))
        n = C.c_int64(0)
        self._ck(self.L.sio_mesh_surface_sample(self.h, V.ctypes.data, len(V), tp, len(T), float(res), out.ctypes.data, bound, C.byref(n)))
        return out[:n.value].copy()

    # -- SIP-step QPs --
    def sip_step_qp(self, W, xdes, offs, rows, depths, xmin=None, xmax=None, reg=1e-2, inflation=1e-3, eps_abs=1e-5, max_iter=4000):
        """dx (B,n), status (B,) of one lock-step iteration's strict QPs; status 3 = solve on the host."""
        W = _f64(W)
        xdes = _f64(xdes)
        B, n = W.shape
        off = np.ascontiguousarray(offs, dtype=np.int64)
        rows = _f64(rows).reshape(-1, n)
        depths = _f64(depths)
        assert len(off) == B + 1 and off[-1] == len(rows) == len(depths)
        lo = None if xmin is None else _f64(xmin)
        hi = None if xmax is None else _f64(xmax)
        dx = np.zeros((B, n))
        status = np.empty(B, dtype=np.int32)
        self._ck(self.L.sio_sip_step_qp(self.h, n, B, W.ctypes.data, xdes.ctypes.data, _ptr(lo), _ptr(hi), off.ctypes.data,
                                        rows.ctypes.data if len(rows) else None, depths.ctypes.data if len(depths) else None,
                                        float(reg), float(inflation), float(eps_abs), int(max_iter), dx.ctypes.data, status.ctypes.data))
        return dx, status

    def last_qp_iters(self, detail=False):
        n, mx, lim = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self._ck(self.L.sio_ctx_last_qp_iters(self.h, C.byref(n), C.byref(mx), C.byref(lim)))
        return (n.value, mx.value, lim.value) if detail else n.value

    # -- K4 --
    def robot_create(self, parent, tparent, axis, jtype=None):
        par = np.ascontiguousarray(parent, dtype=np.int32)
        n = len(par)
        tp = _f64(tparent).reshape(n, 12)
        ax = _f64(axis).reshape(n, 3)
        jt = None if jtype is None else np.ascontiguousarray(jtype, dtype=np.int32)
        out = C.c_int32(-1)
        self._ck(self.L.sio_robot_create(self.h, n, par.ctypes.data, _ptr(jt), tp.ctypes.data, ax.ctypes.data, C.byref(out)))
        return out.value

    def robot_destroy(self, r):
        self._ck(self.L.sio_robot_destroy(self.h, r))

    def fk(self, robot, q, n_links):
        q = _f64(q).reshape(-1, n_links)
        out = np.empty((len(q), n_links, 12))
        self._ck(self.L.sio_fk(self.h, robot, q.ctypes.data, len(q), out.ctypes.data))
        return out

    def chain_rows(self, robot, n_links, link_grids, q, link_of_pt, pts, cfg_of_pt=None, flags=SIO_F64, want_rows=True):
        if cfg_of_pt is None and np.isscalar(link_of_pt) and len(pts) <= self._POSE_ROWS_CAP and len(q) == 1 and n_links <= 16:
            # ONE configuration, one link, a handful of points (what a link constraint of a robot-pose problem asks three
            # times per iteration): inputs into arrays this context keeps, their addresses fetched once (see pose_rows)
            key = (int(robot), int(n_links), tuple(int(g) for g in link_grids))
            sc = getattr(self, "_cr", None)
            if sc is None or sc[0] != key:
                cap = self._POSE_ROWS_CAP
                arrs = (np.array(key[2], dtype=np.int32), np.empty((1, n_links)), np.zeros(cap, dtype=np.int32), np.empty((cap, 3)),
                        np.empty(cap), np.empty((cap, n_links)))
                sc = self._cr = (key,) + arrs + tuple(a.ctypes.data for a in arrs)
            _, lg, qq, lop, p, d, rows, plg, pq, plop, pp, pd, prows = sc
            P = len(pts)
            qq[0, :] = q[0]
            lop[:P] = int(link_of_pt)
            if P:
                p[:P] = pts
            self._ck(self.L.sio_chain_rows(self.h, robot, plg, pq, 1, None, plop, pp, P, int(flags), pd, prows if want_rows else None))
            return d[:P].copy(), (rows[:P].copy() if want_rows else None)
        lg = np.ascontiguousarray(link_grids, dtype=np.int32)
        q = _f64(q).reshape(-1, n_links)
        p = _f64(pts).reshape(-1, 3)
        lop = np.ascontiguousarray(np.broadcast_to(link_of_pt, (len(p),)), dtype=np.int32)
        cop = None if cfg_of_pt is None else np.ascontiguousarray(cfg_of_pt, dtype=np.int32)
        d = np.empty(len(p))
        rows = np.empty((len(p), n_links)) if want_rows else None
        self._ck(self.L.sio_chain_rows(self.h, robot, lg.ctypes.data, q.ctypes.data, len(q), _ptr(cop), lop.ctypes.data,
                                       p.ctypes.data, len(p), int(flags), d.ctypes.data, _ptr(rows)))
        return d, rows

    def robot_min_distance(self, robot, n_links, link_grids, links, cloud, T_cloud, q, bound=None, flags=SIO_F32):
        lg = np.ascontiguousarray(link_grids, dtype=np.int32)
        ls = np.ascontiguousarray(links, dtype=np.int32)
        q = _f64(q).reshape(-1, n_links)
        S, nsel = len(q), len(ls)
        Tc = _f64(T_cloud).reshape(12)
        bnd = None if bound is None else _f64(np.broadcast_to(bound, (S, nsel)))
        dmin = np.empty((S, nsel))
        arg = np.empty((S, nsel), dtype=np.int64)
        nu = np.zeros((S, nsel), dtype=np.int32)
        self._ck(self.L.sio_robot_min_distance(self.h, robot, lg.ctypes.data, ls.ctypes.data, nsel, cloud, Tc.ctypes.data,
                                               q.ctypes.data, S, _ptr(bnd), int(flags), dmin.ctypes.data, arg.ctypes.data,
                                               nu.ctypes.data))
        return dmin, arg, nu

    def robot_pairs_min_distance(self, robot, n_links, link_grids, cloud, T_cloud, q, link_of_pair, bound=None, flags=SIO_F32):
        """(dmin, argmin) of explicit (configuration, link) pairs: q (P, n_links), link_of_pair (P,)."""
        lg = np.ascontiguousarray(link_grids, dtype=np.int32)
        q = _f64(q).reshape(-1, n_links)
        lp = np.ascontiguousarray(link_of_pair, dtype=np.int32)
        assert len(lp) == len(q)
        Tc = _f64(T_cloud).reshape(12)
        bnd = None if bound is None else _f64(np.broadcast_to(bound, (len(q),)))
        dmin = np.empty(len(q))
        arg = np.empty(len(q), dtype=np.int64)
        self._ck(self.L.sio_robot_pairs_min_distance(self.h, robot, lg.ctypes.data, cloud, Tc.ctypes.data, q.ctypes.data,
                                                     lp.ctypes.data, len(q), _ptr(bnd), int(flags), dmin.ctypes.data,
                                                     arg.ctypes.data, None))
        return dmin, arg

    def last_bulk_ms(self):
        out = C.c_double(0)
        self._ck(self.L.sio_ctx_last_bulk_ms(self.h, C.byref(out)))
        return out.value

    def last_prune_stats(self):
        ev, tot = C.c_int64(0), C.c_int64(0)
        self._ck(self.L.sio_ctx_last_prune_stats(self.h, C.byref(ev), C.byref(tot)))
        return ev.value, tot.value

    def bench_gather(self, n_elems, n_gathers, iters=10):
        out = C.c_double(0)
        self._ck(self.L.sio_bench_gather(self.h, int(n_elems), int(n_gathers), int(iters), C.byref(out)))
        return out.value


class MinDistancePlan:
    """`Context.min_distance` for a fixed (grids, cloud, B, flags), with everything the interpreter has to do per call done
    once: the plan owns the input and output arrays and holds their addresses, so a call is two small array copies and one
    foreign call.  `min_distance` itself spends 13 us per call in numpy / ctypes glue (six `ndarray.ctypes.data` at 1.7 us,
    array conversions and allocations) -- 8 % of a 64-pose x 2^20-point sweep, half of a single-pose query.
        plan = ctx.min_distance_plan(grid, cloud, B)
        dmin, argmin, n_under = plan(T_rel, bound)        # arrays OWNED BY THE PLAN: valid until its next call
    Same entry point (sio_min_distance), same copies on the device side, same results."""

    def __init__(self, ctx, grids, cloud, B, flags=SIO_F32, with_bound=True, want_under=True):
        self.ctx, self.B = ctx, int(B)
        self._ga = np.ascontiguousarray([grids] if np.isscalar(grids) else grids, dtype=np.int32)
        self.T = np.empty((self.B, 12))
        self.bound = np.empty(self.B) if with_bound else None
        self.dmin, self.argmin = np.empty(self.B), np.empty(self.B, dtype=np.int64)
        self.n_under = np.zeros(self.B, dtype=np.int32) if want_under else None
        self._fn = ctx.L.sio_min_distance
        self._args = (ctx.h, self._ga.ctypes.data, len(self._ga), None, int(cloud), self.T.ctypes.data, _ptr(self.bound), self.B,
                      int(flags), self.dmin.ctypes.data, self.argmin.ctypes.data, _ptr(self.n_under))

    def __call__(self, T_rel, bound=None):
        self.T[...] = np.reshape(T_rel, (self.B, 12))
        if self.bound is not None:
            self.bound[...] = bound
        rc = self._fn(*self._args)
        if rc != SIO_OK:
            raise SioError(rc, self.ctx.L.sio_last_error().decode())
        return self.dmin, self.argmin, self.n_under


_default_ctx = {}


def default_context(device=None):
    """Process-wide context per device (created on first use).  device=None: the GPU of this process -- LOCAL_RANK
    under torchrun (one process per GPU), else device 0."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
