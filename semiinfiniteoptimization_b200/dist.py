"""Multi-GPU plumbing: one process per GPU, `torch.distributed` (NCCL on the GPU box, gloo in the
CPU tests) -- SURVEY.md 8e.

The oracle path has NO data-path collective: independent problems (C5) or blocks of trajectory
time samples (C4) are sharded across ranks with geometry replicated, every rank runs its own
kernels, and only COMPACT results cross the links:

  * `gather_records`   all-gather of fixed-size per-problem records (C5: pose + status + counts);
  * `argmin_allreduce` the global (value, payload) minimum over ranks (C4: the deepest
                       (link, env, t, point) over all time blocks), by gathering one small record
                       per rank and reducing it on every rank identically (deterministic tie-break:
                       value, then the integer payload in order).

Payloads are tens of bytes per problem, so latency -- not NVLink bandwidth -- is what matters; no
custom kernel is warranted here (the compute kernels have no collective to fuse with).
"""
import numpy as np


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


def world():
    d = _dist()
    return (d.get_rank(), d.get_world_size()) if d else (0, 1)


def shard_range(n, rank=None, world_size=None):
    """Contiguous block [lo, hi) of n units owned by `rank`; blocks differ in size by at most one."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _device_for_backend():
    import torch
    d = _dist()
    if d is not None and d.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


_bufs = {}


def _buffers(dev, ws, nmax, k):
    """Device staging + gathered buffers and a pinned host landing buffer, kept per shape: a solve gathers once, but
    allocating (and page-locking) 8 MB per call costs more than the exchange itself."""
    import torch
    key = (str(dev), ws, nmax, k)
    b = _bufs.get(key)
    if b is None:
        send = torch.zeros((nmax, k), dtype=torch.float64, device=dev)
        recv = torch.empty((ws * nmax, k), dtype=torch.float64, device=dev)
        pin = dev.type == "cuda"
        h_send = torch.zeros((nmax, k), dtype=torch.float64, pin_memory=pin)
        h_recv = torch.empty((ws * nmax, k), dtype=torch.float64, pin_memory=pin)
        b = _bufs[key] = (send, recv, h_send, h_recv)
    return b


def gather_records(local, n_total=None, copy=True):
    """All-gather a (n_local, k) float64 array of per-unit records from every rank, in rank order (= unit order for
    `shard_range` blocks).  ONE collective (all_gather_into_tensor into a preallocated buffer) and one copy each way
    through pinned memory; with `n_total` given the per-rank counts follow from shard_range and nothing else is
    exchanged.  Ranks may own different counts (blocks are padded to the largest).
    copy=False returns a view of the pinned landing buffer when every rank owns the same count -- valid until the next
    gather of the same shape; the default hands out fresh memory (measured at 2 ranks x 32,768 records x 16 doubles:
    3.0 ms, most of it first-touch page faults of that 8 MB copy)."""
    import torch
    local = np.ascontiguousarray(local, dtype=np.float64)
    if local.ndim == 1:
        local = local[:, None]
    d = _dist()
    if d is None:
        return local.copy()
    dev = _device_for_backend()
    rank, ws = d.get_rank(), d.get_world_size()
    k = local.shape[1]
    if n_total is not None:
        counts = [shard_range(n_total, r, ws) for r in range(ws)]
        counts = [hi - lo for lo, hi in counts]
        assert counts[rank] == local.shape[0], (counts, rank, local.shape)
    else:
        c = torch.zeros(ws, dtype=torch.int64, device=dev)
        c[rank] = local.shape[0]
        d.all_reduce(c)
        counts = [int(v) for v in c.cpu()]
    nmax = max(counts) if counts else 0
    if nmax == 0:
        return np.zeros((0, k))
    send, recv, h_send, h_recv = _buffers(dev, ws, nmax, k)
    n = local.shape[0]
    h_send[:n].copy_(torch.from_numpy(local))
    send.copy_(h_send, non_blocking=True)
    d.all_gather_into_tensor(recv, send)
    h_recv.copy_(recv, non_blocking=True)
    if dev.type == "cuda":
        torch.cuda.current_stream().synchronize()
    allr = h_recv.numpy().reshape(ws, nmax, k)
    if all(cn == nmax for cn in counts):
        res = allr.reshape(ws * nmax, k)
        if copy:
            res = res.copy()
    else:
        res = np.concatenate([allr[r, :cn] for r, cn in enumerate(counts)], axis=0)
    if n_total is not None:
        assert res.shape[0] == n_total, (res.shape, n_total)
    return res


class ResultBoards:
    """The exchange of the compact K2 results as peer stores over NVLink (include/sio_abi.h "result boards"): every rank
    owns a board of `world` slots; K2 calls made with SIO_PUSH_BOARDS store this rank's (dmin, argmin, n_under) records
    into slot `rank` of every rank's board.  After the ranks' streams have drained and a barrier, `read()` returns every
    rank's latest results from the LOCAL board -- no collective launch in between.  Single process: one slot."""

    def __init__(self, ctx, B):
        self.ctx, self.B = ctx, int(B)
        self.rank, self.world = world()
        self.slot_bytes = ((20 * self.B + 63) // 64) * 64
        self.local, handle = ctx.board_create(self.slot_bytes * self.world)
        self.ptrs = [None] * self.world
        self.ptrs[self.rank] = self.local
        d = _dist()
        if d is not None and self.world > 1:
            handles = [None] * self.world
            d.all_gather_object(handles, handle)
            for r, h in enumerate(handles):
                if r != self.rank:
                    self.ptrs[r] = ctx.board_open(h)
        ctx.set_boards(self.rank, self.ptrs, self.slot_bytes)

    def read(self):
        return self.ctx.board_read(self.world, self.slot_bytes, self.B)

    def close(self, collective=True):
        """collective=False: this rank alone gives its boards up (another rank failed to set its own up), no barriers."""
        if self.local is None:
            return
        self.ctx.set_boards(0, [], 0)
        if collective:
            barrier()                              # nobody may still be storing into a board that is about to go
        for r, p in enumerate(self.ptrs):
            if r != self.rank and p is not None:
                self.ctx.board_close(p, False)
        if collective:
            barrier()
        self.ctx.board_close(self.local, True)
        self.local = None


def argmin_allreduce(value, payload):
    """Global minimum over ranks of (value, payload): returns the winning (value, payload) on every
    rank.  `payload` is a fixed-length sequence of floats (ids, time, point ...).  A rank with
    nothing to offer passes value = +inf.  Ties on value break on the payload, lexicographically,
    so every rank picks the same winner."""
    rec = np.concatenate([[float(value)], np.asarray(payload, dtype=np.float64)])
    allrec = gather_records(rec[None, :], n_total=world()[1])
    order = np.lexsort(tuple(allrec[:, c] for c in range(allrec.shape[1] - 1, -1, -1)))
    best = allrec[order[0]]
    return float(best[0]), best[1:]


def barrier():
    d = _dist()
    if d is not None:
        d.barrier()
