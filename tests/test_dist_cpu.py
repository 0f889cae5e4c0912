"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: contiguous sharding of independent
units, the end-of-solve gather of per-problem records (C5) and the (value, payload) arg-min
exchange over time blocks (C4) -- SURVEY.md 8e."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_everything_once():
    from semiinfiniteoptimization_b200.dist import shard_range
    for n in (0, 1, 7, 64, 65536, 65537):
        for ws in (1, 2, 3, 8):
            blocks = [shard_range(n, r, ws) for r in range(ws)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_fallbacks():
    from semiinfiniteoptimization_b200 import dist
    assert dist.world() == (0, 1)
    a = np.arange(12.0).reshape(4, 3)
    assert np.array_equal(dist.gather_records(a), a)
    v, p = dist.argmin_allreduce(0.5, [1, 2, 3.5])
    assert v == 0.5 and list(p) == [1, 2, 3.5]


WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    sys.path.insert(0, %r)
    import torch.distributed as dist
    from semiinfiniteoptimization_b200 import dist as sd
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"],
                            rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    rank, ws = sd.world()
    n = 11                                    # ragged: 6 + 5
    lo, hi = sd.shard_range(n)
    local = np.stack([np.arange(lo, hi) * 1.0, np.arange(lo, hi) ** 2.0], axis=1)
    allrec = sd.gather_records(local, n_total=n)
    assert np.array_equal(allrec[:, 0], np.arange(n)) and np.array_equal(allrec[:, 1], np.arange(n) ** 2.0)
    # equal shards, result handed out as a view of the landing buffer (copy=False): right values, and the next gather of
    # the same shape overwrites it -- the documented lifetime
    mine = np.full((4, 3), float(rank))
    view = sd.gather_records(mine, n_total=8, copy=False)
    assert view.shape == (8, 3) and np.array_equal(view[:4], np.zeros((4, 3))) and np.array_equal(view[4:], np.ones((4, 3)))
    kept = view.copy()
    again = sd.gather_records(mine + 10.0, n_total=8, copy=False)
    assert np.array_equal(again, kept + 10.0) and np.array_equal(view, again)
    fresh = sd.gather_records(mine, n_total=8)
    assert np.array_equal(fresh, kept) and not np.shares_memory(fresh, again)
    # empty shard on one rank
    e = sd.gather_records(np.zeros((0, 2)) if rank == 1 else np.ones((3, 2)))
    assert e.shape == (3, 2)
    # arg-min exchange: rank 1 holds the deepest sample; ties break on the payload
    v, p = sd.argmin_allreduce([0.25, -0.125][rank], [[4, 0, 0.5], [2, 0, 0.75]][rank])
    assert v == -0.125 and list(p) == [2, 0, 0.75]
    v, p = sd.argmin_allreduce(1.0, [[3, 1], [3, 0]][rank])
    assert v == 1.0 and list(p) == [3, 0]
    v, p = sd.argmin_allreduce(float("inf") if rank == 0 else 2.0, [0, 0] if rank == 0 else [7, 7])
    assert v == 2.0 and list(p) == [7, 7]
    sd.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""") % ROOT


def test_gloo_world2_gather_and_argmin(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", PORT=port, MASTER_ADDR="127.0.0.1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, o)
        assert "ok" in o
