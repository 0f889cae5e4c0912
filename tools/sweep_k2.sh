#!/bin/bash
# times the bench workload's K2 bulk kernel for several transforms-per-CTA settings (tuning aid)
for c in 4 8 16 32; do
  SIO_MIN_CHUNK=$c python - <<'PY'
import os, sys
sys.path.insert(0, '.')
import numpy as np, torch, bench
from semiinfiniteoptimization_b200 import _abi
vg, cloud, T_rel, bound = bench.build_workload(0)
ctx = _abi.Context(0); g = ctx.grid_create(vg.values, vg.bmin, vg.bmax); c = ctx.cloud_create(cloud)
B = len(T_rel); dT = torch.from_numpy(T_rel).cuda(); db = torch.from_numpy(bound).cuda()
dd = torch.empty(B, dtype=torch.float64, device="cuda"); da = torch.empty(B, dtype=torch.int64, device="cuda"); dn = torch.empty(B, dtype=torch.int32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ms = []
for i in range(30):
    flush.zero_(); torch.cuda.synchronize()
    ctx.min_distance_dev([g], c, dT.data_ptr(), db.data_ptr(), B, _abi.SIO_F32, dd.data_ptr(), da.data_ptr(), d_n_under_ptr=dn.data_ptr())
    ctx.sync(); ms.append(ctx.last_bulk_ms())
print("chunk", os.environ["SIO_MIN_CHUNK"], "bulk ms median %.4f min %.4f" % (np.median(ms[5:]), min(ms[5:])), float(dd[0]), int(da[0]))
PY
done
