"""One schedule() of the default workload with the host-side debug lines (REVS_DEBUG_HOST=1 REVS_DEBUG_E2E=1)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import revs_admm_b200 as R
K = int(sys.argv[1]) if len(sys.argv) > 1 else 3
trees, hm, cost, sizes, T = bench.make_rank_problem("synthetic-refshape-125k-homes-per-gpu-x96", 0)
H = sum(sizes)
keep, hm_p, out_p = [], {}, {}
for k, v in hm.items():
    hm_p[k], t = bench.pinned_like(v); keep.append(t)
for k, shape, dt in (("P_sch", (H, T), np.float64), ("mask", (H, (T + 63) // 64), np.uint64), ("diff", (15, H), np.float64)):
    out_p[k], t = bench.pinned_like(np.empty(shape, dtype=dt)); keep.append(t)
s = R.PipelinedSolver(sizes, T, pipelines=K)
for i in range(3):
    print("---- schedule", i, file=sys.stderr, flush=True)
    s.schedule(trees, hm_p, cost, out=out_p, compact=True, **bench.ADMM)
s.close()
