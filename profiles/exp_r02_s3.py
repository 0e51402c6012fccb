"""Round 2, session 3 experiment: pipelines per GPU with captured loops (device-resident and end-to-end), one
pipeline's share run alone (critical path vs throughput), phase marks of schedule().

    REVS_DEBUG_E2E=1 python profiles/exp_r02_s3.py [K ...]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import revs_admm_b200 as R  # noqa: E402
import torch  # noqa: E402

wl = "synthetic-refshape-125k-homes-per-gpu-x96"
Ks = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4, 6, 8]
trees, hm, cost, sizes, T = bench.make_rank_problem(wl, 0)
H = sum(sizes)
keep, hm_p, out_p = [], {}, {}
for k, v in hm.items():
    hm_p[k], t = bench.pinned_like(v)
    keep.append(t)
for k, shape, dt in (("P_sch", (H, T), np.float64), ("mask", (H, (T + 63) // 64), np.uint64), ("diff", (15, H), np.float64)):
    out_p[k], t = bench.pinned_like(np.empty(shape, dtype=dt))
    keep.append(t)


def timed(fn, n=4, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / n
    e1.record()
    torch.cuda.synchronize()
    return max(wall, e0.elapsed_time(e1) / n)


trace = os.environ.pop("REVS_DEBUG_E2E", None)
for K in Ks:
    s = R.PipelinedSolver(sizes, T, pipelines=K)
    s.set_feeder_trees(trees)
    s.set_homes(**hm_p)
    s.set_tariff(cost)
    dev = timed(lambda: s.solve_admm(**bench.ADMM))
    spans = [round(p.stats()["total_ms"], 2) for p in s.parts]
    alone = None
    if K > 1:
        alone = [round(timed(lambda: s.parts[k].solve_admm(**bench.ADMM), n=2, warm=1), 2) for k in (0, K - 1)]
    e2e = timed(lambda: s.schedule(trees, hm_p, cost, out=out_p, compact=True, **bench.ADMM))
    print(f"K={K}: device {dev:.2f} ms  e2e {e2e:.2f} ms  per-pipeline device spans {spans}  first/last pipeline alone {alone}", flush=True)
    if trace and K in (3, 4, 6):
        os.environ["REVS_DEBUG_E2E"] = "1"
        s.schedule(trees, hm_p, cost, out=out_p, compact=True, **bench.ADMM)
        os.environ.pop("REVS_DEBUG_E2E")
    s.close()
