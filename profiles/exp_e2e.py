"""Where the end-to-end leg spends its time: upload / solve / results, pinned host buffers."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import revs_admm_b200 as R
import torch
wl = "synthetic-multifeeder-125k-homes-per-gpu-x96"
trees, hm, cost, sizes, T = bench.make_rank_problem(wl, 0)
H = sum(sizes)
keep = []
hm_p = {}
for k, v in hm.items():
    hm_p[k], t = bench.pinned_like(v); keep.append(t)
out_p = {}
for k, shape in (("P_sch", (H, T)), ("P_ev", (H, T)), ("SOC", (H, T + 1)), ("diff", (15, H))):
    out_p[k], t = bench.pinned_like(np.empty(shape)); keep.append(t)
for K in (1, 3):
    s = R.PipelinedSolver(sizes, T, pipelines=K)
    def once():
        t0 = time.perf_counter(); s.set_feeder_trees(trees)
        t1 = time.perf_counter(); s.set_homes(**hm_p); s.set_tariff(cost)
        t2 = time.perf_counter(); s.solve_admm(**bench.ADMM)
        t3 = time.perf_counter(); s.results(out=out_p)
        t4 = time.perf_counter()
        return [1e3 * (b - a) for a, b in ((t0, t1), (t1, t2), (t2, t3), (t3, t4))]
    once(); once()
    r = np.mean([once() for _ in range(4)], axis=0)
    print(f"K={K}: trees {r[0]:.2f} ms, homes+tariff {r[1]:.2f} ms, solve {r[2]:.2f} ms, results {r[3]:.2f} ms, total {r.sum():.2f} ms", flush=True)
    s.close()
    s = R.PipelinedSolver(sizes, T, pipelines=K)
    for _ in range(2): s.schedule(trees, hm_p, cost, out=out_p, **bench.ADMM)
    t0 = time.perf_counter()
    for _ in range(4): s.schedule(trees, hm_p, cost, out=out_p, **bench.ADMM)
    print(f"K={K}: schedule() {1e3 * (time.perf_counter() - t0) / 4:.2f} ms per step (upload -> solve -> download per pipeline, overlapped)", flush=True)
    s.close()
