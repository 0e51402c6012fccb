"""Experiment: K independent solver pipelines on one GPU (zones split K ways, one host thread each)."""
import os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import revs_admm_b200 as R
import torch

wl = "synthetic-multifeeder-125k-homes-per-gpu-x96"
trees, hm, cost, sizes, T = bench.make_rank_problem(wl, 0)
off = np.concatenate([[0], np.cumsum(sizes)])
for K in (1, 2, 3, 4):
    nz = len(sizes)
    cuts = [nz * k // K for k in range(K + 1)]
    solvers = []
    for k in range(K):
        a, b = cuts[k], cuts[k + 1]
        s = R.Solver(sizes[a:b], T)
        s.set_feeder_trees(trees[a:b])
        s.set_homes(**{kk: np.ascontiguousarray(v[off[a]:off[b]]) for kk, v in hm.items()})
        s.set_tariff(cost)
        solvers.append(s)
    def work(s):
        s.solve_admm(**bench.ADMM)
    def step():
        th = [threading.Thread(target=work, args=(s,)) for s in solvers]
        for t in th: t.start()
        for t in th: t.join()
    for _ in range(3): step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n): step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n * 1e3
    print(f"K={K}: {dt:.2f} ms per schedule of all {sum(sizes)} homes; per-pipeline device spans {[round(s.stats()['total_ms'],2) for s in solvers]}", flush=True)
    for s in solvers: s.close()
