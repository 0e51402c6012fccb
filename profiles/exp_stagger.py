"""Experiment: staggered start of the pipelines' captured loops (pipeline k starts k * d ms after pipeline 0)."""
import os, sys, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
import revs_admm_b200 as R
trees, hm, cost, sizes, T = bench.make_rank_problem("synthetic-refshape-125k-homes-per-gpu-x96", 0)
s = R.PipelinedSolver(sizes, T, pipelines=4)
s.set_feeder_trees(trees); s.set_homes(**hm); s.set_tariff(cost)
def step(d):
    t0 = time.perf_counter()
    def work(k):
        while (time.perf_counter() - t0) * 1e3 < k * d:
            pass
        s.parts[k].solve_admm(**bench.ADMM)
    th = [threading.Thread(target=work, args=(k,)) for k in range(len(s.parts))]
    for t in th: t.start()
    for t in th: t.join()
for d in (0.0, 0.2, 0.5, 1.0, 2.0, 0.0, 0.5):
    for _ in range(2): step(d)
    torch.cuda.synchronize()
    res = []
    for rep in range(3):
        t0 = time.perf_counter()
        for _ in range(4): step(d)
        torch.cuda.synchronize()
        res.append((time.perf_counter() - t0) * 1e3 / 4)
    print(f"stagger {d:.1f} ms: " + " ".join(f"{x:.2f}" for x in res) + "  spans " + str([round(p.stats()['total_ms'], 2) for p in s.parts]), flush=True)
s.close()
