"""Wall time of every ADMM iteration (host sync per iteration, single pipeline) and its QP statistics."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import revs_admm_b200 as R
trees, hm, cost, sizes, T = bench.make_rank_problem("synthetic-multifeeder-125k-homes-per-gpu-x96", 0)
with R.Solver(sizes, T) as s:
    s.set_feeder_trees(trees); s.set_homes(**hm); s.set_tariff(cost)
    for rep in range(3):
        s.admm_begin(**bench.ADMM)
        ts, prev = [], dict(qp_columns=0, qp_newton_iterations=0, qp_outer_iterations=0, qp_ms=0.0, qp_warp_ms=0.0)
        rows = []
        for k in range(bench.ADMM["iter_max"]):
            t0 = time.perf_counter(); s.admm_step(); dt = (time.perf_counter() - t0) * 1e3
            st = s.stats()
            rows.append((k, round(dt, 3), st["qp_columns"] - prev["qp_columns"], st["qp_newton_iterations"] - prev["qp_newton_iterations"],
                         st["qp_outer_iterations"] - prev["qp_outer_iterations"], round(st["qp_warp_ms"] - prev["qp_warp_ms"], 3), st["max_working_set"]))
            prev = {k2: st[k2] for k2 in prev}
    print("iter  ms  qp_columns  newton_its  rounds  warp_ms  max_ws")
    for r in rows: print(*r)
    print("sum ms", round(sum(r[1] for r in rows), 2))
