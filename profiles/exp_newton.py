"""Per-iteration statistics of the tree-Newton path on the 10k-home radial zone (REVS_DEBUG lines on stderr)."""
import os, sys, time
os.environ["REVS_DEBUG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import revs_admm_b200 as R
from revs_admm_b200.feeder import population
trees, hm, cost, sizes, T = population("radial10k", 1, seed=0)
with R.Solver(sizes, T) as s:
    s.set_feeder_trees(trees)
    s.set_homes(**hm)
    s.set_tariff(cost)
    for rep in range(2):
        t0 = time.perf_counter()
        done = s.solve_admm(kappa=5.0, iter_max=15, vset=1.03, vlow=0.95, vhigh=1.05)
        print("solve", rep, "host ms", (time.perf_counter() - t0) * 1e3, s.stats()["total_ms"], file=sys.stderr)
