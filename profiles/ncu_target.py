"""Target of the ncu captures: host-driven (graph = 0) ADMM solves of one GPU's share of the default workload, single
pipeline, so that every kernel is an ordinary launch ncu can name.  Usage (see profiles/README_r02.md):

  ncu --set full --clock-control none --import-source on --launch-skip 280 --launch-count 16 -o gpurun_out/r02_full \
      python profiles/ncu_target.py [workload] [feeders]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import revs_admm_b200 as R  # noqa: E402
from revs_admm_b200.feeder import population, POPULATIONS  # noqa: E402

pop = sys.argv[1] if len(sys.argv) > 1 else "refshape"
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 125
trees, hm, cost, sizes, T = population(pop, nf, seed=0)
with R.Solver(sizes, T) as s:
    s.set_option("graph", 0)
    s.set_option("overlap_home", 0)
    s.set_feeder_trees(trees)
    s.set_homes(**hm)
    s.set_tariff(cost)
    for rep in range(2):
        done = s.solve_admm(kappa=5.0, iter_max=15, vset=1.03, vlow=0.95, vhigh=1.05)
    st = s.stats()
    print("ms", st["total_ms"], "launches", st["kernel_launches"], "rounds", st["qp_outer_iterations"])
