import os, sys
sys.path.insert(0, "/root/repo")
import bench
import revs_admm_b200 as R
trees, hm, cost, sizes, T = bench.make_rank_problem("synthetic-refshape-125k-homes-per-gpu-x96", 0)
with R.Solver(sizes, T) as s:
    s.set_feeder_trees(trees); s.set_homes(**hm); s.set_tariff(cost)
    s.solve_admm(**bench.ADMM)
    print(s.stats())
