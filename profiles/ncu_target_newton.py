"""Target of the ncu capture of tree_newton_kernel: config 3 (one 10k-home radial zone x 96 hours), host-driven loop.
  ncu --set full --clock-control none --import-source on -k regex:tree_newton -s 20 -c 1 -o gpurun_out/r02_s3_newton python profiles/ncu_target_newton.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import revs_admm_b200 as R
trees, hm, cost, sizes, T = bench.make_rank_problem("synthetic-radial-10k-homes-x96", 0, split=False)
with R.Solver(sizes, T) as s:
    s.set_option("graph", 0)
    s.set_option("overlap_home", 0)
    s.set_feeder_trees(trees); s.set_homes(**hm); s.set_tariff(cost)
    for rep in range(2):
        s.solve_admm(**bench.ADMM)
    st = s.stats()
    print("ms", st["total_ms"], "launches", st["kernel_launches"])
