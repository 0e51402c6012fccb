"""Per-kernel-family device time of one schedule (host-driven loop, CUDA-event spans; REVS_DEBUG_HOST=1 prints them).
    REVS_DEBUG_HOST=1 [REVS_LIB=...] python profiles/dbg_spans.py [overlap_home]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import revs_admm_b200 as R
trees, hm, cost, sizes, T = bench.make_rank_problem("synthetic-refshape-125k-homes-per-gpu-x96", 0)
with R.Solver(sizes, T) as s:
    s.set_option("graph", 0)
    s.set_option("overlap_home", int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    s.set_feeder_trees(trees); s.set_homes(**hm); s.set_tariff(cost)
    for _ in range(3):
        s.solve_admm(**bench.ADMM)
    st = s.stats()
    print({k: round(v, 3) for k, v in st.items() if k.endswith("_ms")})
