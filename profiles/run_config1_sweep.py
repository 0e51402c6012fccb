"""BASELINE.json config 2 ("121144 feeder, all five communities at 30/60/90 % adoption, 3600 and
4800 W"): the reference's feeder and input files through this package's REVS fixture and
lpsolver.solve_ADMM on one GPU, every point checked against the CPU oracle.

    python profiles/run_config1_sweep.py [out.json]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import revs_oracle as O  # noqa: E402  (checker only)
from revs_admm_b200.lpsolver import solve_ADMM  # noqa: E402
from revs_admm_b200.revs_fixture import REVS  # noqa: E402

KW = dict(kappa=5.0, iter_max=15, vset=1.03, vlow=0.95, vhigh=1.05)      # revs_config.yaml


def main():
    rows = []
    inp = os.path.join(ROOT, "tests", "golden", "input")
    for com in (1, 2, 3, 4, 5):
        fx = REVS(data_path=inp, out_path="/tmp/o", grb_path="/tmp/g", fig_path="/tmp/f", comunityID=com, optimizer_mode="distributed")
        for adoption in (30, 60, 90):
            for rating in (3600, 4800):
                tariff, homes, dist, saved = fx.read_inputs(adoption=adoption, rating=rating, seed=1234)
                solve_ADMM(homes, dist, tariff, None, **KW)                       # warm-up (library load, first launch)
                t0 = time.perf_counter()
                diff, P, S, C, st = solve_ADMM(homes, dist, tariff, None, return_stats=True, **KW)
                wall = time.perf_counter() - t0
                t0 = time.perf_counter()
                do, Po, So, Co = O.solve_ADMM(homes, dist, tariff, None, **KW)
                cpu = time.perf_counter() - t0
                c = np.asarray(tariff)
                rows.append(dict(community=com, adoption=adoption, rating_w=rating, ev_homes=len(saved["ev_homes"]), homes=len(homes),
                                 device_ms=round(st["total_ms"], 3), call_ms=round(1e3 * wall, 2), oracle_s=round(cpu, 2),
                                 max_working_set=st["max_working_set"],
                                 max_abs_dP_kw=float(max(np.abs(P[h] - Po[h]).max() for h in Po)),
                                 charging_hours_identical=bool(all(np.array_equal(S[h], So[h]) for h in So)),
                                 homes_with_other_hours=int(sum(not np.array_equal(S[h], So[h]) for h in So)),
                                 same_number_of_hours=bool(all((S[h] > 0).sum() == (So[h] > 0).sum() for h in So)),
                                 max_abs_ddiff=float(max(abs(diff[k][h] - do[k][h]) for k in do for h in do[k])),
                                 cost=float(sum(c @ P[h] for h in P))))
                print(json.dumps(rows[-1]), flush=True)
    # charging hours must be identical in all 30 points (hour costs are compared on a 2^-20 grid)
    ok = all(r["max_abs_ddiff"] <= 1e-7 and r["charging_hours_identical"] and r["max_abs_dP_kw"] <= 1e-4 for r in rows)
    print("SWEEP", "OK" if ok else "FAILED", len(rows), "points")
    if len(sys.argv) > 1:
        json.dump(rows, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
