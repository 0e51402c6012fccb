"""Dump the operator target z of ADMM iteration `it` of the 10k-home radial zone (for the host model of the kernel)."""
import os, sys
os.environ["REVS_DEBUG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import revs_admm_b200 as R
from revs_admm_b200.feeder import population
it = int(sys.argv[1]) if len(sys.argv) > 1 else 3
trees, hm, cost, sizes, T = population("radial10k", 1, seed=0)
kw = dict(kappa=5.0, vset=1.03, vlow=0.95, vhigh=1.05)
with R.Solver(sizes, T) as s:
    s.set_feeder_trees(trees)
    s.set_homes(**hm)
    s.set_tariff(cost)
    done = s.solve_admm(iter_max=it, **kw)
    out = s.results(done)
    p_est, gamma = s.estimate()
    z = (p_est + out["P_sch"]) / 2.0 - gamma / kw["kappa"]
    print("---- utility_step from these iterates, cold multipliers", file=sys.stderr)
    g, lam = s.utility_step(p_est, out["P_sch"], gamma, **kw)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "newton_z_it%d.npz" % it), z=z.astype(np.float64), g=g, nact=(lam > 0).sum(axis=0))
