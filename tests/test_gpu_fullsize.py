"""Full-size checks (BASELINE.json config: 125k homes x 96 steps per GPU) through size-independent
properties of the path -- the CPU oracle cannot run this size in seconds.  See
profiles/check_properties.py for what is checked."""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tool():
    spec = importlib.util.spec_from_file_location("check_properties", os.path.join(ROOT, "profiles", "check_properties.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("seed", [0, 5])
def test_projection_is_feasible_at_full_size(gpu_lib, seed):
    cp = _tool()
    trees, pe, gm, res, st = cp.run("synthetic-multifeeder-125k-homes-per-gpu-x96", seed)
    assert pe.min() >= 0.0
    assert cp.feasibility(trees, pe, zones=120) <= 1e-9           # R P_est <= u in every sampled zone
    assert st["max_working_set"] >= 8                              # the limits bind
    # schedule bookkeeping: P_sch = load + P_ev, P_ev in {0, rating}, SOC recursion and final SOC window
    assert np.all((res["P_ev"] == 0.0) | (res["P_ev"] > 0.0))
    soc_end = res["SOC"][:, -1]
    ev = res["P_ev"].max(axis=1) > 0
    assert np.all(soc_end[ev] >= 0.9 - 1e-9) and np.all(soc_end <= 1.0 + 1e-9)


def test_modes_agree_bitwise_at_full_size(gpu_lib):
    cp = _tool()
    wl = "synthetic-multifeeder-125k-homes-per-gpu-x96"
    _, pe, gm, res, _ = cp.run(wl, 0)
    _, pe2, gm2, res2, _ = cp.run(wl, 0, stepwise=True)
    _, pe3, gm3, res3, _ = cp.run(wl, 0, screen=0)
    for a, b in ((pe, pe2), (gm, gm2), (res["P_sch"], res2["P_sch"]), (pe, pe3), (gm, gm3), (res["P_sch"], res3["P_sch"])):
        assert np.array_equal(a, b)


def test_radial_10k_single_zone_kkt(gpu_lib):
    """BASELINE.json config 3: ONE radial feeder of 10 000 residences (a single dense voltage zone), 96 steps, limits
    that bind on hundreds of rows per hour.  No dense matrix exists for it anywhere (device or host): the operator
    estimate after a few ADMM iterations is checked through its KKT certificate with R applied on the tree
    (FeederTree.drop = R_res @ x in O(n), lpsolver.py:192)."""
    from revs_admm_b200.feeder import population
    trees, hm, cost, sizes, T = population("radial10k", 1, seed=0)
    assert sizes == [10000] and T == 96
    tr = trees[0]
    kw = dict(kappa=5.0, vset=1.03, vlow=0.95, vhigh=1.05)
    u = kw["vhigh"] ** 2 - kw["vset"] ** 2
    with gpu_lib.Solver(sizes, T) as s:
        s.set_feeder_trees(trees)
        s.set_homes(**hm)
        s.set_tariff(cost)
        done = s.solve_admm(iter_max=4, **kw)
        out = s.results(done)
        p_est, gamma = s.estimate()
        st = s.stats()
        # one more operator step from the final iterates, with its multipliers
        g, lam = s.utility_step(p_est, out["P_sch"], gamma, **kw)
    assert st["max_working_set"] > 128                      # far beyond what the dense kernels hold
    z = (p_est + out["P_sch"]) / 2.0 - gamma / kw["kappa"]
    assert g.min() >= 0.0 and lam.min() >= 0.0
    v = tr.drop(g)
    assert (v - u).max() <= 1e-9
    assert np.abs(g - np.maximum(z - tr.drop(lam), 0.0)).max() <= 1e-8
    assert np.abs(lam * (u - v)).max() <= 1e-7
    assert (lam > 0).sum(axis=0).max() > 128
    assert (tr.drop(p_est) - u).max() <= 1e-9               # the estimate of the loop itself is voltage-feasible
