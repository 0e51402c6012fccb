"""Solver-independent certificates for the CUDA operator QP, and oracle comparisons at the
benchmark's own population.

The oracle's Utility QP and the kernels share an algorithm, so oracle == GPU proves the port, not
the maths.  Here the output of ``revs_utility_step`` (g and the voltage-row multipliers) is checked
against the KKT conditions of the reference's program (lpsolver.py:163-238: min kappa/2 |g|^2 +
<g, Gamma - kappa/2 (P_est + P_sch)>  s.t.  R g <= u, g >= 0) with nothing but numpy matmuls:
    g >= 0,  R g <= u,  lam >= 0,  g = [z - R lam]_+,  lam_i > 0  =>  (R g)_i = u.
A point with these properties is THE solution of the strictly convex QP, whatever found it.
"""
import os

import numpy as np
import pytest

import revs_oracle as O

pytestmark = pytest.mark.gpu

KW = dict(kappa=5.0, vset=1.03, vlow=0.95, vhigh=1.05)
U = KW["vhigh"] ** 2 - KW["vset"] ** 2


def kkt_check(Rblocks, z, g, lam, u, what=""):
    """Assert the KKT conditions zone by zone; returns the largest multiplier and the number of active rows."""
    off, worst_lam, n_active = 0, 0.0, 0
    for R in Rblocks:
        n = R.shape[0]
        sl = slice(off, off + n)
        gz, lz, zz = g[sl], lam[sl], z[sl]
        v = R @ gz
        assert gz.min() >= 0.0, what
        assert (v - u).max() <= 1e-9, (what, (v - u).max())
        assert lz.min() >= 0.0, what
        stat = np.abs(gz - np.maximum(zz - R @ lz, 0.0)).max()
        assert stat <= 1e-9, (what, stat)
        act = lz > 0
        if act.any():
            assert np.abs(v[act] - u).max() <= 1e-9, (what, np.abs(v[act] - u).max())
            worst_lam = max(worst_lam, lz.max())
            n_active += int(act.sum())
        off += n
    return worst_lam, n_active


def _oracle_history(Rb, hm, cost, iters):
    return O.solve_ADMM_arrays(Rb, load=hm["load"], cost=cost, ev_mask=hm["has_ev"].astype(bool), rating=hm["rating"],
                               capacity=hm["capacity"], initial=hm["initial"], start=hm["start"], end=hm["end"],
                               kappa=KW["kappa"], iter_max=iters, vset=KW["vset"], vlow=KW["vlow"], vhigh=KW["vhigh"],
                               return_history=True)


def test_kkt_certificate_real_feeder_every_iteration(gpu_lib, case121144):
    """The operator QP of every one of the 15 ADMM iterations of the reference's own case."""
    from revs_admm_b200.feeder import split_zones, tree_from_graph
    from revs_admm_b200.lpsolver import _home_arrays
    tree = tree_from_graph(case121144["dist"])
    zones = split_zones(tree)
    perm = np.concatenate([h for _, h in zones])
    hm = _home_arrays(case121144["homes"], [tree.res_ids[i] for i in perm])
    cost = np.asarray(case121144["tariff"], float)
    Rb = [z.rmat()[np.ix_(z.res_node, z.res_node)] for z, _ in zones]
    hist = _oracle_history(Rb, hm, cost, 15)["history"]
    H, T = hm["load"].shape
    states = [(np.zeros((H, T)),) * 3] + [h for h in hist[:-1]]      # inputs of iterations 1..15
    total_active = 0
    with gpu_lib.Solver([z.n_res for z, _ in zones], T) as s:
        s.set_feeder_trees([z for z, _ in zones])
        lam = None
        for k, (pe, ps, gm) in enumerate(states):
            g, lam = s.utility_step(pe, ps, gm, lam0=lam, **KW)            # warm start, as the loop does
            z = (pe + ps) / 2.0 - gm / KW["kappa"]
            _, na = kkt_check(Rb, z, g, lam, U, f"iteration {k + 1}")
            total_active += na
            assert np.abs(g - hist[k][0]).max() <= 1e-6                     # and the oracle agrees
    assert total_active > 1000          # the limits bind: hundreds of active rows per iteration


@pytest.mark.parametrize("name,first", [("refshape", 0), ("refshape", 57), ("laterals", 3)])
def test_kkt_certificate_bench_population(gpu_lib, name, first):
    """Feeders of the benchmark populations (bench.py), states of ADMM iterations 2, 4, 9 and 15."""
    from revs_admm_b200.feeder import population
    trees, hm, cost, sizes, T = population(name, 3, seed=0, first_feeder=first)
    Rb = [t.rmat()[np.ix_(t.res_node, t.res_node)] for t in trees]
    hist = _oracle_history(Rb, hm, cost, 15)["history"]
    with gpu_lib.Solver(sizes, T) as s:
        s.set_feeder_trees(trees)
        for k in (1, 3, 8, 14):
            pe, ps, gm = hist[k - 1]
            g, lam = s.utility_step(pe, ps, gm, **KW)                       # cold start
            z = (pe + ps) / 2.0 - gm / KW["kappa"]
            lmax, na = kkt_check(Rb, z, g, lam, U, f"{name} iteration {k + 1}")
            assert na > 0 and lmax > 0.0


@pytest.mark.parametrize("name", ["refshape", "laterals"])
def test_bench_population_first_feeders_match_oracle(gpu_lib, name):
    """Oracle vs GPU on the benchmark workload itself: the first 4 feeders (4000 homes, 96 steps),
    the reference's 15 iterations -- same asserts as test_admm_multi_feeder_synthetic_96."""
    from revs_admm_b200.feeder import population
    trees, hm, cost, sizes, T = population(name, 4, seed=0)
    Rb = [t.rmat()[np.ix_(t.res_node, t.res_node)] for t in trees]
    ref = _oracle_history(Rb, hm, cost, 15)
    with gpu_lib.Solver(sizes, T) as s:
        s.set_feeder_trees(trees)
        s.set_homes(**hm)
        s.set_tariff(cost)
        done = s.solve_admm(iter_max=15, **KW)
        out = s.results(done)
        P_est, Gam = s.estimate()
        st = s.stats()
    assert done == 15
    assert np.array_equal(out["P_ev"], ref["P_ev"])
    assert np.abs(out["P_sch"] - ref["P_sch"]).max() <= 1e-4
    assert np.abs(P_est - ref["P_est"]).max() <= 1e-4
    assert np.abs(Gam - ref["Gamma"]).max() <= 1e-4
    assert np.abs(out["diff"] - ref["diff"]).max() <= 1e-7
    assert np.allclose(out["SOC"], ref["SOC"], atol=1e-12)
    assert st["max_working_set"] >= 2
    # the final estimate is voltage-feasible in every zone (R P_est <= u), whatever the schedule does
    off = 0
    for R in Rb:
        n = R.shape[0]
        assert (R @ P_est[off:off + n]).max() <= U + 1e-9
        off += n


def test_iteration2_of_the_reference_file_through_the_gpu(gpu_lib, golden):
    """The tightest pin of the operator QP to reference-held vectors (tests/golden/reconstruct_ties.py):
    with the reconstructed iteration-1 choices, diff[2] = ||proj(P_sch[1]) - P_sch[1]|| / T of the 267 EV
    homes, computed by the CUDA QP, is within 2e-4 (median) of the reference's own file."""
    import sys
    from conftest import GOLDEN
    sys.path.insert(0, GOLDEN)
    import reconstruct_ties as RT
    g, arr, T, cost, evrow, (ztree, zhomes) = RT.load_case()
    ch = np.load(os.path.join(GOLDEN, "tie_choices_iter1_121144_com2.npz"))
    loc = {int(i): j for j, i in enumerate(zhomes)}
    ev = np.array([loc[int(i)] for i in evrow])
    P1 = arr["load"][zhomes].copy()
    for gi, j in enumerate(ev):
        P1[j, ch["hours"][gi]] += 4.8
    zero = np.zeros_like(P1)
    gamma1 = -(KW["kappa"] / 2.0) * P1                  # Gamma[1] = 0 + kappa/2 (P_est[1] - P_sch[1]), P_est[1] = 0
    with gpu_lib.Solver([ztree.n_res], T) as s:
        s.set_feeder_tree(0, ztree.parent, ztree.r, ztree.res_node)
        pe2, lam = s.utility_step(zero, P1, gamma1, **KW)
    R = O.rmat_from_tree(ztree.parent, ztree.r)[np.ix_(ztree.res_node, ztree.res_node)]
    kkt_check([R], P1, pe2, lam, U, "iteration 2 of the reference run")
    d2 = np.sqrt(((pe2 - P1) ** 2).sum(axis=1)) / T
    err = np.abs(d2[ev] - golden["distributed_diff"][:, 1])
    assert np.median(err) < 2e-4 and err.max() < 2e-2, (np.median(err), err.max())
