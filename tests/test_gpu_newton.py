"""The tree-Newton operator kernel (csrc/tree_newton.cu): large radial zones, any number of binding voltage rows.

Reference: class Utility (lpsolver.py:163-238) puts R_res g <= vhigh^2 - vset^2 over ALL residences of the graph with no
limit on how many rows bind.  Checked here against the dense CPU oracle (oracle/revs_oracle.py, same QP solved with
dense matrices), by a solver-independent KKT certificate, and against this library's own dense kernels."""
import numpy as np
import pytest

import revs_oracle as O

pytestmark = pytest.mark.gpu

U = 1.05 ** 2 - 1.03 ** 2


def _oracle(trees, hm, cost, kw):
    Rb = [O.rmat_from_tree(t.parent, t.r)[np.ix_(t.res_node, t.res_node)] for t in trees]
    return O.solve_ADMM_arrays(Rb, load=hm["load"], cost=cost, ev_mask=hm["has_ev"].astype(bool), rating=hm["rating"],
                               capacity=hm["capacity"], initial=hm["initial"], start=hm["start"], end=hm["end"], **kw)


def _gpu(lib, sizes, T, trees, hm, cost, kw, newton_min_n=None, graph=1):
    with lib.Solver(sizes, T) as s:
        if newton_min_n is not None:
            s.set_option("newton_min_n", newton_min_n)
        s.set_option("graph", graph)
        s.set_feeder_trees(trees)
        s.set_homes(**hm)
        s.set_tariff(cost)
        done = s.solve_admm(**kw)
        out = s.results(done)
        out["P_est"], out["Gamma"] = s.estimate()
        out["stats"] = s.stats()
    return out


def test_single_large_zone_many_binding_rows_kkt(gpu_lib):
    """One 2500-home radial zone (trunk + laterals: a single dense sensitivity block) under limits the base load alone
    exceeds: several hundred rows bind per hour.  The projection returned by revs_utility_step carries a KKT
    certificate (g = [z - R lam]_+, lam >= 0, R g <= u, complementarity) and equals the dense oracle's."""
    from revs_admm_b200.feeder import radial_feeder, synthetic_homes
    n, T = 2500, 4
    tr = radial_feeder(n, seed=5)
    hm = synthetic_homes(n, 96, seed=11)
    rng = np.random.default_rng(3)
    z = hm["load"][:, [0, 40, 52, 60]] + 4.8 * (rng.random((n, T)) < 0.6)
    zero = np.zeros((n, T))
    with gpu_lib.Solver([n], T) as s:
        s.set_feeder_tree(0, tr.parent, tr.r, tr.res_node)
        g, lam = s.utility_step(2.0 * z, zero, zero, kappa=5.0, vset=1.03, vlow=0.95, vhigh=1.05)   # target (p_est + p_sch) / 2 = z
        st = s.stats()
    R = O.rmat_from_tree(tr.parent, tr.r)[np.ix_(tr.res_node, tr.res_node)]
    assert (lam > 0).sum(axis=0).max() > 128, (lam > 0).sum(axis=0)
    assert st["max_working_set"] > 128
    assert g.min() >= 0.0 and lam.min() >= 0.0
    v = R @ g
    assert (v - U).max() <= 1e-9
    assert np.abs(g - np.maximum(z - R @ lam, 0.0)).max() <= 1e-9
    assert np.abs(lam * (U - v)).max() <= 1e-9
    for t in range(T):
        g0, l0, _ = O.project_voltage(z[:, t], R, U)
        assert np.abs(g[:, t] - g0).max() <= 1e-8


def test_former_overflow_case_matches_oracle(gpu_lib):
    """The 1000-home zone on which the dense kernels stop (more than 128 binding rows per hour, see
    test_gpu_edges.py::test_working_set_overflow_is_loud): given as a tree it runs on the tree-Newton path by default and
    lands on the oracle's schedule."""
    from revs_admm_b200.feeder import synthetic_feeder, synthetic_homes, synthetic_tariff
    n, T = 1000, 24
    t = synthetic_feeder(n, seed=0, laterals=5)
    hm = synthetic_homes(n, T, seed=77)
    cost = synthetic_tariff(T)
    kw = dict(kappa=5.0, iter_max=4, vset=1.03, vlow=0.95, vhigh=1.05)
    out = _gpu(gpu_lib, [n], T, [t], hm, cost, kw)
    ref = _oracle([t], hm, cost, kw)
    assert out["stats"]["max_working_set"] > 128
    assert np.array_equal(out["P_ev"], ref["P_ev"])
    assert np.abs(out["P_sch"] - ref["P_sch"]).max() <= 1e-4
    assert np.abs(out["P_est"] - ref["P_est"]).max() <= 1e-4
    assert np.abs(out["diff"] - ref["diff"]).max() <= 1e-7


@pytest.mark.parametrize("sizes,vhigh,graph", [([100, 200], 1.02, 1), ([297, 157, 257, 320, 129], 1.02, 0), ([600, 90], 1.05, 1), ([1], 1.02, 1)])
def test_newton_path_on_small_zones_matches_oracle(gpu_lib, sizes, vhigh, graph):
    """newton_min_n = 0 sends every zone through the tree-Newton kernel (captured loop and host-driven loop)."""
    from revs_admm_b200.feeder import synthetic_feeder, synthetic_homes, synthetic_tariff
    T = 24
    seed = sum(sizes)
    trees = [synthetic_feeder(n, seed=seed + i, r_secondary=1e-3, laterals=max(1, min(5, n))) for i, n in enumerate(sizes)]
    hm = synthetic_homes(sum(sizes), T, seed=seed)
    cost = synthetic_tariff(T)
    kw = dict(kappa=5.0, iter_max=5, vset=1.0, vlow=0.95, vhigh=vhigh)
    out = _gpu(gpu_lib, sizes, T, trees, hm, cost, kw, newton_min_n=0, graph=graph)
    ref = _oracle(trees, hm, cost, kw)
    assert np.array_equal(out["P_ev"], ref["P_ev"])
    assert np.abs(out["P_sch"] - ref["P_sch"]).max() <= 1e-4
    assert np.abs(out["P_est"] - ref["P_est"]).max() <= 1e-4
    assert np.abs(out["diff"] - ref["diff"]).max() <= 1e-7


def test_mixed_dense_and_newton_zones(gpu_lib):
    """Zones on both paths in one solver (threshold between their sizes) against all-dense and against the oracle."""
    from revs_admm_b200.feeder import synthetic_feeder, synthetic_homes, synthetic_tariff
    sizes, T = [90, 400, 150, 700], 24
    trees = [synthetic_feeder(n, seed=40 + i, r_secondary=1e-3) for i, n in enumerate(sizes)]
    hm = synthetic_homes(sum(sizes), T, seed=9)
    cost = synthetic_tariff(T)
    kw = dict(kappa=5.0, iter_max=5, vset=1.0, vlow=0.95, vhigh=1.03)
    mixed = _gpu(gpu_lib, sizes, T, trees, hm, cost, kw, newton_min_n=300)
    dense = _gpu(gpu_lib, sizes, T, trees, hm, cost, kw, newton_min_n=4096)
    ref = _oracle(trees, hm, cost, kw)
    for out in (mixed, dense):
        assert np.array_equal(out["P_ev"], ref["P_ev"])
        assert np.abs(out["P_est"] - ref["P_est"]).max() <= 1e-4
    assert np.abs(mixed["P_sch"] - dense["P_sch"]).max() <= 1e-6


def test_reference_feeder_unsplit_on_the_newton_path(gpu_lib, case121144):
    """The reference's own case with the whole community graph as ONE zone per connected component of the tree, every
    zone on the tree-Newton path: same schedule as the oracle run of test_gpu_parity."""
    from revs_admm_b200 import lpsolver
    from revs_admm_b200.feeder import tree_from_graph
    import revs_admm_b200 as R
    c = case121144
    tree = tree_from_graph(c["dist"])
    arr = lpsolver._home_arrays(c["homes"], list(tree.res_ids))
    cost = np.asarray(c["tariff"], dtype=float)
    kw = dict(kappa=5.0, iter_max=6, vset=1.03, vlow=0.95, vhigh=1.05)
    T = len(cost)
    with R.Solver([tree.n_res], T) as s:
        s.set_option("newton_min_n", 0)
        s.set_feeder_tree(0, tree.parent, tree.r, tree.res_node)
        s.set_homes(**arr)
        s.set_tariff(cost)
        done = s.solve_admm(**kw)
        out = s.results(done)
        st = s.stats()
    Rm = O.rmat_from_tree(tree.parent, tree.r)[np.ix_(tree.res_node, tree.res_node)]
    ref = O.solve_ADMM_arrays([Rm], load=arr["load"], cost=cost, ev_mask=arr["has_ev"].astype(bool),
                              rating=arr["rating"], capacity=arr["capacity"], initial=arr["initial"], start=arr["start"],
                              end=arr["end"], **kw)
    assert tree.n_res == 1126 and st["max_working_set"] >= 5
    assert np.array_equal(out["P_ev"], ref["P_ev"])
    assert np.abs(out["P_sch"] - ref["P_sch"]).max() <= 1e-4
    assert np.abs(out["diff"] - ref["diff"]).max() <= 1e-7
