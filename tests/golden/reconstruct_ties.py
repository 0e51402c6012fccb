#!/usr/bin/env python
"""Reconstruct the tie-breaks Gurobi made in iteration 1 of the reference's distributed run
(out/121144-com2/distributed/adopt90-rating4800-seed1234.txt) from the convergence values the
reference itself wrote -- an attempt to pin iterations >= 2 of the oracle to reference-held vectors.

Why iteration 1 decides everything.  diff[1][h] = ||P_sch[1][h]|| / T fixes the SUM of the loads in
the three charging hours of home h but not the hours: 143 of the 267 EV homes have several hour
triples with the same load sum whose cost lies within Gurobi's MIPGap (1e-4, relative) of the
optimum, and the reference's MIQP returns any of them.  In iteration 2 the home problem is the same
program again (a = Gamma[1] + kappa/2 (P_est[1] + P_sch[1]) = 0), so S2 = S1 and
    diff[2][h] = || proj(P_sch[1])[h] - P_sch[1][h] || / T
is a function of the iteration-1 choices of ALL homes of the zone through the operator QP.  This
script searches the choices (coordinate descent from several starts, one warm-started QP solve of
the touched hours per trial) for the assignment that reproduces the reference's diff[2].

Result (this container, 3 starts): the earliest-hour rule of the oracle is at median |error| 7.5e-4
of the file; the best assignment found reaches median 3e-5 .. 9e-5 with a handful of homes at 1e-2,
and no home below 1e-9.  The search does not reach an exact match: different starts end in different
assignments of similar quality.  So the ceiling stays "partially pinned": iterations >= 2 of the
reference are reproduced to ~1e-4 in the median with the committed choices
(tests/golden/tie_choices_iter1_121144_com2.npz), not to rounding.

Needs only the committed fixtures (no /root/reference).  Run:  python tests/golden/reconstruct_ties.py [starts]
"""
import itertools
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import revs_oracle as O  # noqa: E402
from revs_admm_b200.feeder import split_zones, tree_from_graph  # noqa: E402
from revs_admm_b200.revs_fixture import REVS  # noqa: E402

KAPPA, VSET, VHIGH, RATING = 5.0, 1.03, 1.05, 4.8
MIPGAP = 1e-4


def load_case():
    g = dict(np.load(os.path.join(HERE, "ref_out_121144_com2.npz")))
    fx = REVS(data_path=os.path.join(HERE, "input"), out_path="/tmp/revs_out", grb_path="/tmp/revs_grb",
              fig_path="/tmp/revs_fig", regionID=121, networkID=121144, comunityID=2, optimizer_mode="distributed")
    tariff, homes, dist, _ = fx.read_inputs(adoption=90, rating=4800, seed=1234, capacity=20, initial_soc=0.2,
                                            start_time=11, end_time=23, shift_time=6)
    tree = tree_from_graph(dist)
    res = tree.res_ids
    arr, T, H = O.homes_to_arrays(homes, res)
    pos = {int(h): i for i, h in enumerate(res)}
    evrow = np.array([pos[int(h)] for h in g["distributed_ev_ids"]])
    zones = split_zones(tree)
    zone = next((z, h) for z, h in zones if set(evrow) <= set(h.tolist()))     # community 2 = one voltage zone
    return g, arr, T, np.asarray(tariff, float), evrow, zone


def candidates(gold1, load, cost, T):
    """Hour triples of the plug-in window that reproduce diff[1] and are optimal within MIPGap."""
    combos = np.array(list(itertools.combinations(range(11, 23), 3)))
    out = []
    for gi in range(len(load)):
        ld = load[gi]
        d = O.home_delta(cost, ld, np.zeros(T), np.zeros(T), np.zeros(T), KAPPA, RATING)
        val = np.sqrt(ld @ ld + 2 * RATING * ld[combos].sum(1) + 3 * RATING * RATING) / T
        obj = d[combos].sum(1)
        full = cost @ ld + 0.5 * KAPPA * (ld @ ld)
        ok = (np.abs(val - gold1[gi]) < 1e-13) & (obj - obj.min() <= MIPGAP * abs(full + obj.min()) + 1e-12)
        cs = [tuple(int(x) for x in c) for c, a in zip(combos, ok) if a]
        cs.sort(key=lambda c: (d[list(c)].sum(), c))
        out.append(cs)
    return out


def main():
    starts = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    g, arr, T, cost, evrow, (ztree, zhomes) = load_case()
    gold = g["distributed_diff"]
    load = arr["load"]
    R = O.rmat_from_tree(ztree.parent, ztree.r)[np.ix_(ztree.res_node, ztree.res_node)]
    rn2 = (R * R).sum(axis=1)
    u = VHIGH ** 2 - VSET ** 2
    loc = {int(i): j for j, i in enumerate(zhomes)}
    evloc = np.array([loc[int(i)] for i in evrow])
    cands = candidates(gold[:, 0], load[evrow], cost, T)
    print("ambiguous homes:", sum(len(c) > 1 for c in cands), "of", len(cands))

    def score(Pe, z):
        d = Pe[evloc] - z[evloc]
        return np.abs(np.sqrt((d * d).sum(1)) / T - gold[:, 1])

    best = None
    for seed in range(starts):
        rng = np.random.default_rng(seed)
        sel = [c[0] if seed == 0 else c[rng.integers(len(c))] for c in cands]
        z = load[zhomes].copy()
        for gi in range(len(cands)):
            z[evloc[gi], list(sel[gi])] += RATING
        Pe, lam = np.maximum(z, 0), np.zeros_like(z)
        for t in range(T):
            Pe[:, t], lam[:, t], _ = O.project_voltage(z[:, t], R, u, None, rn2=rn2)
        J = (score(Pe, z) ** 2).sum()
        print(f"start {seed}: median |err| {np.median(score(Pe, z)):.3e}")
        amb = [gi for gi in range(len(cands)) if len(cands[gi]) > 1]
        for sweep in range(40):
            nchg, t0 = 0, time.time()
            for gi in rng.permutation(amb):
                j, cur = evloc[gi], sel[gi]
                bestJ, beststate = J, None
                for alt in cands[gi]:
                    if alt == cur:
                        continue
                    hrs = sorted(set(alt) ^ set(cur))
                    zz, Pe2, l2 = z.copy(), Pe.copy(), lam[:, hrs].copy()
                    for q, t in enumerate(hrs):
                        zz[j, t] = load[evrow[gi], t] + (RATING if t in alt else 0.0)
                        Pe2[:, t], l2[:, q], _ = O.project_voltage(zz[:, t], R, u, lam[:, t], rn2=rn2)
                    J2 = (score(Pe2, zz) ** 2).sum()
                    if J2 < bestJ:
                        bestJ, beststate = J2, (alt, hrs, zz, Pe2, l2)
                if beststate is not None:
                    sel[gi], hrs, z, Pe, l2 = beststate
                    lam[:, hrs] = l2
                    J = bestJ
                    nchg += 1
            e = score(Pe, z)
            print(f"  sweep {sweep}: J {J:.4e} median {np.median(e):.3e} max {e.max():.3e} changed {nchg} ({time.time() - t0:.0f}s)", flush=True)
            if nchg == 0:
                break
        if best is None or J < best[0]:
            best = (J, [tuple(s) for s in sel])
    np.savez(os.path.join(HERE, "tie_choices_iter1_121144_com2.npz"), ev_ids=g["distributed_ev_ids"],
             hours=np.array(best[1], dtype=np.int32), J=best[0])
    print("best J", best[0])


if __name__ == "__main__":
    main()
