"""Drop-in for the reference's lpsolver.py: same entry points, B200 back end.

Reference                                   here
------------------------------------------  ---------------------------------------------
compute_Rmat(graph)        lpsolver.py:17   same matrix, from the rooted tree (host; setup only)
Home(...).solve()          lpsolver.py:45   Home(...)    -> revs_home_step    (1 warp / home)
Utility(...).solve()       lpsolver.py:163  Utility(...) -> revs_utility_step (working-set QP
                                            + FP64 tensor-core contraction)
solve_ADMM(...)            lpsolver.py:242  whole loop on the device -> revs_solve_admm
solve_residence(...)       lpsolver.py:430  revs_solve_individual
solve_central(...)         lpsolver.py:463  the reference's program in closed form + GPU voltage check

Arguments keep the reference's meaning; ``grbpath`` / ``path`` (Gurobi log directories)
are accepted and ignored.  Everything numerical runs in hand-written CUDA through the C
ABI of include/revs_admm.h; if the library or the GPU is missing these calls raise.
"""
import numpy as np

from . import _cabi
from .feeder import split_zones, tree_from_graph

__all__ = ["compute_Rmat", "Home", "Utility", "solve_ADMM", "solve_residence",
           "solve_residences", "solve_central", "compute_voltage", "compute_flows"]


# ------------------------------------------------------------------ helpers
def _home_arrays(homes, res):
    """dict-of-homes (extract.get_homes_ev_param) -> the arrays of revs_set_homes."""
    H = len(res)
    load = np.array([homes[h]["LOAD"] for h in res], dtype=np.float64)
    has_ev = np.zeros(H, np.uint8)
    rating, cap, init = np.zeros(H), np.ones(H), np.zeros(H)
    start, end = np.zeros(H, np.int32), np.zeros(H, np.int32)
    for i, h in enumerate(res):
        ev = homes[h]["EV"]
        if ev:
            has_ev[i] = 1
            rating[i], cap[i], init[i] = ev["rating"], ev["capacity"], ev["initial"]
            start[i], end[i] = ev["start"], ev["end"]
    return dict(load=load, has_ev=has_ev, rating=rating, capacity=cap, initial=init,
                start=start, end=end)


def _zones(graph):
    """Rooted tree of the feeder, its independent voltage zones (subtrees of the substation's
    children; R is block diagonal over them) and the residence permutation zone order ->
    reference order."""
    tree = tree_from_graph(graph)
    zones = split_zones(tree)
    perm = np.concatenate([h for _, h in zones]) if zones else np.zeros(0, dtype=np.int64)
    return tree, zones, perm


def _solver_for_graph(graph, homes, cost, device=0):
    tree, zones, perm = _zones(graph)
    res = tree.res_ids
    s = _cabi.Solver([len(h) for _, h in zones], len(cost), device=device)
    s.set_feeder_trees([z for z, _ in zones])
    s.set_homes(**_home_arrays(homes, [res[i] for i in perm]))
    s.set_tariff(cost)
    return s, tree, res, perm


def _as_table(tab, res, T):
    if isinstance(tab, dict):
        return np.array([np.asarray(tab[h], dtype=np.float64) for h in res]).reshape(len(res), T)
    return np.asarray(tab, dtype=np.float64).reshape(len(res), T)


# ------------------------------------------------------------------ network
def compute_Rmat(graph):
    """R = 2 F D F^T of lpsolver.py:17-26, rows/cols ordered like
    ``[n for n in graph.nodes if label != 'S']``."""
    tree = tree_from_graph(graph)
    Rt = tree.rmat()
    pos = {n: i for i, n in enumerate(tree.node_ids)}
    perm = [pos[n] for n in graph.nodes if graph.nodes[n]["label"] != "S"]
    return Rt[np.ix_(perm, perm)]


# ------------------------------------------------------------------ sub-problems
class Home:
    """One consumer's charging problem (lpsolver.py:44-160); solved on the GPU."""

    def __init__(self, cost, homedata, p_est, p_sch, gamma, kappa=5.0):
        self.c = list(cost)
        self.T = len(cost)
        self.data = homedata
        self.kappa = kappa
        self._in = [np.asarray(x, dtype=np.float64).reshape(1, self.T) for x in (p_est, p_sch, gamma)]

    def solve(self, grbpath=None):
        ev = self.data["EV"]
        with _cabi.Solver([1], self.T) as s:
            s.set_sensitivity(0, np.zeros((1, 1)))
            s.set_homes(**_home_arrays({0: self.data}, [0]))
            s.set_tariff(self.c)
            g, p = s.home_step(*self._in, kappa=self.kappa)
        self.g_opt = g[0].tolist()
        self.p_opt = p[0].tolist()
        soc = [0.0] * (self.T + 1)
        if ev:
            soc[0] = ev["initial"]
            for t in range(self.T):
                soc[t + 1] = soc[t] + self.p_opt[t] / ev["capacity"]
        self.s_opt = soc
        return


class Utility:
    """The operator's estimate under the voltage limits (lpsolver.py:163-238)."""

    def __init__(self, graph, P_util, P_sch, Gamma, kappa=5.0, vset=1.0, low=0.95, high=1.05):
        self.tree, self.zones, self.perm = _zones(graph)
        self.res = self.tree.res_ids
        self.nodes = [n for n in graph.nodes if graph.nodes[n]["label"] != "S"]
        self.N = len(self.nodes)
        self.T = len(Gamma[self.res[0]])
        self.kappa, self.vset, self.low, self.high = kappa, vset, low, high
        self._in = [_as_table(x, self.res, self.T) for x in (P_util, P_sch, Gamma)]

    def solve(self, grbpath=None):
        with _cabi.Solver([len(h) for _, h in self.zones], self.T) as s:
            s.set_feeder_trees([z for z, _ in self.zones])
            g, lam = s.utility_step(*[x[self.perm] for x in self._in], kappa=self.kappa, vset=self.vset,
                                    vlow=self.low, vhigh=self.high)
        self.g_opt = {self.res[j]: g[k].tolist() for k, j in enumerate(self.perm)}
        self.lam_opt = {self.res[j]: lam[k] for k, j in enumerate(self.perm)}
        return


# ------------------------------------------------------------------ the ADMM loop
def solve_ADMM(homes, graph, cost, grbpath=None, kappa=5.0, iter_max=15,
               vset=1.0, vlow=0.95, vhigh=1.05, tol=0.0, device=0, return_stats=False):
    """Iterative ADMM of lpsolver.py:242-290, entirely on the device.

    Returns ``diff, P_sch, S, C`` exactly like the reference: diff[k][h] for k=1..iter_max,
    and the last iterate's residence profile, EV charger profile and SOC profile per home.
    ``tol`` > 0 (an extension) stops once both ADMM residuals fall below it."""
    s, tree, res, perm = _solver_for_graph(graph, homes, cost, device)
    with s:
        done = s.solve_admm(kappa=kappa, iter_max=iter_max, vset=vset, vlow=vlow, vhigh=vhigh, tol=tol)
        out = s.results(done)
        stats = s.stats()
    inv = np.empty(len(perm), dtype=np.int64)          # reference residence index -> solver row
    inv[perm] = np.arange(len(perm))
    diff = {k + 1: {h: out["diff"][k, inv[i]] for i, h in enumerate(res)} for k in range(done)}
    P = {h: out["P_sch"][inv[i]] for i, h in enumerate(res)}
    S = {h: out["P_ev"][inv[i]] for i, h in enumerate(res)}
    Csoc = {h: out["SOC"][inv[i]] for i, h in enumerate(res)}
    if return_stats:
        return diff, P, S, Csoc, stats
    return diff, P, S, Csoc


# ------------------------------------------------------------------ individual optimum
def solve_residences(tariff, homes, device=0):
    """solve_residence for a whole dict of homes in one launch."""
    res = list(homes)
    with _cabi.Solver([len(res)], len(tariff), device=device) as s:
        s.set_homes(**_home_arrays(homes, res))
        s.set_tariff(tariff)
        out = s.solve_individual()
    return ({h: out["P_ev"][i] for i, h in enumerate(res)},
            {h: out["SOC"][i] for i, h in enumerate(res)},
            {h: out["P_res"][i] for i, h in enumerate(res)})


def solve_residence(tariff, data, path=None):
    """lpsolver.py:430-460: returns p_opt, s_opt, g_opt of one home."""
    p, s, g = solve_residences(tariff, {0: data})
    return p[0], s[0], g[0]


def solve_central(tariff, homes, dist, path=None, vset=1.0, vmin=0.9, vmax=1.05, device=0):
    """Centralized schedule, lpsolver.py:463-502, with the reference's exact semantics.

    The reference's program minimises the tariff cost of g = load + p (objective_centralized,
    lpsolver.py:418-428) over the charger variables of add_home_EV (lpsolver.py:338-379: binary
    status, init <= soc <= 1 -- and, unlike class Home, NO final-SOC row) under the voltage rows of
    network_constraints (lpsolver.py:386-405):  -R g <= vmax^2 - vset^2  (vacuous for g >= 0) and
    -R g >= vmin^2 - vset^2, i.e.  R g <= vset^2 - vmin^2.  Charging only adds cost and voltage drop,
    so with a positive tariff the optimum is p = 0 if the base load satisfies the voltage rows, and the
    program is infeasible otherwise -- the reference's own centralized result file
    (out/121144-com2/centralized) holds exactly that: no charging, SOC constant at its initial value.

    What remains to compute is the reliability check of the base load: R_res @ load on the GPU (FP64
    tensor-core contraction, revs_reliability) against vset^2 - vmin^2 for every residence and hour.
    Returns p_opt, s_opt, g_opt like the reference; raises RuntimeError where the reference prints
    'No solution found' and exits."""
    tariff = np.asarray(tariff, dtype=np.float64)
    if not (tariff > 0).all():
        raise NotImplementedError("solve_central: a non-positive tariff makes charging profitable; the closed form "
                                  "used here (see docstring) needs tariff > 0")
    tree, zones, perm = _zones(dist)
    res = tree.res_ids
    T = len(tariff)
    limit = vset * vset - vmin * vmin
    if zones:
        P = np.array([np.asarray(homes[res[j]]["LOAD"], dtype=np.float64) for j in perm]).reshape(len(perm), T)
        off = np.concatenate([[0], np.cumsum([len(h) for _, h in zones])])
        worst = -np.inf
        with _cabi.Solver([len(h) for _, h in zones], T, device=device) as s:
            s.set_feeder_trees([z for z, _ in zones])
            for f, (z, _) in enumerate(zones):
                drop = s.reliability(f, _cabi.REVS_REL_DROP, z.res_node, P=P[off[f]:off[f + 1]])
                worst = max(worst, float(drop.max()))
        if worst > limit + 1e-9 or not (vmax * vmax - vset * vset >= 0.0):
            raise RuntimeError(f"No solution found: the base load alone drops the squared voltage by {worst:.6g} pu "
                               f"(limit vset^2 - vmin^2 = {limit:.6g}); the reference's centralized program is infeasible")
    p_opt = {h: np.zeros(T) for h in res}
    s_opt = {h: np.full(T + 1, float(homes[h]["EV"]["initial"]) if homes[h]["EV"] else 0.0) for h in res}
    g_opt = {h: np.asarray(homes[h]["LOAD"], dtype=np.float64).copy() for h in res}
    return p_opt, s_opt, g_opt


# ------------------------------------------------------------------ reliability check
def _reliability(graph, p_sch, kind, vset, scale_of_zone, fill, device):
    """Run the contraction zone by zone; nodes of zones without residences get `fill`."""
    tree, zones, perm = _zones(graph)
    T = len(next(iter(p_sch.values())))
    out = {}
    if zones:
        P = np.array([np.asarray(p_sch[tree.res_ids[j]], dtype=np.float64) for j in perm]).reshape(len(perm), T)
        off = np.concatenate([[0], np.cumsum([len(h) for _, h in zones])])
        with _cabi.Solver([len(h) for _, h in zones], T, device=device) as s:
            s.set_feeder_trees([z for z, _ in zones])
            for f, (z, _) in enumerate(zones):
                vals = s.reliability(f, kind, np.arange(z.n_nodes), vset=vset, scale=scale_of_zone(z),
                                     P=P[off[f]:off[f + 1]])
                out[f] = (z, vals)
    return tree, T, out, fill


def compute_voltage(graph, p_sch, vset=1.0, device=0):
    """drawing.py:61-78: {node: voltage profile} for every non-substation node."""
    tree, T, out, _ = _reliability(graph, p_sch, _cabi.REVS_REL_VOLTAGE, vset, lambda z: None, vset, device)
    volt = {n: [vset] * T for n in tree.node_ids}
    for z, V in out.values():
        for i, n in enumerate(z.node_ids):
            volt[n] = V[i].tolist()
    return volt


LINE_RATING_KVA = {  # conductor ampacity x voltage, drawing.py:30-41
    "OH_Voluta": 95 * 0.24, "OH_Periwinkle": 125 * 0.24, "OH_Conch": 165 * 0.24,
    "OH_Neritina": 220 * 0.24, "OH_Runcina": 265 * 0.24, "OH_Zuzara": 350 * 0.24,
    "OH_Swanate": 145 * 12.47, "OH_Sparrow": 185 * 12.47, "OH_Raven": 240 * 12.47,
    "OH_Pegion": 315 * 12.47, "OH_Penguin": 365 * 12.47,
}


def compute_flows(graph, p_sch, device=0):
    """drawing.py:29-59: {edge: signed loading = flow / rating} for every line."""
    def scale(z):
        rating = np.array([np.sqrt(3) * LINE_RATING_KVA[t] for t in z.edge_type])
        return z.edge_sign / rating
    tree, T, out, _ = _reliability(graph, p_sch, _cabi.REVS_REL_FLOW, 1.0, scale, 0.0, device)
    flows = {e: [0.0] * T for e in tree.edge_keys}
    for z, F in out.values():
        for i, e in enumerate(z.edge_keys):
            flows[e] = F[i].tolist()
    return flows
