"""Radial feeders as arrays: what the device needs of the reference's networkx graph.

The reference keeps the feeder as a networkx graph (extract.py:48-80 GetDistNet) and turns
it into dense sensitivity matrices by inverting its incidence matrix (lpsolver.py:17-26).
Here the graph is reduced once, on the host, to a rooted tree in topological order
(parent index, resistance of the edge above each node, residence -> node index); the
dense blocks are then built on the GPU (csrc/feeder_build.cu).
"""
from dataclasses import dataclass, field

import numpy as np


@dataclass
class FeederTree:
    parent: np.ndarray            # int32 [n]  (-1 = substation), parent[i] < i
    r: np.ndarray                 # float64 [n] resistance of the edge above node i
    res_node: np.ndarray          # int32 [n_res] tree index of every residence
    node_ids: list = field(default_factory=list)   # tree index -> graph node id
    res_ids: list = field(default_factory=list)    # residence order of the reference
    edge_keys: list = field(default_factory=list)  # tree index -> (u, v) as in graph.edges
    edge_sign: np.ndarray = None  # +1 if (u, v) points away from the substation
    edge_type: list = field(default_factory=list)

    @property
    def n_nodes(self):
        return len(self.parent)

    @property
    def n_res(self):
        return len(self.res_node)

    def cumulative_r(self):
        c = np.zeros(self.n_nodes)
        for i in range(self.n_nodes):
            c[i] = (c[self.parent[i]] if self.parent[i] >= 0 else 0.0) + self.r[i]
        return c

    def rmat(self):
        """Dense R over all non-substation nodes in tree order (host, O(n^2)); the
        matrix of lpsolver.py:17-26 up to the node permutation ``node_ids``."""
        n = self.n_nodes
        R = np.zeros((n, n))
        for i in range(n):
            p = self.parent[i]
            if p >= 0:
                R[i, :i] = R[p, :i]
                R[:i, i] = R[i, :i]
                R[i, i] = R[p, p] + 2.0 * self.r[i]
            else:
                R[i, i] = 2.0 * self.r[i]
        return R


    def drop(self, P):
        """R_res @ P for a residence schedule P [n_res, T] without forming R (host, O(n T)): subtree sums leaves ->
        root, then the squared-voltage drop root -> leaves (v_k = v_parent + 2 r_k f_k).  The check of lpsolver.py:192."""
        P = np.asarray(P, dtype=np.float64).reshape(self.n_res, -1)
        f = np.zeros((self.n_nodes, P.shape[1]))
        np.add.at(f, self.res_node, P)
        for i in range(self.n_nodes - 1, -1, -1):
            if self.parent[i] >= 0:
                f[self.parent[i]] += f[i]
        v = f
        for i in range(self.n_nodes):
            v[i] = 2.0 * self.r[i] * f[i] + (v[self.parent[i]] if self.parent[i] >= 0 else 0.0)
        return v[self.res_node]


def tree_from_graph(graph):
    """Root the reference's feeder graph at its substation (label 'S')."""
    roots = [n for n in graph.nodes if graph.nodes[n]["label"] == "S"]
    if len(roots) != 1:
        raise ValueError("expected exactly one substation node (label 'S')")
    root = roots[0]
    if graph.number_of_edges() != graph.number_of_nodes() - 1:
        raise ValueError("feeder is not radial (edges != nodes - 1)")
    order, par = [], {}
    index = {root: -1}
    stack = [root]
    while stack:                      # depth-first, parents before children
        u = stack.pop()
        for v in graph.neighbors(u):
            if v in index:
                continue
            index[v] = len(order)
            order.append(v)
            par[v] = u
            stack.append(v)
    if len(order) != graph.number_of_nodes() - 1:
        raise ValueError("feeder graph is not connected")
    n = len(order)
    parent = np.array([index[par[v]] for v in order], dtype=np.int32)
    r = np.array([graph.edges[par[v], v]["r"] for v in order], dtype=np.float64)
    orient = {}
    for (a, b) in graph.edges:
        orient[frozenset((a, b))] = (a, b)
    keys, sign, types = [], np.ones(n), []
    for i, v in enumerate(order):
        a, b = orient[frozenset((par[v], v))]
        keys.append((a, b))
        sign[i] = 1.0 if a == par[v] else -1.0
        types.append(graph.edges[a, b].get("type"))
    res_ids = [n_ for n_ in graph if graph.nodes[n_]["label"] == "H"]
    res_node = np.array([index[h] for h in res_ids], dtype=np.int32)
    return FeederTree(parent=parent, r=r, res_node=res_node, node_ids=order, res_ids=res_ids,
                      edge_keys=keys, edge_sign=sign, edge_type=types)


def split_zones(tree):
    """Split a feeder at the substation into its independent voltage zones.

    Two nodes that leave the substation on different lines share no edge of their root
    paths, so the sensitivity matrix R (lpsolver.py:17-26) is block diagonal over the
    subtrees of the substation's children -- the reference's `<net>-com.txt` communities are
    exactly these blocks.  The operator QP therefore separates over them, and solving the
    zones as separate (smaller) feeders is the same problem with less memory, less work and
    smaller working sets.  Returns [(zone_tree, home_positions)], home_positions indexing the
    residence order of `tree`; zones without residences are dropped."""
    n = tree.n_nodes
    zone = np.empty(n, dtype=np.int64)
    nz = 0
    for i in range(n):
        p = tree.parent[i]
        if p < 0:
            zone[i] = nz
            nz += 1
        else:
            zone[i] = zone[p]
    res_zone = zone[tree.res_node]
    out = []
    for z in range(nz):
        homes = np.nonzero(res_zone == z)[0]
        if len(homes) == 0:
            continue
        nodes = np.nonzero(zone == z)[0]                    # ascending -> still topological
        local = np.full(n, -1, dtype=np.int64)
        local[nodes] = np.arange(len(nodes))
        par = tree.parent[nodes]
        par = np.where(par >= 0, local[np.maximum(par, 0)], -1).astype(np.int32)
        zt = FeederTree(parent=par, r=tree.r[nodes].copy(), res_node=local[tree.res_node[homes]].astype(np.int32),
                        node_ids=[tree.node_ids[i] for i in nodes] if tree.node_ids else [],
                        res_ids=[tree.res_ids[i] for i in homes] if tree.res_ids else [],
                        edge_keys=[tree.edge_keys[i] for i in nodes] if tree.edge_keys else [],
                        edge_sign=tree.edge_sign[nodes].copy() if tree.edge_sign is not None else None,
                        edge_type=[tree.edge_type[i] for i in nodes] if tree.edge_type else [])
        out.append((zt, homes))
    return out


def synthetic_feeder(n_homes, seed=0, laterals=5, homes_per_xfmr=2.5,
                     r_primary=1.0e-6, r_secondary=3.2e-4):
    """Synthetic radial feeder shaped like the reference's 121144 network: a substation,
    ``laterals`` primary chains of transformer nodes, 1-4 residences per transformer on
    secondary lines (the reference has 1126 homes / 462 transformers, primary r ~1e-6,
    secondary r ~3e-4 with a long tail)."""
    rng = np.random.default_rng(seed)
    n_x = max(laterals, int(np.ceil(n_homes / homes_per_xfmr)))
    lat = np.sort(rng.integers(0, laterals, size=n_x))          # lateral of each transformer
    first = np.r_[True, lat[1:] != lat[:-1]]
    parent_x = np.where(first, -1, np.arange(n_x) - 1).astype(np.int32)
    r_x = r_primary * rng.lognormal(0.0, 0.8, size=n_x)
    home_x = np.sort(rng.integers(0, n_x, size=n_homes)).astype(np.int32)
    r_h = r_secondary * rng.lognormal(0.0, 1.0, size=n_homes)
    parent = np.concatenate([parent_x, home_x]).astype(np.int32)
    r = np.concatenate([r_x, r_h])
    res_node = (n_x + np.arange(n_homes)).astype(np.int32)
    return FeederTree(parent=parent, r=r, res_node=res_node, edge_sign=np.ones(len(parent)))


def synthetic_homes(n_homes, T, seed=0, adoption=0.9, rating_kw=4.8, capacity=20.0,
                    initial=0.2, steps_per_hour=None):
    """Synthetic home population of the reference's shape: a daily base-load profile with
    an evening peak, EV adopters with the reference's charger data (revs_config.yaml:12-20),
    plug-in window 17:00 -> 05:00 as in start_time=11,end_time=23 of a day that starts at
    06:00.  With T=96 the capacity is scaled so that the charging energy is unchanged."""
    rng = np.random.default_rng(seed + 7919)
    sph = steps_per_hour or max(1, T // 24)
    hours = (np.arange(T) / sph + 6.0) % 24.0
    shape = 0.6 + 0.5 * np.exp(-0.5 * ((hours - 19.0) / 2.5) ** 2) + 0.25 * np.exp(-0.5 * ((hours - 8.0) / 1.5) ** 2)
    scale = rng.lognormal(0.3, 0.5, size=(n_homes, 1))
    load = scale * shape[None, :] * (1.0 + 0.15 * rng.standard_normal((n_homes, T)))
    load = np.round(np.maximum(load, 0.05), 5)
    has_ev = (rng.random(n_homes) < adoption).astype(np.uint8)
    rating = np.full(n_homes, rating_kw)
    cap = np.full(n_homes, capacity * sph)        # kWh expressed in kW*steps
    init = np.full(n_homes, initial)
    start = np.full(n_homes, 11 * sph, dtype=np.int32)
    end = np.full(n_homes, 23 * sph, dtype=np.int32)
    return dict(load=load, has_ev=has_ev, rating=rating, capacity=cap, initial=init,
                start=start, end=end)


def synthetic_tariff(T):
    """The DVP time-of-use tariff of the reference (input/DVP-tariff.txt) rolled to a day
    that starts at 06:00 (extract.py:23 with shift=6), repeated to T steps."""
    day = np.array([0.07866] * 5 + [0.095111] * 10 + [0.214357] * 3 + [0.095111] * 6)
    day = np.roll(day, -6)
    sph = max(1, T // 24)
    return np.resize(np.repeat(day, sph), T) if T >= 24 else day[:T]


def reference_shaped_feeder(n_homes, seed=0, zone_lo=157, zone_hi=297, homes_per_xfmr=2.4,
                            r_primary=6.0e-7, r_secondary=2.2e-4, sigma_secondary=0.95, r_scale=1.0):
    """Synthetic radial feeder with the anatomy of the reference's 121144 network (tests/golden/input):

    * the substation feeds several independent voltage zones of ``zone_lo..zone_hi`` residences
      (121144: 157, 208, 216, 248, 297),
    * a zone is a primary tree of transformer / road nodes -- mostly a chain, with side branches,
      40..50 levels deep, edges of ~6e-7 ohm-equivalents (121144: median 5.1e-7, max 1e-5),
    * 2..3 residences per transformer on secondary lines (median 2.2e-4, long tail up to ~15x),
      about half of them daisy-chained behind another residence (121144: 99 of 216 residences
      have a residence below them).

    ``r_scale`` multiplies every resistance: 1.0 reproduces the reference's own regime, in which
    the base load alone exceeds the voltage limit u = vhigh^2 - vset^2 by 1.5x..5x in every zone
    (ADMM then never reaches P_est = P_sch); ~0.2 gives a voltage-feasible population whose
    limits bind only when the chargers run."""
    rng = np.random.default_rng(seed)
    sizes = []
    left = int(n_homes)
    while left > 0:
        n = int(rng.integers(zone_lo, zone_hi + 1))
        if left - n < zone_lo:
            n = left if left <= zone_hi else left // 2
        sizes.append(n)
        left -= n
    parent, r, res_node = [], [], []
    for n in sizes:
        base = len(parent)
        n_x = max(2, int(round(n / homes_per_xfmr)))
        # primary tree: chain with side branches
        xp = np.empty(n_x, dtype=np.int64)
        xp[0] = -1
        for i in range(1, n_x):
            xp[i] = i - 1 if rng.random() < 0.78 else int(rng.integers(0, i))
        parent.extend([-1 if p < 0 else base + int(p) for p in xp])
        r.extend((r_scale * r_primary * rng.lognormal(0.0, 0.6, size=n_x)).tolist())
        # residences: every one hangs on a transformer or behind the previous residence of that transformer
        owner = np.sort(rng.integers(0, n_x, size=n))
        rh = r_scale * r_secondary * rng.lognormal(0.0, sigma_secondary, size=n)
        last_of = {}
        for j in range(n):
            x = int(owner[j])
            node = len(parent)
            if x in last_of and rng.random() < 0.55:
                parent.append(last_of[x])              # daisy chain
            else:
                parent.append(base + x)
            r.append(float(rh[j]))
            last_of[x] = node
            res_node.append(node)
    parent = np.asarray(parent, dtype=np.int32)
    # topological order: parents precede children by construction (transformers of a zone first,
    # a residence's parent is an earlier node)
    return FeederTree(parent=parent, r=np.asarray(r, dtype=np.float64), res_node=np.asarray(res_node, dtype=np.int32),
                      edge_sign=np.ones(len(parent)))


# Named synthetic populations of bench.py and of the full-size tests: (homes per feeder, T, generator).
POPULATIONS = {
    # reference-shaped zones (157..297 residences), base load <= ~0.7 u, limits bind when the chargers run
    "refshape": dict(homes=1000, T=96, gen=lambda n, seed: reference_shaped_feeder(n, seed=seed, r_scale=0.7)),
    # round-1 population: zones of 43..165 residences, every lateral on the substation, base load over the limit
    "laterals": dict(homes=1000, T=96, gen=lambda n, seed: synthetic_feeder(n, seed=seed, laterals=max(5, n // 100))),
    # one radial feeder: trunk + laterals, a single voltage zone (BASELINE.json config 3)
    # (r_scale: the base load peaks at ~0.7 u at the end of the feeder, all chargers together would need 2.4 u)
    "radial10k": dict(homes=10000, T=96, gen=lambda n, seed: radial_feeder(n, seed=seed, r_scale=0.07)),
}


def radial_feeder(n_homes, seed=0, homes_per_xfmr=2.5, lateral_len=40, r_trunk=2.0e-7, r_primary=6.0e-7,
                  r_secondary=2.2e-4, r_scale=1.0):
    """A genuinely radial feeder: ONE trunk leaves the substation, laterals of ``lateral_len``
    transformers branch off it, residences hang on the transformers.  Every pair of homes shares
    at least the first trunk segment, so the sensitivity matrix is a single dense block."""
    rng = np.random.default_rng(seed)
    n_x = max(2, int(np.ceil(n_homes / homes_per_xfmr)))
    n_lat = max(1, int(np.ceil(n_x / lateral_len)))
    parent, r = [], []
    for k in range(n_lat):                                   # trunk nodes 0..n_lat-1
        parent.append(k - 1)
        r.append(r_scale * r_trunk * rng.lognormal(0.0, 0.4))
    xf = []
    for k in range(n_lat):
        prev = k
        for i in range(min(lateral_len, n_x - k * lateral_len)):
            parent.append(prev)
            r.append(r_scale * r_primary * rng.lognormal(0.0, 0.6))
            prev = len(parent) - 1
            xf.append(prev)
    owner = np.sort(rng.integers(0, len(xf), size=n_homes))
    rh = r_scale * r_secondary * rng.lognormal(0.0, 0.95, size=n_homes)
    res_node = []
    for j in range(n_homes):
        parent.append(xf[int(owner[j])])
        r.append(float(rh[j]))
        res_node.append(len(parent) - 1)
    return FeederTree(parent=np.asarray(parent, dtype=np.int32), r=np.asarray(r, dtype=np.float64),
                      res_node=np.asarray(res_node, dtype=np.int32), edge_sign=np.ones(len(parent)))


def population(name, n_feeders, seed=0, split=True, first_feeder=0):
    """Feeders ``first_feeder .. first_feeder + n_feeders - 1`` of population ``name`` with draw ``seed``:
    (trees, homes dict, tariff, zone sizes, T).  With ``split`` every feeder is handed over as its
    independent voltage zones (what lpsolver.solve_ADMM of this package does for a networkx feeder)
    and the homes are reordered accordingly.  Feeder f of a population is the same whatever slice it
    is generated in, so a test can compare the first feeders of the benchmark population with the oracle."""
    spec = POPULATIONS[name]
    n, T = spec["homes"], spec["T"]
    feeders = [spec["gen"](n, 100003 * seed + f) for f in range(first_feeder, first_feeder + n_feeders)]
    parts = [synthetic_homes(n, T, seed=100003 * seed + 77 + f) for f in range(first_feeder, first_feeder + n_feeders)]
    hm = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    if not split:
        return feeders, hm, synthetic_tariff(T), [n] * n_feeders, T
    trees, perm = [], []
    for f, tr in enumerate(feeders):
        for zt, homes in split_zones(tr):
            trees.append(zt)
            perm.append(f * n + homes)
    perm = np.concatenate(perm)
    hm = {k: np.ascontiguousarray(v[perm]) for k, v in hm.items()}
    return trees, hm, synthetic_tariff(T), [t.n_res for t in trees], T
