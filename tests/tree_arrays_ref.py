"""numpy restatement of the zone-tree arrays of the tree-structured operator QP kernel (csrc/tree_qp.cu,
host side: revs_capi.cu:build_zone_arrays).  TEST INFRASTRUCTURE: used by tests/test_host.py to check the
decomposition itself against the dense sensitivity block, independent of any CUDA code.

For the residences of a zone in depth-first order, R[i][j] = 2 cumr(lca(i, j)) = min(c[i..j-1]) (i < j) with
c[p] = 2 cumr(lca(p, p+1)), R[i][i] = d[i].  The Cartesian tree of c has one node per c-position q covering the
leaves [lo_q, hi_q] with weight w_q = c[q] - (value of its parent node), and
    (R g)[p] = sum_{q : lo_q <= p <= hi_q} w_q F_q + e[p] g[p],   F_q = sum_{lo_q <= j <= hi_q} g[j],
e[p] = d[p] - max(c[p-1], c[p]).
"""
import numpy as np


def zone_arrays(parent, r, res_node):
    parent = np.asarray(parent)
    n_nodes, n = len(parent), len(res_node)
    cumr = np.zeros(n_nodes)
    depth = np.zeros(n_nodes, dtype=np.int64)
    children = [[] for _ in range(n_nodes)]
    roots = []
    for i in range(n_nodes):
        p = parent[i]
        if p >= 0:
            cumr[i] = cumr[p] + r[i]
            depth[i] = depth[p] + 1
            children[p].append(i)
        else:
            cumr[i] = r[i]
            roots.append(i)
    homes_at = [[] for _ in range(n_nodes)]
    for h, nd in enumerate(res_node):
        homes_at[nd].append(h)
    perm, leaf_node = [], []
    stack = list(reversed(roots))
    while stack:                                   # depth-first, preorder: a residence precedes the residences below it
        x = stack.pop()
        for h in homes_at[x]:
            perm.append(h)
            leaf_node.append(x)
        stack.extend(reversed(children[x]))
    assert len(perm) == n

    def lca(a, b):
        while a != b:
            if a < 0 or b < 0:
                return -1
            if depth[a] > depth[b]:
                a = parent[a]
            elif depth[b] > depth[a]:
                b = parent[b]
            else:
                a, b = parent[a], parent[b]
        return a

    c = np.zeros(max(n - 1, 0))
    for p in range(n - 1):
        a = lca(leaf_node[p], leaf_node[p + 1])
        c[p] = 0.0 if a < 0 else 2.0 * cumr[a]
    d = np.array([2.0 * cumr[x] for x in leaf_node])
    # Cartesian tree of c: previous strictly smaller, next smaller-or-equal
    m = n - 1
    prev_s, next_se = np.full(m, -1), np.full(m, m)
    st = []
    for q in range(m):
        while st and c[st[-1]] >= c[q]:
            next_se[st.pop()] = q
        prev_s[q] = st[-1] if st else -1
        st.append(q)
    lo = prev_s + 1                                # leaves [lo, hi]
    hi = np.where(next_se < m, next_se, n - 1)
    pv = np.maximum(np.where(prev_s >= 0, c[np.maximum(prev_s, 0)], 0.0), np.where(next_se < m, c[np.minimum(next_se, m - 1)], 0.0)) if m else np.zeros(0)
    w = c - pv
    e = d.copy()
    for p in range(n):
        nb = max(c[p - 1] if p > 0 else 0.0, c[p] if p < n - 1 else 0.0)
        e[p] = d[p] - nb
    return dict(perm=np.array(perm), c=c, d=d, lo=lo, hi=hi, w=w, e=e)


def product(za, g_dfs):
    """(R g) in depth-first order from the arrays (the kernel's three-scan algorithm, written with loops)."""
    n = len(za["d"])
    G = np.concatenate([[0.0], np.cumsum(g_dfs)])
    T = za["w"] * (G[za["hi"] + 1] - G[za["lo"]])
    v = za["e"] * g_dfs
    oa, ob = np.argsort(za["lo"], kind="stable"), np.argsort(za["hi"], kind="stable")
    S1, S2 = np.concatenate([[0.0], np.cumsum(T[oa])]), np.concatenate([[0.0], np.cumsum(T[ob])])
    cnt_lo = np.searchsorted(za["lo"][oa], np.arange(n), side="right")
    cnt_hi = np.searchsorted(za["hi"][ob], np.arange(n), side="left")
    return v + S1[cnt_lo] - S2[cnt_hi]


def row(za, i):
    n = len(za["d"])
    out = np.empty(n)
    out[i] = za["d"][i]
    run = np.inf
    for j in range(i + 1, n):
        run = min(run, za["c"][j - 1])
        out[j] = run
    run = np.inf
    for j in range(i - 1, -1, -1):
        run = min(run, za["c"][j])
        out[j] = run
    return out
