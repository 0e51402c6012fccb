"""world_size-2 gloo tests of the multi-GPU host logic (sharding + global convergence)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def test_shard_feeders_partition():
    from revs_admm_b200.parallel import shard_feeders
    sizes = [1000] * 125
    cover = []
    for w in (1, 2, 4, 8):
        parts = [shard_feeders(sizes, w, r) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == len(sizes)
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        homes = [sum(sizes[a:b]) for a, b in parts]
        assert max(homes) - min(homes) <= 1000
    ragged = [5, 4000, 7, 7, 3000, 1, 900]
    parts = [shard_feeders(ragged, 3, r) for r in range(3)]
    assert sorted(sum((list(range(a, b)) for a, b in parts), [])) == list(range(len(ragged)))


class _OracleStepper:
    """Stand-in for _cabi.Solver on a CPU-only box: same begin/step protocol, oracle maths."""

    def __init__(self, trees, hm, cost):
        import revs_oracle as O
        self.O, self.hm, self.cost = O, hm, cost
        self.Rb = [O.rmat_from_tree(t.parent, t.r)[np.ix_(t.res_node, t.res_node)] for t in trees]

    def admm_begin(self, **kw):
        self.kw, self.k, self.prev = kw, 0, None

    def admm_step(self):
        self.k += 1
        hm = self.hm
        out = self.O.solve_ADMM_arrays(self.Rb, load=hm["load"], cost=self.cost, ev_mask=hm["has_ev"].astype(bool),
                                       rating=hm["rating"], capacity=hm["capacity"], initial=hm["initial"],
                                       start=hm["start"], end=hm["end"], kappa=self.kw["kappa"], iter_max=self.k,
                                       vset=self.kw["vset"], vlow=self.kw["vlow"], vhigh=self.kw["vhigh"])
        prev = np.zeros_like(out["P_sch"]) if self.prev is None else self.prev
        sums = np.array([((out["P_est"] - out["P_sch"]) ** 2).sum(), ((out["P_sch"] - prev) ** 2).sum(),
                         float(out["P_sch"].size)])
        self.prev, self.out = out["P_sch"], out
        return sums


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    from revs_admm_b200.feeder import synthetic_feeder, synthetic_homes, synthetic_tariff
    from revs_admm_b200.parallel import run_admm, shard_feeders
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    sizes, T = [30, 22, 41, 17], 24
    lo, hi = shard_feeders(sizes, world, rank)
    trees = [synthetic_feeder(n, seed=i, r_secondary=1e-5) for i, n in enumerate(sizes)]
    off = np.concatenate([[0], np.cumsum(sizes)])
    hm_all = synthetic_homes(int(off[-1]), T, seed=4)
    hm = {k: v[off[lo]:off[hi]] for k, v in hm_all.items()}
    st = _OracleStepper(trees[lo:hi], hm, synthetic_tariff(T))
    iters, hist = run_admm(st, kappa=5.0, iter_max=30, vset=1.0, vlow=0.95, vhigh=1.05, tol=1e-5)
    q.put((rank, lo, hi, iters, hist, st.out["P_sch"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_global_convergence_matches_single_process():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=300) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, it0, h0, P0), (r1, lo1, hi1, it1, h1, P1) = got
    assert it0 == it1 and h0 == h1            # both ranks stop on the same global residuals
    assert lo0 == 0 and hi0 == lo1 and hi1 == 4
    # single process over all feeders: same iteration count, same schedules
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from revs_admm_b200.feeder import synthetic_feeder, synthetic_homes, synthetic_tariff
    from revs_admm_b200.parallel import run_admm
    sizes, T = [30, 22, 41, 17], 24
    trees = [synthetic_feeder(n, seed=i, r_secondary=1e-5) for i, n in enumerate(sizes)]
    st = _OracleStepper(trees, synthetic_homes(sum(sizes), T, seed=4), synthetic_tariff(T))
    iters, hist = run_admm(st, kappa=5.0, iter_max=30, vset=1.0, vlow=0.95, vhigh=1.05, tol=1e-5)
    assert iters == it0 and iters < 30
    assert np.allclose(hist, h0, rtol=1e-9, atol=1e-14)
    assert np.array_equal(st.out["P_sch"], np.vstack([P0, P1]))


def test_row_blocks_partition_the_rows_by_whole_tiles():
    """Row-partitioned contraction of one feeder (revs_reliability_sharded): the blocks of the ranks tile the rows,
    are contiguous, and never cut a contraction tile."""
    from revs_admm_b200.parallel import row_block
    for n_rows, tile, world in [(2100, 64, 8), (5, 64, 4), (64, 64, 2), (1000, 128, 3), (0, 64, 2), (14100, 64, 16)]:
        blocks = [row_block(n_rows, tile, world, r) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n_rows
        for (lo, hi), (lo2, _) in zip(blocks, blocks[1:]):
            assert hi == lo2 and lo <= hi
        for lo, hi in blocks:
            assert lo % tile == 0 and (hi % tile == 0 or hi == n_rows)
        sizes = [hi - lo for lo, hi in blocks]
        assert max(sizes) - min(sizes) <= tile or n_rows < tile * world
