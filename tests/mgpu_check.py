#!/usr/bin/env python
"""Multi-GPU check of the global stopping rule and of the row-partitioned reliability check (run on a box with
>= 2 B200s; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py

Every rank owns different feeders.  (1) Reference run: iterations one at a time, local residual sums
all-reduced on the host side with NCCL (parallel.run_admm).  (2) Peer-memory run: attach_peers, then ONE
captured loop per rank with tol > 0 -- the all-reduce happens inside dual_update_kernel over NVLink.  Both
must stop after the same number of iterations on every rank, with identical schedules; the per-iteration
global sums of a stepped peer run must equal the NCCL all-reduce of the local sums to rounding.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import revs_admm_b200 as R  # noqa: E402
from revs_admm_b200.feeder import synthetic_feeder, synthetic_homes, synthetic_tariff  # noqa: E402
from revs_admm_b200.parallel import allreduce_sums, attach_gather, attach_peers, residuals, run_admm  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    T = 24
    sizes = [60 + 7 * rank, 45, 33 + rank]
    # loose networks with a different resistance per rank: ranks would stop at different iterations on their own
    trees = [synthetic_feeder(n, seed=100 * rank + i, r_secondary=(1e-5 if rank % 2 else 5e-5)) for i, n in enumerate(sizes)]
    hm = synthetic_homes(sum(sizes), T, seed=rank)
    cost = synthetic_tariff(T)
    kw = dict(kappa=5.0, iter_max=60, vset=1.0, vlow=0.95, vhigh=1.05)
    tol = 1e-5

    def fresh():
        s = R.Solver(sizes, T, device=local)
        s.set_feeder_trees(trees)
        s.set_homes(**hm)
        s.set_tariff(cost)
        return s

    with fresh() as s:                                   # own stopping rule (no exchange)
        own = s.solve_admm(tol=tol, **kw)
    with fresh() as s:                                   # (1) host-side NCCL all-reduce per iteration
        it_ref, hist_ref = run_admm(s, tol=tol, device=dev, **kw)
        ref = s.results(it_ref)
    with fresh() as s:                                   # (2) peer mailboxes, captured loop
        assert attach_peers(s)
        it_peer = s.solve_admm(tol=tol, **kw)
        peer = s.results(it_peer)
        st = s.stats()
        # stepped run with peers: the sums returned are global
        s.admm_begin(**kw)
        sums = [s.admm_step() for _ in range(3)]
        s.comm_detach()
    with fresh() as s:
        s.admm_begin(**kw)
        loc = [allreduce_sums(s.admm_step(), dev) for _ in range(3)]
    # (3) one feeder's reliability check with the rows partitioned over the ranks, all-gather fused into the contraction
    # kernel (revs_reliability_sharded): every rank must end with the complete result of the single-GPU check
    big = synthetic_feeder(1500, seed=4242, r_secondary=2e-4)          # the same feeder on every rank
    Pb = np.random.default_rng(7).random((1500, T)) * 3.0
    rows = np.arange(big.n_nodes, dtype=np.int32)
    with R.Solver([1500], T, device=local) as s:
        s.set_feeder_tree(0, big.parent, big.r, big.res_node)
        full_v = s.reliability(0, R.REVS_REL_VOLTAGE, rows, vset=1.03, P=Pb)
        full_f = s.reliability(0, R.REVS_REL_FLOW, rows, P=Pb)
        assert attach_gather(s, len(rows) * T) == world
        shard_ok = True
        for rep in range(3):                                            # repeated exchanges (alternating payload halves)
            sh_v = s.reliability_sharded(0, R.REVS_REL_VOLTAGE, rows, Pb * (1.0 + 0.0 * rep), vset=1.03)
            sh_f = s.reliability_sharded(0, R.REVS_REL_FLOW, rows, Pb)
            shard_ok = shard_ok and np.array_equal(sh_v, full_v) and np.array_equal(sh_f, full_f)
        few = s.reliability_sharded(0, R.REVS_REL_DROP, rows[:5], Pb)   # fewer tiles than ranks: some ranks only signal and wait
        shard_ok = shard_ok and np.array_equal(few, s.reliability(0, R.REVS_REL_DROP, rows[:5], P=Pb))
        dist.barrier()
    its = [torch.zeros(3, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(its, torch.tensor([own, it_ref, it_peer], dtype=torch.int64, device=dev))
    its = [t.tolist() for t in its]
    checks = [all(t[1] == its[0][1] and t[2] == its[0][1] for t in its),
              all(np.array_equal(ref[k], peer[k]) for k in ("P_sch", "P_ev", "diff")),
              all(np.allclose(a, b, rtol=1e-12, atol=1e-18) for a, b in zip(sums, loc)),      # (iteration 2 of a loose network: sums at rounding level)
              bool(shard_ok)]
    ok = all(checks)
    det = [torch.zeros(4, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(det, torch.tensor([int(c) for c in checks], dtype=torch.int64, device=dev))
    det = [t.tolist() for t in det]
    r, d = residuals(sums[-1], kw["kappa"])
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"mgpu_check": "ok" if int(flag.item()) else "FAILED", "world": world,
                          "iterations_per_rank [own rule, nccl all-reduce, peer mailboxes]": its,
                          "residuals_after_3_iterations": [r, d], "final_residuals": [st["primal_residual"], st["dual_residual"]],
                          "checks_per_rank [same iterations, same schedules, global sums == nccl sums, row-partitioned reliability == single GPU]": det}))
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
