"""Multi-GPU driver: one process per B200, feeders sharded across ranks.

The ADMM loop of the reference couples homes only through their own feeder's sensitivity
block (Utility.network, lpsolver.py:179-190, is block-diagonal over feeders), so whole
feeders are the unit of distribution: every rank owns a contiguous, home-balanced slice of
the feeder list and runs the device loop on it.  The only exchange per iteration is the
global convergence test -- three scalars (residual sums and the home-hour count) summed
with one all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests).  With tol <= 0 (the
reference's fixed iteration count) even that is only needed for reporting.
"""
import numpy as np


def shard_feeders(sizes, world_size, rank):
    """Contiguous slice [lo, hi) of the feeder list for `rank`, balancing homes."""
    sizes = np.asarray(sizes, dtype=np.int64)
    if world_size <= 1:
        return 0, len(sizes)
    cum = np.concatenate([[0], np.cumsum(sizes)])
    total = cum[-1]
    bounds = [int(np.searchsorted(cum, total * r / world_size, side="left")) for r in range(world_size + 1)]
    bounds[0], bounds[-1] = 0, len(sizes)
    for r in range(1, world_size + 1):
        bounds[r] = max(bounds[r], bounds[r - 1])
    return bounds[rank], bounds[rank + 1]


def allreduce_sums(sums, device=None):
    """Sum the per-rank residual sums {sum primal^2, sum dual^2, home-hours} over all ranks."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(sums, dtype=np.float64)
    t = torch.tensor(np.asarray(sums, dtype=np.float64), dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def residuals(total_sums, kappa):
    sp, sd, cnt = total_sums
    return float(np.sqrt(sp / cnt)), float(kappa * np.sqrt(sd / cnt))


def run_admm(stepper, kappa=5.0, iter_max=15, vset=1.0, vlow=0.95, vhigh=1.05, tol=0.0, device=None):
    """Drive `stepper` (a _cabi.Solver, or anything with admm_begin/admm_step) for this
    rank's feeders; stop on the GLOBAL residuals.  Returns (iterations, history)."""
    stepper.admm_begin(kappa=kappa, iter_max=iter_max, vset=vset, vlow=vlow, vhigh=vhigh)
    history = []
    for k in range(iter_max):
        local = stepper.admm_step()
        if tol > 0.0:
            r, s = residuals(allreduce_sums(local, device), kappa)
            history.append((r, s))
            if r < tol and s < tol:
                return k + 1, history
        else:
            history.append(tuple(local[:2]))
    return iter_max, history
