"""End-to-end leg only, under torchrun at N ranks: ms per schedule() (max over ranks), phase marks of rank 0.
    torchrun ... profiles/exp_e2e_n.py [pipelines]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
import revs_admm_b200 as R
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
try:
    import pynvml
    pynvml.nvmlInit(); pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
except Exception:
    pass
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
K = int(sys.argv[1]) if len(sys.argv) > 1 else 4
trees, hm, cost, sizes, T = bench.make_rank_problem("synthetic-refshape-125k-homes-per-gpu-x96", rank)
H = sum(sizes)
keep, hm_p, out_p = [], {}, {}
for k, v in hm.items():
    hm_p[k], t = bench.pinned_like(v); keep.append(t)
for k, shape, dt in (("P_sch", (H, T), np.float64), ("mask", (H, (T + 63) // 64), np.uint64), ("diff", (15, H), np.float64)):
    out_p[k], t = bench.pinned_like(np.empty(shape, dtype=dt)); keep.append(t)
trace = os.environ.pop("REVS_DEBUG_E2E", None)
s = R.PipelinedSolver(sizes, T, device=local, pipelines=K)
for _ in range(2):
    s.schedule(trees, hm_p, cost, out=out_p, compact=True, **bench.ADMM)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 4
for _ in range(n):
    s.schedule(trees, hm_p, cost, out=out_p, compact=True, **bench.ADMM)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) * 1e3 / n
t = torch.tensor([ms], device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"world {world} K={K}: e2e {t.item():.2f} ms per schedule (max over ranks), rank 0 {ms:.2f}, cores {os.cpu_count()}, affinity {len(os.sched_getaffinity(0))}", flush=True)
    if trace:
        os.environ["REVS_DEBUG_E2E"] = "1"
        s.schedule(trees, hm_p, cost, out=out_p, compact=True, **bench.ADMM)
if world > 1:
    dist.barrier()
s.close()
if world > 1:
    dist.destroy_process_group()
