import torch, time
for mb in (16, 96, 288):
    n = mb * 1024 * 1024 // 8
    d = torch.empty(n, dtype=torch.float64, device="cuda"); h = torch.empty(n, dtype=torch.float64).pin_memory()
    for name, fn in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
        for _ in range(2): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): fn()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
        print(f"{name} {mb} MB: {mb / 1024 / dt:.1f} GiB/s ({dt * 1e3:.2f} ms)")
