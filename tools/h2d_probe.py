import torch, time
for mb in (16, 64, 256, 1024):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t = time.perf_counter(); n = max(2, 4096 // mb)
    for _ in range(n): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print("H2D pinned %4d MB: %.1f GB/s" % (mb, n * (mb << 20) / dt / 1e9))
    t = time.perf_counter()
    for _ in range(n): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print("D2H pinned %4d MB: %.1f GB/s" % (mb, n * (mb << 20) / dt / 1e9))
import os; print("cpus", os.cpu_count())
