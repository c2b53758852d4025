"""torchrun --nproc-per-node N tools/exchange_breakdown.py : where the multi-rank SCM step spends its time."""
import os as _os
_os.environ.setdefault("KOVER_B200_PROFILE", "1")   # CUDA events around the scan kernels (info().scan_ms_total)
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch, torch.distributed as dist
import bench
from kover_b200 import _native
from kover_b200.sharding import Comm, merge_packets
from kover_b200.utils import build_row_mask

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
comm = Comm()
n, kl = 2000, 10_000_000
k = kl * world
sh = _native.Shard(lr, n, k, rank * kl, kl); sh.generate_synthetic(72)
steps = bench.step_inputs(n, 72)
nb = _native.n_utility_blocks(k, 1000000); cap = 256
local, gathered = comm.byte_buffers(_native.packet_bytes(nb, cap))
pinned = torch.empty((world, local.numel()), dtype=torch.uint8).pin_memory()
T = {"scan_packet": 0.0, "all_gather(list)": 0.0, "cpu()": 0.0, "merge": 0.0, "all_gather_into_tensor": 0.0, "pinned copy+sync": 0.0, "barrier-only": 0.0}
R = 60
for i in range(R + 5):
    p, pos, neg = steps[i % len(steps)]
    mp, mn = build_row_mask(pos, n, 64), build_row_mask(neg, n, 64)
    t0 = time.perf_counter(); sh.scm_scan_packet(mp, mn, len(pos), len(neg), p, 1000000, cap, local)
    t1 = time.perf_counter(); dist.all_gather(list(gathered.unbind(0)), local)
    t2 = time.perf_counter(); rows = gathered.cpu().numpy()
    t3 = time.perf_counter(); out = merge_packets(rows, nb, cap, p, 1000000)
    t4 = time.perf_counter(); dist.all_gather_into_tensor(gathered.view(-1), local)
    t5 = time.perf_counter(); pinned.copy_(gathered, non_blocking=True); torch.cuda.current_stream().synchronize()
    t6 = time.perf_counter()
    if i >= 5:
        for key, v in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5)):
            T[key] += v
if rank == 0:
    for key, v in T.items():
        print("%-24s %8.1f us" % (key, v / R * 1e6))
    print("kernel %.1f us" % (sh.info().last_kernel_ms * 1e3))
dist.destroy_process_group()
