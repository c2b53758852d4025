"""Where a multi-rank SCM step spends its time beyond the scan (config-2 shards, one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/exchange_breakdown.py
Per rank, averaged over the timed steps: wall time of the step through rc.best_utility_rules, and -- from the device
clock stamps the rank's last CTA leaves in the result header -- the time to publish the packet into the peers'
mailboxes, the time spent waiting for every peer's packet (NVLink latency + how much later the slowest rank finished
its scan: rank skew), and the merge + result publication.  Prints one JSON line per rank."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch
import torch.distributed as dist

from bench import N_GENOMES, N_KMERS_PER_GPU, SEED, UTIL_BLOCK_SIZE, step_inputs
from kover_b200 import _native
from kover_b200.learning.common.rules import KmerRuleClassifications
from kover_b200.sharding import Comm, ShardGroup

rank, lr, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr)
comm = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    comm = Comm()
k_local, k_total = N_KMERS_PER_GPU, N_KMERS_PER_GPU * world
shard = _native.Shard(lr, N_GENOMES, k_total, rank * k_local, k_local)
shard.generate_synthetic(SEED)
group = ShardGroup([shard], N_GENOMES, k_total, comm=comm, rank_slices=[(r * k_local, k_local) for r in range(world)])


class Meta(object):
    shape = ((N_GENOMES + 63) // 64, k_total)
    dtype = np.dtype(np.uint64)
    chunks = (1, 100000)


rc = KmerRuleClassifications(Meta(), N_GENOMES, shard_group=group)
steps = step_inputs(N_GENOMES, SEED)
for i in range(20):
    rc.best_utility_rules(*steps[i % len(steps)], util_block_size=UTIL_BLOCK_SIZE)
n = 200
if comm is not None:
    comm.barrier()
a = shard.info()
t0 = time.perf_counter()
for i in range(n):
    rc.best_utility_rules(*steps[i % len(steps)], util_block_size=UTIL_BLOCK_SIZE)
wall = (time.perf_counter() - t0) / n * 1e6
b = shard.info()
m = max(1, b.exch_steps - a.exch_steps)
shard.set_profiling(True)
c = shard.info()
for i in range(50):
    rc.best_utility_rules(*steps[i % len(steps)], util_block_size=UTIL_BLOCK_SIZE)
kern = (shard.info().scan_ms_total - c.scan_ms_total) / 50 * 1e3
line = {"rank": rank, "world": world, "step_wall_us": round(wall, 1), "kernel_event_us": round(kern, 1),
        "exchange_steps": b.exch_steps - a.exch_steps,
        "publish_us": round((b.exch_publish_us - a.exch_publish_us) / m, 2),
        "wait_for_peers_us": round((b.exch_wait_us - a.exch_wait_us) / m, 2),
        "merge_and_result_us": round((b.exch_merge_us - a.exch_merge_us) / m, 2),
        "p2p": group.p2p_slot_bytes > 0}
lines = [line]
if comm is not None:
    gathered = [None] * world
    dist.all_gather_object(gathered, line)
    lines = gathered
if rank == 0:
    for ln in lines:
        print(json.dumps(ln))
rc.close()
if comm is not None:
    dist.destroy_process_group()
