"""RAM -> HBM upload rate of kv_upload_block for the env knobs KOVER_B200_UPLOAD_THREADS / KOVER_B200_UPLOAD_SLOT_MB."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from kover_b200 import _native
n, k = 2000, 10_000_000
sh = _native.Shard(0, n, k)
sh.generate_synthetic(72)
block = sh.read_block(0, sh.n_words, 0, k // 4)          # 640 MB
sh.upload_block(block, 0, 0)
best = 1e9
for _ in range(3):
    t0 = time.perf_counter(); sh.upload_block(block, 0, 0); best = min(best, time.perf_counter() - t0)
print("%.2f GB/s" % (block.nbytes / best / 1e9))
