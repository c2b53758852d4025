"""Fixed cost of one fused SCM step (kv_scm_best_rules: launch + block merge + tie collection + result copy + one
synchronisation) measured on a shard so small that the kernels take microseconds."""
import os as _os
_os.environ.setdefault("KOVER_B200_PROFILE", "1")   # CUDA events around the scan kernels (info().scan_ms_total)
import sys, time
import numpy as np
sys.path.insert(0, ".")
from kover_b200 import _native
from kover_b200.utils import build_row_mask
for n, k in ((64, 2048), (2000, 2048), (2000, 1_000_000)):
    sh = _native.Shard(0, n, k)
    sh.generate_synthetic(72)
    pos, neg = np.arange(0, n // 2), np.arange(n // 2, n)
    mp, mn = build_row_mask(pos, n, 64), build_row_mask(neg, n, 64)
    for _ in range(50):
        sh.scm_best_rules(mp, mn, len(pos), len(neg), 1.0, 1000000)
    t0 = time.perf_counter()
    R = 2000
    for _ in range(R):
        sh.scm_best_rules(mp, mn, len(pos), len(neg), 1.0, 1000000)
    dt = (time.perf_counter() - t0) / R
    t0 = time.perf_counter()
    for _ in range(R):
        sh.scm_scan(mp, mn, len(pos), len(neg), 1.0, 1000000)
    dt2 = (time.perf_counter() - t0) / R
    print("n=%d k=%d: scm_best_rules %.1f us per call (scan kernel %.1f us); scm_scan %.1f us" % (n, k, dt * 1e6, sh.info().last_kernel_ms * 1e3, dt2 * 1e6))
    sh.close()
