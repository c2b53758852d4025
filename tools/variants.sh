for v in ${VARIANTS:-0 1 2 3 4 5 6 7 8 9 10 11 12 13 14 15}; do
  KOVER_B200_SCAN_VARIANT=$v python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0, '.')
import bench
from kover_b200 import _native
n, k = 2000, 10_000_000
sh = _native.Shard(0, n, k); sh.generate_synthetic(72)
for sorted_labels in (True, False):
    steps = bench.step_inputs(n, 72, sorted_by_label=sorted_labels)
    from kover_b200.utils import build_row_mask
    ms = []
    pre = [(p, build_row_mask(pos, n, 64), build_row_mask(neg, n, 64), len(pos), len(neg)) for p, pos, neg in steps[:12]]
    for i in range(12):
        p, mp, mn, npos, nneg = pre[i]
        out = sh.scm_best_rules(mp, mn, npos, nneg, p, 1000000)
        if i >= 3: ms.append(sh.info().last_kernel_ms)
    sh.timer_start()
    for i in range(12):
        p, mp, mn, npos, nneg = pre[i]
        out = sh.scm_best_rules(mp, mn, npos, nneg, p, 1000000)
    tot = sh.timer_stop() / 12
    print("variant", os.environ["KOVER_B200_SCAN_VARIANT"], "sorted" if sorted_labels else "general", "kernel %.4f ms  %.1f GB/s  step %.4f ms" % (np.mean(ms), 2.56 / np.mean(ms) * 1e3, tot), flush=True)
PY
done
