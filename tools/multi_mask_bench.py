"""k_count_tma: time one pass over a config-2 shard for the mask pairs of a cross-validation (full training set
+ n folds, positives / negatives as separate masks, genomes sorted by label) against one two-mask scan per
pair, and check the counts against single-mask sum_rows passes.  usage: python tools/multi_mask_bench.py [n k]"""
import os as _os
_os.environ.setdefault("KOVER_B200_PROFILE", "1")   # CUDA events around the scan kernels (info().scan_ms_total)
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from kover_b200 import _native
from kover_b200.synthetic import synthetic_labels, train_split
from kover_b200.utils import build_row_mask

n, k = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2000, 10_000_000)
sh = _native.Shard(0, n, k)
sh.generate_synthetic(72)
W = (n + 63) // 64
gb = W * sh.info().ld * 8 / 1e9


def fold_masks(sorted_labels, n_folds):
    y = synthetic_labels(72, n, sorted_by_label=sorted_labels)
    train = train_split(72, n)
    rs = np.random.RandomState(72)
    fold_of = rs.randint(0, n_folds, size=len(train))
    sets = [train] + [train[fold_of != f] for f in range(n_folds)]
    masks = []
    for ex in sets:
        masks.append(build_row_mask(ex[y[ex] == 1], n, 64))
        masks.append(build_row_mask(ex[y[ex] == 0], n, 64))
    return np.stack(masks)


for sorted_labels in (True, False):
    allm = fold_masks(sorted_labels, 7)
    # reference timing: one fused two-mask scan
    pm, nm_ = allm[0], allm[1]
    sh.scm_best_rules(pm, nm_, 10, 10, 1.0, 1000000)
    sh.scm_best_rules(pm, nm_, 10, 10, 1.0, 1000000)
    t_pair = sh.info().last_kernel_ms
    print("labels %s: two-mask scan %.3f ms (%.0f GB/s)" % ("sorted" if sorted_labels else "interleaved", t_pair, gb / t_pair * 1e3))
    single = {}
    for nm in (2, 4, 6, 8, 12, 16):
        masks = np.ascontiguousarray(allm[:nm])
        out = sh.sum_rows_multi(masks)
        out = sh.sum_rows_multi(masks)
        ms = sh.info().last_kernel_ms
        nz = int((masks != 0).sum())
        print("  %2d masks (%2d pairs): %.3f ms = %.0f GB/s streamed; %.3f ms per pair = %.2fx the separate scans;"
              " %d non-zero (row, mask) words -> %.2f T popc32/s"
              % (nm, nm // 2, ms, gb / ms * 1e3, ms / (nm // 2), t_pair * (nm // 2) / ms, nz,
                 nz * 2.0 * sh.info().ld / ms / 1e9))
        for m in range(nm):
            if m not in single:
                o = np.empty(k, dtype=np.uint32)
                sh.sum_rows(masks[m], o)
                single[m] = o
            assert np.array_equal(out[m], single[m]), (nm, m)
print("counts identical to single-mask passes")
