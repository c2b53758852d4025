"""Three passes of k_count_tma over a config-2 shard with the 12 masks of a 5-fold cross-validation + full training
set (genomes sorted by label unless --interleaved): the command the ncu captures under profiles/ are taken on."""
import os as _os
_os.environ.setdefault("KOVER_B200_PROFILE", "1")   # CUDA events around the scan kernels (info().scan_ms_total)
import sys

import numpy as np

sys.path.insert(0, ".")
from kover_b200 import _native
from kover_b200.synthetic import synthetic_labels, train_split
from kover_b200.utils import build_row_mask

n, k = 2000, 10_000_000
sh = _native.Shard(0, n, k)
sh.generate_synthetic(72)
y = synthetic_labels(72, n, sorted_by_label="--interleaved" not in sys.argv)
train = train_split(72, n)
fold_of = np.random.RandomState(72).randint(0, 5, size=len(train))
masks = []
for ex in [train] + [train[fold_of != f] for f in range(5)]:
    masks += [build_row_mask(ex[y[ex] == 1], n, 64), build_row_mask(ex[y[ex] == 0], n, 64)]
masks = np.stack(masks)
for _ in range(3):
    sh.sum_rows_multi(masks)
    print("k_count_tma, 12 masks: %.3f ms" % sh.info().last_kernel_ms)
