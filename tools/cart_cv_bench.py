"""BASELINE configs[2] through the experiment layer: cost-complexity pruning with 5-fold cross-validation
(experiment_cart.py:297-436) -- five fold trees + the master tree grown in lock-step over the resident matrix --
and bound-based pruning (:208-294).  usage: python tools/cart_cv_bench.py [n k max_depth]"""
import os as _os
_os.environ.setdefault("KOVER_B200_PROFILE", "1")   # CUDA events around the scan kernels (info().scan_ms_total)
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from kover_b200 import _native
from kover_b200.learning.common.rules import KmerRuleClassifications, LazyKmerRuleList
from kover_b200.learning.experiments import experiment_cart as cexp
from kover_b200.sharding import ShardGroup
from kover_b200.synthetic import train_split

n, k, depth = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (5000, 20_000_000, 10)
sh = _native.Shard(0, n, k)
sh.generate_synthetic(72)


class Meta:
    shape = ((n + 63) // 64, k)
    dtype = np.dtype(np.uint64)
    chunks = (1, 100000)


rc = KmerRuleClassifications(Meta(), n, shard_group=ShardGroup([sh], n, k))
rules = LazyKmerRuleList(np.arange(k), np.arange(k))
cols = rc.get_columns([11, 4242, 99991])
rs = np.random.RandomState(72)
labels = ((cols[:, 0] == 1) & (cols[:, 1] == 0) | (cols[:, 2] == 1)).astype(np.uint8)      # planted phenotype + noise
flip = rs.rand(n) < 0.05
labels[flip] = 1 - labels[flip]
train = train_split(72, n)
fold_of = rs.randint(0, 5, size=len(train))
folds = [(train[fold_of != f], train[fold_of == f]) for f in range(5)]
hps = {"criterion": "gini", "max_depth": depth, "min_samples_split": 2, "class_importance": {0: 1.0, 1: 1.0}}
for rep in range(2):
    i0 = sh.info()
    t0 = time.perf_counter()
    out_hps, score, tree = cexp.learn_pruned_tree_cv(rc, rules, labels, train, folds, 2, hps, [])
    dt = time.perf_counter() - t0
    i1 = sh.info()
    print("CV pruning  : %.3f s, 6 trees in lock-step, %d matrix passes, %.1f ms of count/scan kernels; CV risk %.4f, "
          "alpha %.5f, %d rules kept" % (dt, i1.scan_calls - i0.scan_calls, i1.scan_ms_total - i0.scan_ms_total, score,
                                         out_hps["pruning_alpha"], len(tree.rules)))
    t0 = time.perf_counter()
    out_hps, bound, tree, _ = cexp.learn_pruned_tree_bound(rc, rules, labels, train, 2, hps, 0.05, 5000000, [])
    dt = time.perf_counter() - t0
    i2 = sh.info()
    print("bound pruning: %.3f s, %d matrix passes; bound %.4f, %d rules kept" % (dt, i2.scan_calls - i1.scan_calls, bound,
                                                                                 len(tree.rules)))
