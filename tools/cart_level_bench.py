"""CART fit (Gini, two classes) on a synthetic matrix resident in HBM: level-batched split search (class masks of
8 nodes per matrix pass) against one scan per node.  usage: python tools/cart_level_bench.py [n k max_depth]"""
import os as _os
_os.environ.setdefault("KOVER_B200_PROFILE", "1")   # CUDA events around the scan kernels (info().scan_ms_total)
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from kover_b200 import _native
from kover_b200.learning.common.rules import KmerRuleClassifications, LazyKmerRuleList
from kover_b200.learning.learners.cart import DecisionTreeClassifier
from kover_b200.sharding import ShardGroup
from kover_b200.synthetic import synthetic_labels, train_split

n, k, depth = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (5000, 20_000_000, 10)
sh = _native.Shard(0, n, k)
sh.generate_synthetic(72)


class Meta:
    shape = ((n + 63) // 64, k)
    dtype = np.dtype(np.uint64)
    chunks = (1, 100000)


rc = KmerRuleClassifications(Meta(), n, shard_group=ShardGroup([sh], n, k))
rules = LazyKmerRuleList(np.arange(k), np.arange(k))
y = synthetic_labels(72, n)
train = train_split(72, n)
ex = {0: train[y[train] == 0], 1: train[y[train] == 1]}
occ = rc.sum_rows(train)


def tiebreaker(idx):                      # experiment_cart.py:82-94: keep the most frequent k-mers
    o = occ[idx]
    return idx[np.isclose(o, o.max())]


from kover_b200.learning.experiments.experiment_cart import OccurrenceTiebreaker
device_tiebreaker = OccurrenceTiebreaker(rc, train)      # the same tiebreaker, applied by the level search on the device


def tree_sig(node):
    if node.rule is None:
        return ("leaf", node.n_examples)
    return (int(node.rule_idx), len(node.equivalent_rules_idx), tree_sig(node.left_child), tree_sig(node.right_child))


import cProfile
import pstats
res = {}
DecisionTreeClassifier("gini", 4, 2, {0: 1.0, 1: 1.0}).fit(rules, rc, ex, tiebreaker=tiebreaker)     # warm-up: allocations
for mode in (("level",) if "--level-only" in sys.argv else ("level+device tiebreak",) if "--device-only" in sys.argv
             else ("level+device tiebreak", "level", "per-node")):
    if mode == "per-node":
        rc.gini_best_splits = lambda nodes, pri, tot, bl=(): [rc.gini_best_split(e, pri, tot, bl) for e in nodes]
    t = DecisionTreeClassifier("gini", depth, 2, {0: 1.0, 1: 1.0})
    levels = []
    i0 = sh.info()
    prof = cProfile.Profile() if "--profile" in sys.argv else None
    t0 = time.perf_counter()
    if prof:
        prof.enable()
    t.fit(rules, rc, ex, tiebreaker=device_tiebreaker if "device" in mode else tiebreaker,
          level_callback=lambda infos: levels.append(time.perf_counter()))
    if prof:
        prof.disable()
        pstats.Stats(prof).sort_stats("tottime").print_stats(12)
    dt = time.perf_counter() - t0
    i1 = sh.info()
    n_split = len(t.decision_tree.rules)
    print("%-22s: fit %.3f s, %d split nodes, depth %d, %d matrix passes, %.1f ms in count/scan kernels"
          % (mode, dt, n_split, len(levels), i1.scan_calls - i0.scan_calls, i1.scan_ms_total - i0.scan_ms_total))
    res[mode] = tree_sig(t.decision_tree)
if "per-node" in res:
    assert res["level"] == res["per-node"] == res["level+device tiebreak"]
    print("identical trees")
