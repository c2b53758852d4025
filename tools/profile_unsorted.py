"""A few SCM steps on config 2 with interleaved labels (both masks touch every packed row): the regime where the scan is
bound by popcounts rather than HBM.  For ncu: ncu --set full -k regex:k_scan_tma -s 3 -c 1 python tools/profile_unsorted.py"""
import sys

import numpy as np

sys.path.insert(0, ".")
from kover_b200 import _native
from kover_b200.learning.common.rules import KmerRuleClassifications
from kover_b200.sharding import ShardGroup
from kover_b200.synthetic import synthetic_labels, train_split

n, k = 2000, 10_000_000
sh = _native.Shard(0, n, k)
sh.generate_synthetic(72)


class Meta:
    shape = (32, k)
    dtype = np.dtype(np.uint64)
    chunks = (1, 100000)


rc = KmerRuleClassifications(Meta(), n, shard_group=ShardGroup([sh], n, k))
y = synthetic_labels(72, n, sorted_by_label=False)
train = train_split(72, n)
pos, neg = train[y[train] == 1], train[y[train] == 0]
for _ in range(6):
    r = rc.best_utility_rules(1.0, pos, neg, util_block_size=1000000)
print(r[0], len(r[1]), sh.info().last_kernel_ms)
