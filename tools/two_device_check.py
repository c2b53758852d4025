"""Several GPUs in ONE process (devices=[0, 1, ...]): the lock-step grid, a tree level with the device tiebreaker and
a whole tree against the single-device results.  usage (2+ GPUs): python tools/two_device_check.py"""
import sys
import numpy as np
sys.path.insert(0, ".")
from kover_b200 import _native
from kover_b200.learning.common.rules import KmerRuleClassifications, LazyKmerRuleList
from kover_b200.learning.experiments.experiment_cart import OccurrenceTiebreaker
from kover_b200.learning.learners.cart import DecisionTreeClassifier
from kover_b200.synthetic import synthetic_block, synthetic_labels, train_split

n_dev = _native.device_count()
assert n_dev >= 2, "needs at least 2 GPUs"
seed, n, k = 13, 1100, 400000
P = synthetic_block(seed, n, 0, k)
y = synthetic_labels(seed, n, sorted_by_label=False)
train = train_split(seed, n)
fold_of = np.random.RandomState(seed).randint(0, 3, size=len(train))
jobs = []
for f in range(4):
    tr = train if f == 3 else train[fold_of != f]
    fp, fn = tr[y[tr] == 1], tr[y[tr] == 0]
    jobs += [(p, a, b) for p in (0.5, 2.0) for a, b in ((fp, fn), (fn, fp))]
ex = {0: train[y[train] == 0], 1: train[y[train] == 1]}
rules = LazyKmerRuleList(["K%d" % i for i in range(k)], np.arange(k))


def run(devices):
    rc = KmerRuleClassifications(P, n, devices=devices)
    grid = rc.best_utility_rules_grid(jobs, util_block_size=99991)
    tb = OccurrenceTiebreaker(rc, train)
    t = DecisionTreeClassifier("gini", 6, 2, {0: 1.0, 1: 1.0})
    t.fit(rules, rc, ex, tiebreaker=tb)
    tb.release()
    sig = [(int(r.kmer_index), float(r.importance)) for r in t.decision_tree.rules]
    counts = rc.sum_rows(train)
    rc.close()
    return grid, sig, counts


one = run([0])
many = run(list(range(n_dev)))
for a, b in zip(one[0], many[0]):
    assert a[0] == b[0] and all(np.array_equal(u, v) for u, v in zip(a[1:], b[1:]))
assert one[1] == many[1] and np.array_equal(one[2], many[2])
print("TWO_DEVICE_CHECK_OK %d devices in one process: grid of %d jobs, %d-split tree, sum_rows identical to one device"
      % (n_dev, len(jobs), len(one[1])))
