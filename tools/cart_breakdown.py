import sys, time
import numpy as np
sys.path.insert(0, ".")
import bench
from kover_b200 import _native
from kover_b200.learning.common.rules import KmerRuleClassifications, LazyKmerRuleList
from kover_b200.learning.learners.cart import DecisionTreeClassifier
from kover_b200.sharding import ShardGroup
from kover_b200.synthetic import synthetic_labels, train_split
n, k = 2000, 10_000_000
sh = _native.Shard(0, n, k); sh.generate_synthetic(72)
class Meta: shape=(32,k); dtype=np.dtype(np.uint64); chunks=(1,100000)
rc = KmerRuleClassifications(Meta(), n, shard_group=ShardGroup([sh], n, k))
y = synthetic_labels(72, n); train = train_split(72, n)
ex = {0: train[y[train]==0], 1: train[y[train]==1]}
rules = LazyKmerRuleList(np.arange(k), np.arange(k))
orig = rc.gini_best_split            # (per-node timing: the level-batched search is switched off for this breakdown)
log = []
def timed(*a, **kw):
    t0=time.perf_counter(); r = orig(*a, **kw); log.append((time.perf_counter()-t0, len(r[1]), sum(len(v) for v in a[0].values()))); return r
rc.gini_best_split = timed
rc.gini_best_splits = lambda nodes, pri, tot, bl=(), tie_tables=None: [timed(e, pri, tot, bl) for e in nodes]
t = DecisionTreeClassifier("gini", 5, 2, {0:1.0,1:1.0})
t0=time.perf_counter(); t.fit(rules, rc, ex); dt=time.perf_counter()-t0
print("fit %.1f ms, %d split nodes" % (dt*1e3, len(t.decision_tree.rules)))
print("gini_best_split total %.1f ms" % (sum(l[0] for l in log)*1e3))
for l in log: print("  %.3f ms  ties=%d  node_examples=%d" % (l[0]*1e3, l[1], l[2]))

# steady state of the largest tie sets: repeat the same node scans
big = sorted(log, key=lambda l: -l[1])[:2]
import itertools
def leaves(node):
    return [node] if node.is_leaf else leaves(node.left_child) + leaves(node.right_child)
def all_nodes(node):
    return [node] + ([] if node.is_leaf else all_nodes(node.left_child) + all_nodes(node.right_child))
n_total = {c: float(len(v)) for c, v in ex.items()}
pri = {c: n_total[c] / sum(n_total.values()) for c in ex}
for node in all_nodes(t.decision_tree):
    if node.rule is not None and len(node.equivalent_rules_idx) > 5000:
        for rep in range(3):
            t0 = time.perf_counter(); m_, cand = orig(node.class_examples_idx, pri, n_total); dt = time.perf_counter() - t0
            print("repeat: %.3f ms ties=%d" % (dt * 1e3, len(cand)))
