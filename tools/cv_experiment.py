"""BASELINE configs[1] end to end on one GPU: p grid {0.1,0.5,1,2,5,10} x {conjunction, disjunction}, 5-fold
cross-validation + the bound selection on the full training set, over a synthetic 2 000 x 10 M matrix resident
in HBM (experiment_scm.py:196-248, 568-629).  Prints where the wall time goes: matrix passes, re-scores, host.
usage: python tools/cv_experiment.py [n_genomes n_kmers max_rules]"""
import os as _os
_os.environ.setdefault("KOVER_B200_PROFILE", "1")   # CUDA events around the scan kernels (info().scan_ms_total)
import cProfile
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from kover_b200 import _native
from kover_b200.learning.common.rules import KmerRuleClassifications, LazyKmerRuleList
from kover_b200.learning.experiments.experiment_scm import bound_selection, cross_validation
from kover_b200.sharding import ShardGroup
from kover_b200.synthetic import synthetic_labels, train_split

n, k, max_rules = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (2000, 10_000_000, 10)
profile = "--profile" in sys.argv
sh = _native.Shard(0, n, k)
sh.generate_synthetic(72)


class Meta:
    shape = ((n + 63) // 64, k)
    dtype = np.dtype(np.uint64)
    chunks = (1, 100000)


rc = KmerRuleClassifications(Meta(), n, shard_group=ShardGroup([sh], n, k))
rules = LazyKmerRuleList(np.arange(k), np.arange(k))
# phenotype planted on three mid-density k-mers + label noise, genomes sorted by label like create.py:190-194
# (the synthetic matrix itself is not re-ordered: the labels are a function of matrix columns, then the example
# indices are what the learners see; sorting would need a row permutation of the resident matrix)
cols = rc.get_columns([11, 4242, 99991])
rs = np.random.RandomState(72)
labels = ((cols[:, 0] == 1) & (cols[:, 1] == 0) | (cols[:, 2] == 1)).astype(np.uint8)
flip = rs.rand(n) < 0.05
labels[flip] = 1 - labels[flip]
if "--sorted" in sys.argv:
    labels = synthetic_labels(72, n, sorted_by_label=True)
train = train_split(72, n)
test = np.setdiff1d(np.arange(n), train)
n_folds = 5
fold_of = rs.randint(0, n_folds, size=len(train))
p_values = [0.1, 0.5, 1.0, 2.0, 5.0, 10.0]
model_types = ["conjunction", "disjunction"]


def risk_table(train_idx):
    pos, neg = train_idx[labels[train_idx] == 1], train_idx[labels[train_idx] == 0]
    _, by_kmer, by_anti = rc.kmer_risk_tables(pos, neg)
    return np.hstack((by_kmer, by_anti))


t0 = time.perf_counter()
folds = []
for f in range(n_folds):
    tr, te = train[fold_of != f], train[fold_of == f]
    folds.append((tr, te, risk_table(tr)))
full_risks = risk_table(train)
t_risk = time.perf_counter() - t0
i0 = sh.info()

prof = cProfile.Profile() if profile else None
t0 = time.perf_counter()
if prof:
    prof.enable()
score, hp, per_hp = cross_validation(rc, rules, labels, folds, model_types, p_values, max_rules, [])
if prof:
    prof.disable()
t_cv = time.perf_counter() - t0
i1 = sh.info()
t0 = time.perf_counter()
b_score, b_hp, model, _, _, _ = bound_selection(rc, rules, labels, train, full_risks, model_types, p_values, max_rules,
                                                10000, [], 0.05, 5000000, np.random.RandomState(42))
t_bound = time.perf_counter() - t0
i2 = sh.info()
print("positives %d / %d; risk tables of %d folds + train: %.3f s" % (labels.sum(), n, n_folds, t_risk))
print("cross-validation: %d fits (12 HP x %d folds), %.3f s; %d matrix passes (%.1f ms of scan kernels), %d re-scores; "
      "best HP %s score %.5f" % (12 * n_folds, n_folds, t_cv, i1.scan_calls - i0.scan_calls,
                                 i1.scan_ms_total - i0.scan_ms_total, i1.rescore_calls - i0.rescore_calls, hp, score))
print("bound selection : 12 fits, %.3f s; %d matrix passes, %d re-scores; best HP %s bound %.5f; model %s"
      % (t_bound, i2.scan_calls - i1.scan_calls, i2.rescore_calls - i1.rescore_calls, b_hp, b_score,
         " ".join(str(r) for r in model)))
if prof:
    pstats.Stats(prof).sort_stats("cumulative").print_stats(25)
