"""Where the time of a lock-step grid iteration goes: the 72-job cross-validation grid of bench.py, timed at the
C ABI (kv_scm_queue_run) and through KmerRuleClassifications.best_utility_rules_grid."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench
from kover_b200 import _native
from kover_b200.learning.common.rules import KmerRuleClassifications
from kover_b200.sharding import ShardGroup

n, k = 2000, 10_000_000
sh = _native.Shard(0, n, k)
sh.generate_synthetic(72)


class Meta:
    shape = (32, k)
    dtype = np.dtype(np.uint64)
    chunks = (1, 100000)


rc = KmerRuleClassifications(Meta(), n, shard_group=ShardGroup([sh], n, k))
_, pos, neg = bench.step_inputs(n, 72)[0]
rs = np.random.RandomState(72)
train_all = np.sort(np.concatenate((pos, neg)))
fold_of = rs.randint(0, 5, size=len(train_all))
is_pos = np.isin(train_all, pos)
jobs = []
for f in range(6):
    keep = np.ones(len(train_all), dtype=bool) if f == 5 else fold_of != f
    fp, fn = train_all[keep & is_pos], train_all[keep & ~is_pos]
    jobs += [(pp, a, b) for pp in bench.P_GRID for a, b in ((fp, fn), (fn, fp))]


def timed(fn, reps=10):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    return (time.perf_counter() - t0) / reps * 1e3, r


ms, res = timed(lambda: rc.best_utility_rules_grid(jobs, util_block_size=1000000))
print("best_utility_rules_grid, %d jobs: %.3f ms; tie-set sizes %s" % (len(jobs), ms, sorted(len(r[1]) for r in res)[-5:]))
orig = sh.scm_queue
captured = {}


def spy(q, ubs):
    captured["q"] = q
    return orig(q, ubs)


sh.scm_queue = spy
rc.best_utility_rules_grid(jobs, util_block_size=1000000)
sh.scm_queue = orig
q = captured["q"]
ms, _ = timed(lambda: sh.scm_queue(q, 1000000))
print("Shard.scm_queue (ctypes wrapper)   : %.3f ms" % ms)
nj, r = len(q), sh.QUEUE_SLOT_RECS
masks = np.empty((nj, 2, sh.n_words), dtype=np.uint64)
for j, job in enumerate(q):
    masks[j, 0], masks[j, 1] = job[1], job[2]
n_pos = np.array([job[3] for job in q], dtype=np.int64)
n_neg = np.array([job[4] for job in q], dtype=np.int64)
ps = np.array([job[0] for job in q], dtype=np.float64)
reuse = np.array([job[5] for job in q], dtype=np.int32)
best, found = np.empty(nj), np.empty(nj, dtype=np.int64)
idx, pe, nc = (np.empty((nj, r), dtype=np.int64) for _ in range(3))
P = _native.ptr
ms, _ = timed(lambda: sh._lib.kv_scm_queue_run(sh._h, nj, P(masks), P(n_pos), P(n_neg), P(ps), P(reuse), 1000000, P(best),
                                               P(found), P(idx), P(pe), P(nc)))
print("kv_scm_queue_run (C ABI)           : %.3f ms; launches %d; largest tie set %d" % (ms, sh.info().last_launches, found.max()))
