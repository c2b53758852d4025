"""Where a single-GPU SCM step spends its time (host pieces timed with perf_counter, kernels with events)."""
import os as _os
_os.environ.setdefault("KOVER_B200_PROFILE", "1")   # CUDA events around the scan kernels (info().scan_ms_total)
import sys, time
import numpy as np
sys.path.insert(0, ".")
import bench
from kover_b200 import _native
from kover_b200.learning.common.rules import KmerRuleClassifications
from kover_b200.sharding import ShardGroup

n, k = 2000, 10_000_000
sh = _native.Shard(0, n, k); sh.generate_synthetic(72)
class Meta: shape=(32,k); dtype=np.dtype(np.uint64); chunks=(1,100000)
rc = KmerRuleClassifications(Meta(), n, shard_group=ShardGroup([sh], n, k))
steps = bench.step_inputs(n, 72)
def t(f, reps=200):
    f(); t0=time.perf_counter()
    for _ in range(reps): f()
    return (time.perf_counter()-t0)/reps*1e6
p,pos,neg = steps[0]
mp, mn = rc.row_mask(pos), rc.row_mask(neg)
print("row_mask x2           %7.1f us" % t(lambda: (rc.row_mask(pos), rc.row_mask(neg))))
print("set_rule_blacklist    %7.1f us" % t(lambda: rc.set_rule_blacklist(())))
print("shard.scm_best_rules  %7.1f us  (kernel %.1f us)" % (t(lambda: sh.scm_best_rules(mp,mn,len(pos),len(neg),p,1000000)), sh.info().last_kernel_ms*1e3))
print("rc.best_utility_rules %7.1f us" % t(lambda: rc.best_utility_rules(p,pos,neg)))
print("shard.scm_scan        %7.1f us" % t(lambda: sh.scm_scan(mp,mn,len(pos),len(neg),p,1000000)))
out=np.empty(k,dtype=np.uint16)
print("shard.sum_rows u16    %7.1f us  (kernel %.1f us)" % (t(lambda: sh.sum_rows(mp,out), 20), sh.info().last_kernel_ms*1e3))
out=np.empty(k,dtype=np.uint32)
print("shard.sum_rows u32    %7.1f us" % t(lambda: sh.sum_rows(mp,out), 20))
print("rc.sum_rows           %7.1f us" % t(lambda: rc.sum_rows(pos), 20))
print("rc.get_columns(int)   %7.1f us" % t(lambda: rc.get_columns(12345), 100))
ex={0:neg,1:pos}
from oracle import kover_oracle as ko
n_total, altered = ko.cart_altered_priors(ex,{0:1.0,1:1.0})
print("rc.gini_best_split    %7.1f us  (count kernel %.1f us)" % (t(lambda: rc.gini_best_split(ex,altered,n_total), 50), sh.info().last_kernel_ms*1e3))
rs = np.random.RandomState(0)
for nm in (1, 3, 4, 8):
    y3 = rs.randint(0, nm, size=n)
    sets = [np.where(y3 == c)[0] for c in range(nm)]
    masks = np.stack([rc.row_mask(r) for r in sets])
    sh.sum_rows_multi(masks[:, :])  # warm
    t0 = time.perf_counter(); sh._lib.kv_sum_rows_multi  # noqa
    out_ = sh.sum_rows_multi(masks)
    dt = time.perf_counter() - t0
    print("k_count_multi<%d>  kernel %.1f us = %.0f GB/s   (call incl. D2H of %d x K u32: %.1f ms)" % (nm, sh.info().last_kernel_ms * 1e3, 2.56 / sh.info().last_kernel_ms * 1e3, nm, dt * 1e3))
ex3 = {c: np.where(rs.randint(0, 3, size=n) == c)[0] for c in range(3)}
nt3 = {c: float(len(v)) for c, v in ex3.items()}
pr3 = {c: nt3[c] / n for c in ex3}
print("rc.gini_best_split 3 classes %7.1f us" % t(lambda: rc.gini_best_split(ex3, pr3, nt3), 20))
