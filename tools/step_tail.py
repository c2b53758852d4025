"""What a single-shard SCM step costs around its scan kernel at config 2 (2 000 genomes x 10 M k-mers): the step as ONE
kernel (fused tail, results through mapped host memory) against the scan + select + ties launches, cooperative against
plain launch, index arrays against masks.  usage: python tools/step_tail.py [n_kmers]"""
import os
import subprocess
import sys

MODES = [("fused, cooperative launch", "1", "1"), ("fused, plain launch", "1", "0"), ("scan + select + ties", "0", "1")]

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import time
    import numpy as np
    sys.path.insert(0, ".")
    from kover_b200 import _native
    from kover_b200.learning.common.rules import KmerRuleClassifications
    from kover_b200.sharding import ShardGroup
    from kover_b200.synthetic import synthetic_labels, train_split
    n, k = 2000, int(sys.argv[2])
    shard = _native.Shard(0, n, k)
    shard.generate_synthetic(72)

    class Meta(object):
        shape = (32, k)
        dtype = np.dtype(np.uint64)
        chunks = (1, 100000)
    rc = KmerRuleClassifications(Meta(), n, shard_group=ShardGroup([shard], n, k))
    y = synthetic_labels(72, n)
    train = train_split(72, n)
    pos, neg = train[y[train] == 1], train[y[train] == 0]
    mp, mn = rc.row_mask(pos), rc.row_mask(neg)

    def timed(fn, reps=300):
        for _ in range(20):
            fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps * 1e6
    api = timed(lambda: rc.best_utility_rules(1.0, pos, neg))
    abi = timed(lambda: shard.scm_best_rules(mp, mn, len(pos), len(neg), 1.0, 1000000))
    shard.set_profiling(True)
    a = shard.info().scan_ms_total
    for _ in range(50):
        shard.scm_best_rules(mp, mn, len(pos), len(neg), 1.0, 1000000)
    kern = (shard.info().scan_ms_total - a) / 50 * 1e3
    print("  rc.best_utility_rules %7.1f us | C ABI (masks) %7.1f us | kernel (CUDA events) %7.1f us | "
          "kernel share of the API call %.3f" % (api, abi, kern, kern / api))
    sys.exit(0)

k = sys.argv[1] if len(sys.argv) > 1 else "10000000"
for name, fused, coop in MODES:
    print("%s:" % name, flush=True)
    env = dict(os.environ, KOVER_B200_FUSED=fused, KOVER_B200_COOP=coop)
    subprocess.run([sys.executable, __file__, "--child", k], env=env, check=False)
