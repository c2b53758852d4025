"""HDF5 -> HBM: write a Kover-layout kmer_matrix (chunks (1, 100000), gzip, like kl_64.cpp:229-263 / create.py:224-230)
with the test writer, then time KmerRuleClassifications(dataset.kmer_matrix, n) with the chunks inflated ON THE DEVICE
(kv_upload_deflate) and on the host cores (KOVER_B200_DEVICE_INFLATE=0: chunk reads + zlib thread pool + pinned
staging + H2D), and check the resident images.  usage: python tools/hdf5_upload_bench.py [n k gzip [density]]"""
import os
import sys
import tempfile
import time

import numpy as np

sys.path[:0] = [".", "tests"]
from hdf5_lite_writer import Writer
from kover_b200.dataset import hdf5_lite
from kover_b200.learning.common.rules import KmerRuleClassifications
from kover_b200.synthetic import synthetic_block

n, k, gz = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (2000, 2_000_000, 4)
P = synthetic_block(72, n, 0, k)
if len(sys.argv) > 4:                          # sparser than the bench matrix: AND with random masks of that density
    rs = np.random.RandomState(1)
    keep = rs.rand(*P.shape) < float(sys.argv[4])
    P = np.where(keep, P, np.uint64(0))
path = os.path.join(tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None), "bench.kover")
t0 = time.perf_counter()
with Writer(path) as w:
    w.create_dataset("kmer_matrix", P, chunks=(1, 100000), compression=gz)
t_write = time.perf_counter() - t0
size = os.path.getsize(path)
print("wrote %.2f GB packed as %.2f GB on disk (gzip %d) in %.1f s" % (P.nbytes / 1e9, size / 1e9, gz, t_write))
KmerRuleClassifications(P[:, :4096], n, devices=[0]).close()          # CUDA context + library load out of the timing
images = {}
for mode, label in (("1", "device inflate"), ("0", "host inflate  "), ("1", "device inflate")):
    os.environ["KOVER_B200_DEVICE_INFLATE"] = mode
    f = hdf5_lite.File(path)
    d = f["kmer_matrix"]
    if os.environ.get("KOVER_B200_INFLATE_STATS"):          # where the time of the constructor goes
        from kover_b200 import _native
        t0 = time.perf_counter()
        sh_probe = _native.Shard(0, n, k)
        t_create = time.perf_counter() - t0
        sh_probe.close()
        print("  (kv_create of a %d x %d shard alone: %.1f ms)" % (n, k, t_create * 1e3), flush=True)
    t0 = time.perf_counter()
    rc = KmerRuleClassifications(d, n, devices=[0])
    dt = time.perf_counter() - t0
    print("KmerRuleClassifications(hdf5 dataset), %s: %.3f s = %.2f GB/s HDF5 -> HBM (packed bytes)"
          % (label, dt, P.nbytes / dt / 1e9), flush=True)
    sh = rc._group.shards[0]
    images[mode] = sh.read_block(0, sh.n_words, 0, min(k, 300_000))
    rc.close()
    f.close()
assert np.array_equal(images["1"], P[:, :images["1"].shape[1]]) and np.array_equal(images["0"], images["1"])
print("resident images agree with the matrix that was written")
os.remove(path)
