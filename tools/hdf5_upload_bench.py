"""HDF5 -> HBM: write a Kover-layout kmer_matrix (chunks (1, 100000), gzip, like kl_64.cpp:229-263 / create.py:224-230)
with the built-in writer, then time KmerRuleClassifications(dataset.kmer_matrix, n) -- chunk reads + inflate (thread
pool) + pinned staging + H2D -- and check the resident image.  usage: python tools/hdf5_upload_bench.py [n k gzip]"""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, ".")
from kover_b200.dataset import hdf5_lite
from kover_b200.dataset.hdf5_lite_writer import Writer
from kover_b200.learning.common.rules import KmerRuleClassifications
from kover_b200.synthetic import synthetic_block

n, k, gz = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (2000, 2_000_000, 4)
P = synthetic_block(72, n, 0, k)
path = os.path.join(tempfile.mkdtemp(), "bench.kover")
t0 = time.perf_counter()
with Writer(path) as w:
    w.create_dataset("kmer_matrix", P, chunks=(1, 100000), compression=gz)
t_write = time.perf_counter() - t0
size = os.path.getsize(path)
print("wrote %.2f GB packed as %.2f GB on disk (gzip %d) in %.1f s" % (P.nbytes / 1e9, size / 1e9, gz, t_write))
f = hdf5_lite.File(path)
d = f["kmer_matrix"]
for threads in (1, 4, 8, 16):
    t0 = time.perf_counter()
    got = d._read_box_raw(((0, d.shape[0]), (0, d.shape[1])), threads=threads)
    dt = time.perf_counter() - t0
    print("read + inflate, %2d threads: %.2f s = %.2f GB/s (packed bytes)" % (threads, dt, P.nbytes / dt / 1e9))
assert np.array_equal(got, P)
KmerRuleClassifications(P[:, :4096], n, devices=[0]).close()          # CUDA context + library load out of the timing
t0 = time.perf_counter()
rc = KmerRuleClassifications(d, n, devices=[0])
dt = time.perf_counter() - t0
print("KmerRuleClassifications(hdf5 dataset): %.2f s = %.2f GB/s HDF5 -> HBM" % (dt, P.nbytes / dt / 1e9))
t0 = time.perf_counter()
rc2 = KmerRuleClassifications(P, n, devices=[0])
dt2 = time.perf_counter() - t0
print("KmerRuleClassifications(numpy array) : %.2f s = %.2f GB/s RAM -> HBM" % (dt2, P.nbytes / dt2 / 1e9))
rows = np.arange(0, n, 3)
assert np.array_equal(rc.sum_rows(rows), rc2.sum_rows(rows))
print("resident images agree")
