"""Per-chunk latency of the device inflate (k_inflate: one warp per chunk) by kind of DEFLATE stream: literals only
(Z_HUFFMAN_ONLY), the bench matrix at gzip 1 / 4 / 9, a sparse k-mer-like matrix (long matches), stored blocks.
Run with KOVER_B200_INFLATE_STATS=1 to see the kernel time on stderr.  usage: python tools/inflate_kernel_bench.py [n_chunks]"""
import os
import sys
import time
import zlib

import numpy as np

sys.path.insert(0, ".")
if not os.environ.get("KV_NO_STATS"):                                # (the phase report serialises the slabs)
    os.environ.setdefault("KOVER_B200_INFLATE_STATS", "1")
from kover_b200 import _native
from kover_b200.synthetic import synthetic_block

n_chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 640
cols = 100_000
rows = synthetic_block(72, 64 * 16, 0, cols)                      # 16 distinct chunk rows of the bench matrix
rs = np.random.RandomState(0)
sparse = np.where(rs.rand(16, cols) < 0.02, rows, np.uint64(0))   # ~2 % of the words non-zero


def streams(data, level=4, strategy=zlib.Z_DEFAULT_STRATEGY):
    out = []
    for r in range(data.shape[0]):
        c = zlib.compressobj(level, zlib.DEFLATED, 15, 8, strategy)
        out.append(c.compress(data[r].tobytes()) + c.flush())
    return out


cases = [("bench matrix, gzip 4", streams(rows, 4)), ("bench matrix, gzip 1", streams(rows, 1)),
         ("bench matrix, gzip 9", streams(rows, 9)), ("bench matrix, literals only", streams(rows, 6, zlib.Z_HUFFMAN_ONLY)),
         ("bench matrix, stored", streams(rows, 0)), ("sparse (2 % non-zero words), gzip 4", streams(sparse, 4))]
only = os.environ.get("KV_CASES")                                 # e.g. KV_CASES="sparse" or "gzip 4,literals"
if only:
    cases = [c for c in cases if any(k.strip() in c[0] for k in only.split(","))]
shard = _native.Shard(0, 64 * n_chunks, cols)
for name, ss in cases:
    bufs = [np.frombuffer(s + b"\x00" * 8, dtype=np.uint8).copy() for s in ss]
    addr = np.array([bufs[i % 16].ctypes.data for i in range(n_chunks)], dtype=np.uint64)
    nbytes = np.array([len(ss[i % 16]) for i in range(n_chunks)], dtype=np.int64)
    ratio = sum(len(s) for s in ss) / (16 * cols * 8.0)
    print("%-40s deflated/packed = %.3f" % (name, ratio), flush=True)
    for rep in range(3):
        t0 = time.perf_counter()
        st = shard.upload_deflate(64, addr, nbytes, np.arange(n_chunks), np.zeros(n_chunks, dtype=np.int64), 1, cols)
        dt = time.perf_counter() - t0
        assert not st.any(), st
    print("    call %.1f ms = %.2f GB/s of packed matrix" % (dt * 1e3, n_chunks * cols * 8 / dt / 1e9), flush=True)
shard.close()
