"""GPU: HDF5 -> HBM with the gzip chunks inflated ON THE DEVICE (kv_upload_deflate, csrc/kv_inflate.cuh) against the
host path (hdf5_lite + zlib + kv_upload_block) and against the matrix that was written: the resident image read back
with kv_read_block must be bit-identical.  Streams of every DEFLATE block type (stored, fixed Huffman, dynamic
Huffman at levels 1 / 4 / 9), 64- and 32-bit packing, chunk shapes incl. Kover's (1, 100000), edge chunks, shard
boundaries inside a chunk, highly compressible rows (long overlapping matches) and incompressible rows; corrupt
streams must surface as errors through the host path.  Reference: rules.py:247-262 reads + inflates the same chunks
through h5py on every sum_rows call; create.py:222-230 / kl_64.cpp:229-263 define the chunking."""
import os
import zlib

import numpy as np
import pytest

from hdf5_lite_writer import Writer
from kover_b200.dataset import hdf5_lite
from kover_b200.learning.common.rules import KmerRuleClassifications

pytestmark = pytest.mark.gpu


def matrix(seed, n_rows, n_cols, dtype, kind):
    rs = np.random.RandomState(seed)
    bits = np.dtype(dtype).itemsize * 8
    if kind == "random":
        hi = rs.randint(0, 2 ** 32, size=(n_rows, n_cols), dtype=np.uint64)
        m = (hi << np.uint64(32) | rs.randint(0, 2 ** 32, size=(n_rows, n_cols), dtype=np.uint64)) if bits == 64 else hi
    elif kind == "two_values":                   # bytes 0x00 (90 %) / 0xFF, and a few others: very short Huffman codes
        b = np.where(rs.rand(n_rows, n_cols * (bits // 8)) < 0.9, 0, 255).astype(np.uint8)
        b[rs.rand(*b.shape) < 0.01] = 7
        m = b.view(np.uint64 if bits == 64 else np.uint32).astype(np.uint64)
    elif kind == "sparse":                       # k-mer-like: most words zero, runs of identical words
        m = np.zeros((n_rows, n_cols), dtype=np.uint64)
        idx = rs.rand(n_rows, n_cols) < 0.03
        m[idx] = rs.randint(0, 2 ** 16, size=int(idx.sum())).astype(np.uint64) << np.uint64(rs.randint(0, bits - 16))
        m[:, n_cols // 3:n_cols // 3 + 900] = np.uint64(0x0102030405060708 & (2 ** bits - 1))
    else:                                        # "mixed": rows of different character
        m = np.zeros((n_rows, n_cols), dtype=np.uint64)
        for r in range(n_rows):
            sub = matrix(seed * 100 + r, 1, n_cols, np.uint64 if bits == 64 else np.uint32, "random" if r % 3 == 0 else "sparse")
            m[r] = sub[0]
        m[n_rows // 2] = 0                        # an all-zero chunk row
    return (m & np.uint64(2 ** bits - 1)).astype(dtype)


def raw_deflate(strategy=zlib.Z_DEFAULT_STRATEGY, level=6):
    def make(data):
        c = zlib.compressobj(level, zlib.DEFLATED, 15, 8, strategy)
        return c.compress(data) + c.flush()
    return make


def resident(rc, n_words):
    return np.hstack([s.read_block(0, n_words, 0, s.n_cols) for s in rc._group.shards])


def expect64(M, pack_bits):
    if pack_bits == 64:
        return M
    rows = M.shape[0] + (M.shape[0] & 1)
    P = np.zeros((rows, M.shape[1]), dtype=np.uint64)
    P[:M.shape[0]] = M
    return (P[0::2] << np.uint64(32)) | P[1::2]


CASES = [
    # id, dtype, rows, cols, chunks, compression, data kind
    ("kover_1x100000_gz4", np.uint64, 5, 250_123, (1, 100_000), 4, "mixed"),
    ("gz1_sparse", np.uint64, 7, 40_000, (1, 10_000), 1, "sparse"),
    ("gz9_mixed", np.uint64, 6, 30_011, (1, 4_096), 9, "mixed"),
    ("stored_blocks", np.uint64, 4, 20_000, (1, 7_000), 0, "random"),
    ("fixed_huffman", np.uint64, 4, 20_000, (1, 7_000), raw_deflate(zlib.Z_FIXED), "mixed"),
    ("huffman_only", np.uint64, 4, 20_000, (1, 7_000), raw_deflate(zlib.Z_HUFFMAN_ONLY), "mixed"),
    ("rle", np.uint64, 4, 20_000, (1, 7_000), raw_deflate(zlib.Z_RLE), "sparse"),
    ("multi_row_chunks", np.uint64, 9, 12_345, (2, 700), 4, "mixed"),
    ("pack32_gz4", np.uint32, 11, 33_333, (1, 5_000), 4, "mixed"),
    ("pack32_multi_row", np.uint32, 8, 9_999, (3, 1_000), 6, "sparse"),
    # two byte values only, no matches: 1- and 2-bit literal codes, up to 32 symbols inside one 32-bit window
    ("skewed_literals", np.uint64, 3, 16_000, (1, 8_000), raw_deflate(zlib.Z_HUFFMAN_ONLY), "two_values"),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("devices", [[0], [0, 0, 0]], ids=["1shard", "3shards"])
@pytest.mark.parametrize("decoder", ["4", "3", "2", "1"], ids=["all_symbols_speculative", "two_codes_per_load", "warp_parallel_literals", "single_lane"])
def test_device_inflate_matches_host_path(tmp_path, monkeypatch, case, devices, decoder):
    name, dtype, n_rows, n_cols, chunks, compression, kind = case
    monkeypatch.setenv("KOVER_B200_INFLATE_V", decoder)
    pack_bits = np.dtype(dtype).itemsize * 8
    M = matrix(len(name), n_rows, n_cols, dtype, kind)
    path = str(tmp_path / "m.h5")
    with Writer(path) as w:
        w.create_dataset("kmer_matrix", M, chunks=chunks, compression=compression)
    n_genomes = n_rows * pack_bits - 3                        # ragged last packed row
    want = expect64(M, pack_bits)
    images = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("KOVER_B200_DEVICE_INFLATE", mode)
        f = hdf5_lite.File(path)
        d = f["kmer_matrix"]
        assert d.deflate_chunk_table(0, n_rows, 0, n_cols) is not None
        rc = KmerRuleClassifications(d, n_genomes, devices=devices)
        launches = sum(s.info().total_launches for s in rc._group.shards)
        images[mode] = resident(rc, want.shape[0])
        rc.close()
        f.close()
        if mode == "1":
            assert launches >= len(devices)                   # the inflate kernel ran on every shard
    assert np.array_equal(images["1"], want)
    assert np.array_equal(images["0"], want)


def test_h5py_like_dataset_takes_the_device_inflate_path(tmp_path, monkeypatch):
    """A dataset object in h5py's style (shape / dtype / chunks / name / file.filename, slicing inflates on the host) --
    what a Kover installation passes to KmerRuleClassifications -- is uploaded through the raw chunks of its file and
    the device decoder, without a single read through the object itself; with the device path switched off the same
    object is read slab by slab.  Both images equal the written matrix."""
    from h5py_like import H5pyLikeDataset
    M = matrix(5, 6, 120_000, np.uint64, "mixed")
    path = str(tmp_path / "m.h5")
    with Writer(path) as w:
        w.create_dataset("kmer_matrix", M, chunks=(1, 50_000), compression=4)
    n_genomes = 6 * 64 - 5
    for mode, devices in (("1", [0]), ("1", [0, 0, 0]), ("0", [0])):
        monkeypatch.setenv("KOVER_B200_DEVICE_INFLATE", mode)
        like = H5pyLikeDataset(path, "kmer_matrix")
        rc = KmerRuleClassifications(like, n_genomes, devices=devices)
        assert (like.reads == 0) == (mode == "1")
        if mode == "1":
            assert all(s.upload_deflate_stats()["total_ms"] > 0 for s in rc._group.shards)
        assert np.array_equal(resident(rc, 6), M)
        rc.close()


def test_corrupt_stream_is_reported_through_the_host_path(tmp_path):
    """A flipped byte inside one chunk: the device marks the chunk (bad code / distance / checksum), the host path
    re-inflates it and raises zlib's own error; the other chunks are unaffected."""
    M = matrix(5, 4, 30_000, np.uint64, "mixed")
    path = str(tmp_path / "c.h5")
    with Writer(path) as w:
        w.create_dataset("kmer_matrix", M, chunks=(1, 10_000), compression=4)
    f = hdf5_lite.File(path)
    tab = f["kmer_matrix"].deflate_chunk_table(0, 4, 0, 30_000)
    victim = 5
    offset = int(tab["address"][victim] - f._buf.ctypes.data) + int(tab["nbytes"][victim]) // 2
    f.close()
    data = bytearray(open(path, "rb").read())
    data[offset] ^= 0x5A
    open(path, "wb").write(bytes(data))
    f = hdf5_lite.File(path)
    with pytest.raises(zlib.error):
        KmerRuleClassifications(f["kmer_matrix"], 4 * 64, devices=[0])
    f.close()


@pytest.mark.parametrize("decoder", ["4", "3", "2", "1"], ids=["all_symbols_speculative", "two_codes_per_load", "warp_parallel_literals", "single_lane"])
def test_status_codes_of_malformed_streams(monkeypatch, decoder):
    """kv_upload_deflate's per-chunk status on hand-made streams (C ABI level): good, truncated, wrong checksum,
    wrong size, bad header."""
    from kover_b200 import _native
    monkeypatch.setenv("KOVER_B200_INFLATE_V", decoder)
    n_cols = 4096
    shard = _native.Shard(0, 64 * 8, n_cols)
    rs = np.random.RandomState(1)
    rows = rs.randint(0, 2 ** 62, size=(8, n_cols), dtype=np.int64).astype(np.uint64)
    rows[1] = 7
    good = [zlib.compress(rows[r].tobytes(), 4) for r in range(8)]
    streams = list(good)
    streams[2] = good[2][:len(good[2]) // 2]                                      # truncated
    streams[3] = good[3][:-1] + bytes([good[3][-1] ^ 1])                          # Adler-32 mismatch
    streams[4] = zlib.compress(rows[4, :100].tobytes(), 4)                        # inflates to fewer bytes
    streams[5] = b"\x79" + good[5][1:]                                            # bad CMF / FLG check
    streams[6] = zlib.compress(np.concatenate((rows[6], rows[6])).tobytes(), 4)   # inflates to more bytes
    bufs = [np.frombuffer(s + b"\x00" * 8, dtype=np.uint8).copy() for s in streams]
    addr = np.array([b.ctypes.data for b in bufs], dtype=np.uint64)
    nbytes = np.array([len(s) for s in streams], dtype=np.int64)
    status = shard.upload_deflate(64, addr, nbytes, np.arange(8), np.zeros(8, dtype=np.int64), 1, n_cols)
    assert status[0] == 0 and status[1] == 0 and status[7] == 0
    assert all(status[i] != 0 for i in (2, 3, 4, 5, 6)), status
    img = shard.read_block(0, 8, 0, n_cols)
    for r in (0, 1, 7):
        assert np.array_equal(img[r], rows[r])
    shard.close()


def _fuzz_row(rs, n_bytes):
    """One chunk of a random character: the byte statistics and match structure DEFLATE sees vary from pure noise to
    long runs, short and long match distances, skewed alphabets (long Huffman codes) and periodic patterns."""
    kind = rs.randint(0, 9)
    if kind == 0:
        b = rs.randint(0, 256, size=n_bytes)
    elif kind == 1:                                  # long zero runs broken by noise
        b = np.where(rs.rand(n_bytes) < rs.choice([0.001, 0.02, 0.2]), rs.randint(0, 256, size=n_bytes), 0)
    elif kind == 2:                                  # a periodic pattern (period 1 .. 40) with mutations: overlapping matches
        period = rs.randint(1, 41)
        b = np.tile(rs.randint(0, 256, size=period), n_bytes // period + 1)[:n_bytes]
        b = np.where(rs.rand(n_bytes) < 0.01, rs.randint(0, 256, size=n_bytes), b)
    elif kind == 3:                                  # skewed alphabet: geometric byte frequencies -> codes of 1 .. 15 bits
        b = np.minimum(rs.geometric(rs.choice([0.5, 0.2, 0.05]), size=n_bytes) - 1, 255)
    elif kind == 4:                                  # far matches: copies of earlier segments at distances up to 32 KB
        b = rs.randint(0, 256, size=n_bytes)
        for _ in range(n_bytes // 300):
            ln, dist = rs.randint(3, 300), rs.randint(1, 32768)
            at = rs.randint(0, n_bytes - ln)
            if at - dist >= 0:
                b[at:at + ln] = b[at - dist:at - dist + ln]
    elif kind == 5:                                  # k-mer-like packed words: rare / core / mid densities
        dens = rs.choice([0.01, 0.5, 0.99], size=n_bytes // 8 + 1, p=[0.6, 0.3, 0.1])
        bits = rs.rand(n_bytes // 8 + 1, 64) < dens[:, None]
        b = np.packbits(bits, axis=1).reshape(-1)[:n_bytes]
    elif kind == 6:
        b = np.full(n_bytes, rs.randint(0, 256))     # one byte value
    elif kind == 7:                                  # few distinct values
        b = rs.choice(rs.randint(0, 256, size=rs.randint(2, 6)), size=n_bytes)
    else:                                            # text-like
        words = [bytes(rs.randint(97, 123, size=rs.randint(2, 9)).astype(np.uint8)) for _ in range(50)]
        b = np.frombuffer(b" ".join(words[i] for i in rs.randint(0, 50, size=n_bytes // 3)), dtype=np.uint8)[:n_bytes]
        b = np.concatenate([b, np.zeros(n_bytes - b.shape[0], dtype=np.uint8)])
    return np.asarray(b, dtype=np.uint8)


@pytest.mark.parametrize("decoder", ["4", "3", "2"], ids=["all_symbols_speculative", "two_codes_per_load", "warp_parallel_literals"])
@pytest.mark.parametrize("n_cols", [1000, 40000], ids=["8KB", "320KB"])
def test_randomized_streams_against_zlib(monkeypatch, decoder, n_cols):
    """64 chunks of random character x every zlib level / strategy / window size / memory level, inflated on the device
    straight into the image: bit-identical to the bytes zlib compressed, status 0 for every chunk.  (zlib.decompress of
    the same streams is the oracle: the reference inflates through libhdf5 -> zlib, rules.py:253.)"""
    from kover_b200 import _native
    monkeypatch.setenv("KOVER_B200_INFLATE_V", decoder)
    rs = np.random.RandomState(1234 + n_cols)
    n_chunks = 64
    shard = _native.Shard(0, 64 * n_chunks, n_cols)
    rows, streams = [], []
    strategies = [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]
    for i in range(n_chunks):
        raw = _fuzz_row(rs, n_cols * 8)
        level = int(rs.randint(0, 10))
        c = zlib.compressobj(level, zlib.DEFLATED, int(rs.choice([9, 12, 15])), int(rs.randint(1, 10)),
                             strategies[rs.randint(0, len(strategies))])
        parts = []
        at = 0
        while at < raw.shape[0]:                     # a few sync flushes: empty stored blocks inside the stream
            step = int(rs.randint(1, raw.shape[0] + 1))
            parts.append(c.compress(raw[at:at + step].tobytes()))
            at += step
            if at < raw.shape[0] and rs.rand() < 0.5:
                parts.append(c.flush(zlib.Z_SYNC_FLUSH if rs.rand() < 0.7 else zlib.Z_FULL_FLUSH))
        parts.append(c.flush())
        stream = b"".join(parts)
        assert zlib.decompress(stream) == raw.tobytes()
        rows.append(raw.view(np.uint64))
        streams.append(stream)
    bufs = [np.frombuffer(st + b"\x00" * 8, dtype=np.uint8).copy() for st in streams]
    addr = np.array([b.ctypes.data for b in bufs], dtype=np.uint64)
    nbytes = np.array([len(st) for st in streams], dtype=np.int64)
    status = shard.upload_deflate(64, addr, nbytes, np.arange(n_chunks), np.zeros(n_chunks, dtype=np.int64), 1, n_cols)
    assert not status.any(), status
    img = shard.read_block(0, n_chunks, 0, n_cols)
    for r in range(n_chunks):
        assert np.array_equal(img[r], rows[r]), r
    shard.close()
