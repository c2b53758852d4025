"""Host-side layout helpers: the bit layout of the packed k-mer matrix and of the row masks.

Mirrors (names and argument meaning) reference kover/utils.py:117-187 and the nested
build_row_mask of learning/common/rules.py:210-222.  These run on W = ceil(n/64) words or on a
handful of gathered columns, never on the matrix; they are numpy-vectorised re-statements of the
reference's Python loops.
"""
from math import ceil

import numpy as np


def _minimum_uint_size(max_value):
    """Smallest unsigned integer dtype that can store ``max_value`` (utils.py:117-130)."""
    if max_value <= np.iinfo(np.uint8).max:
        return np.uint8
    elif max_value <= np.iinfo(np.uint16).max:
        return np.uint16
    elif max_value <= np.iinfo(np.uint32).max:
        return np.uint32
    elif max_value <= np.iinfo(np.uint64).max:
        return np.uint64
    raise ValueError("No unsigned integer type can store %r" % (max_value,))


def _pack_type(pack_size):
    if pack_size == 64:
        return np.uint64
    elif pack_size == 32:
        return np.uint32
    raise ValueError("Supported data types are 32-bit and 64-bit integers.")


def _pack_binary_bytes_to_ints(a, pack_size):
    """uint8 [n, k] of 0/1 -> packed [ceil(n/pack_size), k]; example i is bit pack_size-1-(i % pack_size) of
    word i // pack_size (utils.py:133-156)."""
    t = _pack_type(pack_size)
    a = np.asarray(a)
    n, k = a.shape
    n_words = int(ceil(1.0 * n / pack_size))
    out = np.zeros((n_words, k), dtype=t)
    for w in range(n_words):
        rows = a[w * pack_size:(w + 1) * pack_size]
        shifts = (pack_size - 1 - np.arange(rows.shape[0])).astype(t).reshape(-1, 1)
        out[w] = np.bitwise_or.reduce(np.left_shift(rows.astype(t), shifts), axis=0)
    return out


def _unpack_binary_bytes_from_ints(a):
    """packed [W] or [W, k] -> uint8 [W * pack_size] or [W * pack_size, k] (utils.py:159-187)."""
    a = np.asarray(a)
    if a.dtype == np.uint32:
        pack_size = 32
    elif a.dtype == np.uint64:
        pack_size = 64
    else:
        raise ValueError("Supported data types are 32-bit and 64-bit integers.")
    # big-endian bytes list the most significant bit first, which is example 0 of the word
    be = np.ascontiguousarray(a.astype(a.dtype.newbyteorder(">")))
    if a.ndim == 1:
        return np.unpackbits(be.view(np.uint8))
    n_words, k = a.shape
    bits = np.unpackbits(be.view(np.uint8).reshape(n_words, k, pack_size // 8), axis=2)   # [W, k, pack_size]
    return np.ascontiguousarray(bits.transpose(0, 2, 1)).reshape(n_words * pack_size, k)


def build_row_mask(example_idx, n_examples, mask_n_bits):
    """Row mask that keeps only ``example_idx`` (rules.py:210-222); duplicates count once."""
    if mask_n_bits not in [8, 16, 32, 64, 128]:
        raise ValueError("Unsupported mask format. Use 8, 16, 32, 64 or 128 bits.")
    if mask_n_bits == 128:
        raise ValueError("128-bit masks have no numpy dtype.")
    nbytes = mask_n_bits // 8
    n_masks = int(ceil(float(n_examples) / mask_n_bits))
    idx = np.asarray(example_idx, dtype=np.int64).reshape(-1)
    bits = np.zeros(n_masks * mask_n_bits, dtype=np.uint8)
    if idx.size:
        if idx.min() < 0 or idx.max() >= bits.shape[0]:
            raise IndexError("list index out of range")
        bits[idx] = 1
    # packbits is MSB-first, so example i lands on bit mask_n_bits-1-(i % mask_n_bits) of word i // mask_n_bits
    masks = np.packbits(bits).view(">u%d" % nbytes).astype("u%d" % nbytes)
    return masks
