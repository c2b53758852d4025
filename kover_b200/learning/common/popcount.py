"""Drop-in for the reference's Cython module kover.learning.common.popcount (popcount.pyx:52-67,
97-112): same two functions, same in-place contract, computed by k_popc_inplace on the GPU.

The learners in this package do not call these (the boundary moved up to sum_rows granularity, see
include/kover_b200.h); they exist so that code written against the reference FFI keeps working.
"""
import numpy as np

from ... import _native

DEVICE = 0


def inplace_popcount_32(arr, row_mask):
    """arr[i, j] <- popcount(arr[i, j] & row_mask[i]) for every non-zero arr[i, j]; arr is uint32 [m, n]."""
    arr = _as_writable(arr, np.uint32)
    _native.inplace_popcount(arr, row_mask, DEVICE)


def inplace_popcount_64(arr, row_mask):
    """arr[i, j] <- popcount(arr[i, j] & row_mask[i]) for every non-zero arr[i, j]; arr is uint64 [m, n]."""
    arr = _as_writable(arr, np.uint64)
    _native.inplace_popcount(arr, row_mask, DEVICE)


def _as_writable(arr, dtype):
    if not isinstance(arr, np.ndarray) or arr.dtype != dtype:
        raise ValueError("Buffer dtype mismatch, expected '%s'" % np.dtype(dtype).name)
    if not arr.flags.writeable:
        raise ValueError("buffer source array is read-only")
    return arr
