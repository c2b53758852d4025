"""CPU: host-side layout helpers of the package against the golden vectors minted from the reference."""
import os

import numpy as np
import pytest

from kover_b200 import utils
from kover_b200.sharding import partition_columns
from kover_b200.synthetic import synthetic_block, synthetic_labels, train_split


def test_pack_unpack_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "pack.npz"))
    assert np.array_equal(utils._pack_binary_bytes_to_ints(g["X"], 64), g["p64"])
    assert np.array_equal(utils._pack_binary_bytes_to_ints(g["X"], 32), g["p32"])
    assert np.array_equal(utils._unpack_binary_bytes_from_ints(g["p64"]), g["u64"])
    assert np.array_equal(utils._unpack_binary_bytes_from_ints(g["p32"]), g["u32"])
    assert np.array_equal(utils._unpack_binary_bytes_from_ints(g["p64"][:, 0].copy()), g["u64"][:, 0])
    sizes = [np.dtype(utils._minimum_uint_size(v)).itemsize for v in (0, 255, 256, 65535, 65536, 2 ** 32)]
    assert sizes == g["min_uint_itemsize"].tolist()
    with pytest.raises(ValueError):
        utils._pack_binary_bytes_to_ints(g["X"], 16)
    with pytest.raises(ValueError):
        utils._unpack_binary_bytes_from_ints(g["p64"].astype(np.uint16))


def test_build_row_mask():
    m = utils.build_row_mask([0, 63, 64, 129], 130, 64)
    assert m.dtype == np.uint64 and m.tolist() == [(1 << 63) | 1, 1 << 63, 1 << 62]
    assert utils.build_row_mask([], 130, 64).tolist() == [0, 0, 0]
    assert utils.build_row_mask([5, 5], 8, 8).tolist() == [4]
    with pytest.raises(ValueError):
        utils.build_row_mask([0], 10, 12)


@pytest.mark.parametrize("k,parts", [(1000000, 8), (5000, 2), (513, 2), (100, 4), (100000000, 8)])
def test_partition_columns(k, parts):
    sl = partition_columns(k, parts)
    assert len(sl) == parts and sl[0][0] == 0 and sum(n for _, n in sl) == k
    for (a, n), (b, _) in zip(sl[:-1], sl[1:]):
        assert a + n == b and b % 512 == 0


def test_synthetic_is_deterministic_and_sliceable():
    a = synthetic_block(3, 200, 0, 3000)
    b = synthetic_block(3, 200, 1000, 500)
    assert np.array_equal(a[:, 1000:1500], b)
    assert a.shape == (4, 3000) and (a[-1] & np.uint64((1 << 56) - 1)).max() == 0   # 200 % 64 = 8 valid bits
    dens = np.unpackbits(a[:3].view(np.uint8), axis=0).mean()
    assert 0.3 < dens < 0.6
    y = synthetic_labels(3, 201)
    assert y[:100].sum() == 0 and y[100:].sum() == 101
    t = train_split(72, 300)
    assert len(t) == 200 and np.all(np.diff(t) > 0)
