"""CPU: the C-ABI library loads without a GPU, exports every symbol include/kover_b200.h declares
(and nothing in the binding table is missing from the header), its host-only entry points work, and
everything that needs the device fails loudly instead of falling back to the CPU."""
import os
import re

import numpy as np
import pytest

from kover_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "kover_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(kv_[a-z0-9_]+)\s*\(", src))


def test_header_and_binding_table_agree():
    assert header_symbols() == set(_native.SYMBOLS.keys())


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert lib.kv_abi_version() == 1


def test_n_blocks():
    assert _native.n_utility_blocks(5000, 1000000) == 1
    assert _native.n_utility_blocks(600000, 1000000) == 2
    assert _native.n_utility_blocks(5000, 777) == 13


def np_select(block_max):
    """scm.py:270-286 restated with numpy's own allclose on the block maxima."""
    best, sel = -np.inf, np.zeros(len(block_max), dtype=np.uint8)
    for b, m in enumerate(block_max):
        if m > best or np.allclose(best, m):
            if np.allclose(m, best):
                sel[b] = 1
            else:
                best = m
                sel[:] = 0
                sel[b] = 1
    return best, sel


@pytest.mark.parametrize("seed", range(20))
def test_select_blocks_matches_numpy_allclose(seed):
    rs = np.random.RandomState(seed)
    n = rs.randint(1, 40)
    base = rs.choice([10.0, 149.0, -999990.0, 0.0, 1e-9, -3.5])
    bm = base + rs.choice([0.0, 1e-9, 1e-6, -1e-6, 1e-4, -1e-4, 5.0, -5.0, 9.99, -9.99, 1.0], size=n)
    bm[rs.rand(n) < 0.2] = -np.inf
    best, sel = _native.select_blocks(bm)
    wbest, wsel = np_select(bm)
    assert best == wbest and np.array_equal(sel, wsel)


def test_select_blocks_all_blacklisted_head():
    # a fully blacklisted first block ties -inf with -inf and is appended (last quirk)
    best, sel = _native.select_blocks(np.array([-np.inf, -np.inf, 3.0, 3.0, 2.0]))
    assert best == 3.0 and sel.tolist() == [0, 0, 1, 1, 0]
    best, sel = _native.select_blocks(np.array([-np.inf, -np.inf]))
    assert best == -np.inf and sel.tolist() == [1, 1]


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="needs a machine WITHOUT a GPU")
def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _native.Shard(0, 100, 1000)
    a = np.ones((2, 3), dtype=np.uint64)
    with pytest.raises(RuntimeError):
        _native.inplace_popcount(a, np.ones(2, dtype=np.uint64))
    from kover_b200.learning.common.rules import KmerRuleClassifications
    with pytest.raises(RuntimeError):
        KmerRuleClassifications(np.zeros((2, 600), dtype=np.uint64), 100)


def test_constructor_errors_precede_device_use():
    from kover_b200.learning.common.rules import KmerRuleClassifications
    P = np.zeros((2, 600), dtype=np.uint64)
    with pytest.raises(ValueError):
        KmerRuleClassifications(P.astype(np.uint16), 100)
    with pytest.raises(ValueError):
        KmerRuleClassifications(P, 100, block_size=(1, 2, 3))
    with pytest.raises(ValueError):
        KmerRuleClassifications(P, 100, block_size=(1.0, 2))
