"""ctypes binding of libkoverb200.so (C ABI: include/kover_b200.h).

Nothing here computes: every function forwards to the CUDA library and turns its integer status
into the exception type the reference raises for the same mistake (ValueError for bad arguments,
RuntimeError otherwise).  If the library is missing it is built once with nvcc; if that is not
possible, or no GPU is usable, calls fail loudly -- there is no CPU fallback.
"""
import ctypes
import os

import numpy as np

from . import build as _build

KV_OK, KV_E_INVALID, KV_E_CUDA, KV_E_NOMEM, KV_E_OVERFLOW, KV_E_INDEX = 0, -1, -2, -3, -4, -5

_c_i64 = ctypes.c_int64
_c_p = ctypes.c_void_p


class ShardInfo(ctypes.Structure):
    _fields_ = [("device", ctypes.c_int32), ("sm_count", ctypes.c_int32), ("n_examples", _c_i64),
                ("n_words", _c_i64), ("n_kmers_total", _c_i64), ("col0", _c_i64), ("n_cols", _c_i64),
                ("ld", _c_i64), ("matrix_bytes", _c_i64), ("last_kernel_ms", ctypes.c_double),
                ("last_launches", _c_i64), ("total_launches", _c_i64), ("scan_ms_total", ctypes.c_double),
                ("scan_calls", _c_i64), ("rescore_calls", _c_i64), ("exch_steps", _c_i64),
                ("exch_publish_us", ctypes.c_double), ("exch_wait_us", ctypes.c_double),
                ("exch_merge_us", ctypes.c_double), ("stream_steps", _c_i64), ("stream_redos", _c_i64),
                ("level_nodes_counted", _c_i64), ("level_nodes_derived", _c_i64)]


# every exported symbol with (restype, argtypes); tests/test_abi.py checks this table against the header
SYMBOLS = {
    "kv_last_error": (ctypes.c_char_p, []),
    "kv_abi_version": (ctypes.c_int, []),
    "kv_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "kv_create": (ctypes.c_int, [ctypes.c_int, _c_i64, _c_i64, _c_i64, _c_i64, ctypes.POINTER(_c_p)]),
    "kv_destroy": (ctypes.c_int, [_c_p]),
    "kv_get_info": (ctypes.c_int, [_c_p, ctypes.POINTER(ShardInfo)]),
    "kv_upload_block": (ctypes.c_int, [_c_p, ctypes.c_int, _c_i64, _c_i64, _c_i64, _c_i64, _c_p, _c_i64]),
    "kv_upload_deflate": (ctypes.c_int, [_c_p, ctypes.c_int, _c_i64, _c_p, _c_p, _c_p, _c_p, _c_i64, _c_i64, _c_p]),
    "kv_upload_deflate_stats": (ctypes.c_int, [_c_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                               ctypes.POINTER(ctypes.c_double)]),
    "kv_generate_synthetic": (ctypes.c_int, [_c_p, ctypes.c_uint64]),
    "kv_read_block": (ctypes.c_int, [_c_p, _c_i64, _c_i64, _c_i64, _c_i64, _c_p, _c_i64]),
    "kv_timer_start": (ctypes.c_int, [_c_p]),
    "kv_timer_stop": (ctypes.c_int, [_c_p, ctypes.POINTER(ctypes.c_double)]),
    "kv_inplace_popcount_64": (ctypes.c_int, [ctypes.c_int, _c_p, _c_i64, _c_i64, _c_p]),
    "kv_inplace_popcount_32": (ctypes.c_int, [ctypes.c_int, _c_p, _c_i64, _c_i64, _c_p]),
    "kv_sum_rows": (ctypes.c_int, [_c_p, _c_p, _c_p, ctypes.c_int]),
    "kv_sum_rows_full": (ctypes.c_int, [_c_p, _c_p, _c_i64, _c_p, _c_p, ctypes.c_int]),
    "kv_sum_rows_multi": (ctypes.c_int, [_c_p, ctypes.c_int, _c_p, _c_p]),
    "kv_risk_errors": (ctypes.c_int, [_c_p, _c_p, _c_p, _c_i64, _c_i64, _c_p]),
    "kv_risk_map": (ctypes.c_int, [_c_p, _c_p, _c_p, _c_i64, ctypes.c_int, _c_p, _c_p]),
    "kv_get_columns": (ctypes.c_int, [_c_p, _c_p, _c_i64, _c_p]),
    "kv_set_blacklist": (ctypes.c_int, [_c_p, _c_p, _c_i64]),
    "kv_scm_n_blocks": (_c_i64, [_c_i64, _c_i64]),
    "kv_scm_reuse_counts": (ctypes.c_int, [_c_p, ctypes.c_int]),
    "kv_scm_scan": (ctypes.c_int, [_c_p, _c_p, _c_p, _c_i64, _c_i64, ctypes.c_double, _c_i64, _c_p, _c_i64]),
    "kv_scm_select_blocks": (ctypes.c_int, [_c_p, _c_i64, ctypes.POINTER(ctypes.c_double), _c_p]),
    "kv_scm_ties": (ctypes.c_int, [_c_p, _c_p, _c_p, _c_i64, _c_i64, _c_p, _c_p, _c_p, ctypes.POINTER(_c_i64)]),
    "kv_scm_queue_push": (ctypes.c_int, [_c_p, _c_i64, _c_i64, _c_p, _c_p, _c_i64, _c_i64, ctypes.c_double, _c_i64,
                                         ctypes.c_int]),
    "kv_scm_queue_collect": (ctypes.c_int, [_c_p, _c_i64, _c_p, _c_p, _c_p, _c_p, _c_p]),
    "kv_scm_queue_run": (ctypes.c_int, [_c_p, _c_i64, _c_p, _c_p, _c_p, _c_p, _c_p, _c_i64, _c_p, _c_p, _c_p, _c_p, _c_p]),
    "kv_scm_queue_packets": (ctypes.c_int, [_c_p, _c_i64, _c_p, _c_p, _c_p, _c_p, _c_p, _c_i64, _c_i64, _c_p, ctypes.c_int]),
    "kv_scm_packet_bytes": (_c_i64, [_c_i64, _c_i64]),
    "kv_scm_scan_packet": (ctypes.c_int, [_c_p, _c_p, _c_p, _c_i64, _c_i64, ctypes.c_double, _c_i64, _c_i64, _c_i64,
                                          _c_p, ctypes.c_int]),
    "kv_scm_merge_packets": (ctypes.c_int, [_c_p, _c_i64, _c_i64, _c_i64, ctypes.c_double, _c_i64,
                                            ctypes.POINTER(ctypes.c_double), _c_i64, _c_p, _c_p, _c_p,
                                            ctypes.POINTER(_c_i64), _c_p, _c_p, ctypes.POINTER(ctypes.c_int)]),
    "kv_p2p_create": (ctypes.c_int, [_c_p, ctypes.c_int, ctypes.c_int, _c_i64, _c_p]),
    "kv_p2p_connect": (ctypes.c_int, [_c_p, _c_p]),
    "kv_scm_best_rules_p2p": (ctypes.c_int, [_c_p, _c_p, _c_p, _c_i64, _c_i64, ctypes.c_double, _c_i64, _c_i64,
                                             ctypes.POINTER(ctypes.c_double), _c_i64, _c_p, _c_p, _c_p,
                                             ctypes.POINTER(_c_i64), _c_p, _c_p, ctypes.POINTER(ctypes.c_int)]),
    "kv_scm_best_rules_p2p_idx": (ctypes.c_int, [_c_p, _c_p, _c_i64, _c_p, _c_i64, ctypes.c_double, _c_i64, _c_i64,
                                                 ctypes.POINTER(ctypes.c_double), _c_i64, _c_p, _c_p, _c_p,
                                                 ctypes.POINTER(_c_i64), _c_p, _c_p, ctypes.POINTER(ctypes.c_int)]),
    "kv_scm_best_rules": (ctypes.c_int, [_c_p, _c_p, _c_p, _c_i64, _c_i64, ctypes.c_double, _c_i64,
                                         ctypes.POINTER(ctypes.c_double), _c_i64, _c_p, _c_p, _c_p,
                                         ctypes.POINTER(_c_i64)]),
    "kv_scm_best_rules_idx": (ctypes.c_int, [_c_p, _c_p, _c_i64, _c_p, _c_i64, ctypes.c_double, _c_i64,
                                             ctypes.POINTER(ctypes.c_double), _c_i64, _c_p, _c_p, _c_p,
                                             ctypes.POINTER(_c_i64)]),
    "kv_set_profiling": (ctypes.c_int, [_c_p, ctypes.c_int]),
    "kv_scm_state_set": (ctypes.c_int, [_c_p, _c_p, _c_p, ctypes.POINTER(_c_i64), ctypes.POINTER(_c_i64)]),
    "kv_mask_and_column": (ctypes.c_int, [_c_p, _c_i64, _c_p, ctypes.c_int, _c_p, ctypes.POINTER(_c_i64),
                                          ctypes.POINTER(_c_i64)]),
    "kv_gini_scan": (ctypes.c_int, [_c_p, ctypes.c_int, _c_p, _c_p, _c_p, _c_p, ctypes.POINTER(ctypes.c_double)]),
    "kv_gini_best_split": (ctypes.c_int, [_c_p, ctypes.c_int, _c_p, _c_p, _c_p, _c_p, ctypes.POINTER(ctypes.c_double),
                                          _c_i64, _c_p, ctypes.POINTER(_c_i64)]),
    "kv_gini_ties": (ctypes.c_int, [_c_p, ctypes.c_double, _c_i64, _c_p, ctypes.POINTER(_c_i64)]),
    "kv_gini_scores": (ctypes.c_int, [_c_p, _c_p]),
    "kv_host_alloc": (ctypes.c_int, [_c_i64, ctypes.POINTER(ctypes.c_void_p)]),
    "kv_host_free": (ctypes.c_int, [_c_p]),
    "kv_gini_level": (ctypes.c_int, [_c_p, _c_i64, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p]),
    "kv_occ_table_set": (ctypes.c_int, [_c_p, ctypes.c_int, _c_p]),
    "kv_occ_table_gather": (ctypes.c_int, [_c_p, ctypes.c_int, _c_p, _c_i64, _c_p]),
    "kv_gini_level_family": (ctypes.c_int, [_c_p, _c_i64, _c_p, _c_p]),
    "kv_gini_level_fetch": (ctypes.c_int, [_c_p, _c_i64, _c_p, _c_i64]),
    "kv_gini_level_result": (ctypes.c_int, [_c_p, _c_i64, ctypes.POINTER(ctypes.POINTER(ctypes.c_int64)),
                                            ctypes.POINTER(_c_i64)]),
}

_lib = None


def library_path():
    return _build.LIB


def load():
    """The loaded library (built on first use if absent).  Raises RuntimeError if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.isfile(path):
        try:
            _build.build_native()
        except Exception as e:  # noqa: BLE001
            raise RuntimeError("libkoverb200.so is missing and could not be built with nvcc (%s). kover_b200 has "
                               "no CPU fallback: build it with `python -m kover_b200.build`." % e)
    lib = ctypes.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        f = getattr(lib, name)
        f.restype = res
        f.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().kv_last_error().decode("utf-8", "replace")


def check(rc):
    if rc == KV_OK:
        return
    msg = last_error()
    if rc == KV_E_INVALID:
        raise ValueError(msg)
    if rc == KV_E_NOMEM:
        raise MemoryError(msg)
    if rc == KV_E_INDEX:
        raise IndexError(msg)
    raise RuntimeError(msg)


def ptr(a):
    """Data pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags.c_contiguous
    return a.ctypes.data_as(ctypes.c_void_p)


# ---- pinned result arrays ------------------------------------------------------------------------
_PINNED_FREE = {}            # rounded size -> [addresses] of page-locked blocks nobody references any more
_PINNED_MIN = 1 << 20        # smaller results stay in ordinary numpy memory
_PINNED_CACHE_LIMIT = 4 << 30
_pinned_cached_bytes = 0


def _pinned_release(addr, size):
    global _pinned_cached_bytes
    if _pinned_cached_bytes + size > _PINNED_CACHE_LIMIT:
        try:
            load().kv_host_free(ctypes.c_void_p(addr))
        except Exception:  # noqa: BLE001 (interpreter shutdown)
            pass
        return
    _PINNED_FREE.setdefault(size, []).append(addr)
    _pinned_cached_bytes += size


def result_empty(shape, dtype):
    """np.empty for RESULT arrays that a device-to-host copy fills: page-locked memory from a recycling pool when
    the array is large (a block returns to the pool when the last view of the array dies), plain numpy otherwise
    or when pinning fails.  The caller owns the array like any numpy array."""
    global _pinned_cached_bytes
    import weakref
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    if nbytes < _PINNED_MIN:
        return np.empty(shape, dtype=dtype)
    size = 1 << (nbytes - 1).bit_length() if nbytes < (64 << 20) else -(-nbytes // (16 << 20)) * (16 << 20)
    free = _PINNED_FREE.get(size)
    if free:
        addr = free.pop()
        _pinned_cached_bytes -= size
    else:
        p = ctypes.c_void_p()
        if load().kv_host_alloc(size, ctypes.byref(p)) != 0:
            return np.empty(shape, dtype=dtype)
        addr = p.value
    block = (ctypes.c_char * size).from_address(addr)
    weakref.finalize(block, _pinned_release, addr, size)       # every numpy view keeps `block` alive through .base
    return np.frombuffer(block, dtype=dtype, count=nbytes // dtype.itemsize).reshape(shape)


def device_count():
    n = ctypes.c_int(0)
    check(load().kv_device_count(ctypes.byref(n)))
    return n.value


class Shard(object):
    """One contiguous k-mer-column slice of the packed matrix resident in the HBM of one GPU."""

    def __init__(self, device, n_examples, n_kmers_total, col0=0, n_cols=None):
        self._lib = load()
        self._h = _c_p()
        if n_cols is None:
            n_cols = n_kmers_total - col0
        check(self._lib.kv_create(int(device), int(n_examples), int(n_kmers_total), int(col0), int(n_cols),
                                  ctypes.byref(self._h)))
        self.device = int(device)
        self.n_examples = int(n_examples)
        self.n_words = (self.n_examples + 63) // 64
        self.n_kmers_total = int(n_kmers_total)
        self.col0 = int(col0)
        self.n_cols = int(n_cols)
        # reusable output buffers of the per-step calls (their contents are copied out before returning)
        self._tie_cap = 1 << 16
        self._tie_buf = [np.empty(self._tie_cap, dtype=np.int64) for _ in range(3)]
        self._tie_ptr = [a.ctypes.data for a in self._tie_buf]      # (.ctypes builds an object per access: cached)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.kv_destroy(self._h)
            self._h = _c_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def info(self):
        i = ShardInfo()
        check(self._lib.kv_get_info(self._h, ctypes.byref(i)))
        return i

    # -- upload
    def upload_block(self, block, row0, gcol0):
        block = np.asarray(block)
        if block.dtype == np.uint64:
            bits = 64
        elif block.dtype == np.uint32:
            bits = 32
        else:
            raise ValueError("Unsupported data type for packed attribute classifications array. The supported data"
                             " types are np.uint32 and np.uint64.")
        if block.ndim == 1:
            block = block.reshape(1, -1)
        if block.strides[1] != block.itemsize or block.strides[0] % block.itemsize or block.strides[0] < 0:
            block = np.ascontiguousarray(block)
        ld = block.strides[0] // block.itemsize if block.shape[0] > 1 else max(block.shape[1], 1)
        check(self._lib.kv_upload_block(self._h, bits, int(row0), block.shape[0], int(gcol0), block.shape[1],
                                        ctypes.c_void_p(block.ctypes.data), int(ld)))

    def upload_deflate(self, pack_bits, addresses, nbytes, row0, gcol0, chunk_rows, chunk_cols):
        """Raw zlib streams of HDF5 chunks -> the resident image, inflated on the device (kv_upload_deflate).
        addresses: host addresses of the streams (uint64 array; the memory must stay mapped during the call).
        -> int32 status per chunk (0 = placed; non-zero = malformed stream, redo that chunk on the host)."""
        addresses = np.ascontiguousarray(addresses, dtype=np.uint64)
        nbytes = np.ascontiguousarray(nbytes, dtype=np.int64)
        row0 = np.ascontiguousarray(row0, dtype=np.int64)
        gcol0 = np.ascontiguousarray(gcol0, dtype=np.int64)
        n = addresses.shape[0]
        assert nbytes.shape == row0.shape == gcol0.shape == (n,)
        status = np.zeros(n, dtype=np.int32)
        check(self._lib.kv_upload_deflate(self._h, int(pack_bits), n, ptr(addresses), ptr(nbytes), ptr(row0), ptr(gcol0),
                                          int(chunk_rows), int(chunk_cols), ptr(status)))
        return status

    def upload_deflate_stats(self):
        """{setup_ms, gather_ms, total_ms} of the last upload_deflate on this shard (host clock)."""
        a, b, c = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_double(0)
        check(self._lib.kv_upload_deflate_stats(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return {"setup_ms": a.value, "gather_ms": b.value, "total_ms": c.value}

    def generate_synthetic(self, seed):
        check(self._lib.kv_generate_synthetic(self._h, ctypes.c_uint64(int(seed))))

    def read_block(self, row0, n_rows, lcol0, n_cols, out=None):
        if out is None:
            out = np.empty((n_rows, n_cols), dtype=np.uint64)
        assert out.dtype == np.uint64 and out.shape == (n_rows, n_cols) and out.flags.c_contiguous
        check(self._lib.kv_read_block(self._h, int(row0), int(n_rows), int(lcol0), int(n_cols), ptr(out), int(n_cols)))
        return out

    def timer_start(self):
        check(self._lib.kv_timer_start(self._h))

    def timer_stop(self):
        ms = ctypes.c_double(0.0)
        check(self._lib.kv_timer_stop(self._h, ctypes.byref(ms)))
        return ms.value

    # -- a3
    def sum_rows(self, mask, out):
        mask = np.ascontiguousarray(mask, dtype=np.uint64)
        assert mask.shape == (self.n_words,) and out.shape == (self.n_cols,)
        check(self._lib.kv_sum_rows(self._h, ptr(mask), ptr(out), out.itemsize))
        return out

    def sum_rows_full(self, mask, n_rows, out_presence, out_absence):
        mask = np.ascontiguousarray(mask, dtype=np.uint64)
        assert mask.shape == (self.n_words,) and out_presence.shape == out_absence.shape == (self.n_cols,)
        assert out_presence.dtype == out_absence.dtype
        check(self._lib.kv_sum_rows_full(self._h, ptr(mask), int(n_rows), ptr(out_presence), ptr(out_absence),
                                         out_presence.itemsize))

    def sum_rows_multi(self, masks):
        masks = np.ascontiguousarray(masks, dtype=np.uint64)
        assert masks.ndim == 2 and masks.shape[1] == self.n_words
        out = result_empty((masks.shape[0], self.n_cols), np.uint32)
        check(self._lib.kv_sum_rows_multi(self._h, masks.shape[0], ptr(masks), ptr(out)))
        return out

    # -- dataset split risk tables
    def risk_errors(self, pos_mask, neg_mask, n_pos, n_total, present):
        pos_mask = np.ascontiguousarray(pos_mask, dtype=np.uint64)
        neg_mask = np.ascontiguousarray(neg_mask, dtype=np.uint64)
        assert present.dtype == np.uint8 and present.shape == (n_total + 1,)
        check(self._lib.kv_risk_errors(self._h, ptr(pos_mask), ptr(neg_mask), int(n_pos), int(n_total), ptr(present)))

    def risk_map(self, lut_kmer, lut_anti, out_kmer, out_anti):
        lk = np.ascontiguousarray(lut_kmer, dtype=np.uint32)
        la = np.ascontiguousarray(lut_anti, dtype=np.uint32)
        assert out_kmer.shape == out_anti.shape == (self.n_cols,) and out_kmer.dtype == out_anti.dtype
        check(self._lib.kv_risk_map(self._h, ptr(lk), ptr(la), lk.shape[0], out_kmer.itemsize, ptr(out_kmer),
                                    ptr(out_anti)))

    # -- a4
    def get_columns(self, local_cols):
        cols = np.ascontiguousarray(local_cols, dtype=np.int64)
        out = np.empty((cols.shape[0], self.n_words), dtype=np.uint64)
        check(self._lib.kv_get_columns(self._h, ptr(cols), cols.shape[0], ptr(out)))
        return out

    # -- a6
    def set_blacklist(self, rule_idx):
        idx = np.ascontiguousarray(rule_idx, dtype=np.int64)
        check(self._lib.kv_set_blacklist(self._h, ptr(idx), idx.shape[0]))

    def scm_reuse_counts(self, swap_roles=False):
        """Arm the next scm_* call to re-score the resident counts instead of scanning (kv_scm_reuse_counts)."""
        check(self._lib.kv_scm_reuse_counts(self._h, int(bool(swap_roles))))

    def scm_scan(self, pos_mask, neg_mask, n_pos, n_neg, p, util_block_size):
        nb = n_utility_blocks(self.n_kmers_total, util_block_size)
        bm = np.empty(nb, dtype=np.float64)
        pos_mask = np.ascontiguousarray(pos_mask, dtype=np.uint64)
        neg_mask = np.ascontiguousarray(neg_mask, dtype=np.uint64)
        assert pos_mask.shape == (self.n_words,) and neg_mask.shape == (self.n_words,)
        check(self._lib.kv_scm_scan(self._h, ptr(pos_mask), ptr(neg_mask), int(n_pos), int(n_neg), float(p),
                                    int(util_block_size), ptr(bm), nb))
        return bm

    def scm_ties(self, block_max, selected, capacity=4096):
        bm = np.ascontiguousarray(block_max, dtype=np.float64)
        sel = np.ascontiguousarray(selected, dtype=np.uint8)
        while True:
            idx = np.empty(capacity, dtype=np.int64)
            pe = np.empty(capacity, dtype=np.int64)
            nc = np.empty(capacity, dtype=np.int64)
            n = _c_i64(0)
            rc = self._lib.kv_scm_ties(self._h, ptr(bm), ptr(sel), bm.shape[0], capacity, ptr(idx), ptr(pe), ptr(nc),
                                       ctypes.byref(n))
            if rc == KV_E_OVERFLOW:
                capacity = int(n.value)
                continue
            check(rc)
            return idx[:n.value], pe[:n.value], nc[:n.value]

    def scm_scan_packet(self, pos_mask, neg_mask, n_pos, n_neg, p, util_block_size, capacity, packet):
        """Fill ``packet`` (a torch uint8 tensor on this shard's GPU, or a host numpy uint8 array) with the block
        maxima + local tie candidates of one scan (kv_scm_scan_packet)."""
        pos_mask = np.ascontiguousarray(pos_mask, dtype=np.uint64)
        neg_mask = np.ascontiguousarray(neg_mask, dtype=np.uint64)
        nb = n_utility_blocks(self.n_kmers_total, util_block_size)
        need = packet_bytes(nb, capacity)
        if isinstance(packet, np.ndarray):
            assert packet.dtype == np.uint8 and packet.size >= need and packet.flags.c_contiguous
            addr, on_device = packet.ctypes.data, 0
        else:
            assert packet.numel() * packet.element_size() >= need and packet.is_contiguous()
            addr, on_device = packet.data_ptr(), int(packet.is_cuda)
        check(self._lib.kv_scm_scan_packet(self._h, ptr(pos_mask), ptr(neg_mask), int(n_pos), int(n_neg), float(p),
                                           int(util_block_size), nb, int(capacity), ctypes.c_void_p(addr), on_device))

    # -- peer mailbox (multi-rank, one process per GPU)
    def p2p_create(self, rank, world, slot_bytes):
        handle = np.zeros(64, dtype=np.uint8)
        check(self._lib.kv_p2p_create(self._h, int(rank), int(world), int(slot_bytes), ptr(handle)))
        return handle

    def p2p_connect(self, all_handles):
        h = np.ascontiguousarray(all_handles, dtype=np.uint8)
        check(self._lib.kv_p2p_connect(self._h, ptr(h)))

    def scm_best_rules_p2p(self, pos_mask, neg_mask, n_pos, n_neg, p, util_block_size, packet_capacity,
                           capacity=4096):
        """-> ((best, idx, pos_err, neg_cover) or None on packet overflow, merged block maxima, best, selected)."""
        pos_mask = np.ascontiguousarray(pos_mask, dtype=np.uint64)
        neg_mask = np.ascontiguousarray(neg_mask, dtype=np.uint64)
        nb = n_utility_blocks(self.n_kmers_total, util_block_size)
        bm, sel = np.empty(nb, dtype=np.float64), np.empty(nb, dtype=np.uint8)
        reuse = capacity <= self._tie_cap              # the step's scratch arrays are allocated once per shard
        idx, pe, nc = self._tie_buf if reuse else (np.empty(capacity, dtype=np.int64) for _ in range(3))
        best, n, over = ctypes.c_double(0.0), _c_i64(0), ctypes.c_int(0)
        rc = self._lib.kv_scm_best_rules_p2p(self._h, ptr(pos_mask), ptr(neg_mask), int(n_pos), int(n_neg), float(p),
                                             int(util_block_size), int(packet_capacity), ctypes.byref(best), capacity,
                                             ptr(idx), ptr(pe), ptr(nc), ctypes.byref(n), ptr(bm), ptr(sel),
                                             ctypes.byref(over))
        if rc == KV_E_OVERFLOW:       # more exact ties than `capacity` (cannot exceed world * packet_capacity)
            over.value = 1
        else:
            check(rc)
        if over.value:
            return None, bm, best.value, sel
        m = n.value
        if reuse:
            return (best.value, idx[:m].copy(), pe[:m].copy(), nc[:m].copy()), bm, best.value, sel
        return (best.value, idx[:m], pe[:m], nc[:m]), bm, best.value, sel

    def scm_best_rules_p2p_idx(self, pos_idx, neg_idx, p, util_block_size, packet_capacity):
        """The multi-rank step from example index arrays (int64, contiguous) -> the 4-tuple, or None when a packet or
        the merged result overflowed (the caller redoes the step through the mask path and its two-phase fallback)."""
        if getattr(self, "_p2p_ubs", None) != util_block_size:         # (per util_block_size: block tables of the merge)
            nb = n_utility_blocks(self.n_kmers_total, util_block_size)
            self._p2p_ubs, self._p2p_bm, self._p2p_sel = util_block_size, np.empty(nb, dtype=np.float64), np.empty(nb, dtype=np.uint8)
            self._p2p_ptr = (self._p2p_bm.ctypes.data, self._p2p_sel.ctypes.data)
        capacity, (idx, pe, nc), (p_idx, p_pe, p_nc) = self._tie_cap, self._tie_buf, self._tie_ptr
        best, n, over = ctypes.c_double(0.0), _c_i64(0), ctypes.c_int(0)
        rc = self._lib.kv_scm_best_rules_p2p_idx(self._h, pos_idx.__array_interface__["data"][0], pos_idx.shape[0],
                                                 neg_idx.__array_interface__["data"][0], neg_idx.shape[0], p,
                                                 util_block_size, packet_capacity, ctypes.byref(best), capacity, p_idx,
                                                 p_pe, p_nc, ctypes.byref(n), self._p2p_ptr[0], self._p2p_ptr[1],
                                                 ctypes.byref(over))
        if rc == KV_E_OVERFLOW or (rc == KV_OK and over.value):
            return None
        if rc:
            check(rc)
        m = n.value
        return best.value, idx[:m].copy(), pe[:m].copy(), nc[:m].copy()

    QUEUE_SLOT_RECS = 256

    def scm_queue(self, jobs, util_block_size):
        """jobs: [(p, pos_mask, neg_mask, n_pos, n_neg, reuse)] with reuse 0 = scan, 1 = re-score the previous
        scan's counts, 2 = same with positives and negatives exchanged.  Everything is enqueued back to back and
        collected with one synchronisation; jobs whose tie set did not fit their slot are redone alone.
        -> [(best, idx, pos_err, neg_cover)] in job order."""
        n = len(jobs)
        r = self.QUEUE_SLOT_RECS
        masks = _stack_job_masks(jobs, self.n_words)
        ps, _, _, n_pos, n_neg, reuse = zip(*jobs) if n else ((), (), (), (), (), ())
        n_pos, n_neg = np.array(n_pos, dtype=np.int64), np.array(n_neg, dtype=np.int64)
        ps, reuse = np.array(ps, dtype=np.float64), np.array(reuse, dtype=np.int32)
        best, found = np.empty(n, dtype=np.float64), np.empty(n, dtype=np.int64)
        idx, pe, nc = (np.empty((n, r), dtype=np.int64) for _ in range(3))
        check(self._lib.kv_scm_queue_run(self._h, n, ptr(masks), ptr(n_pos), ptr(n_neg), ptr(ps), ptr(reuse),
                                         int(util_block_size), ptr(best), ptr(found), ptr(idx), ptr(pe), ptr(nc)))
        # (the result arrays are this call's own: the per-job results are views into them, not copies)
        best_l, found_l = best.tolist(), found.tolist()
        out = []
        for j in range(n):
            m = found_l[j]
            if m > r:
                p, pm, nm, jp, jn, _ = jobs[j]
                out.append(self.scm_best_rules(pm, nm, jp, jn, p, util_block_size, capacity=m))
            else:
                out.append((best_l[j], idx[j, :m], pe[j, :m], nc[j, :m]))
        return out

    def scm_queue_packets(self, jobs, util_block_size, capacity, packets):
        """The lock-step grid on a partial shard (kv_scm_queue_packets).  jobs as in scm_queue; ``packets``: a host
        numpy uint8 array or a torch uint8 tensor on this shard's GPU of n_jobs * packet_bytes bytes, filled with one
        packet per job."""
        n = len(jobs)
        masks = _stack_job_masks(jobs, self.n_words)
        n_pos = np.array([job[3] for job in jobs], dtype=np.int64)
        n_neg = np.array([job[4] for job in jobs], dtype=np.int64)
        ps = np.array([job[0] for job in jobs], dtype=np.float64)
        reuse = np.array([job[5] for job in jobs], dtype=np.int32)
        need = n * packet_bytes(n_utility_blocks(self.n_kmers_total, util_block_size), capacity)
        if isinstance(packets, np.ndarray):
            assert packets.dtype == np.uint8 and packets.size >= need and packets.flags.c_contiguous
            addr, on_device = packets.ctypes.data, 0
        else:
            assert packets.numel() * packets.element_size() >= need and packets.is_contiguous()
            addr, on_device = packets.data_ptr(), int(packets.is_cuda)
        check(self._lib.kv_scm_queue_packets(self._h, n, ptr(masks), ptr(n_pos), ptr(n_neg), ptr(ps), ptr(reuse),
                                             int(util_block_size), int(capacity), ctypes.c_void_p(addr), on_device))

    def scm_best_rules(self, pos_mask, neg_mask, n_pos, n_neg, p, util_block_size, capacity=None):
        if pos_mask.dtype != np.uint64 or neg_mask.dtype != np.uint64:
            pos_mask, neg_mask = pos_mask.astype(np.uint64), neg_mask.astype(np.uint64)
        assert pos_mask.shape == (self.n_words,) and neg_mask.shape == (self.n_words,)
        if capacity is None:
            capacity, (idx, pe, nc) = self._tie_cap, self._tie_buf
        else:
            idx, pe, nc = (np.empty(capacity, dtype=np.int64) for _ in range(3))
        best = ctypes.c_double(0.0)
        n = _c_i64(0)
        rc = self._lib.kv_scm_best_rules(self._h, pos_mask.ctypes.data, neg_mask.ctypes.data, int(n_pos), int(n_neg),
                                         float(p), int(util_block_size), ctypes.byref(best), capacity,
                                         idx.ctypes.data, pe.ctypes.data, nc.ctypes.data, ctypes.byref(n))
        if rc == KV_E_OVERFLOW:
            # rare (huge equivalent-rule set): redo through the three-phase path with a buffer of the right size
            bm = self.scm_scan(pos_mask, neg_mask, n_pos, n_neg, p, util_block_size)
            best_v, sel = select_blocks(bm)
            idx, pe, nc = self.scm_ties(bm, sel, int(n.value))
            return best_v, idx, pe, nc
        check(rc)
        m = n.value
        return best.value, idx[:m].copy(), pe[:m].copy(), nc[:m].copy()

    def scm_best_rules_idx(self, pos_idx, neg_idx, p, util_block_size):
        """scm_best_rules from example index arrays (int64, contiguous): the row masks are built in the library."""
        capacity, (idx, pe, nc), (p_idx, p_pe, p_nc) = self._tie_cap, self._tie_buf, self._tie_ptr
        best = ctypes.c_double(0.0)
        n = _c_i64(0)
        # (the per-iteration call of the SCM: pointers straight from the array interface, no ctypes objects)
        rc = self._lib.kv_scm_best_rules_idx(self._h, pos_idx.__array_interface__["data"][0], pos_idx.shape[0],
                                             neg_idx.__array_interface__["data"][0], neg_idx.shape[0], p,
                                             util_block_size, ctypes.byref(best), capacity, p_idx, p_pe, p_nc,
                                             ctypes.byref(n))
        if rc == KV_E_OVERFLOW:
            return None
        if rc:
            check(rc)
        m = n.value
        return best.value, idx[:m].copy(), pe[:m].copy(), nc[:m].copy()

    # -- device-resident greedy state (scm.py:133-139)
    def scm_state_set(self, pos_mask, neg_mask):
        pos_mask = np.ascontiguousarray(pos_mask, dtype=np.uint64)
        neg_mask = np.ascontiguousarray(neg_mask, dtype=np.uint64)
        assert pos_mask.shape == (self.n_words,) and neg_mask.shape == (self.n_words,)
        n_pos, n_neg = _c_i64(0), _c_i64(0)
        check(self._lib.kv_scm_state_set(self._h, ptr(pos_mask), ptr(neg_mask), ctypes.byref(n_pos), ctypes.byref(n_neg)))
        return n_pos.value, n_neg.value

    def mask_and_column(self, local_col=None, words=None, invert=False, want_words=False):
        """AND the resident masks with a column of this shard (local_col) or with the given W words.
        -> (n_pos, n_neg, column words or None)."""
        n_pos, n_neg = _c_i64(0), _c_i64(0)
        out = np.empty(self.n_words, dtype=np.uint64) if want_words else None
        if local_col is None:
            words = np.ascontiguousarray(words, dtype=np.uint64)
            assert words.shape == (self.n_words,)
        check(self._lib.kv_mask_and_column(self._h, -1 if local_col is None else int(local_col),
                                           ptr(words) if local_col is None else None, int(bool(invert)),
                                           ptr(out), ctypes.byref(n_pos), ctypes.byref(n_neg)))
        return n_pos.value, n_neg.value, out

    def scm_best_rules_state(self, p, util_block_size):
        """scm_best_rules over the resident greedy state.  -> the 4-tuple, or None on tie-buffer overflow."""
        capacity, (idx, pe, nc) = self._tie_cap, self._tie_buf
        best, n = ctypes.c_double(0.0), _c_i64(0)
        rc = self._lib.kv_scm_best_rules(self._h, None, None, 0, 0, float(p), int(util_block_size), ctypes.byref(best),
                                         capacity, idx.ctypes.data, pe.ctypes.data, nc.ctypes.data, ctypes.byref(n))
        if rc == KV_E_OVERFLOW:
            return None
        check(rc)
        m = n.value
        return best.value, idx[:m].copy(), pe[:m].copy(), nc[:m].copy()

    def scm_best_rules_p2p_state(self, p, util_block_size, packet_capacity):
        """The multi-rank step over the resident greedy state -> the 4-tuple, or None on overflow."""
        nb = n_utility_blocks(self.n_kmers_total, util_block_size)
        if getattr(self, "_p2p_nb", None) != nb:
            self._p2p_nb, self._p2p_bm, self._p2p_sel = nb, np.empty(nb, dtype=np.float64), np.empty(nb, dtype=np.uint8)
        capacity, (idx, pe, nc) = self._tie_cap, self._tie_buf
        best, n, over = ctypes.c_double(0.0), _c_i64(0), ctypes.c_int(0)
        rc = self._lib.kv_scm_best_rules_p2p(self._h, None, None, 0, 0, float(p), int(util_block_size),
                                             int(packet_capacity), ctypes.byref(best), capacity, ptr(idx), ptr(pe), ptr(nc),
                                             ctypes.byref(n), ptr(self._p2p_bm), ptr(self._p2p_sel), ctypes.byref(over))
        if rc == KV_E_OVERFLOW or (rc == KV_OK and over.value):
            return None
        check(rc)
        m = n.value
        return best.value, idx[:m].copy(), pe[:m].copy(), nc[:m].copy()

    def set_profiling(self, on=True):
        """CUDA events around the scan kernels (info().last_kernel_ms / scan_ms_total); off by default."""
        check(self._lib.kv_set_profiling(self._h, int(bool(on))))

    # -- a9 / a11
    def gini_scan(self, class_masks, altered_prior, n_total, n_node):
        masks = np.ascontiguousarray(class_masks, dtype=np.uint64)
        assert masks.ndim == 2 and masks.shape[1] == self.n_words
        pri = np.ascontiguousarray(altered_prior, dtype=np.float64)
        tot = np.ascontiguousarray(n_total, dtype=np.float64)
        nn = np.ascontiguousarray(n_node, dtype=np.int64)
        assert pri.shape[0] == tot.shape[0] == nn.shape[0] == masks.shape[0]
        out = ctypes.c_double(0.0)
        check(self._lib.kv_gini_scan(self._h, masks.shape[0], ptr(masks), ptr(pri), ptr(tot), ptr(nn),
                                     ctypes.byref(out)))
        return out.value

    def gini_best_split(self, class_masks, altered_prior, n_total, n_node, capacity=None):
        """Single-shard fused call -> (min score, rule indices achieving it, ascending)."""
        masks = np.ascontiguousarray(class_masks, dtype=np.uint64)
        assert masks.ndim == 2 and masks.shape[1] == self.n_words
        pri = np.ascontiguousarray(altered_prior, dtype=np.float64)
        tot = np.ascontiguousarray(n_total, dtype=np.float64)
        nn = np.ascontiguousarray(n_node, dtype=np.int64)
        if capacity is None:
            capacity, idx = self._tie_cap, self._tie_buf[0]
        else:
            idx = np.empty(capacity, dtype=np.int64)
        m, n = ctypes.c_double(0.0), _c_i64(0)
        rc = self._lib.kv_gini_best_split(self._h, masks.shape[0], ptr(masks), ptr(pri), ptr(tot), ptr(nn),
                                          ctypes.byref(m), capacity, ptr(idx), ctypes.byref(n))
        if rc == KV_E_OVERFLOW:      # the counts and the minimum are still on the device
            return m.value, self.gini_ties(m.value, int(n.value))
        check(rc)
        return m.value, idx[:n.value].copy()

    def gini_ties(self, min_score, capacity=4096):
        while True:
            idx = np.empty(capacity, dtype=np.int64)
            n = _c_i64(0)
            rc = self._lib.kv_gini_ties(self._h, float(min_score), capacity, ptr(idx), ctypes.byref(n))
            if rc == KV_E_OVERFLOW:
                capacity = int(n.value)
                continue
            check(rc)
            return idx[:n.value]

    def occ_table_set(self, table, mask):
        """Device-resident occurrence table ``table`` = presence counts under ``mask`` (None frees it)."""
        if mask is None:
            check(self._lib.kv_occ_table_set(self._h, int(table), None))
            return
        mask = np.ascontiguousarray(mask, dtype=np.uint64)
        assert mask.shape == (self.n_words,)
        check(self._lib.kv_occ_table_set(self._h, int(table), ptr(mask)))

    def occ_table_gather(self, table, rule_idx):
        idx = np.ascontiguousarray(rule_idx, dtype=np.int64)
        out = np.empty(idx.shape[0], dtype=np.uint32)
        check(self._lib.kv_occ_table_gather(self._h, int(table), ptr(idx), idx.shape[0], ptr(out)))
        return out

    def gini_level_family(self, node_keys=None, parent_keys=None, free=False):
        """Arms the next gini_level (kv_gini_level_family): sibling nodes whose parent was searched by the previous
        armed call are counted once -- the smaller one -- and the other is parent - sibling.  No keys: forget the
        resident level (free=True also returns the arena's memory)."""
        if node_keys is None:
            check(self._lib.kv_gini_level_family(self._h, -1 if free else 0, None, None))
            return
        keys = np.ascontiguousarray(node_keys, dtype=np.int64)
        parents = np.ascontiguousarray(parent_keys, dtype=np.int64)
        assert keys.shape == parents.shape and keys.ndim == 1
        check(self._lib.kv_gini_level_family(self._h, keys.shape[0], ptr(keys), ptr(parents)))

    def gini_level(self, node_masks, altered_prior, n_total, n_node, occ_tables=None, family=None):
        """Split search of several two-class nodes, up to 8 nodes per matrix pass.  node_masks [n_nodes, 2, W],
        n_node [n_nodes, 2] -> [(min score over this shard, rule indices achieving it, ascending)] per node.
        occ_tables [n_nodes] (table id per node, -1 = none): the occurrence tiebreaker is applied on the device; the
        result entries then carry a third element, the largest occurrence among the node's ties on this shard.
        family = (node keys, parent keys): see gini_level_family."""
        masks = np.ascontiguousarray(node_masks, dtype=np.uint64)
        assert masks.ndim == 3 and masks.shape[1] == 2 and masks.shape[2] == self.n_words
        if family is not None:
            assert len(family[0]) == masks.shape[0]
            self.gini_level_family(family[0], family[1])
        nn = np.ascontiguousarray(n_node, dtype=np.int64)
        # priors / class totals: one pair for all the nodes (one tree) or one pair per node (several trees)
        pri = np.ascontiguousarray(np.broadcast_to(np.asarray(altered_prior, dtype=np.float64), nn.shape))
        tot = np.ascontiguousarray(np.broadcast_to(np.asarray(n_total, dtype=np.float64), nn.shape))
        assert nn.shape == (masks.shape[0], 2)
        mins = np.empty(masks.shape[0], dtype=np.float64)
        found = np.empty(masks.shape[0], dtype=np.int64)
        occ, occ_max = None, None
        if occ_tables is not None:
            occ = np.ascontiguousarray(occ_tables, dtype=np.int32)
            assert occ.shape == (masks.shape[0],)
            occ_max = np.zeros(masks.shape[0], dtype=np.uint32)
        check(self._lib.kv_gini_level(self._h, masks.shape[0], ptr(masks), ptr(pri), ptr(tot), ptr(nn),
                                      ptr(occ) if occ is not None else None, ptr(mins), ptr(found),
                                      ptr(occ_max) if occ_max is not None else None))
        # read-only views into the shard's pinned result arena, valid until the next gini_level on this shard
        out = []
        for i in range(masks.shape[0]):
            if found[i] == 0:
                idx = np.zeros(0, dtype=np.int64)
            else:
                p_idx, n = ctypes.POINTER(ctypes.c_int64)(), _c_i64(0)
                check(self._lib.kv_gini_level_result(self._h, i, ctypes.byref(p_idx), ctypes.byref(n)))
                idx = np.ctypeslib.as_array(p_idx, shape=(int(n.value),))
                idx.flags.writeable = False
            out.append((float(mins[i]), idx) if occ is None else (float(mins[i]), idx, int(occ_max[i])))
        return out

    def gini_scores(self):
        out = np.empty(self.n_cols, dtype=np.float64)
        check(self._lib.kv_gini_scores(self._h, ptr(out)))
        return out


def _stack_job_masks(jobs, n_words):
    """uint64 [n_jobs, 2, W] of the jobs' (positive, negative) row masks; the jobs of a grid share a handful of mask
    arrays, which are stacked once and gathered by index."""
    n = len(jobs)
    slot, distinct = {}, []
    index = np.empty((n, 2), dtype=np.intp)
    for j, job in enumerate(jobs):
        for h in (0, 1):
            a = job[1 + h]
            k = slot.get(id(a))
            if k is None:
                k = slot[id(a)] = len(distinct)
                distinct.append(a)
            index[j, h] = k
    if not distinct:
        return np.empty((0, 2, n_words), dtype=np.uint64)
    table = np.ascontiguousarray(np.stack(distinct), dtype=np.uint64)
    assert table.shape[1] == n_words
    return table[index]


def n_utility_blocks(n_kmers_total, util_block_size):
    return int(load().kv_scm_n_blocks(int(n_kmers_total), int(util_block_size)))


def packet_bytes(n_blocks, capacity):
    return int(load().kv_scm_packet_bytes(int(n_blocks), int(capacity)))


TIE_FLAG_NEG_INF = 1 << 62
PACKET_REC_DTYPE = np.dtype([("idx", "<i8"), ("pe", "<u4"), ("nc", "<u4")])


def parse_packet(buf, n_blocks, capacity):
    """buf: uint8 [packet_bytes] -> (block_max f64[n_blocks], count, threshold, records[min(count, capacity)])."""
    bm = buf[:8 * n_blocks].view(np.float64)
    count = int(buf[8 * n_blocks:8 * n_blocks + 8].view(np.uint64)[0])
    thr = float(buf[8 * n_blocks + 8:8 * n_blocks + 16].view(np.float64)[0])
    recs = buf[8 * n_blocks + 16:8 * n_blocks + 16 + 16 * capacity].view(PACKET_REC_DTYPE)
    return bm, count, thr, recs[:min(count, capacity)]


def merge_packets(rows, n_blocks, packet_capacity, p, util_block_size):
    """kv_scm_merge_packets over uint8 [n_packets, packet_bytes].
    -> ((best, idx, pos_err, neg_cover) or None on packet overflow, merged block maxima, best, selected)."""
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    assert rows.ndim == 2 and rows.shape[1] == packet_bytes(n_blocks, packet_capacity)
    capacity = max(1, rows.shape[0] * packet_capacity)
    bm, sel = np.empty(n_blocks, dtype=np.float64), np.empty(n_blocks, dtype=np.uint8)
    idx, pe, nc = (np.empty(capacity, dtype=np.int64) for _ in range(3))
    best, n, over = ctypes.c_double(0.0), _c_i64(0), ctypes.c_int(0)
    check(load().kv_scm_merge_packets(ptr(rows), rows.shape[0], int(n_blocks), int(packet_capacity), float(p),
                                      int(util_block_size), ctypes.byref(best), capacity, ptr(idx), ptr(pe), ptr(nc),
                                      ctypes.byref(n), ptr(bm), ptr(sel), ctypes.byref(over)))
    if over.value:
        return None, bm, best.value, sel
    return (best.value, idx[:n.value], pe[:n.value], nc[:n.value]), bm, best.value, sel


def select_blocks(block_max):
    """Replay of the reference's sequential utility-block merge (scm.py:270-286) -> (best, selected)."""
    bm = np.ascontiguousarray(block_max, dtype=np.float64)
    sel = np.zeros(bm.shape[0], dtype=np.uint8)
    best = ctypes.c_double(0.0)
    check(load().kv_scm_select_blocks(ptr(bm), bm.shape[0], ctypes.byref(best), ptr(sel)))
    return best.value, sel


def inplace_popcount(arr, row_mask, device=0):
    lib = load()
    if arr.dtype == np.uint64:
        fn, t = lib.kv_inplace_popcount_64, np.uint64
    elif arr.dtype == np.uint32:
        fn, t = lib.kv_inplace_popcount_32, np.uint32
    else:
        raise ValueError("Buffer dtype mismatch: expected uint32 or uint64")
    if arr.ndim != 2:
        raise ValueError("Buffer has wrong number of dimensions (expected 2, got %d)" % arr.ndim)
    mask = np.ascontiguousarray(row_mask, dtype=t)
    if mask.ndim != 1 or mask.shape[0] < arr.shape[0]:
        raise ValueError("row_mask must have one entry per row of arr")
    if arr.flags.c_contiguous:
        check(fn(int(device), ptr(arr), arr.shape[0], arr.shape[1], ptr(mask)))
    else:  # typed memoryviews accept strided buffers; round-trip through a contiguous copy
        tmp = np.ascontiguousarray(arr)
        check(fn(int(device), ptr(tmp), tmp.shape[0], tmp.shape[1], ptr(mask)))
        arr[...] = tmp
