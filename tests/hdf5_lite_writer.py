"""Minimal HDF5 writer: produces files in the same on-disk structures the HDF5 library's default ("earliest")
format uses for Kover datasets -- version-0 superblock, version-1 object headers, symbol-table groups (v1 B-tree +
local heap + SNOD nodes), contiguous or chunked + deflate datasets indexed by a v1 chunk B-tree, version-1
attribute messages, fixed-length and variable-length (global heap) strings -- following the HDF5 File Format
Specification version 3.0.

It exists because this image has neither h5py nor libhdf5: it lets the tests, the bench and the tools put a
Kover-layout dataset on disk (reference dataset/create.py:160-240, split.py:150-230) and read it back through
``hdf5_lite`` / the upload path.  Files it writes are plain HDF5 (`h5dump`-able); with h5py installed use h5py.

    w = Writer("example.kover")
    w.attrs("/", {"uuid": "...", "compression": "gzip (level 4)"})
    w.create_dataset("kmer_matrix", packed, chunks=(1, 100000), compression=4)
    w.create_dataset("splits/s1/train_genome_idx", idx, attrs={...})
    w.close()
"""
import struct
import zlib

import numpy as np

_UNDEF = 0xFFFFFFFFFFFFFFFF
_GROUP_LEAF_K, _GROUP_INTERNAL_K, _CHUNK_K = 4, 16, 32


def _pad8(b):
    return b + b"\x00" * (-len(b) % 8)


def _msg(mtype, body, flags=0):
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


class _Node(object):
    def __init__(self):
        self.children = {}          # name -> _Node (groups only)
        self.attrs = {}
        self.dataset = None         # dict(data=..., chunks=..., compression=...)


class Writer(object):
    def __init__(self, path):
        self.path = path
        self.root = _Node()
        self.buf = bytearray()
        self._gheap = []            # (index, bytes) pending global-heap objects

    # ---- building the tree
    def _node(self, name, create=True):
        node = self.root
        for part in [q for q in name.split("/") if q]:
            if part not in node.children:
                if not create:
                    raise KeyError(name)
                node.children[part] = _Node()
            node = node.children[part]
        return node

    def create_group(self, name):
        return self._node(name)

    def attrs(self, name, values):
        self._node(name).attrs.update(values)

    def create_dataset(self, name, data, chunks=None, compression=None, attrs=None):
        node = self._node(name)
        node.dataset = dict(data=data, chunks=chunks, compression=compression)
        if attrs:
            node.attrs.update(attrs)

    # ---- allocation
    def _alloc(self, data):
        self.buf += b"\x00" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    # ---- datatypes / dataspaces / values
    @staticmethod
    def _dataspace(shape):
        rank = len(shape)
        return struct.pack("<BBB5x", 1, rank, 0) + b"".join(struct.pack("<Q", int(n)) for n in shape)

    @staticmethod
    def _datatype(dt):
        if dt == "vlen_str":
            base = struct.pack("<BBBBI", 0x10 | 3, 0x10, 0, 0, 1)                    # 1-byte UTF-8 string, null-terminated
            return struct.pack("<BBBBI", 0x10 | 9, 0x01, 0x01, 0, 16) + base        # vlen of string, UTF-8
        dt = np.dtype(dt)
        if dt.kind in "iu":
            b0 = (1 if dt.byteorder == ">" else 0) | (0x08 if dt.kind == "i" else 0)
            return struct.pack("<BBBBIHH", 0x10 | 0, b0, 0, 0, dt.itemsize, 0, dt.itemsize * 8)
        if dt.kind == "f":
            exp_bits, mant_bits, bias = {2: (5, 10, 15), 4: (8, 23, 127), 8: (11, 52, 1023)}[dt.itemsize]
            b0 = (1 if dt.byteorder == ">" else 0) | 0x20
            return struct.pack("<BBBBIHHBBBBI", 0x10 | 1, b0, dt.itemsize * 8 - 1, 0, dt.itemsize, 0, dt.itemsize * 8,
                               mant_bits, exp_bits, 0, mant_bits, bias)
        if dt.kind == "S":
            return struct.pack("<BBBBI", 0x10 | 3, 0x01, 0, 0, dt.itemsize)          # null-padded ASCII
        raise TypeError("hdf5_lite_writer: unsupported dtype %s" % dt)

    def _encode(self, value):
        """-> (datatype message body, shape, raw bytes) for an attribute value or dataset array."""
        if isinstance(value, str):
            return self._datatype("vlen_str"), (), self._vlen([value])
        if isinstance(value, bytes):
            return self._datatype(np.dtype("S%d" % max(1, len(value)))), (), value.ljust(max(1, len(value)), b"\x00")
        a = np.asarray(value)
        if a.dtype.kind == "U" or (a.dtype == object and all(isinstance(x, str) for x in a.reshape(-1))):
            return self._datatype("vlen_str"), a.shape, self._vlen([str(x) for x in a.reshape(-1)])
        if a.dtype == bool:
            a = a.astype(np.uint8)
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("<"))
        return self._datatype(a.dtype), a.shape, np.ascontiguousarray(a).tobytes()

    def _vlen(self, strings):
        out = b""
        for s in strings:
            raw = s.encode("utf-8")
            self._gheap.append(raw)
            out += struct.pack("<IQI", len(raw), 0, len(self._gheap))            # heap address patched at flush
        return out

    def _flush_gheap(self, blobs):
        """Write the pending vlen strings as ONE global heap collection; patch its address into the descriptors
        (every descriptor in `blobs` -- a list of bytearrays -- whose heap address is still 0)."""
        if not self._gheap:
            return
        body = b""
        for i, raw in enumerate(self._gheap):
            body += struct.pack("<HHIQ", i + 1, 1, 0, len(raw)) + _pad8(raw)
        total = max(4096, 16 + len(body) + 16)
        free = total - 16 - len(body)
        coll = b"GCOL" + struct.pack("<B3xQ", 1, total) + body + struct.pack("<HHIQ", 0, 0, 0, free)
        coll += b"\x00" * (total - len(coll))
        addr = self._alloc(coll)
        for blob in blobs:
            for p in range(0, len(blob), 16):
                blob[p + 4:p + 12] = struct.pack("<Q", addr)
        self._gheap = []

    def _attr_msgs(self, attrs):
        msgs = []
        for name in sorted(attrs):
            dt, shape, raw = self._encode(attrs[name])
            raw = bytearray(raw)
            if dt[0] & 0x0F == 9:
                self._flush_gheap([raw])
            nm = name.encode("utf-8") + b"\x00"
            ds = self._dataspace(shape)
            body = struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + bytes(raw)
            msgs.append(_msg(0x0C, body))
        return msgs

    def _object_header(self, msgs):
        body = b"".join(msgs)
        return self._alloc(struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body)

    # ---- datasets
    def _write_dataset(self, node):
        d = node.dataset
        big = None
        if isinstance(d["data"], np.ndarray) and d["data"].dtype.kind in "ui" and d["chunks"] is not None \
                and d["data"].dtype.byteorder != ">":
            # a large numeric matrix written in chunks: no intermediate copies of the whole array
            big = np.ascontiguousarray(d["data"])
            dt, shape, raw = self._datatype(big.dtype), big.shape, b""
        else:
            dt, shape, raw = self._encode(d["data"])
        raw = bytearray(raw)
        if dt[0] & 0x0F == 9:
            self._flush_gheap([raw])
        item = struct.unpack_from("<I", dt, 4)[0]
        msgs = [_msg(0x01, self._dataspace(shape)), _msg(0x03, dt, flags=1),
                _msg(0x05, struct.pack("<BBBB", 2, 2, 2, 0))]                      # fill value: v2, alloc late, never written
        chunks = d["chunks"]
        if chunks is None and d["compression"] is not None:
            chunks = tuple(max(1, n) for n in shape)
        if chunks is None:
            addr = self._alloc(bytes(raw)) if raw else _UNDEF
            msgs.append(_msg(0x08, struct.pack("<BBQQ", 3, 1, addr, len(raw))))
        else:
            level = d["compression"]
            if level is not None:
                msgs.append(_msg(0x0B, struct.pack("<BB6x", 1, 1) +
                                 struct.pack("<HHHH", 1, 8, 1, 1) + _pad8(b"deflate\x00") + struct.pack("<I4x", 4 if callable(level) else int(level))))
            arr = big if big is not None else np.frombuffer(bytes(raw), dtype=np.dtype("V%d" % item)).reshape(shape)
            entries = []
            grid = [range(0, max(n, 1), c) for n, c in zip(shape, chunks)]
            offsets = [()]
            for r in grid:
                offsets = [t + (o,) for t in offsets for o in r]
            def stored(offs):
                block = np.zeros(chunks, dtype=arr.dtype)
                src = tuple(slice(o, min(o + c, n)) for o, c, n in zip(offs, chunks, shape))
                part = arr[src]
                block[tuple(slice(0, s) for s in part.shape)] = part
                data = block.tobytes()
                if callable(level):                     # a caller-made zlib stream (stored / fixed-Huffman blocks, ...)
                    data = level(data)
                elif level is not None:
                    data = zlib.compress(data, int(level))
                return data

            if len(offsets) > 16:                        # zlib releases the GIL: deflate the chunks on all the cores
                import os
                from concurrent.futures import ThreadPoolExecutor
                with ThreadPoolExecutor(max_workers=max(2, min(32, os.cpu_count() or 4))) as pool:
                    blobs = list(pool.map(stored, offsets))
            else:
                blobs = [stored(offs) for offs in offsets]
            for offs, data in zip(offsets, blobs):
                entries.append((offs, self._alloc(data), len(data)))
            btree = self._chunk_btree(entries, shape, chunks, item)
            msgs.append(_msg(0x08, struct.pack("<BBB", 3, 2, len(shape) + 1) + struct.pack("<Q", btree) +
                             b"".join(struct.pack("<I", c) for c in chunks) + struct.pack("<I", item)))
        msgs += self._attr_msgs(node.attrs)
        return self._object_header(msgs)

    def _chunk_btree(self, entries, shape, chunks, item):
        rank = len(shape)

        def key(size, offs):
            return struct.pack("<II", size, 0) + b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<Q", 0)

        if not entries:
            return _UNDEF
        end_key = key(0, tuple(((n + c - 1) // c) * c for n, c in zip(shape, chunks)))
        level, nodes = 0, [(offs, addr, size) for offs, addr, size in entries]      # (first offsets, child address, size)
        while True:
            parents = []
            for g in range(0, len(nodes), 2 * _CHUNK_K):
                group = nodes[g:g + 2 * _CHUNK_K]
                body = b"TREE" + struct.pack("<BBHQQ", 1, level, len(group), _UNDEF, _UNDEF)
                for offs, addr, size in group:
                    body += key(size, offs) + struct.pack("<Q", addr)
                nxt = nodes[g + 2 * _CHUNK_K][0] if g + 2 * _CHUNK_K < len(nodes) else None
                body += key(0, nxt) if nxt is not None else end_key
                parents.append((group[0][0], self._alloc(body), 0))
            if len(parents) == 1:
                return parents[0][1]
            # sibling pointers are left undefined: readers that walk the tree from the root do not need them
            nodes, level = parents, level + 1

    # ---- groups
    def _write_group(self, node):
        names = sorted(node.children, key=lambda s: s.encode("utf-8"))
        child_addr = {}
        for nm in names:
            ch = node.children[nm]
            child_addr[nm] = self._write_dataset(ch) if ch.dataset is not None else self._write_group(ch)[0]
        # local heap: offset 0 holds the empty string; names are 8-byte aligned
        heap, name_off = bytearray(b"\x00" * 8), {}
        for nm in names:
            name_off[nm] = len(heap)
            heap += _pad8(nm.encode("utf-8") + b"\x00")
        heap_size = max(len(heap) + 16, 88)
        free_off = len(heap)
        heap += struct.pack("<QQ", 1, heap_size - free_off)                          # one free block: next = 1 (none), size
        heap += b"\x00" * (heap_size - len(heap))
        data_addr = self._alloc(bytes(heap))
        heap_addr = self._alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, heap_size, free_off, data_addr))
        # symbol nodes of up to 2K entries under ONE B-tree node
        snods = []
        for g in range(0, max(len(names), 1), 2 * _GROUP_LEAF_K):
            group = names[g:g + 2 * _GROUP_LEAF_K]
            body = b"SNOD" + struct.pack("<BxH", 1, len(group))
            for nm in group:
                body += struct.pack("<QQI4x16x", name_off[nm], child_addr[nm], 0)
            body += b"\x00" * (8 + 2 * _GROUP_LEAF_K * 40 - len(body))
            snods.append((name_off[group[-1]] if group else 0, self._alloc(body)))
        if len(snods) > 2 * _GROUP_INTERNAL_K:
            raise NotImplementedError("hdf5_lite_writer: more than %d links in one group" % (4 * _GROUP_LEAF_K * _GROUP_INTERNAL_K))
        body = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), _UNDEF, _UNDEF) + struct.pack("<Q", 0)
        for last_name_off, addr in snods:
            body += struct.pack("<QQ", addr, last_name_off)
        body += b"\x00" * (24 + 8 + 2 * _GROUP_INTERNAL_K * 16 - len(body))
        btree_addr = self._alloc(body)
        msgs = [_msg(0x11, struct.pack("<QQ", btree_addr, heap_addr))] + self._attr_msgs(node.attrs)
        return self._object_header(msgs), btree_addr, heap_addr

    def close(self):
        self.buf = bytearray(b"\x00" * 96)                                           # superblock placeholder
        root_addr, btree, heap = self._write_group(self.root)
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, _GROUP_LEAF_K, _GROUP_INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, _UNDEF, len(self.buf), _UNDEF)
        sb += struct.pack("<QQI4xQQ", 0, root_addr, 1, btree, heap)
        self.buf[:len(sb)] = sb
        with open(self.path, "wb") as fh:
            fh.write(self.buf)

    def __enter__(self):
        return self

    def __exit__(self, exc_type, *exc):
        if exc_type is None:
            self.close()
