"""Minimal read-only HDF5 reader for Kover datasets on machines without h5py.

A Kover dataset (reference dataset/create.py:160-240, split.py:150-230, tools/kmer_tools/src/kl_64.cpp:229-263) is
an HDF5 file written with the library's default ("earliest") format: version-0 superblock, version-1 object
headers, groups as symbol tables (v1 B-tree + local heap), the packed ``kmer_matrix`` as a chunked, gzip-filtered
2-D uint64 / uint32 dataset indexed by a v1 B-tree, small datasets contiguous or compact, attributes and
``kmer_sequences`` / ``genome_identifiers`` as fixed- or variable-length strings (global heap).  This module reads
exactly that subset of the HDF5 File Format Specification (version 3.0, sections III.A-III.E, IV.A) and exposes
it with the slice of the h5py API the reference uses: ``File[...]``, ``Group.keys() / items() / attrs /
__contains__``, ``Dataset.shape / dtype / chunks / compression / attrs / __getitem__ / __len__``.

It is host-side I/O only -- the counterpart of ``h5py.File`` in reference utils.py:29-47 -- and is used when h5py
is not importable; with h5py installed ``kover_b200.dataset.ds`` opens files with h5py.  Not supported (raises
``NotImplementedError``): version-2 object headers / new-style groups, fractal heaps, v2 B-trees, virtual and
external storage, filters other than deflate / shuffle / fletcher32, compound / array / enum / reference types.
"""
import os
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_SIGNATURE = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF
_POOL = None


def _pool():
    """One inflate pool per process (zlib releases the GIL): chunk reads of a slab run on all the host cores."""
    global _POOL
    if _POOL is None:
        _POOL = ThreadPoolExecutor(max_workers=max(2, min(32, os.cpu_count() or 8)))
    return _POOL


class _VlenStr(object):
    """Marker dtype: variable-length string stored in the global heap."""

    def __init__(self, utf8):
        self.utf8 = utf8


def _align8(n):
    return (n + 7) & ~7


class File(object):
    """``File(path)`` -> the root group.  Read-only; the file is memory-mapped."""

    def __init__(self, path, mode="r", **_ignored):
        if mode not in ("r",):
            raise ValueError("hdf5_lite is read-only")
        self.filename = path
        self._fh = open(path, "rb")
        size = os.fstat(self._fh.fileno()).st_size
        self._buf = np.memmap(self._fh, dtype=np.uint8, mode="r") if size else np.zeros(0, dtype=np.uint8)
        self._mv = memoryview(self._buf)
        self._gheaps = {}
        self._locate_superblock(size)
        self._root = Group(self, self._root_header, "/")

    # ---- low level
    def _read(self, addr, n):
        a = self._base + addr
        if addr == _UNDEF or a + n > len(self._mv):
            raise ValueError("HDF5: address out of range (truncated or unsupported file)")
        return bytes(self._mv[a:a + n])

    def _locate_superblock(self, size):
        off = 0
        while off + 8 <= size:                       # the superblock sits at 0, 512, 1024, 2048, ... (user block)
            if bytes(self._mv[off:off + 8]) == _SIGNATURE:
                break
            off = 512 if off == 0 else off * 2
        else:
            raise ValueError("not an HDF5 file: %s" % self.filename)
        version = self._mv[off + 8]
        self._base = 0
        if version in (0, 1):
            size_off, size_len = self._mv[off + 13], self._mv[off + 14]
            if size_off != 8 or size_len != 8:
                raise NotImplementedError("HDF5: only 8-byte offsets and lengths are supported")
            p = off + 24 + (4 if version == 1 else 0)
            base, _free, _eof, _drv = struct.unpack_from("<QQQQ", self._mv, p)
            self._base = base if base != _UNDEF else 0
            # root group symbol table entry: link name offset, object header address, cache type, reserved, scratch
            self._root_header = struct.unpack_from("<Q", self._mv, p + 32 + 8)[0]
        elif version in (2, 3):
            if self._mv[off + 9] != 8 or self._mv[off + 10] != 8:
                raise NotImplementedError("HDF5: only 8-byte offsets and lengths are supported")
            base, _ext, _eof, root = struct.unpack_from("<QQQQ", self._mv, off + 12)
            self._base = base if base != _UNDEF else 0
            self._root_header = root
        else:
            raise NotImplementedError("HDF5: unknown superblock version %d" % version)

    def _global_heap_object(self, addr, index):
        heap = self._gheaps.get(addr)
        if heap is None:
            head = self._read(addr, 16)
            if head[:4] != b"GCOL":
                raise ValueError("HDF5: bad global heap collection")
            total = struct.unpack_from("<Q", head, 8)[0]
            raw = self._read(addr, total)
            heap, p = {}, 16
            while p + 16 <= total:
                idx, _refs, _res, n = struct.unpack_from("<HHIQ", raw, p)
                if idx == 0:
                    break
                heap[idx] = raw[p + 16:p + 16 + n]
                p += 16 + _align8(n)
            self._gheaps[addr] = heap
        return heap[index]

    # ---- h5py-like surface
    def __getitem__(self, name):
        return self._root[name]

    def __contains__(self, name):
        return name in self._root

    def keys(self):
        return self._root.keys()

    def items(self):
        return self._root.items()

    @property
    def attrs(self):
        return self._root.attrs

    def close(self):
        self._mv.release()
        del self._buf
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class _Object(object):
    """Object header (version 1): the list of (type, flags, body) messages, continuation blocks followed."""

    def __init__(self, f, addr, name):
        self._f, self.name = f, name
        self._messages = self._read_header(addr)
        self._attrs = None

    def _read_header(self, addr):
        f = self._f
        head = f._read(addr, 16)
        if head[:4] == b"OHDR":
            raise NotImplementedError("HDF5: version-2 object headers (libver='latest' files) are not supported")
        version, _res, n_msgs, _refs, hdr_size = struct.unpack_from("<BBHII", head, 0)
        if version != 1:
            raise NotImplementedError("HDF5: object header version %d" % version)
        msgs, blocks = [], [(addr + 16, hdr_size)]
        while blocks and len(msgs) < n_msgs:
            a, n = blocks.pop(0)
            raw, p = f._read(a, n), 0
            while p + 8 <= n and len(msgs) < n_msgs:
                mtype, msize, mflags = struct.unpack_from("<HHB", raw, p)
                body = raw[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x10:                                  # object header continuation
                    blocks.append(struct.unpack_from("<QQ", body, 0))
                msgs.append((mtype, mflags, body))
        return msgs

    def _first(self, mtype):
        for t, _fl, body in self._messages:
            if t == mtype:
                return body
        return None

    @property
    def attrs(self):
        if self._attrs is None:
            self._attrs = {}
            for t, _fl, body in self._messages:
                if t == 0x0C:
                    k, v = self._parse_attribute(body)
                    self._attrs[k] = v
        return self._attrs

    def _parse_attribute(self, body):
        version = body[0]
        if version == 1:
            name_size, dt_size, ds_size = struct.unpack_from("<HHH", body, 2)
            p = 8
            pad = _align8
        elif version in (2, 3):
            name_size, dt_size, ds_size = struct.unpack_from("<HHH", body, 2)
            p = 8 + (1 if version == 3 else 0)
            pad = lambda n: n
            if body[1] & 0x03:
                raise NotImplementedError("HDF5: shared attribute datatypes / dataspaces")
        else:
            raise NotImplementedError("HDF5: attribute message version %d" % version)
        name = body[p:p + name_size].split(b"\x00", 1)[0].decode("utf-8")
        p += pad(name_size)
        dtype = _parse_datatype(body[p:p + dt_size])
        p += pad(dt_size)
        shape = _parse_dataspace(body[p:p + ds_size])
        p += pad(ds_size)
        value = _decode_values(self._f, dtype, shape, body[p:])
        return name, value


class Group(_Object):
    def __init__(self, f, addr, name):
        super(Group, self).__init__(f, addr, name)
        self._links = None

    def _load_links(self):
        if self._links is not None:
            return self._links
        links = {}
        stab = self._first(0x11)
        if stab is not None:
            btree, heap = struct.unpack_from("<QQ", stab, 0)
            heap_data = self._local_heap(heap)
            self._walk_group_btree(btree, heap_data, links)
        elif self._first(0x02) is not None or self._first(0x06) is not None:
            raise NotImplementedError("HDF5: new-style groups (link messages) are not supported")
        self._links = links
        return links

    def _local_heap(self, addr):
        head = self._f._read(addr, 32)
        if head[:4] != b"HEAP":
            raise ValueError("HDF5: bad local heap")
        size, _free, data_addr = struct.unpack_from("<QQQ", head, 8)
        return self._f._read(data_addr, size)

    def _walk_group_btree(self, addr, heap, links):
        f = self._f
        head = f._read(addr, 24)
        if head[:4] == b"SNOD":
            n = struct.unpack_from("<H", head, 6)[0]
            raw = f._read(addr + 8, n * 40)
            for i in range(n):
                name_off, obj_addr = struct.unpack_from("<QQ", raw, i * 40)
                links[heap[name_off:heap.index(b"\x00", name_off)].decode("utf-8")] = obj_addr
            return
        if head[:4] != b"TREE" or head[4] != 0:
            raise ValueError("HDF5: bad group B-tree node")
        n = struct.unpack_from("<H", head, 6)[0]
        raw = f._read(addr + 24, (2 * n + 1) * 8)
        for i in range(n):                                   # key0 child0 key1 child1 ... keyN
            self._walk_group_btree(struct.unpack_from("<Q", raw, (2 * i + 1) * 8)[0], heap, links)

    def keys(self):
        return sorted(self._load_links().keys())             # h5py iterates groups in name order

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self._load_links())

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    iteritems = items

    def values(self):
        return [self[k] for k in self.keys()]

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __getitem__(self, name):
        node = self
        for part in [q for q in name.split("/") if q]:
            if not isinstance(node, Group):
                raise KeyError(name)
            links = node._load_links()
            if part not in links:
                raise KeyError("Unable to open object (object '%s' doesn't exist)" % part)
            node = _open_object(node._f, links[part], (node.name.rstrip("/") + "/" + part))
        return node


def _open_object(f, addr, name):
    obj = _Object(f, addr, name)
    if obj._first(0x08) is not None:                           # data layout message -> dataset
        d = Dataset.__new__(Dataset)
        d.__dict__.update(obj.__dict__)
        d._setup()
        return d
    g = Group.__new__(Group)
    g.__dict__.update(obj.__dict__)
    g._links = None
    return g


# ------------------------------------------------------------------------------------------------ messages
def _parse_dataspace(body):
    version, rank, flags = body[0], body[1], body[2]
    if version == 1:
        p = 8
    elif version == 2:
        p = 4
        if body[3] == 2:                                       # null dataspace
            return None
    else:
        raise NotImplementedError("HDF5: dataspace version %d" % version)
    return tuple(struct.unpack_from("<%dQ" % rank, body, p)) if rank else ()


def _parse_datatype(body):
    cls, version = body[0] & 0x0F, body[0] >> 4
    b0, b1 = body[1], body[2]
    size = struct.unpack_from("<I", body, 4)[0]
    if cls == 0:                                               # fixed point
        order = ">" if b0 & 1 else "<"
        return np.dtype("%s%s%d" % (order, "i" if b0 & 0x08 else "u", size))
    if cls == 1:                                               # floating point (IEEE layouts only)
        order = ">" if b0 & 1 else "<"
        if size not in (2, 4, 8):
            raise NotImplementedError("HDF5: %d-byte floats" % size)
        return np.dtype("%sf%d" % (order, size))
    if cls == 3:                                               # fixed-length string
        return np.dtype("S%d" % size)
    if cls == 9:                                               # variable length
        if (b0 & 0x0F) != 1:
            raise NotImplementedError("HDF5: variable-length sequences (only strings are supported)")
        return _VlenStr(utf8=(b1 & 0x0F) == 1)
    raise NotImplementedError("HDF5: datatype class %d (version %d)" % (cls, version))


def _decode_values(f, dtype, shape, raw):
    if shape is None:
        return None
    count = int(np.prod(shape, dtype=np.int64)) if shape else 1
    if isinstance(dtype, _VlenStr):
        out = []
        for i in range(count):
            n, addr, idx = struct.unpack_from("<IQI", raw, i * 16)
            data = f._global_heap_object(addr, idx)[:n] if n else b""
            out.append(data.decode("utf-8") if dtype.utf8 else data.decode("latin-1"))
        if not shape:
            return out[0]
        return np.array(out, dtype=object).reshape(shape)
    arr = np.frombuffer(raw, dtype=dtype, count=count)
    if not shape:
        v = arr[0]
        return v if dtype.kind != "S" else bytes(v)
    return arr.reshape(shape).copy()


# ------------------------------------------------------------------------------------------------ datasets
class Dataset(_Object):
    def _setup(self):
        self.shape = _parse_dataspace(self._first(0x01))
        self._dtype = _parse_datatype(self._first(0x03))
        self._filters = self._parse_filters(self._first(0x0B))
        layout = self._first(0x08)
        version = layout[0]
        self.chunks = None
        if version == 3:
            cls = layout[1]
            if cls == 0:                                       # compact: data in the header
                size = struct.unpack_from("<H", layout, 2)[0]
                self._kind, self._compact = "compact", layout[4:4 + size]
            elif cls == 1:
                self._kind = "contiguous"
                self._addr, self._size = struct.unpack_from("<QQ", layout, 2)
            elif cls == 2:
                rank = layout[2]
                self._kind = "chunked"
                self._btree = struct.unpack_from("<Q", layout, 3)[0]
                dims = struct.unpack_from("<%dI" % rank, layout, 11)
                self.chunks = tuple(dims[:-1])                 # the last "dimension" is the element size
                self._index = self._arrays = None
            else:
                raise NotImplementedError("HDF5: data layout class %d" % cls)
        elif version in (1, 2):
            rank, cls = layout[1], layout[2]
            p = 8
            if cls == 0:
                dims = struct.unpack_from("<%dI" % rank, layout, p)
                size = struct.unpack_from("<I", layout, p + 4 * rank)[0]
                self._kind, self._compact = "compact", layout[p + 4 * rank + 4:p + 4 * rank + 4 + size]
            else:
                addr = struct.unpack_from("<Q", layout, p)[0]
                dims = struct.unpack_from("<%dI" % rank, layout, p + 8)
                if cls == 1:
                    self._kind, self._addr = "contiguous", addr
                else:
                    self._kind, self._btree, self.chunks, self._index = "chunked", addr, tuple(dims[:-1]), None
                    self._arrays = None
        else:
            raise NotImplementedError("HDF5: data layout message version %d" % version)

    @staticmethod
    def _parse_filters(body):
        if body is None:
            return []
        version, n = body[0], body[1]
        p = 8 if version == 1 else 2
        out = []
        for _ in range(n):
            fid = struct.unpack_from("<H", body, p)[0]
            if version == 1 or fid >= 256:
                name_len, _flags, n_vals = struct.unpack_from("<HHH", body, p + 2)
                p += 8
                p += _align8(name_len) if version == 1 else name_len
            else:
                _flags, n_vals = struct.unpack_from("<HH", body, p + 2)
                p += 6
            vals = struct.unpack_from("<%dI" % n_vals, body, p)
            p += 4 * n_vals
            if version == 1 and n_vals % 2:
                p += 4
            out.append((fid, vals))
        return out

    # ---- h5py-like surface
    @property
    def dtype(self):
        if isinstance(self._dtype, _VlenStr):
            return np.dtype(object)
        return self._dtype

    @property
    def compression(self):
        return "gzip" if any(fid == 1 for fid, _ in self._filters) else None

    @property
    def compression_opts(self):
        for fid, vals in self._filters:
            if fid == 1:
                return vals[0] if vals else None
        return None

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64))

    def __len__(self):
        return self.shape[0]

    def _item_size(self):
        return 16 if isinstance(self._dtype, _VlenStr) else self._dtype.itemsize

    def _unfilter(self, raw, mask=0):
        for i in range(len(self._filters) - 1, -1, -1):        # undo the pipeline in reverse
            if mask & (1 << i):
                continue
            fid, vals = self._filters[i]
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:
                width = vals[0] if vals else self._item_size()
                a = np.frombuffer(raw, dtype=np.uint8)
                n = len(a) // width
                raw = a[:n * width].reshape(width, n).T.tobytes() + a[n * width:].tobytes()
            elif fid == 3:
                raw = raw[:-4]
            else:
                raise NotImplementedError("HDF5: filter %d" % fid)
        return raw

    def _chunk_arrays(self):
        """The v1 chunk B-tree as arrays: (chunk offsets int64 [n, rank], address uint64 [n], stored size [n],
        filter mask [n]).  Leaf nodes are parsed with one numpy view each (a Kover matrix of 50 000 genomes x 100 M
        k-mers has 782 000 chunks)."""
        if getattr(self, "_arrays", None) is None:
            parts = []
            if self._btree != _UNDEF:
                self._walk_chunk_btree(self._btree, parts)
            rank = len(self.shape)
            if parts:
                leaf = np.concatenate(parts)
                self._arrays = (leaf["offs"][:, :rank].astype(np.int64), leaf["child"].astype(np.uint64),
                                leaf["size"].astype(np.int64), leaf["mask"].astype(np.uint32))
            else:
                self._arrays = (np.zeros((0, rank), dtype=np.int64), np.zeros(0, dtype=np.uint64),
                                np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.uint32))
        return self._arrays

    def _chunk_index(self):
        """{chunk offset tuple: (address, stored size, filter mask)} from the v1 chunk B-tree."""
        if self._index is None:
            offs, addr, size, mask = self._chunk_arrays()
            self._index = dict(zip(map(tuple, offs.tolist()), zip(addr.tolist(), size.tolist(), mask.tolist())))
        return self._index

    def _walk_chunk_btree(self, addr, parts):
        f = self._f
        rank = len(self.shape)
        head = f._read(addr, 24)
        if head[:4] != b"TREE" or head[4] != 1:
            raise ValueError("HDF5: bad chunk B-tree node")
        level, n = head[5], struct.unpack_from("<H", head, 6)[0]
        key_size = 8 + 8 * (rank + 1)
        raw = f._read(addr + 24, n * (key_size + 8) + key_size)
        entry = np.dtype([("size", "<u4"), ("mask", "<u4"), ("offs", "<u8", (rank + 1,)), ("child", "<u8")])
        nodes = np.frombuffer(raw, dtype=entry, count=n)
        if level == 0:
            parts.append(nodes)
        else:
            for child in nodes["child"].tolist():
                self._walk_chunk_btree(child, parts)

    def deflate_chunk_table(self, row0, row1, col0, col1):
        """The raw gzip streams of the chunks that intersect rows [row0, row1) x columns [col0, col1) of a 2-D chunked
        dataset whose ONLY filter is deflate (what `kover dataset create` writes: create.py:222-230, kl_64.cpp:229-263),
        so that they can be inflated on the device instead of on the host.
        -> None when the dataset is not of that kind, else a dict of arrays, one entry per allocated chunk:
        ``address`` (host address of the stream inside the memory-mapped file), ``nbytes``, ``row0``, ``col0``, and
        ``raw`` = True for chunks stored with the filter skipped (filter mask bit 0): those need the host path."""
        if self._kind != "chunked" or len(self.shape) != 2 or [fid for fid, _ in self._filters] != [1]:
            return None
        if isinstance(self._dtype, _VlenStr) or self._dtype not in (np.dtype("<u8"), np.dtype("<u4")):
            return None
        offs, addr, nbytes, mask = self._chunk_arrays()       # (unallocated chunks are absent: fill value, zero)
        cr, cc = self.chunks
        r1, c1 = min(row1, self.shape[0]), min(col1, self.shape[1])
        hit = (offs[:, 0] + cr > row0) & (offs[:, 0] < r1) & (offs[:, 1] + cc > col0) & (offs[:, 1] < c1)
        sel = np.flatnonzero(hit)
        sel = sel[np.lexsort((offs[sel, 1], offs[sel, 0]))]   # row-major order, like the reference's block loop
        base = self._f._buf.ctypes.data + self._f._base
        if sel.size and int((addr[sel].astype(np.int64) + nbytes[sel]).max()) + self._f._base > len(self._f._mv):
            raise ValueError("HDF5: address out of range (truncated or unsupported file)")
        return {"address": addr[sel] + np.uint64(base), "nbytes": nbytes[sel],
                "row0": offs[sel, 0].copy(), "col0": offs[sel, 1].copy(),
                "raw": (mask[sel] & 1).astype(bool), "chunks": (cr, cc)}

    def _read_chunk(self, offs):
        entry = self._chunk_index().get(offs)
        n_items = int(np.prod(self.chunks, dtype=np.int64))
        if entry is None:                                      # unallocated chunk: fill value (zero)
            return np.zeros(self.chunks, dtype=self._storage_dtype())
        addr, size, mask = entry
        raw = self._unfilter(self._f._read(addr, size), mask)
        return np.frombuffer(raw, dtype=self._storage_dtype(), count=n_items).reshape(self.chunks)

    def _storage_dtype(self):
        return np.dtype("V16") if isinstance(self._dtype, _VlenStr) else self._dtype

    def _read_all_raw(self):
        sd = self._storage_dtype()
        if self.size == 0:
            return np.zeros(self.shape, dtype=sd)
        if self._kind == "compact":
            return np.frombuffer(self._compact, dtype=sd, count=self.size).reshape(self.shape)
        if self._kind == "contiguous":
            if self._addr == _UNDEF:
                return np.zeros(self.shape, dtype=sd)
            return np.frombuffer(self._f._read(self._addr, self.size * sd.itemsize), dtype=sd).reshape(self.shape)
        return self._read_box_raw(tuple((0, n) for n in self.shape))

    def _read_box_raw(self, box, threads=None):
        """The hyper-rectangle ``box`` = ((lo, hi), ...) of a chunked dataset, chunks inflated by a thread pool
        (zlib releases the GIL).  threads: None = the process-wide pool, 1 = inline, n = a private pool."""
        out = np.zeros(tuple(hi - lo for lo, hi in box), dtype=self._storage_dtype())
        ranges = [range((lo // c) * c, hi, c) for (lo, hi), c in zip(box, self.chunks)]
        todo = [()]
        for r in ranges:
            todo = [t + (o,) for t in todo for o in r]

        def place(offs):
            chunk = self._read_chunk(offs)
            src, dst = [], []
            for o, c, (lo, hi), n in zip(offs, self.chunks, box, self.shape):
                a, b = max(o, lo), min(o + c, hi, n)
                src.append(slice(a - o, b - o))
                dst.append(slice(a - lo, b - lo))
            out[tuple(dst)] = chunk[tuple(src)]

        if len(todo) > 4 and threads is None:
            list(_pool().map(place, todo))
        elif len(todo) > 4 and threads > 1:
            with ThreadPoolExecutor(max_workers=threads) as pool:
                list(pool.map(place, todo))
        else:
            for offs in todo:
                place(offs)
        return out

    def _finish(self, raw):
        if isinstance(self._dtype, _VlenStr):
            flat = raw.reshape(-1)
            vals = _decode_values(self._f, self._dtype, (flat.shape[0],), flat.tobytes())
            return vals.reshape(raw.shape)
        return raw

    def __getitem__(self, key):
        """numpy-style basic indexing (integers, slices, Ellipsis) plus ONE list / array index (h5py's fancy
        indexing on a single axis, which the reference uses for ``kmer_matrix[:, columns]``, rules.py:160)."""
        if self.shape is None:
            raise ValueError("null dataspace")
        if not isinstance(key, tuple):
            key = (key,)
        if Ellipsis in key:
            i = key.index(Ellipsis)
            key = key[:i] + (slice(None),) * (len(self.shape) - len(key) + 1) + key[i + 1:]
        key = key + (slice(None),) * (len(self.shape) - len(key))
        if len(key) != len(self.shape):
            raise IndexError("too many indices")
        if len(self.shape) == 0:
            return self._finish(self._read_all_raw())[()]
        box, post = [], []
        for k, n in zip(key, self.shape):
            if isinstance(k, slice):
                lo, hi, step = k.indices(n)
                if step != 1:
                    raise NotImplementedError("hdf5_lite: strided slices")
                hi = max(hi, lo)
                box.append((lo, hi))
                post.append(slice(None))
            elif isinstance(k, (list, tuple, np.ndarray)):
                idx = np.asarray(k)
                if idx.dtype == bool:
                    idx = np.flatnonzero(idx)
                idx = np.where(idx < 0, idx + n, idx).astype(np.int64)
                if idx.size and (idx.min() < 0 or idx.max() >= n):
                    raise IndexError("index out of range")
                lo, hi = (int(idx.min()), int(idx.max()) + 1) if idx.size else (0, 0)
                box.append((lo, hi))
                post.append(idx - lo)
            else:
                i = int(k)
                if i < 0:
                    i += n
                if not 0 <= i < n:
                    raise IndexError("index out of range")
                box.append((i, i + 1))
                post.append(0)
        if self._kind == "chunked":
            raw = self._read_box_raw(tuple(box))
        else:
            raw = self._read_all_raw()[tuple(slice(lo, hi) for lo, hi in box)]
        whole = all(isinstance(q, slice) for q in post)
        out = self._finish(np.asarray(raw))
        out = out if whole else out[tuple(post)]
        if isinstance(out, np.ndarray):
            if self._kind == "chunked" and whole and out.flags.owndata:
                return out                                     # assembled from inflated chunks: already a private copy
            return np.array(out)                               # own the memory (the file may be closed later)
        if isinstance(out, np.bytes_):
            return bytes(out)
        return out

    def __array__(self, dtype=None, copy=None):
        a = self[...]
        return a.astype(dtype) if dtype is not None else a

    def iter_chunk_rows(self):
        """Row slabs aligned to the chunk grid: yields (row0, row1) -- the upload loop reads whole chunk rows."""
        step = self.chunks[0] if self.chunks else max(1, self.shape[0])
        for r0 in range(0, self.shape[0], step):
            yield r0, min(r0 + step, self.shape[0])
