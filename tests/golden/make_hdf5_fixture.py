"""Assembles tests/golden/hand_assembled_chunked.h5 byte by byte from the HDF5 File Format Specification (v1.1 / the
"earliest" format libhdf5 1.8 writes, which is what h5py 2.x produced for Kover datasets): no code is shared with
tests/hdf5_lite_writer.py or with the reader.  It deliberately contains what the library writes and our own writer
does not:

  * a root symbol-table entry with cache type 1 (B-tree / heap addresses cached in the scratch pad);
  * object headers with NIL padding messages, a modification-time message (0x12), an old-style fill value message
    (0x04) next to the new one (0x05) and an object-header continuation block (0x10);
  * B-tree nodes allocated at their full 2K+1 key capacity with garbage in the unused slots;
  * a chunk B-tree with TWO levels (an internal node over two leaves), keys carrying the element-size dimension;
  * a version-1 filter pipeline message with the filter NAME ("deflate") and the odd-client-data padding;
  * an edge chunk (250 columns in chunks of 100: the last chunk is stored full size);
  * a dataspace with maximum dimensions.

Layout of the dataset: `kmer_matrix` uint64 [3, 250], chunks (1, 100), gzip 4 (Kover: create.py:224-230,
kl_64.cpp:229-263); plus `phenotype` uint8 [6], contiguous.

    python tests/golden/make_hdf5_fixture.py      (rewrites the committed file; zlib's output is deterministic)
"""
import os
import struct
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
UNDEF = 0xFFFFFFFFFFFFFFFF
OUT = os.path.join(HERE, "hand_assembled_chunked.h5")


def matrix():
    rs = np.random.RandomState(20240229)
    m = rs.randint(0, 2 ** 62, size=(3, 250), dtype=np.int64).astype(np.uint64)
    m[1, 100:200] = 0                      # an all-zero chunk (deflates to a few bytes)
    m[2, :] &= np.uint64(0xFF00FF00FF00FF00)
    return m


def phenotype():
    return np.array([0, 0, 1, 1, 0, 1], dtype=np.uint8)


def pad8(b):
    return b + b"\x00" * (-len(b) % 8)


def message(mtype, body, flags=0):
    body = pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def object_header(messages, total_size=None):
    """Version-1 object header: 12 bytes + 4 alignment + messages (III.C of the specification)."""
    blob = b"".join(messages)
    if total_size is not None:                       # NIL message filling the allocation, like the library leaves
        spare = total_size - len(blob)
        assert spare >= 8
        messages = messages + [message(0x0000, b"\x00" * (spare - 8))]
        blob = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(blob)) + blob


class Image(object):
    def __init__(self):
        self.buf = bytearray()

    def here(self):
        return len(self.buf)

    def put(self, b, align=8):
        self.buf += b"\x00" * (-len(self.buf) % align)
        addr = len(self.buf)
        self.buf += b
        return addr

    def patch(self, addr, b):
        self.buf[addr:addr + len(b)] = b


def build():
    img = Image()
    m, ph = matrix(), phenotype()

    # ---- superblock, version 0 (II.A); the root symbol-table entry is patched in at the end
    img.put(b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
            + struct.pack("<QQQQ", 0, UNDEF, 0, UNDEF) + b"\x00" * 40)
    assert img.here() == 96

    # ---- chunks first (the library allocates raw data wherever the free-space manager puts it)
    chunks = {}
    for r in range(3):
        for c0 in (0, 100, 200):
            full = np.zeros(100, dtype="<u8")
            seg = m[r, c0:c0 + 100]
            full[:seg.shape[0]] = seg
            z = zlib.compress(full.tobytes(), 4)
            chunks[(r, c0)] = (img.put(z, align=1), len(z))
        img.put(b"\xde\xad\xbe\xef" * 3, align=1)      # unrelated bytes between allocations

    # ---- chunk B-tree, two levels (III.A.1, node type 1).  Key: chunk size, filter mask, rank+1 element offsets.
    K_CHUNK = 32
    key_size = 8 + 8 * 3
    node_size = 24 + (2 * K_CHUNK + 1) * key_size + 2 * K_CHUNK * 8

    def chunk_key(size, offs):
        return struct.pack("<II3Q", size, 0, offs[0], offs[1], 0)

    def node(level, entries, last_key, left, right):
        body = b""
        for key, child in entries:
            body += key + struct.pack("<Q", child)
        body += last_key
        head = b"TREE" + struct.pack("<BBHQQ", 1, level, len(entries), left, right)
        garbage = (b"\xa5" * node_size)[:node_size - len(head) - len(body)]      # unused key / child slots
        return head + body + garbage

    order = sorted(chunks.keys())
    leaves = [order[:5], order[5:]]
    leaf_addr = [img.here() + (-img.here() % 8), 0]
    leaf_addr[1] = leaf_addr[0] + node_size + 40
    for i, members in enumerate(leaves):
        entries = [(chunk_key(chunks[o][1], o), chunks[o][0]) for o in members]
        nxt = leaves[i + 1][0] if i + 1 < len(leaves) else (3, 0)                 # upper bound: one past the dataset
        last_key = chunk_key(0, nxt)
        a = img.put(node(0, entries, last_key, leaf_addr[0] if i == 1 else UNDEF, leaf_addr[1] if i == 0 else UNDEF))
        assert a == leaf_addr[i]
        img.put(b"\x00" * 40)
    root_entries = [(chunk_key(chunks[members[0]][1], members[0]), leaf_addr[i]) for i, members in enumerate(leaves)]
    btree_root = img.put(node(1, root_entries, chunk_key(0, (3, 0)), UNDEF, UNDEF))

    # ---- dataset object header for kmer_matrix, with a continuation block
    dataspace = struct.pack("<BBB5x", 1, 2, 1) + struct.pack("<QQ", 3, 250) + struct.pack("<QQ", 3, 250)
    datatype = struct.pack("<BBBBI", 0x10, 0x00, 0x00, 0x00, 8) + struct.pack("<HH", 0, 64)
    fill_old = struct.pack("<I", 0)                                                # 0x04: size 0 = undefined
    fill_new = struct.pack("<BBBB", 2, 3, 0, 0)                                    # 0x05 v2: incremental, never, undefined
    layout = struct.pack("<BBB", 3, 2, 3) + struct.pack("<Q", btree_root) + struct.pack("<III", 1, 100, 8)
    name = b"deflate\x00"
    pipeline = struct.pack("<BB6x", 1, 1) + struct.pack("<HHHH", 1, len(name), 1, 1) + name + struct.pack("<I", 4) \
        + b"\x00" * 4
    mtime = struct.pack("<B3xI", 1, 1456704000)
    cont_block_msgs = [message(0x000B, pipeline, flags=1), message(0x0012, mtime), message(0x0000, b"\x00" * 16)]
    cont_blob = b"".join(cont_block_msgs)
    cont_addr = img.put(cont_blob)
    first_msgs = [message(0x0001, dataspace), message(0x0003, datatype, flags=1), message(0x0004, fill_old),
                  message(0x0005, fill_new, flags=1), message(0x0010, struct.pack("<QQ", cont_addr, len(cont_blob))),
                  message(0x0008, layout)]
    blob = b"".join(first_msgs)
    hdr = struct.pack("<BBHII4x", 1, 0, len(first_msgs) + len(cont_block_msgs), 1, len(blob)) + blob
    kmer_matrix_addr = img.put(hdr)

    # ---- phenotype: contiguous uint8 [6]
    ph_data = img.put(ph.tobytes())
    ph_hdr = object_header([
        message(0x0001, struct.pack("<BBB5x", 1, 1, 0) + struct.pack("<Q", 6)),
        message(0x0003, struct.pack("<BBBBI", 0x10, 0x00, 0x00, 0x00, 1) + struct.pack("<HH", 0, 8), flags=1),
        message(0x0005, struct.pack("<BBBB", 2, 2, 2, 0), flags=1),
        message(0x0008, struct.pack("<BB", 3, 1) + struct.pack("<QQ", ph_data, 6)),
        message(0x0012, mtime),
    ], total_size=272)
    phenotype_addr = img.put(ph_hdr)

    # ---- root group: local heap (III.D), symbol table node (III.B), group B-tree (node type 0), object header
    names = [b"kmer_matrix", b"phenotype"]
    heap_data = bytearray(pad8(b"\x00"))
    name_off = []
    for nme in names:
        name_off.append(len(heap_data))
        heap_data += pad8(nme + b"\x00")
    free_off = len(heap_data)
    heap_size = 88
    heap_data += struct.pack("<QQ", 1, heap_size - free_off)                       # free block: next = 1 (none), size
    heap_data += b"\x00" * (heap_size - len(heap_data))
    heap_data_addr = img.put(bytes(heap_data))
    heap_addr = img.put(b"HEAP" + struct.pack("<B3xQQQ", 0, heap_size, free_off, heap_data_addr))

    def symbol_entry(name_offset, header_addr):
        return struct.pack("<QQII16x", name_offset, header_addr, 0, 0)

    snod = b"SNOD" + struct.pack("<BBH", 1, 0, 2) + symbol_entry(name_off[0], kmer_matrix_addr) \
        + symbol_entry(name_off[1], phenotype_addr) + b"\x5a" * (40 * 6)             # 2 of 2*4 entries used
    snod_addr = img.put(snod)
    K_GROUP = 16
    gnode = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_addr, name_off[1])
    gnode += b"\xc3" * (24 + (2 * K_GROUP + 1) * 8 + 2 * K_GROUP * 8 - len(gnode))
    group_btree = img.put(gnode)
    root_hdr = img.put(object_header([message(0x0011, struct.pack("<QQ", group_btree, heap_addr))], total_size=40))

    eof = img.put(b"", align=8)
    img.patch(40, struct.pack("<Q", eof))
    img.patch(56, struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", group_btree, heap_addr))
    return bytes(img.buf)


if __name__ == "__main__":
    data = build()
    with open(OUT, "wb") as fh:
        fh.write(data)
    print("wrote %s (%d bytes)" % (OUT, len(data)))
