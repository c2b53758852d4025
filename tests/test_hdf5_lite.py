"""hdf5_lite (the h5py stand-in of machines without h5py) and the dataset adapter -- CPU.

Pinned against a file written by the real HDF5 library: scipy ships ``testhdf5_7.4_GLNX86.mat`` (MATLAB v7.3 =
HDF5 behind a 512-byte user block: version-0 superblock with a base address, symbol-table root group, version-1
object header, contiguous float64 dataset, fixed-string attribute).  The chunk B-tree, deflate, global-heap and
multi-node group paths have no library-written file in this image: they are checked against this package's writer,
which follows the same format specification (stated as such in DESIGN.md), and against a file assembled by hand from
the specification in libhdf5 1.8's layout (tests/golden/make_hdf5_fixture.py), plus malformed variants of it."""
import os

import numpy as np
import pytest

from kover_b200.dataset import hdf5_lite
from kover_b200.dataset.ds import KoverDataset
from hdf5_lite_writer import Writer


def test_reads_a_file_written_by_the_hdf5_library():
    scipy_io = pytest.importorskip("scipy.io")
    path = os.path.join(os.path.dirname(scipy_io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.isfile(path):
        pytest.skip("scipy test data not installed")
    f = hdf5_lite.File(path)
    assert f.keys() == ["testdouble"]
    d = f["testdouble"]
    assert d.shape == (9, 1) and d.dtype == np.dtype("<f8") and d.chunks is None and d.compression is None
    assert d.attrs["MATLAB_class"] == b"double"
    # scipy's own expectation for this file (scipy/io/matlab/tests/test_mio.py): linspace(0, 2 pi, 9)
    assert np.allclose(d[...].reshape(-1), np.linspace(0, 2 * np.pi, 9), rtol=0, atol=1e-15)
    assert d[3, 0] == d[...][3, 0] and np.array_equal(d[2:5], d[...][2:5])
    f.close()


@pytest.mark.parametrize("dtype", [np.uint64, np.uint32])
@pytest.mark.parametrize("chunks,gzip", [((1, 1000), 4), ((2, 700), 1), ((1, 100000), 9), (None, None)])
def test_packed_matrix_round_trip(tmp_path, dtype, chunks, gzip):
    rs = np.random.RandomState(3)
    M = rs.randint(0, 2 ** 32, size=(7, 2503)).astype(dtype) * dtype(2 ** 16 + 1)
    path = str(tmp_path / "m.h5")
    with Writer(path) as w:
        w.create_dataset("kmer_matrix", M, chunks=chunks, compression=gzip)
    d = hdf5_lite.File(path)["kmer_matrix"]
    assert d.shape == M.shape and d.dtype == np.dtype(dtype)
    assert d.chunks == (None if chunks is None else chunks) and d.compression == (None if gzip is None else "gzip")
    assert np.array_equal(d[...], M) and np.array_equal(np.asarray(d), M)
    assert np.array_equal(d[1:6, 900:2101], M[1:6, 900:2101])                    # rules.py:253-style block reads
    assert np.array_equal(d[:, [3, 1700, 2502]], M[:, [3, 1700, 2502]])          # rules.py:160: dataset[:, columns]
    assert np.array_equal(d[4], M[4]) and d[6, 2502] == M[6, 2502] and d[-1, -1] == M[-1, -1]
    assert d[0:0].shape == (0, 2503)
    with pytest.raises(IndexError):
        d[7]


def test_groups_strings_attributes_round_trip(tmp_path):
    path = str(tmp_path / "g.h5")
    ids = np.array(["genome_%d" % i for i in range(300)], dtype=object)
    with Writer(path) as w:
        w.attrs("/", {"uuid": "abc-123", "created": 1.5, "n": np.int64(7), "raw": b"xyz", "vec": np.arange(3, dtype=np.uint16)})
        w.create_dataset("genome_identifiers", ids)
        w.create_dataset("kmer_sequences", np.array([b"ACGT", b"TTTT", b"GGCA"], dtype="S4"), chunks=(2,), compression=4)
        w.create_dataset("big", np.arange(100000, dtype=np.uint32), chunks=(100,), compression=4)   # 1 000 chunks: 2-level B-tree
        w.create_dataset("scalar", np.float64(2.5))
        w.create_dataset("empty", np.zeros(0, dtype=np.uint8))
        for i in range(12):                                                                       # > 8 links: two symbol nodes
            w.create_dataset("splits/s/folds/fold_%d/train_genome_idx" % (i + 1), np.arange(i + 3, dtype=np.uint16))
        w.attrs("splits/s", {"train_proportion": 0.5, "random_seed": np.int64(42)})
    f = hdf5_lite.File(path)
    assert f.attrs["uuid"] == "abc-123" and f.attrs["created"] == 1.5 and f.attrs["n"] == 7 and f.attrs["raw"] == b"xyz"
    assert np.array_equal(f.attrs["vec"], [0, 1, 2])
    g = f["genome_identifiers"]
    assert g.shape == (300,) and g[299] == "genome_299" and g[[5, 7]].tolist() == ["genome_5", "genome_7"]
    assert g[...].tolist() == ids.tolist()
    ks = f["kmer_sequences"]
    assert ks[0] == b"ACGT" and len(ks[0]) == 4 and ks[...].tolist() == [b"ACGT", b"TTTT", b"GGCA"] and len(ks) == 3
    assert np.array_equal(f["big"][...], np.arange(100000)) and np.array_equal(f["big"][12345:54321], np.arange(12345, 54321))
    assert f["scalar"][()] == 2.5 and f["empty"][...].shape == (0,)
    # h5py iterates a group in name order: fold_1, fold_10, fold_11, fold_12, fold_2, ...
    assert list(f["splits/s/folds"].keys()) == sorted("fold_%d" % (i + 1) for i in range(12))
    for name, fold in f["splits/s/folds"].items():
        assert np.array_equal(fold["train_genome_idx"][...], np.arange(int(name.split("_")[1]) + 2))
    assert "splits" in f and "nope" not in f and "folds" in f["splits/s"] and "splits/s/folds/fold_3" in f
    assert f["splits"]["s"].attrs["random_seed"] == 42 and f["splits/s"].attrs["train_proportion"] == 0.5
    with pytest.raises(KeyError):
        f["splits/other"]


def test_not_hdf5(tmp_path):
    p = tmp_path / "x.bin"
    p.write_bytes(b"definitely not hdf5" * 100)
    with pytest.raises(ValueError):
        hdf5_lite.File(str(p))


@pytest.mark.parametrize("fixed_strings", [False, True])
def test_kover_dataset_adapter(tmp_path, fixed_strings):
    from kover_file import write_kover_file
    from kover_b200.utils import _pack_binary_bytes_to_ints
    rs = np.random.RandomState(5)
    n, k = 150, 3000
    X = (rs.rand(n, k) < 0.3).astype(np.uint8)
    packed = _pack_binary_bytes_to_ints(X, 64)
    y = (rs.rand(n) < 0.5).astype(np.uint8)
    perm = rs.permutation(n)
    train, test = np.sort(perm[:100]), np.sort(perm[100:])
    fold_of = rs.randint(0, 3, size=len(train))
    folds = [(train[fold_of != f], train[fold_of == f]) for f in range(3)]
    path = str(tmp_path / "d.kover")
    write_kover_file(path, packed, n, y, train, test, folds, fixed_strings=fixed_strings)
    d = KoverDataset(path)
    assert d.genome_count == n and d.kmer_count == k and d.kmer_length == 12 and d.classification_type == "binary"
    assert d.compression == "gzip (level 4)" and d.uuid == "00000000-test" and d.genome_source_type == "tsv"
    assert d.kmer_matrix.shape == packed.shape and d.kmer_matrix.chunks == (1, 1000) and np.array_equal(d.kmer_matrix[...], packed)
    assert np.array_equal(d.phenotype.metadata[...], y) and d.phenotype.description == "synthetic phenotype"
    assert [str(t) for t in d.phenotype.tags[...].tolist()] == ["0", "1"]
    s = d.get_split("split1")
    assert [sp.name for sp in d.splits] == ["split1"] and len(s.folds) == 3 and s.random_seed == 42
    assert [f.name for f in s.folds] == ["fold_1", "fold_2", "fold_3"]
    assert np.array_equal(s.train_genome_idx[...], train) and np.array_equal(s.test_genome_idx[...], test)
    assert abs(s.train_proportion - 100.0 / 150) < 1e-12 and "Folds: 3" in str(s)
    assert s.unique_risk_by_kmer.shape == (k,) and s.unique_risk_by_anti_kmer.shape == (k,)
    assert np.array_equal(s.folds[1].train_genome_idx[...], np.sort(folds[1][0]))
    d.close()


def test_randomized_round_trips(tmp_path):
    """Random shapes / chunk shapes / element types / compression levels / slices: what the writer lays out, the
    reader returns (chunk grid edges, partial edge chunks, multi-level chunk B-trees, 1-D and 2-D)."""
    rs = np.random.RandomState(99)
    dtypes = [np.uint8, np.uint16, np.uint32, np.uint64, np.int64, np.float64, np.float32]
    for trial in range(25):
        ndim = 1 + trial % 2
        shape = tuple(int(rs.randint(1, 40 if ndim == 2 else 5000)) if d == 0 else int(rs.randint(1, 700)) for d in range(ndim))
        dt = dtypes[trial % len(dtypes)]
        data = (rs.rand(*shape) * 1000).astype(dt)
        chunks = tuple(int(rs.randint(1, max(2, s // 2 + 2))) for s in shape) if trial % 5 else None
        gz = [None, 1, 4, 9][trial % 4] if chunks is not None else None
        path = str(tmp_path / ("r%d.h5" % trial))
        with Writer(path) as w:
            w.create_dataset("g/sub/data", data, chunks=chunks, compression=gz, attrs={"trial": np.int64(trial)})
        d = hdf5_lite.File(path)["g/sub/data"]
        assert d.shape == shape and d.dtype == np.dtype(dt) and d.attrs["trial"] == trial
        assert np.array_equal(d[...], data)
        for _ in range(6):
            key = tuple(slice(*sorted(rs.randint(0, s + 1, size=2))) for s in shape)
            assert np.array_equal(d[key], data[key]), (trial, key)
        i = tuple(int(rs.randint(0, s)) for s in shape)
        assert d[i] == data[i]


# ------------------------------------------------------------------------------------------------ hand-assembled fixture
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _fixture_bytes():
    with open(os.path.join(GOLDEN, "hand_assembled_chunked.h5"), "rb") as fh:
        return fh.read()


def test_hand_assembled_chunked_deflate_file():
    """A chunked + deflate dataset assembled byte by byte from the format specification by
    tests/golden/make_hdf5_fixture.py -- independent of hdf5_lite_writer.py -- laid out the way libhdf5 1.8 / h5py 2.x
    wrote Kover files: cached root symbol-table entry, NIL / modification-time / old fill-value messages, an
    object-header continuation block, B-tree nodes at full capacity with garbage in unused slots, a TWO-level chunk
    B-tree, a named version-1 deflate filter with its padding word, an edge chunk, maximum dimensions."""
    import make_hdf5_fixture as mk
    f = hdf5_lite.File(os.path.join(GOLDEN, "hand_assembled_chunked.h5"))
    assert f.keys() == ["kmer_matrix", "phenotype"]
    d = f["kmer_matrix"]
    assert d.shape == (3, 250) and d.dtype == np.dtype("<u8") and d.chunks == (1, 100)
    assert d.compression == "gzip" and d.compression_opts == 4
    M = mk.matrix()
    assert np.array_equal(d[...], M)
    assert np.array_equal(d[1:3, 50:240], M[1:3, 50:240]) and np.array_equal(d[:, [3, 249, 100]], M[:, [3, 249, 100]])
    assert np.array_equal(d[2, 200:], M[2, 200:])                   # the edge chunk
    assert len(d._chunk_index()) == 9                               # both leaves of the two-level B-tree
    assert np.array_equal(f["phenotype"][...], mk.phenotype())
    f.close()


def test_hand_assembled_fixture_is_reproducible():
    import make_hdf5_fixture as mk
    assert mk.build() == _fixture_bytes()


def _patched(tmp_path, name, edit):
    data = bytearray(_fixture_bytes())
    data = edit(data)
    path = str(tmp_path / name)
    with open(path, "wb") as fh:
        fh.write(bytes(data))
    return path


def test_truncated_chunk_is_an_error(tmp_path):
    """A file cut in the middle of its chunk data: reading the dataset must fail loudly, not return zeros."""
    f0 = hdf5_lite.File(os.path.join(GOLDEN, "hand_assembled_chunked.h5"))
    last_chunk_end = max(a + n for a, n, _ in f0["kmer_matrix"]._chunk_index().values())
    f0.close()
    # keep the metadata (it lies after the chunks in this file) but blank out and shorten nothing else: emulate a
    # truncated download by zeroing the tail of the last chunk's deflate stream
    def corrupt(data):
        data[last_chunk_end - 40:last_chunk_end] = b"\x00" * 40
        return data
    d = hdf5_lite.File(_patched(tmp_path, "corrupt.h5", corrupt))["kmer_matrix"]
    with pytest.raises(Exception) as e:
        d[...]
    assert "zlib" in type(e.value).__module__ or "Error" in type(e.value).__name__
    # and a file that simply ends early
    with pytest.raises(ValueError, match="out of range|not an HDF5"):
        f = hdf5_lite.File(_patched(tmp_path, "short.h5", lambda data: data[:len(data) // 2]))
        f["kmer_matrix"][...]


def test_unknown_filter_is_not_implemented(tmp_path):
    def edit(data):
        i = data.index(b"deflate\x00")
        assert bytes(data[i - 8:i - 6]) == b"\x01\x00"            # filter id 1 = deflate
        data[i - 8:i - 6] = (32001).to_bytes(2, "little")         # a registered third-party filter (Blosc)
        return data
    d = hdf5_lite.File(_patched(tmp_path, "filter.h5", edit))["kmer_matrix"]
    assert d.compression is None
    with pytest.raises(NotImplementedError, match="filter 32001"):
        d[...]


def test_version_2_object_header_is_not_implemented(tmp_path):
    """libver='latest' files start object headers with "OHDR": documented as unsupported, raised as such."""
    f0 = hdf5_lite.File(os.path.join(GOLDEN, "hand_assembled_chunked.h5"))
    addr = f0._root._load_links()["kmer_matrix"]
    f0.close()

    def edit(data):
        data[addr:addr + 4] = b"OHDR"
        return data
    f = hdf5_lite.File(_patched(tmp_path, "v2.h5", edit))
    with pytest.raises(NotImplementedError, match="version-2 object headers"):
        f["kmer_matrix"]
    assert np.array_equal(f["phenotype"][...], np.array([0, 0, 1, 1, 0, 1], dtype=np.uint8))


def test_raw_chunk_view_of_an_h5py_like_dataset(tmp_path):
    """rules._raw_chunk_view: an h5py-style dataset (file name + object name, no raw-chunk access) is opened a second
    time with the built-in reader so that its gzip chunks can go to the device as they are; anything that does not
    look like the same dataset, or that the reader cannot parse, falls back to reading through the caller's object."""
    from h5py_like import H5pyLikeDataset
    from kover_b200.learning.common.rules import _raw_chunk_view
    rs = np.random.RandomState(8)
    M = rs.randint(0, 2 ** 62, size=(5, 2300)).astype(np.uint64)
    path = str(tmp_path / "k.h5")
    with Writer(path) as w:
        w.create_dataset("kmer_matrix", M, chunks=(1, 1000), compression=4)
        w.create_dataset("plain", M)                                   # contiguous: no chunks to hand out
    like = H5pyLikeDataset(path, "kmer_matrix")
    twin = _raw_chunk_view(like)
    assert twin is not None and twin is not like and tuple(twin.shape) == M.shape
    tab = twin.deflate_chunk_table(0, 5, 0, 2300)
    assert tab is not None and tab["address"].shape[0] == 5 * 3 and np.array_equal(twin[...], M)
    direct = hdf5_lite.File(path)["kmer_matrix"]
    assert _raw_chunk_view(direct) is direct                           # the built-in reader's own datasets: as they are
    assert _raw_chunk_view(M) is None                                  # numpy arrays / memmaps: nothing to open
    assert _raw_chunk_view(H5pyLikeDataset(path, "kmer_matrix", shape=(5, 2299))) is None      # not the same object
    assert _raw_chunk_view(H5pyLikeDataset(path, "plain")) is None
    gone = H5pyLikeDataset(path, "kmer_matrix")
    gone.file.filename = str(tmp_path / "moved_away.h5")
    assert _raw_chunk_view(gone) is None
    with open(str(tmp_path / "not.h5"), "wb") as f:
        f.write(b"definitely not HDF5" * 100)
    gone.file.filename = str(tmp_path / "not.h5")
    assert _raw_chunk_view(gone) is None
