"""Rule objects and the device-resident ``KmerRuleClassifications``.

Mirrors reference kover/learning/common/rules.py: same class names, constructor signature,
methods (``get_columns``, ``remove_rows``, ``shape``, ``sum_rows``), return dtypes and exceptions
(rules.py:27-267).  What changes is where the matrix lives: the constructor streams the packed
``kmer_matrix`` into HBM once (column-sharded over the visible GPUs) and every later call is a
kernel over that resident image instead of a loop of h5py chunk reads + Cython popcount
(rules.py:247-262).

Additions the learners of this package use (not in the reference): ``best_utility_rules`` (the
fused SCM scan, scm.py:238-288) and ``gini_best_split`` (cart.py:112-161, 231-238).
"""
import itertools
import os

import numpy as np

from ... import _native
from ...sharding import Comm, ShardGroup, partition_columns
from ...utils import _minimum_uint_size, _unpack_binary_bytes_from_ints, build_row_mask


class KmerRule(object):
    """Presence ("presence") or absence ("absence") of one k-mer, as a boolean rule on unpacked genomes
    (interface of rules.py:27-54)."""

    def __init__(self, kmer_index, kmer_sequence, type):
        self.kmer_index, self.kmer_sequence, self.type = kmer_index, kmer_sequence, type

    def classify(self, X):
        wanted = 0 if self.type == "absence" else 1
        return (X[:, self.kmer_index] == wanted).astype(np.uint8)

    def inverse(self):
        flipped = "presence" if self.type == "absence" else "absence"
        return KmerRule(kmer_index=self.kmer_index, kmer_sequence=self.kmer_sequence, type=flipped)

    def __str__(self):
        return "%s(%s)" % ("Absence" if self.type == "absence" else "Presence", self.kmer_sequence)


class LazyKmerRuleList(object):
    """The 2K rules of a dataset, materialised on demand: rule i < K is the presence of the k-mer of matrix
    column i, rule K + i its absence (interface and index convention of rules.py:57-79)."""

    def __init__(self, kmer_sequences, kmer_by_rule):
        self.kmer_sequences = kmer_sequences
        self.kmer_by_rule = kmer_by_rule
        self.n_rules = 2 * kmer_by_rule.shape[0]

    def __len__(self):
        return self.n_rules

    def __getitem__(self, idx):
        if idx >= self.n_rules:
            raise ValueError("Index %d is out of range for list of size %d" % (idx, self.n_rules))
        n_kmers = len(self.kmer_sequences)
        column = idx % n_kmers
        kind = "absence" if idx >= n_kmers else "presence"
        sequence = self.kmer_sequences[self.kmer_by_rule[column]]
        if isinstance(sequence, bytes):          # fixed-length HDF5 strings come back as bytes
            sequence = sequence.decode("ascii")
        return KmerRule(column, sequence, kind)


class BaseRuleClassifications(object):
    """The operator interface the learners consume (rules.py:81-96)."""

    def get_columns(self, columns):
        raise NotImplementedError()

    def remove_rows(self, rows):
        raise NotImplementedError()

    @property
    def shape(self):
        raise NotImplementedError()

    def sum_rows(self, rows):
        raise NotImplementedError()


def _flat_int64(a):
    """a as a contiguous 1-D int64 array -- itself when it already is one (the learners' index arrays)."""
    if a.dtype == np.int64 and a.ndim == 1 and a.flags.c_contiguous:
        return a
    return np.ascontiguousarray(a, dtype=np.int64).reshape(-1)


def _raw_chunk_view(dataset):
    """The dataset as something that can hand out the RAW gzip streams of its chunks (``deflate_chunk_table``), or None.
    hdf5_lite datasets do so themselves.  An h5py dataset -- what a Kover installation passes in (utils.py:58-73,
    experiment_scm.py:830) -- inflates every chunk on one host core inside ``dataset[...]`` (rules.py:253); the same file
    is therefore opened a second time, read-only, with the built-in reader, which only parses the chunk index (the
    reference writes its files with h5py's default "earliest" format: version-0 superblock, version-1 object headers,
    version-1 chunk B-tree).  Anything the reader does not understand -- or a dataset that does not look like the same
    object -- falls back to reading through the caller's dataset."""
    if getattr(dataset, "deflate_chunk_table", None) is not None:
        return dataset
    try:
        path, name = dataset.file.filename, dataset.name
    except AttributeError:
        return None
    if not isinstance(path, str) or not isinstance(name, str) or getattr(dataset, "chunks", None) is None:
        return None
    try:
        from ...dataset import hdf5_lite
        twin = hdf5_lite.File(path, "r")[name]
        if getattr(twin, "deflate_chunk_table", None) is None:
            return None
        same = (tuple(twin.shape) == tuple(dataset.shape) and np.dtype(twin.dtype) == np.dtype(dataset.dtype)
                and tuple(twin.chunks or ()) == tuple(dataset.chunks))
        return twin if same else None
    except Exception:                       # (NotImplementedError for newer file formats, OSError, KeyError, ...)
        return None


_level_keys = itertools.count(1)


class NodeExamples(dict):
    """class -> example indices of a tree node, tagged for the level search: ``level_key`` names the node,
    ``parent_key`` the node it was split from (-1: the root).  The children of a split partition their parent's
    examples (cart.py:300-312), which is what allows parent - sibling (kv_gini_level_family)."""

    def __init__(self, members, parent=None):
        dict.__init__(self, members)
        self.level_key = next(_level_keys)
        self.parent_key = getattr(parent, "level_key", None)
        if self.parent_key is None:
            self.parent_key = -1


class DeviceOccurrences(object):
    """k-mer occurrence counts of a set of genomes (the input of the reference's CART tiebreaker,
    experiment_cart.py:82-94, 264), resident on the device when the shards support it."""

    def __init__(self, rule_classifications, rows):
        self._rc = rule_classifications
        self.rows = np.asarray(rows)
        self._host = None
        self.table = None
        group = rule_classifications._group
        if group.occurrence_table_supported() and rule_classifications._free_tables:
            self.table = rule_classifications._free_tables.pop()
            group.occurrence_table_set(self.table, rule_classifications.row_mask(self.rows))

    def host(self):
        if self._host is None:
            self._host = self._rc.sum_rows(self.rows)
        return self._host

    def release(self):
        if self.table is not None and self._rc._group.shards:
            self._rc._group.occurrence_table_set(self.table, None)
            self._rc._free_tables.append(self.table)
        self.table = None


class ScmGreedyState(object):
    """The remaining positive / negative examples of one SCM fit, kept as row masks ON THE DEVICE (scm.py:133-139
    keeps two index lists on the host and filters them with a fetched column after every pick).  ``scan`` runs the
    utility scan over the resident masks, ``apply`` ANDs them with the picked rule's column (kv_mask_and_column) and
    returns how many examples are left.  The index lists are kept as well, lazily: they are only rebuilt (from one
    get_columns call per applied rule) if a scan has to fall back to the index-array path."""

    def __init__(self, rule_classifications, positive_example_idx, negative_example_idx):
        self._rc = rule_classifications
        self._pos = np.asarray(positive_example_idx)
        self._neg = np.asarray(negative_example_idx)
        self._pending = []                                  # rules applied on the device but not yet to the index lists
        self.n_pos, self.n_neg = rule_classifications._group.scm_state_set(
            rule_classifications.row_mask(self._pos), rule_classifications.row_mask(self._neg))

    def index_lists(self):
        for rule in self._pending:
            accepted = self._rc.get_columns(rule)
            self._pos = self._pos[accepted[self._pos] != 0]
            self._neg = self._neg[accepted[self._neg] != 0]
        self._pending = []
        return self._pos, self._neg

    def scan(self, p, rule_blacklist, util_block_size):
        self._rc.set_rule_blacklist(rule_blacklist)
        out = self._rc._group.scm_best_rules_state(float(p), int(util_block_size))
        if out is None:                                     # huge equivalent-rule set: the general path sizes its buffers
            pos, neg = self.index_lists()
            out = self._rc.best_utility_rules(p, pos, neg, rule_blacklist=rule_blacklist, util_block_size=util_block_size)
        return out

    def apply(self, rule_idx):
        n_kmers = self._rc._n_kmers
        rule_idx = int(rule_idx)
        self.n_pos, self.n_neg = self._rc._group.scm_state_apply(rule_idx % n_kmers, invert=rule_idx >= n_kmers)
        self._pending.append(rule_idx)
        return self.n_pos, self.n_neg


def _default_devices():
    env = os.environ.get("KOVER_B200_DEVICES")
    if env:
        return [int(x) for x in env.split(",") if x.strip() != ""]
    return [int(os.environ.get("LOCAL_RANK", "0"))]


class KmerRuleClassifications(BaseRuleClassifications):
    """Methods involving columns account for presence and absence rules.

    dataset: the packed ``kmer_matrix`` -- anything with ``shape``, ``dtype``, numpy-style
        ``__getitem__`` and optionally ``chunks`` (an h5py dataset, a numpy array or memmap).
    n_rows: the number of genomes.
    block_size: kept for signature compatibility (rules.py:104); it sizes the upload slabs only.

    Keyword-only extras: ``devices`` (CUDA ordinals; default $KOVER_B200_DEVICES, else $LOCAL_RANK,
    else 0), ``comm`` (a sharding.Comm: this process holds rank ``comm.rank``'s column slice),
    ``shard_group`` (adopt an existing, already filled group).
    """

    def __init__(self, dataset, n_rows, block_size=None, devices=None, comm=None, shard_group=None):
        self.dataset = dataset
        self.dataset_initial_n_rows = self.dataset_n_rows = n_rows
        self.dataset_removed_rows = []
        self.dataset_removed_rows_mask = np.zeros(n_rows, dtype=bool)

        # block_size: accepted and validated like the reference (rules.py:112-120); it only sizes upload slabs here
        if block_size is not None:
            if len(block_size) != 2 or not all(isinstance(b, int) for b in block_size):
                raise ValueError("The block size must be a tuple of 2 integers.")
            self.block_size = block_size
        else:
            chunks = getattr(dataset, "chunks", None)
            self.block_size = chunks if chunks is not None else (1, dataset.shape[1])

        pack_size = {np.dtype(np.uint32): 32, np.dtype(np.uint64): 64}.get(np.dtype(dataset.dtype))
        if pack_size is None:
            raise ValueError("Unsupported data type for packed attribute classifications array. The supported data" +
                             " types are np.uint32 and np.uint64.")
        self.dataset_pack_size = pack_size

        self._n_kmers = int(dataset.shape[1])
        self._blacklist_key = None
        self._free_tables = list(range(63, -1, -1))         # device occurrence tables (DeviceOccurrences)
        if shard_group is not None:
            self._group = shard_group
        else:
            self._group = self._build_group(devices, comm)
            self._upload()

    # ------------------------------------------------------------------ device image
    def _build_group(self, devices, comm):
        n, k = int(self.dataset_initial_n_rows), self._n_kmers
        if n > self.dataset.shape[0] * self.dataset_pack_size:
            raise ValueError("n_rows exceeds the number of examples the packed matrix can hold.")
        devices = list(devices) if devices is not None else _default_devices()
        if comm is not None:
            slices = partition_columns(k, comm.world_size)
            col0, n_cols = slices[comm.rank]
            if n_cols == 0:
                raise ValueError("more ranks than 512-column tiles")
            return ShardGroup([_native.Shard(devices[0], n, k, col0, n_cols)], n, k, comm=comm, rank_slices=slices)
        slices = [s for s in partition_columns(k, len(devices)) if s[1] > 0]
        return ShardGroup([_native.Shard(d, n, k, c0, nc) for d, (c0, nc) in zip(devices, slices)], n, k)

    def _upload_device_inflate(self):
        """HDF5 -> HBM without inflating on the host: the raw gzip streams of the chunks (rules.py:253 inflates them
        through libhdf5 on every read) go to the GPU, one warp inflates one chunk (kv_upload_deflate).  Used when the
        dataset can hand out its raw chunks (hdf5_lite: chunked, deflate only) and the shards live on a device;
        KOVER_B200_DEVICE_INFLATE=0 switches it off.  A chunk the device rejects (malformed stream) or that was stored
        unfiltered is read through the host path, which raises the proper error if it really is corrupt.
        -> False when this path does not apply."""
        if os.environ.get("KOVER_B200_DEVICE_INFLATE", "1") == "0":
            return False
        ds = _raw_chunk_view(self.dataset)
        if ds is None:
            return False
        table_of = ds.deflate_chunk_table
        if not all(hasattr(s, "upload_deflate") for s in self._group.shards):
            return False
        needed_rows = min(-(-int(self.dataset_initial_n_rows) // self.dataset_pack_size), int(ds.shape[0]))
        stats = os.environ.get("KOVER_B200_INFLATE_STATS") is not None
        import time
        for s in self._group.shards:
            t0 = time.perf_counter()
            tab = table_of(0, needed_rows, s.col0, s.col0 + s.n_cols)
            if stats:
                import sys
                sys.stderr.write("deflate_chunk_table: %.1f ms\n" % ((time.perf_counter() - t0) * 1e3))
            if tab is None:
                return False
            cr, cc = tab["chunks"]
            on_device = np.flatnonzero(~tab["raw"])
            status = np.zeros(tab["raw"].shape[0], dtype=np.int32)
            if on_device.size:
                status[on_device] = s.upload_deflate(self.dataset_pack_size, tab["address"][on_device],
                                                     tab["nbytes"][on_device], tab["row0"][on_device],
                                                     tab["col0"][on_device], cr, cc)
            for i in np.flatnonzero(tab["raw"] | (status != 0)):
                r0, c0 = int(tab["row0"][i]), int(tab["col0"][i])
                chunk = ds._read_chunk((r0, c0))                       # host inflate: raises on a corrupt stream
                s.upload_block(np.ascontiguousarray(chunk[:needed_rows - r0, :int(ds.shape[1]) - c0]), r0, c0)
        return True

    def _upload(self, slab_bytes=128 << 20):
        """Stream the packed matrix into HBM: row slabs x the column span of the local shards.  The next slab is
        read (HDF5: chunk reads + inflate on the host cores) while the current one is staged and copied."""
        from concurrent.futures import ThreadPoolExecutor
        if self._upload_device_inflate():
            return
        ds = self.dataset
        itemsize = np.dtype(ds.dtype).itemsize
        n_packed_rows = int(ds.shape[0])
        needed_rows = -(-int(self.dataset_initial_n_rows) // self.dataset_pack_size)
        chunk_rows = (getattr(ds, "chunks", None) or (1,))[0] or 1
        slabs = []
        for s in self._group.shards:
            c0, c1 = s.col0, s.col0 + s.n_cols
            rows_per_slab = max(1, min(needed_rows, slab_bytes // max(1, s.n_cols * itemsize)))
            rows_per_slab = max(chunk_rows, rows_per_slab // chunk_rows * chunk_rows)     # whole chunk rows per read
            for r0 in range(0, min(needed_rows, n_packed_rows), rows_per_slab):
                slabs.append((s, r0, min(r0 + rows_per_slab, needed_rows, n_packed_rows), c0, c1))
        if not slabs:
            return
        read = lambda slab: np.asarray(ds[slab[1]:slab[2], slab[3]:slab[4]])
        with ThreadPoolExecutor(max_workers=1) as prefetch:
            pending = prefetch.submit(read, slabs[0])
            for i, (s, r0, r1, c0, c1) in enumerate(slabs):
                block = pending.result()
                if i + 1 < len(slabs):
                    pending = prefetch.submit(read, slabs[i + 1])
                s.upload_block(block, r0, c0)

    def close(self):
        self._group.close()

    # ------------------------------------------------------------------ reference interface
    def get_columns(self, columns):
        """0/1 classifications of the given rules on every (non-removed) genome: uint8 [n_genomes] for an integer
        (anything with ``__index__``), [n_genomes, len(columns)] for a list / tuple / array; rule indices >= K are
        absence rules, i.e. the inverted column (rules.py:135-171)."""
        n_kmers = self._n_kmers
        if hasattr(columns, "__index__"):
            c = columns.__index__()
            if 0 <= c < 2 * n_kmers and not self.dataset_removed_rows:
                # the learners' per-iteration call (scm.py:133, cart.py:245): one rule, nothing removed
                packed = self._group.get_packed_columns(np.array([c % n_kmers], dtype=np.int64))[0]
                bits = np.unpackbits(packed.astype(">u8").view(np.uint8))[:self.dataset_initial_n_rows]
                return (1 - bits) if c >= n_kmers else bits
            single, wanted = True, [c]
        else:
            single = False
            wanted = columns.tolist() if isinstance(columns, np.ndarray) else list(columns)

        is_absence = np.array([c >= n_kmers for c in wanted], dtype=bool)
        kmer_cols = np.array([c if c < n_kmers else c % n_kmers for c in wanted], dtype=np.int64).reshape(-1)
        if kmer_cols.size and (kmer_cols.min() < -n_kmers or kmer_cols.max() >= n_kmers):
            raise IndexError("column index out of range")
        kmer_cols = np.where(kmer_cols < 0, kmer_cols + n_kmers, kmer_cols)
        distinct, position = np.unique(kmer_cols, return_inverse=True)

        packed = self._group.get_packed_columns(distinct)                            # [n_distinct, W64]
        unpacked = _unpack_binary_bytes_from_ints(np.ascontiguousarray(packed.T))    # [W64 * 64, n_distinct]
        alive = np.zeros(unpacked.shape[0], dtype=bool)
        alive[:self.dataset_initial_n_rows] = True                                   # drop the padding rows ...
        alive[self.dataset_removed_rows] = False                                     # ... and the removed genomes
        result = unpacked[alive][:, position]
        result[:, is_absence] = 1 - result[:, is_absence]
        return result.reshape(-1) if single else result

    def _dataset_relative_rows(self, rows):
        """The (r+1)-th row that has not been removed, for each r (rules.py:177-184, 227-236)."""
        rows = np.asarray(rows, dtype=np.int64).reshape(-1)
        if len(self.dataset_removed_rows) == 0:
            if rows.size and (rows.min() < 0 or rows.max() >= self.dataset_initial_n_rows):
                raise IndexError("index out of bounds")
            return rows
        alive = np.flatnonzero(~self.dataset_removed_rows_mask)
        return alive[rows]

    def remove_rows(self, rows):
        """Hide genomes from every later call; ``rows`` index the genomes still visible (rules.py:173-194)."""
        newly_removed = self._dataset_relative_rows(rows).tolist()
        if newly_removed:
            self.dataset_removed_rows = sorted(set(self.dataset_removed_rows).union(newly_removed))
            hidden = np.zeros(self.dataset_initial_n_rows, dtype=bool)
            hidden[self.dataset_removed_rows] = True
            self.dataset_removed_rows_mask = hidden
            self.dataset_n_rows = self.dataset_initial_n_rows - len(self.dataset_removed_rows)

    @property
    def shape(self):
        return self.dataset_n_rows, self._n_kmers * 2

    def row_mask(self, rows):
        """64-bit row mask over the packed rows for example indices ``rows`` (rules.py:210-239)."""
        # (the reference sorts `rows` first, rules.py:225; the mask does not depend on the order)
        return build_row_mask(self._dataset_relative_rows(rows), self.dataset_initial_n_rows, 64)

    def sum_rows(self, rows):
        """
        Note: Assumes that the rows argument does not contain duplicate elements. Rows will not be considered more than once.
        """
        rows = np.asarray(rows)
        result_dtype = _minimum_uint_size(rows.shape[0])
        n_kmers = self._n_kmers
        result = _native.result_empty(n_kmers * 2, result_dtype)      # (page-locked when large: D2H at the PCIe rate)
        mask = self.row_mask(rows)
        # presence counts and, for the absence rules, len(rows) - presence (rules.py:265): both halves are
        # produced on the device in the result dtype
        self._group.sum_rows_into(mask, len(rows), result)
        return result

    # ------------------------------------------------------------------ fused scans (this package's learners)
    def set_rule_blacklist(self, rule_blacklist):
        if self._blacklist_key == b"" and isinstance(rule_blacklist, (list, tuple, np.ndarray)) and len(rule_blacklist) == 0:
            return                                   # (the per-iteration call with nothing blacklisted, still nothing)
        bl = np.asarray(rule_blacklist, dtype=np.int64).reshape(-1)
        key = bl.tobytes()
        if key != self._blacklist_key:
            self._group.set_blacklist(bl)
            self._blacklist_key = key

    def best_utility_rules(self, p, positive_example_idx, negative_example_idx, rule_blacklist=(),
                           util_block_size=1000000):
        """scm.py:238-288 in one pass over the matrix.  -> (best_utility, rule idx ascending,
        pos_error_counts, neg_cover_counts) for the rules whose utility ties the maximum."""
        positive_example_idx = np.asarray(positive_example_idx)
        negative_example_idx = np.asarray(negative_example_idx)
        self.set_rule_blacklist(rule_blacklist)
        if not self.dataset_removed_rows:
            # index arrays straight into the library (masks built there): one kernel launch per call and rank
            out = self._group.scm_best_rules_idx(_flat_int64(positive_example_idx), _flat_int64(negative_example_idx),
                                                 float(p), int(util_block_size))
            if out is not None:
                return out
        pos_mask = self.row_mask(positive_example_idx)
        neg_mask = self.row_mask(negative_example_idx)
        return self._group.scm_best_rules(pos_mask, neg_mask, positive_example_idx.shape[0],
                                          negative_example_idx.shape[0], float(p), int(util_block_size))

    def scm_state(self, positive_example_idx, negative_example_idx):
        """A device-resident greedy state for one SCM fit (ScmGreedyState), or None when the shards cannot keep one
        (several shards in one process, checker-backed shards, removed rows)."""
        if self.dataset_removed_rows or not self._group.scm_state_supported():
            return None
        return ScmGreedyState(self, positive_example_idx, negative_example_idx)

    def best_utility_rules_grid(self, jobs, rule_blacklist=(), util_block_size=1000000):
        """``best_utility_rules`` for several (p, positive_example_idx, negative_example_idx) jobs at once: jobs
        whose example sets coincide (every p of a grid; conjunction and disjunction, whose labels are swapped)
        share ONE pass over the matrix.  Returns the results in job order."""
        self.set_rule_blacklist(rule_blacklist)
        packed, masks = [], {}

        def mask_of(idx):                 # jobs of a grid share their index arrays: pack each array once
            key = id(idx)
            if key not in masks:
                masks[key] = (idx, self.row_mask(idx))
            return masks[key][1]
        for p, pos, neg in jobs:
            pos, neg = np.asarray(pos), np.asarray(neg)
            packed.append((float(p), mask_of(pos), mask_of(neg), pos.shape[0], neg.shape[0]))
        return self._group.scm_best_rules_grid(packed, int(util_block_size))

    def best_utility_rules_grid_masks(self, jobs, rule_blacklist=(), util_block_size=1000000):
        """``best_utility_rules_grid`` for jobs given as (p, positive row mask, negative row mask, n_pos, n_neg): the
        lock-step learners keep their remaining examples as packed masks (learners.scm.fit_many)."""
        self.set_rule_blacklist(rule_blacklist)
        packed = [(float(p), pm, nm, int(n_pos), int(n_neg)) for p, pm, nm, n_pos, n_neg in jobs]
        return self._group.scm_best_rules_grid(packed, int(util_block_size))

    def packed_rule_columns(self, rule_indices):
        """The packed matrix columns (uint64 [W] each, presence bits, nothing inverted) of the k-mers behind the given
        rule indices, fetched with one gather -> list in the order asked."""
        cols = np.asarray([int(r) % self._n_kmers for r in rule_indices], dtype=np.int64)
        distinct, position = np.unique(cols, return_inverse=True)
        packed = self._group.get_packed_columns(distinct)
        return [packed[j] for j in position]

    def gini_best_split(self, example_idx, altered_priors, n_total_class_examples, rule_blacklist=()):
        """cart.py:112-161 + 231-238.  example_idx: dict class -> example indices of the node, iterated in
        dict order like the reference.  -> (min score, presence-rule indices achieving it, ascending)."""
        self.set_rule_blacklist(rule_blacklist)
        classes = list(example_idx.keys())
        masks = np.stack([self.row_mask(example_idx[c]) for c in classes])
        prior = np.array([altered_priors[c] for c in classes], dtype=np.float64)
        total = np.array([n_total_class_examples[c] for c in classes], dtype=np.float64)
        n_node = np.array([len(example_idx[c]) for c in classes], dtype=np.int64)
        return self._group.gini_best_split(masks, prior, total, n_node)

    def gini_best_splits(self, example_idx_by_node, altered_priors, n_total_class_examples, rule_blacklist=(),
                         tie_tables=None):
        """``gini_best_split`` for several nodes at once -- the nodes of a tree level (cart.py:260-339 scans them one
        after the other), or of the same level of SEVERAL trees (the fold trees and the master tree of a
        cross-validation): with two classes, the class masks of up to 8 nodes are counted per pass over the matrix.
        altered_priors / n_total_class_examples: one dict for all the nodes, or one dict per node.
        tie_tables: one DeviceOccurrences (``occurrence_table``) or None per node -- the candidates of such a node are
        then already reduced by the reference's occurrence tiebreaker (experiment_cart.py:82-94), on the device
        whenever the level search runs there.
        -> [(min score, presence-rule indices achieving it, ascending)] in node order."""
        nodes = list(example_idx_by_node)
        per_node = isinstance(altered_priors, (list, tuple))
        priors = list(altered_priors) if per_node else [altered_priors] * len(nodes)
        totals = list(n_total_class_examples) if per_node else [n_total_class_examples] * len(nodes)
        tables = list(tie_tables) if tie_tables is not None else [None] * len(nodes)
        any_table = any(t is not None for t in tables)
        two_class = all(len(ex) == 2 and list(ex.keys()) == list(nodes[0].keys()) for ex in nodes)
        # nodes that carry a (level_key, parent_key) -- the tree learner's NodeExamples -- let the level search count only
        # the smaller of two siblings and derive the other one from the parent's resident counts
        family = None
        if all(getattr(ex, "level_key", None) is not None for ex in nodes):
            family = (np.array([ex.level_key for ex in nodes], dtype=np.int64),
                      np.array([ex.parent_key for ex in nodes], dtype=np.int64))
        on_device = two_class and (len(nodes) > 1 or any_table or family is not None) and \
            (not any_table or self._group.occurrence_table_supported())
        if not on_device:
            found = [self.gini_best_split(ex, pr, tt, rule_blacklist) for ex, pr, tt in zip(nodes, priors, totals)]
            out = []
            for (score, idx), table in zip(found, tables):
                if table is not None and len(idx) > 0:
                    occ = table.host()[idx]
                    idx = idx[np.isclose(occ, occ.max())]
                out.append((score, idx))
            return out
        self.set_rule_blacklist(rule_blacklist)
        classes = list(nodes[0].keys())
        masks = np.stack([np.stack([self.row_mask(ex[c]) for c in classes]) for ex in nodes])
        prior = np.array([[pr[c] for c in classes] for pr in priors], dtype=np.float64)
        total = np.array([[tt[c] for c in classes] for tt in totals], dtype=np.float64)
        n_node = np.array([[len(ex[c]) for c in classes] for ex in nodes], dtype=np.int64)
        occ_ids = np.array([-1 if t is None else t.table for t in tables], dtype=np.int32) if any_table else None
        return self._group.gini_best_splits(masks, prior, total, n_node, occ_ids, family=family)

    def occurrence_table(self, rows):
        """Occurrence counts of every k-mer in the genomes ``rows`` (= ``sum_rows(rows)[:K]``), kept on the device
        for the tiebreaker of the tree learner; ``.host()`` fetches them.  Up to 64 tables at a time; ``release()``
        returns a table."""
        return DeviceOccurrences(self, rows)

    def kmer_risk_tables(self, train_pos_idx, train_neg_idx):
        """The individual k-mer risk tables `kover dataset split` stores per split / fold (split.py:171-188):
        -> (unique_risks, unique_risk_by_kmer, unique_risk_by_anti_kmer).  One pass over the matrix instead of
        two sum_rows passes + six numpy passes over 2K floats + an np.unique over 2K values."""
        train_pos_idx, train_neg_idx = np.asarray(train_pos_idx), np.asarray(train_neg_idx)
        return self._group.kmer_risk_tables(self.row_mask(train_pos_idx), self.row_mask(train_neg_idx),
                                            train_pos_idx.shape[0], train_pos_idx.shape[0] + train_neg_idx.shape[0])

    def gini_scores(self):
        """rules_criterion of the last gini_best_split call (cart.py:228-231), float64 [n_kmers]."""
        return self._group.gini_scores()

    def sum_rows_multi(self, row_sets):
        """Presence counts for several row sets in one pass (the reference's TODO at rules.py:200).
        -> uint32 [len(row_sets), n_kmers]."""
        group = self._group
        if group.comm is not None or not all(hasattr(s, "sum_rows_multi") for s in group.shards):
            # multi-rank groups (sum_rows all-gathers its result) and checker-backed shards: one pass per set
            return np.stack([self.sum_rows(r)[:self._n_kmers].astype(np.uint32) for r in row_sets])
        return group.sum_rows_multi(np.stack([self.row_mask(r) for r in row_sets]))
