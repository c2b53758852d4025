"""Column sharding of the packed matrix and the small per-scan merge.

K-mer columns are independent units: every shard holds all W packed rows of a contiguous column
slice, scans it alone, and only tiny results cross shards -- per-utility-block maxima (one double
per 10^6 rules), tie lists, one selected column.  ``ShardGroup`` runs that protocol over

  * the shards this process owns (one per local GPU; usually exactly one), and
  * optionally a ``Comm`` over torch.distributed (one process per GPU, NCCL on the GPU box, gloo in
    the CPU tests) for the shards other ranks own.

The group is written against a small shard interface (scm_scan / scm_ties / gini_scan / gini_ties /
sum_rows / get_columns / set_blacklist, see _native.Shard) so the host-side merge logic can be
exercised without a GPU by the world_size-2 gloo tests, which plug in a checker-backed shard.
"""
import numpy as np

from . import _native

TILE_COLS = 512


def partition_columns(n_kmers, n_parts):
    """Contiguous column slices [(col0, n_cols)] of near-equal size, boundaries on 512-column tiles.
    Parts may be empty only when n_kmers < n_parts * 512; empty parts are dropped by the callers."""
    n_tiles = (n_kmers + TILE_COLS - 1) // TILE_COLS
    bounds = [min(n_kmers, ((n_tiles * i) // n_parts) * TILE_COLS) for i in range(n_parts)] + [n_kmers]
    return [(bounds[i], bounds[i + 1] - bounds[i]) for i in range(n_parts)]


class Comm(object):
    """Numpy-level collectives over a torch.distributed process group (NCCL or gloo)."""

    def __init__(self, group=None, device=None):
        import torch
        import torch.distributed as dist
        self._torch, self._dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group)
        self.world_size = dist.get_world_size(group)
        backend = dist.get_backend(group)
        self.device = device if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu"))

    def _t(self, a):
        return self._torch.from_numpy(np.ascontiguousarray(a)).to(self.device)

    def allreduce_max(self, a):
        t = self._t(a)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX, group=self.group)
        return t.cpu().numpy()

    def allreduce_sum(self, a):
        t = self._t(a)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def allreduce_min(self, a):
        t = self._t(a)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MIN, group=self.group)
        return t.cpu().numpy()

    def allgather_concat(self, a):
        """Concatenate 1-D int64 arrays of different lengths from every rank, in rank order.  ONE collective in the usual
        case: every rank sends [its length, its data] padded to a shared capacity; only when some rank's array does not
        fit -- every rank sees the same lengths, so all of them decide alike -- the capacity grows (it is remembered) and
        the gather is repeated."""
        a = np.ascontiguousarray(a, dtype=np.int64)
        cap = self.__dict__.get("_concat_capacity", 1024)
        while True:
            buf = np.zeros(cap + 1, dtype=np.int64)
            buf[0] = a.shape[0]
            buf[1:1 + min(a.shape[0], cap)] = a[:cap]
            t = self._t(buf)
            out = [self._torch.empty_like(t) for _ in range(self.world_size)]
            self._dist.all_gather(out, t, group=self.group)
            rows = self._torch.stack(out).cpu().numpy()
            sizes = rows[:, 0]
            largest = int(sizes.max())
            # the next call starts from twice what this one needed (tie lists of a tree's small nodes run to millions,
            # those of the next fit to a handful)
            self._concat_capacity = max(1024, 1 << max(0, 2 * largest - 1).bit_length())
            if largest <= cap:
                return np.concatenate([rows[r, 1:1 + int(sizes[r])] for r in range(self.world_size)])
            cap = self._concat_capacity

    def byte_buffers(self, nbytes):
        """(local [nbytes], gathered [world, nbytes]) uint8 tensors on the collective's device, cached by size."""
        cache = self.__dict__.setdefault("_byte_buffers", {})
        if nbytes not in cache:
            cache[nbytes] = (self._torch.zeros(nbytes, dtype=self._torch.uint8, device=self.device),
                             self._torch.zeros((self.world_size, nbytes), dtype=self._torch.uint8, device=self.device))
        return cache[nbytes]

    def allgather_bytes(self, local, gathered):
        """One fixed-size all-gather; returns the gathered bytes as a host numpy array [world, nbytes]."""
        self._dist.all_gather(list(gathered.unbind(0)), local, group=self.group)
        return gathered.cpu().numpy()

    def broadcast(self, a, src):
        t = self._t(a)
        self._dist.broadcast(t, src=self._dist.get_global_rank(self.group, src) if self.group is not None else src,
                             group=self.group)
        return t.cpu().numpy()

    def barrier(self):
        self._dist.barrier(group=self.group)


def merge_packets(rows, n_blocks, capacity, p, util_block_size):
    """rows: uint8 [n_shards, packet_bytes] of kv_scm_scan_packet packets.  Max-merge of the block maxima, the
    reference's sequential block merge (scm.py:270-286) and the exact isclose filter (scm.py:273) on the
    gathered candidates, all in the library's host code (kv_scm_merge_packets).
    -> ((best, idx, pos_err, neg_cover) or None if a packet overflowed, bm, best, sel)."""
    return _native.merge_packets(rows, n_blocks, capacity, p, util_block_size)


class ShardGroup(object):
    """Shards owned by this process (+ optional Comm to the other ranks) acting as one matrix."""

    def __init__(self, shards, n_examples, n_kmers, comm=None, rank_slices=None):
        self.shards = list(shards)
        self.n_examples = int(n_examples)
        self.n_words = (self.n_examples + 63) // 64
        self.n_kmers = int(n_kmers)
        self.comm = comm
        # rank_slices[r] = (col0, n_cols) of rank r (to find the owner of a column)
        self.rank_slices = rank_slices
        if comm is None:
            covered = sum(s.n_cols for s in self.shards)
            if covered != self.n_kmers:
                raise ValueError("local shards cover %d of %d columns" % (covered, self.n_kmers))
        self._single = comm is None and len(self.shards) == 1
        self.exchange_stats = {"one_exchange": 0, "two_phase_fallback": 0}
        self.p2p_slot_bytes = 0
        if comm is not None and len(self.shards) == 1 and comm.world_size > 1:
            self._try_p2p()

    def _try_p2p(self):
        """Peer mailbox over CUDA IPC / NVLink (kv_p2p_*): used when every rank could map every peer.  Disabled
        with KOVER_B200_P2P=0 or when the shards are not device shards; any failure on any rank falls back
        to the NCCL all-gather for all ranks."""
        import os
        shard = self.shards[0]
        ok = int(hasattr(shard, "p2p_create") and os.environ.get("KOVER_B200_P2P", "1") != "0"
                 and str(self.comm.device).startswith("cuda"))
        slot = _native.packet_bytes(max(1, _native.n_utility_blocks(self.n_kmers, 1000000)), self.PACKET_CAPACITY)
        handle = np.zeros(64, dtype=np.uint8)
        if ok:
            try:
                handle = shard.p2p_create(self.comm.rank, self.comm.world_size, slot)
            except Exception:  # noqa: BLE001
                ok = 0
        if int(self.comm.allreduce_min(np.array([ok], dtype=np.int64))[0]) == 0:
            return
        local, gathered = self.comm.byte_buffers(64)
        local.copy_(local.new_tensor(handle))
        handles = self.comm.allgather_bytes(local, gathered)
        try:
            shard.p2p_connect(handles)
        except Exception:  # noqa: BLE001
            ok = 0
        if int(self.comm.allreduce_min(np.array([ok], dtype=np.int64))[0]) == 1:
            self.p2p_slot_bytes = slot

    # ------------------------------------------------------------------ upload
    def upload_block(self, block, row0, gcol0):
        for s in self.shards:
            s.upload_block(block, row0, gcol0)

    # ------------------------------------------------------------------ a3
    def sum_rows_into(self, mask, n_rows, out):
        """The reference's sum_rows result (rules.py:201-267) into out[2K]: presence counts of the masked examples
        in out[:K], n_rows - presence in out[K:], both written by the shards in the output dtype."""
        k = self.n_kmers
        for s in self.shards:
            if hasattr(s, "sum_rows_full"):
                s.sum_rows_full(mask, n_rows, out[s.col0:s.col0 + s.n_cols], out[k + s.col0:k + s.col0 + s.n_cols])
            else:
                s.sum_rows(mask, out[s.col0:s.col0 + s.n_cols])
                out[k + s.col0:k + s.col0 + s.n_cols] = n_rows - out[s.col0:s.col0 + s.n_cols]
        if self.comm is not None:
            # every rank needs the full vector to honour the reference's return value
            lo = min(s.col0 for s in self.shards)
            hi = max(s.col0 + s.n_cols for s in self.shards)
            full = self.comm.allgather_concat(out[lo:hi].astype(np.int64))
            out[:k] = full.astype(out.dtype)
            out[k:] = n_rows - out[:k]
        return out

    # ------------------------------------------------------------------ dataset split risk tables
    def kmer_risk_tables(self, pos_mask, neg_mask, n_pos, n_total):
        """dataset/split.py:178-188 -> (unique_risks float64, unique_risk_by_kmer, unique_risk_by_anti_kmer),
        the two index arrays in _minimum_uint_size(len(unique_risks)) like split.py:186-187."""
        from .utils import _minimum_uint_size
        if self.comm is not None:
            raise NotImplementedError("kmer_risk_tables runs in the single process that writes the split")
        present = np.zeros(n_total + 1, dtype=np.uint8)
        for s in self.shards:
            s.risk_errors(pos_mask, neg_mask, n_pos, n_total, present)
        e = np.flatnonzero(present)
        # the reference's float arithmetic, on the distinct error counts only (it is elementwise)
        kmer_risks = e.astype(float)
        kmer_risks /= n_total
        np.round(kmer_risks, 5, out=kmer_risks)
        anti_kmer_risks = 1.0 - kmer_risks
        np.round(anti_kmer_risks, 5, out=anti_kmer_risks)
        unique_risks, inverse = np.unique(np.hstack((kmer_risks, anti_kmer_risks)), return_inverse=True)
        lut_kmer = np.zeros(n_total + 1, dtype=np.uint32)
        lut_anti = np.zeros(n_total + 1, dtype=np.uint32)
        lut_kmer[e] = inverse[:e.shape[0]]
        lut_anti[e] = inverse[e.shape[0]:]
        dtype = _minimum_uint_size(len(unique_risks))
        by_kmer = _native.result_empty(self.n_kmers, dtype)
        by_anti = _native.result_empty(self.n_kmers, dtype)
        for s in self.shards:
            s.risk_map(lut_kmer, lut_anti, by_kmer[s.col0:s.col0 + s.n_cols], by_anti[s.col0:s.col0 + s.n_cols])
        return unique_risks, by_kmer, by_anti

    # ------------------------------------------------------------------ a4
    def get_packed_columns(self, cols):
        """cols: presence-column indices (global).  -> uint64 [len(cols), W]."""
        cols = np.asarray(cols, dtype=np.int64)
        out = np.zeros((cols.shape[0], self.n_words), dtype=np.uint64)
        for s in self.shards:
            sel = np.flatnonzero((cols >= s.col0) & (cols < s.col0 + s.n_cols))
            if sel.size:
                out[sel] = s.get_columns(cols[sel] - s.col0)
        if self.comm is not None:
            # exactly one rank owns each column and the others hold zeros: a sum-reduce is a broadcast
            out = self.comm.allreduce_sum(out.view(np.int64)).view(np.uint64)
        return out

    # ------------------------------------------------------------------ a6
    def set_blacklist(self, rule_idx):
        for s in self.shards:
            s.set_blacklist(rule_idx)

    PACKET_CAPACITY = 256      # tie candidates per rank that travel with the block maxima

    def _scm_best_rules_one_exchange(self, pos_mask, neg_mask, n_pos, n_neg, p, util_block_size):
        """Multi-rank scan with ONE collective: every rank all-gathers {its block maxima, a local superset of
        the ties}; the merge (scm.py:270-286) and the exact isclose filter (scm.py:273) then run redundantly
        on every host.  Returns None when some rank's candidates overflowed the packet."""
        shard, cap = self.shards[0], self.PACKET_CAPACITY
        nb = _native.n_utility_blocks(self.n_kmers, util_block_size)
        local, gathered = self.comm.byte_buffers(_native.packet_bytes(nb, cap))
        shard.scm_scan_packet(pos_mask, neg_mask, n_pos, n_neg, p, util_block_size, cap, local)
        return merge_packets(self.comm.allgather_bytes(local, gathered), nb, cap, p, util_block_size)

    # ------------------------------------------------------------------ device-resident greedy state (scm.py:133-139)
    def scm_state_supported(self):
        return all(hasattr(s, "mask_and_column") for s in self.shards) and \
            (self._single or (self.comm is not None and len(self.shards) == 1 and self.p2p_slot_bytes > 0))

    def scm_state_set(self, pos_mask, neg_mask):
        return self.shards[0].scm_state_set(pos_mask, neg_mask)

    def scm_state_apply(self, kmer_col, invert):
        """AND the resident masks of every shard / rank with matrix column ``kmer_col`` (global), inverted for an
        absence rule.  The owner reads the column from its matrix; the other ranks receive its W words.
        -> (n_pos, n_neg) left."""
        shard = self.shards[0]
        if self.comm is None:
            n_pos, n_neg, _ = shard.mask_and_column(local_col=kmer_col - shard.col0, invert=invert)
            return n_pos, n_neg
        owner = next(r for r, (c0, nc) in enumerate(self.rank_slices) if c0 <= kmer_col < c0 + nc)
        if owner == self.comm.rank:
            n_pos, n_neg, words = shard.mask_and_column(local_col=kmer_col - shard.col0, invert=invert, want_words=True)
            self.comm.broadcast(words.view(np.int64), owner)
            return n_pos, n_neg
        words = self.comm.broadcast(np.zeros(self.n_words, dtype=np.int64), owner).view(np.uint64)
        n_pos, n_neg, _ = shard.mask_and_column(words=words, invert=invert)
        return n_pos, n_neg

    def scm_best_rules_state(self, p, util_block_size):
        """The scan over the resident state -> the 4-tuple, or None (overflow: the caller falls back to index arrays)."""
        shard = self.shards[0]
        if self._single:
            return shard.scm_best_rules_state(p, util_block_size)
        nb = _native.n_utility_blocks(self.n_kmers, util_block_size)
        if self.p2p_slot_bytes >= _native.packet_bytes(nb, self.PACKET_CAPACITY):
            out = shard.scm_best_rules_p2p_state(p, util_block_size, self.PACKET_CAPACITY)
            if out is not None:
                self.exchange_stats["one_exchange"] += 1
            return out
        return None

    def scm_best_rules_idx(self, pos_idx, neg_idx, p, util_block_size):
        """scm_best_rules from contiguous int64 example index arrays, when the shards can take them directly (one
        device shard, or one process per GPU with the peer mailbox up): the masks are built inside the library and
        the whole step is one kernel launch.  -> the 4-tuple, or None: use scm_best_rules with masks."""
        shard = self.shards[0]
        if self._single:
            if hasattr(shard, "scm_best_rules_idx"):
                return shard.scm_best_rules_idx(pos_idx, neg_idx, p, util_block_size)
            return None
        if self.comm is not None and len(self.shards) == 1 and hasattr(shard, "scm_best_rules_p2p_idx"):
            fits = self.__dict__.setdefault("_p2p_fits", {})          # (two library calls per step otherwise)
            if util_block_size not in fits:
                nb = _native.n_utility_blocks(self.n_kmers, util_block_size)
                fits[util_block_size] = self.p2p_slot_bytes >= _native.packet_bytes(nb, self.PACKET_CAPACITY)
            if fits[util_block_size]:
                out = shard.scm_best_rules_p2p_idx(pos_idx, neg_idx, p, util_block_size, self.PACKET_CAPACITY)
                if out is not None:
                    self.exchange_stats["one_exchange"] += 1
                return out
        return None

    def scm_best_rules(self, pos_mask, neg_mask, n_pos, n_neg, p, util_block_size):
        """-> (best_utility, rule_idx ascending int64, pos_err int64, neg_cover int64); scm.py:238-288."""
        if self._single:
            return self.shards[0].scm_best_rules(pos_mask, neg_mask, n_pos, n_neg, p, util_block_size)
        if self.comm is not None and len(self.shards) == 1 and hasattr(self.shards[0], "scm_scan_packet"):
            nb = _native.n_utility_blocks(self.n_kmers, util_block_size)
            if self.p2p_slot_bytes >= _native.packet_bytes(nb, self.PACKET_CAPACITY):
                out, bm, best, sel = self.shards[0].scm_best_rules_p2p(pos_mask, neg_mask, n_pos, n_neg, p,
                                                                        util_block_size, self.PACKET_CAPACITY)
            else:
                out, bm, best, sel = self._scm_best_rules_one_exchange(pos_mask, neg_mask, n_pos, n_neg, p,
                                                                       util_block_size)
            if out is not None:
                self.exchange_stats["one_exchange"] += 1
                return out
            self.exchange_stats["two_phase_fallback"] += 1
            # a huge equivalent-rule set: the counts are still on the device, collect exactly in a second exchange
            idx, pe, nc = self.shards[0].scm_ties(bm, sel)
            flat = self.comm.allgather_concat(np.stack([idx, pe, nc], axis=1).reshape(-1)).reshape(-1, 3)
            order = np.argsort(flat[:, 0], kind="stable")
            return best, flat[order, 0], flat[order, 1], flat[order, 2]
        bm = None
        for s in self.shards:
            b = s.scm_scan(pos_mask, neg_mask, n_pos, n_neg, p, util_block_size)
            bm = b if bm is None else np.maximum(bm, b)
        if self.comm is not None:
            bm = self.comm.allreduce_max(bm)
        best, sel = _native.select_blocks(bm)
        parts = [s.scm_ties(bm, sel) for s in self.shards]
        idx = np.concatenate([q[0] for q in parts])
        pe = np.concatenate([q[1] for q in parts])
        nc = np.concatenate([q[2] for q in parts])
        if self.comm is not None:
            flat = self.comm.allgather_concat(np.stack([idx, pe, nc], axis=1).reshape(-1)).reshape(-1, 3)
            idx, pe, nc = flat[:, 0], flat[:, 1], flat[:, 2]
        order = np.argsort(idx, kind="stable")
        return best, idx[order], pe[order], nc[order]

    def scm_best_rules_grid(self, jobs, util_block_size):
        """Several SCM scans that may share example masks (north-star item 4: the p grid x both model types x
        the fits of a cross-validation in lock-step).  jobs: [(p, pos_mask, neg_mask, n_pos, n_neg)].  The
        matrix is scanned once per DISTINCT unordered mask pair; every other job re-scores the counts that scan
        left on the device (kv_scm_reuse_counts: another p, or positives and negatives exchanged as the
        disjunction does, scm.py:69-73).  Results in job order, each as scm_best_rules returns it."""
        groups, content = {}, {}

        def content_of(mask):            # (the jobs of a grid share a handful of mask arrays: one tobytes per array)
            b = content.get(id(mask))
            if b is None:
                b = content[id(mask)] = mask.tobytes()
            return b
        for i, (p, pm, nm, n_pos, n_neg) in enumerate(jobs):
            key = (content_of(pm), content_of(nm))
            if key in groups:
                groups[key].append((i, False))
            elif (key[1], key[0]) in groups:
                groups[(key[1], key[0])].append((i, True))
            else:
                groups[key] = [(i, False)]
        results = [None] * len(jobs)
        if self._single and hasattr(self.shards[0], "scm_queue"):
            # one device: enqueue every job back to back, synchronise once
            order, queued = [], []
            for members in groups.values():
                for rank_in_group, (i, swap) in enumerate(members):
                    p, pm, nm, n_pos, n_neg = jobs[i]
                    order.append(i)
                    queued.append((p, pm, nm, n_pos, n_neg, 0 if rank_in_group == 0 else (2 if swap else 1)))
            for i, res in zip(order, self.shards[0].scm_queue(queued, util_block_size)):
                results[i] = res
            return results
        if all(hasattr(s, "scm_queue_packets") for s in self.shards) and (self.comm is None or len(self.shards) == 1):
            return self._scm_grid_packets(jobs, groups, util_block_size)
        for members in groups.values():
            for rank_in_group, (i, swap) in enumerate(members):
                p, pm, nm, n_pos, n_neg = jobs[i]
                if rank_in_group > 0:
                    for s in self.shards:
                        s.scm_reuse_counts(swap)
                results[i] = self.scm_best_rules(pm, nm, n_pos, n_neg, p, util_block_size)
        return results

    def _scm_grid_packets(self, jobs, groups, util_block_size):
        """The lock-step grid over partial shards (several per process, or one process per GPU): every shard runs
        the whole grid locally (kv_scm_queue_packets: distinct mask pairs counted several per matrix pass, every job
        scored from the resident counts) and emits one packet per job; ONE exchange gathers the packets of all the
        jobs; every host then merges them job by job (max-merge of the block maxima, the reference's sequential block
        merge, exact isclose filter).  A job whose candidates overflowed a packet is redone alone."""
        cap = self.PACKET_CAPACITY
        nb = _native.n_utility_blocks(self.n_kmers, util_block_size)
        pb = _native.packet_bytes(nb, cap)
        order, queued = [], []
        for members in groups.values():
            for rank_in_group, (i, swap) in enumerate(members):
                p, pm, nm, n_pos, n_neg = jobs[i]
                order.append(i)
                queued.append((p, pm, nm, n_pos, n_neg, 0 if rank_in_group == 0 else (2 if swap else 1)))
        n = len(queued)
        if self.comm is not None:
            local, gathered = self.comm.byte_buffers(n * pb)
            self.shards[0].scm_queue_packets(queued, util_block_size, cap, local)
            rows = self.comm.allgather_bytes(local, gathered).reshape(self.comm.world_size, n, pb)
        else:
            rows = np.empty((len(self.shards), n, pb), dtype=np.uint8)
            for k, s in enumerate(self.shards):
                s.scm_queue_packets(queued, util_block_size, cap, rows[k])
        self.exchange_stats["grid_exchanges"] = self.exchange_stats.get("grid_exchanges", 0) + 1
        results = [None] * len(jobs)
        for slot, i in enumerate(order):
            out, _bm, _best, _sel = merge_packets(np.ascontiguousarray(rows[:, slot]), nb, cap, queued[slot][0],
                                                  util_block_size)
            if out is None:                  # a huge equivalent-rule set on some shard: this job alone, two-phase
                p, pm, nm, n_pos, n_neg = jobs[i]
                self.exchange_stats["two_phase_fallback"] += 1
                out = self.scm_best_rules(pm, nm, n_pos, n_neg, p, util_block_size)
            results[i] = out
        return results

    # ------------------------------------------------------------------ a9 / a11
    def gini_best_split(self, class_masks, altered_prior, n_total, n_node):
        """-> (min_score, presence-rule indices with score == min_score, ascending); cart.py:112-161, 231-238."""
        if self._single and hasattr(self.shards[0], "gini_best_split"):
            return self.shards[0].gini_best_split(class_masks, altered_prior, n_total, n_node)
        mins = [s.gini_scan(class_masks, altered_prior, n_total, n_node) for s in self.shards]
        m = min(mins)
        if self.comm is not None:
            m = float(self.comm.allreduce_min(np.array([m], dtype=np.float64))[0])
        if m == np.inf:
            return m, np.zeros(0, dtype=np.int64)
        idx = np.concatenate([s.gini_ties(m) if mn == m else np.zeros(0, dtype=np.int64)
                              for s, mn in zip(self.shards, mins)])
        if self.comm is not None:
            idx = self.comm.allgather_concat(idx)
        return m, np.sort(idx)

    # ------------------------------------------------------------------ occurrence tables (CART tiebreaker on the device)
    def occurrence_table_supported(self):
        return all(hasattr(s, "occ_table_set") for s in self.shards) and (self.comm is None or len(self.shards) == 1)

    def occurrence_table_set(self, table, mask):
        for s in self.shards:
            s.occ_table_set(table, mask)

    def gini_best_splits(self, node_masks, altered_prior, n_total, n_node, occ_tables=None, family=None):
        """gini_best_split for several two-class nodes (a tree level): node_masks [n_nodes, 2, W], n_node
        [n_nodes, 2].  Every shard counts the class masks of up to 8 nodes per matrix pass (kv_gini_level); the
        minima are merged over shards / ranks and the tie lists of the shards that hold a node's minimum are
        concatenated in column order.  occ_tables [n_nodes] (local shards only): the reference's occurrence
        tiebreaker (experiment_cart.py:82-94) is applied on the device, the lists hold its survivors.
        family = (node keys, parent keys) [n_nodes] each: sibling subtraction across levels (kv_gini_level_family)."""
        n_node = np.asarray(n_node, dtype=np.int64)
        altered_prior = np.broadcast_to(np.asarray(altered_prior, dtype=np.float64), n_node.shape)
        n_total = np.broadcast_to(np.asarray(n_total, dtype=np.float64), n_node.shape)
        if not hasattr(self.shards[0], "gini_level"):
            assert occ_tables is None
            return [self.gini_best_split(m, pr, tt, nn) for m, pr, tt, nn in zip(node_masks, altered_prior, n_total, n_node)]
        fam = {"family": family} if family is not None and hasattr(self.shards[0], "gini_level_family") else {}
        if occ_tables is not None and self.comm is not None:
            return self._gini_best_splits_ranks_tiebroken(node_masks, altered_prior, n_total, n_node, occ_tables, fam)
        if occ_tables is not None:
            per_shard = [s.gini_level(node_masks, altered_prior, n_total, n_node, occ_tables, **fam) for s in self.shards]
            if self._single:
                return [(m, idx) for m, idx, _ in per_shard[0]]
            out = []
            for i in range(len(per_shard[0])):
                m = min(res[i][0] for res in per_shard)
                holders = [(s, res[i]) for s, res in zip(self.shards, per_shard) if res[i][0] == m and m != np.inf]
                if not holders:
                    out.append((float(m), np.zeros(0, dtype=np.int64)))
                    continue
                # a shard's survivors are relative to ITS largest occurrence: they stand when that is the global one;
                # a smaller maximum within the tolerance (only possible from 100 000 occurrences up) is re-filtered
                g = max(r[2] for _, r in holders)
                tol = 1e-8 + 1e-5 * abs(float(g))
                parts = []
                for s, (_, idx, local_max) in holders:
                    if local_max == g or occ_tables[i] < 0:
                        parts.append(np.array(idx))
                    elif abs(float(local_max) - float(g)) <= tol:
                        o = s.occ_table_gather(int(occ_tables[i]), idx).astype(np.float64)
                        parts.append(np.array(idx)[np.abs(o - float(g)) <= tol])
                out.append((float(m), np.sort(np.concatenate(parts))))
            return out
        per_shard = [s.gini_level(node_masks, altered_prior, n_total, n_node, **fam) for s in self.shards]
        if self._single:
            return per_shard[0]         # (the tie lists are read-only views, valid until the next level search)
        n_nodes = len(per_shard[0])
        mins = np.array([[res[i][0] for i in range(n_nodes)] for res in per_shard], dtype=np.float64).min(axis=0)
        if self.comm is not None:
            mins = self.comm.allreduce_min(mins)
        local = []
        for i in range(n_nodes):
            m = float(mins[i])
            parts = [res[i][1] for res in per_shard if res[i][0] == m and m != np.inf]
            local.append(np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64))
        if self.comm is None:
            return [(float(mins[i]), np.sort(local[i])) for i in range(n_nodes)]
        # one exchange for the whole level: every rank's per-node tie counts, then all the tie indices at once
        sizes = np.array([len(a) for a in local], dtype=np.int64)
        all_sizes = self.comm.allgather_concat(sizes).reshape(self.comm.world_size, n_nodes)
        flat = self.comm.allgather_concat(np.concatenate(local) if n_nodes else np.zeros(0, dtype=np.int64))
        starts = np.concatenate(([0], np.cumsum(all_sizes.reshape(-1))))
        out = []
        for i in range(n_nodes):
            parts = [flat[starts[r * n_nodes + i]:starts[r * n_nodes + i + 1]] for r in range(self.comm.world_size)]
            out.append((float(mins[i]), np.sort(np.concatenate(parts))))
        return out

    def _gini_best_splits_ranks_tiebroken(self, node_masks, altered_prior, n_total, n_node, occ_tables, fam={}):
        """One process per GPU, occurrence tiebreaker on the device: every rank reports, per node, its minimum, the
        largest occurrence among ITS ties and the ties that survive relative to that; two small reductions find the
        global minimum and the global largest occurrence, the ranks whose survivors stand send them in one exchange
        for the whole level -- the millions of ties of small nodes never cross NVLink."""
        shard = self.shards[0]
        res = shard.gini_level(node_masks, altered_prior, n_total, n_node, occ_tables, **fam)
        n_nodes = len(res)
        local_min = np.array([r[0] for r in res], dtype=np.float64)
        mins = self.comm.allreduce_min(local_min)
        holder = (local_min == mins) & (mins != np.inf)
        local_max = np.array([float(r[2]) if h and occ_tables[i] >= 0 else 0.0
                              for i, (r, h) in enumerate(zip(res, holder))], dtype=np.float64)
        g = self.comm.allreduce_max(local_max)
        local = []
        for i in range(n_nodes):
            if not holder[i]:
                local.append(np.zeros(0, dtype=np.int64))
                continue
            idx = np.array(res[i][1])
            if occ_tables[i] >= 0 and local_max[i] != g[i]:
                tol = 1e-8 + 1e-5 * abs(g[i])
                if abs(local_max[i] - g[i]) <= tol:      # only possible from 100 000 occurrences up
                    o = shard.occ_table_gather(int(occ_tables[i]), idx).astype(np.float64)
                    idx = idx[np.abs(o - g[i]) <= tol]
                else:
                    idx = np.zeros(0, dtype=np.int64)
            local.append(idx)
        sizes = np.array([len(a) for a in local], dtype=np.int64)
        all_sizes = self.comm.allgather_concat(sizes).reshape(self.comm.world_size, n_nodes)
        flat = self.comm.allgather_concat(np.concatenate(local) if n_nodes else np.zeros(0, dtype=np.int64))
        starts = np.concatenate(([0], np.cumsum(all_sizes.reshape(-1))))
        out = []
        for i in range(n_nodes):
            parts = [flat[starts[r * n_nodes + i]:starts[r * n_nodes + i + 1]] for r in range(self.comm.world_size)]
            out.append((float(mins[i]), np.sort(np.concatenate(parts))))
        return out

    def gini_scores(self):
        if self.comm is not None:
            raise NotImplementedError("gini_scores is a single-process debugging aid")
        out = np.empty(self.n_kmers, dtype=np.float64)
        for s in self.shards:
            out[s.col0:s.col0 + s.n_cols] = s.gini_scores()
        return out

    def sum_rows_multi(self, masks):
        """uint32 [n_masks, K] presence counts of several masks in one pass (local shards only)."""
        if self.comm is not None:
            raise NotImplementedError("sum_rows_multi is per-process")
        masks = np.ascontiguousarray(masks, dtype=np.uint64)
        if len(self.shards) == 1:
            return self.shards[0].sum_rows_multi(masks)
        out = np.empty((masks.shape[0], self.n_kmers), dtype=np.uint32)
        for s in self.shards:
            out[:, s.col0:s.col0 + s.n_cols] = s.sum_rows_multi(masks)
        return out

    def close(self):
        for s in self.shards:
            s.close()
        self.shards = []
