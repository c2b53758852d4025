"""TEST INFRASTRUCTURE: a CPU stand-in for _native.Shard built on the checker (oracle/), with the same
method surface, so the host-side multi-shard / multi-rank merge logic of kover_b200.sharding can be
exercised without a GPU (world_size-2 gloo tests).  Never imported by the product."""
import numpy as np

from oracle import kover_oracle as ko


class OracleShard(object):
    def __init__(self, packed_full, n_examples, col0, n_cols):
        self.P = np.ascontiguousarray(packed_full[:, col0:col0 + n_cols])
        self.n_examples, self.n_words = n_examples, self.P.shape[0]
        self.n_kmers_total, self.col0, self.n_cols = packed_full.shape[1], col0, n_cols
        self.black = np.zeros(0, dtype=np.int64)

    def close(self):
        pass

    def _counts(self, mask):
        acc = np.zeros(self.n_cols, dtype=np.uint64)
        ko._clib().oracle_masked_column_counts_64(self.P.ctypes.data, self.n_words, self.n_cols, self.n_cols,
                                                  np.ascontiguousarray(mask, dtype=np.uint64).ctypes.data, acc.ctypes.data)
        return acc.astype(np.int64)

    def sum_rows(self, mask, out):
        out[:] = self._counts(mask).astype(out.dtype)
        return out

    def get_columns(self, local_cols):
        return np.ascontiguousarray(self.P[:, np.asarray(local_cols, dtype=np.int64)].T)

    def set_blacklist(self, rule_idx):
        self.black = np.asarray(rule_idx, dtype=np.int64)

    def scm_reuse_counts(self, swap_roles=False):
        pass          # the checker-backed shard recomputes from the masks it is given

    def _utilities(self):
        k, c0, n = self.n_kmers_total, self.col0, self.n_cols
        pres, absn = self._cnt
        rules = np.concatenate([np.arange(c0, c0 + n), k + np.arange(c0, c0 + n)])
        neg_cover = np.concatenate([self._n_neg - pres[1], pres[1]])
        pos_err = np.concatenate([self._n_pos - pres[0], pres[0]])
        u = neg_cover - self._p * pos_err.astype(np.float64)
        u[np.isin(rules, self.black % (2 * k))] = -np.inf
        return rules, u, pos_err, neg_cover

    def scm_scan(self, pos_mask, neg_mask, n_pos, n_neg, p, util_block_size):
        self._cnt = ((self._counts(pos_mask), self._counts(neg_mask)), None)
        self._n_pos, self._n_neg, self._p, self._ubs = n_pos, n_neg, float(p), util_block_size
        rules, u, _, _ = self._utilities()
        nb = -(-2 * self.n_kmers_total // util_block_size)
        bm = np.full(nb, -np.inf)
        np.maximum.at(bm, rules // util_block_size, u)
        return bm

    def scm_scan_packet(self, pos_mask, neg_mask, n_pos, n_neg, p, util_block_size, capacity, packet):
        """The packet kv_scm_scan_packet writes (include/kover_b200.h), into a CPU torch tensor / numpy array."""
        from kover_b200 import _native
        bm = self.scm_scan(pos_mask, neg_mask, n_pos, n_neg, p, util_block_size)
        rules, u, pe, nc = self._utilities()
        lmax = bm.max()
        thr = lmax - 5.0 * (1e-8 + 1e-5 * abs(lmax)) if np.isfinite(lmax) else -np.inf
        keep = u >= thr
        nb = bm.shape[0]
        buf = packet if isinstance(packet, np.ndarray) else packet.numpy()
        buf[:8 * nb] = bm.view(np.uint8)
        buf[8 * nb:8 * nb + 8] = np.array([keep.sum()], dtype=np.uint64).view(np.uint8)
        buf[8 * nb + 8:8 * nb + 16] = np.array([thr], dtype=np.float64).view(np.uint8)
        recs = np.zeros(min(int(keep.sum()), capacity), dtype=_native.PACKET_REC_DTYPE)
        m = recs.shape[0]
        flagged = rules[keep] | np.where(np.isneginf(u[keep]), np.int64(_native.TIE_FLAG_NEG_INF), np.int64(0))
        recs["idx"], recs["pe"], recs["nc"] = flagged[:m], pe[keep][:m], nc[keep][:m]
        buf[8 * nb + 16:8 * nb + 16 + 16 * m] = recs.view(np.uint8)

    def scm_ties(self, block_max, selected, capacity=4096):
        rules, u, pe, nc = self._utilities()
        b = rules // self._ubs
        keep = np.asarray(selected, dtype=bool)[b] & np.isclose(u, np.asarray(block_max)[b])
        order = np.argsort(rules[keep])
        return rules[keep][order], pe[keep][order], nc[keep][order]

    def gini_scan(self, class_masks, altered_prior, n_total, n_node):
        classes = range(len(n_node))
        left = {c: self._counts(class_masks[c]).astype(float) for c in classes}
        right = {c: float(n_node[c]) - left[c] for c in classes}
        pri = {c: altered_prior[c] for c in classes}
        tot = {c: n_total[c] for c in classes}
        g = ko.gini_impurity(left, pri, tot, True)
        g = g + ko.gini_impurity(right, pri, tot, True)
        g[sum(left.values()) == 0] = np.inf
        g[sum(right.values()) == 0] = np.inf
        bl = self.black[(self.black >= self.col0) & (self.black < self.col0 + self.n_cols)] - self.col0
        g[bl] = np.inf
        self._gini = g
        return float(g.min())

    def gini_ties(self, min_score, capacity=4096):
        return np.where(self._gini == min_score)[0].astype(np.int64) + self.col0

    def gini_scores(self):
        return self._gini
