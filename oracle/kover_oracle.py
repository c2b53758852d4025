"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy + a small C kernel) of Kover's
rule-scoring hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg may import this module.  The product (``kover_b200/``) never
does: it fails loudly when the CUDA library is missing.

Every function cites the reference lines it follows (paths relative to
``/root/reference/core/kover``).  Parity pinning: the reference ships no tests or golden
vectors, so this oracle is pinned against OUTPUTS OF THE REFERENCE
ITSELF, run in the build container through ``oracle/ref_loader.py``:
``tests/golden/*.npz`` (minted by ``tests/golden/make_golden.py``, committed) and the live
comparison in ``tests/test_oracle_vs_reference.py``.
"""
import ctypes
import os
import subprocess
from collections import deque
from math import ceil

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_popcount.so")
_lib = None

UTIL_BLOCK_SIZE = 1000000          # learners/scm.py:29
GINI_BLOCK_SIZE = 100000           # learners/cart.py:142


def build_oracle_lib(force=False):
    src = os.path.join(_HERE, "popcount_oracle.c")
    if force or not os.path.isfile(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        # the reference builds its kernel with -march=native (core/setup.py:30-31)
        subprocess.check_call(["gcc", "-O2", "-march=native", "-shared", "-fPIC", src, "-o", _LIB_PATH])
    return _LIB_PATH


def _clib():
    global _lib
    if _lib is None:
        build_oracle_lib()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_inplace_popcount_64_2d.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_long, ctypes.c_void_p]
        _lib.oracle_inplace_popcount_32_2d.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_long, ctypes.c_void_p]
        _lib.oracle_masked_column_counts_64.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_long, ctypes.c_long,
                                                        ctypes.c_void_p, ctypes.c_void_p]
        for f in (_lib.oracle_inplace_popcount_64_2d, _lib.oracle_inplace_popcount_32_2d,
                  _lib.oracle_masked_column_counts_64):
            f.restype = None
    return _lib


# --------------------------------------------------------------------------------------
# popcount kernel -- learning/common/popcount.pyx:52-67, 97-112
# --------------------------------------------------------------------------------------
def inplace_popcount_64(arr, row_mask):
    assert arr.dtype == np.uint64 and arr.ndim == 2 and arr.flags.c_contiguous
    row_mask = np.ascontiguousarray(row_mask, dtype=np.uint64)
    assert row_mask.shape[0] >= arr.shape[0]
    _clib().oracle_inplace_popcount_64_2d(arr.ctypes.data, arr.shape[0], arr.shape[1], row_mask.ctypes.data)


def inplace_popcount_32(arr, row_mask):
    assert arr.dtype == np.uint32 and arr.ndim == 2 and arr.flags.c_contiguous
    row_mask = np.ascontiguousarray(row_mask, dtype=np.uint32)
    assert row_mask.shape[0] >= arr.shape[0]
    _clib().oracle_inplace_popcount_32_2d(arr.ctypes.data, arr.shape[0], arr.shape[1], row_mask.ctypes.data)


# --------------------------------------------------------------------------------------
# layout helpers -- utils.py:117-187
# --------------------------------------------------------------------------------------
def minimum_uint_size(max_value):
    """utils.py:117-130"""
    for t in (np.uint8, np.uint16, np.uint32, np.uint64):
        if max_value <= np.iinfo(t).max:
            return t
    raise ValueError("value too large")


def pack_binary_bytes_to_ints(a, pack_size):
    """utils.py:133-156 -- example i -> word i // pack_size, bit pack_size-1-(i % pack_size)."""
    if pack_size not in (32, 64):
        raise ValueError("Supported data types are 32-bit and 64-bit integers.")
    t = np.uint64 if pack_size == 64 else np.uint32
    a = np.asarray(a)
    n, k = a.shape
    out = np.zeros((int(ceil(1.0 * n / pack_size)), k), dtype=t)
    for i in range(n):
        out[i // pack_size] |= a[i].astype(t) << t(pack_size - 1 - (i % pack_size))
    return out


def unpack_binary_bytes_from_ints(a):
    """utils.py:159-187"""
    if a.dtype == np.uint32:
        pack_size = 32
    elif a.dtype == np.uint64:
        pack_size = 64
    else:
        raise ValueError("Supported data types are 32-bit and 64-bit integers.")
    t = a.dtype.type
    n_rows = a.shape[0] * pack_size
    out = np.zeros((n_rows, a.shape[1]) if a.ndim > 1 else n_rows, dtype=np.uint8)
    for i in range(n_rows):
        w = i // pack_size
        bit = t(pack_size - 1 - (i - pack_size * w))
        out[i] = ((a[w] >> bit) & t(1)) > 0
    return out


def build_row_mask(example_idx, n_examples, mask_n_bits):
    """rules.py:210-222"""
    if mask_n_bits not in (8, 16, 32, 64, 128):
        raise ValueError("Unsupported mask format. Use 8, 16, 32, 64 or 128 bits.")
    n_masks = int(ceil(float(n_examples) / mask_n_bits))
    masks = [0] * n_masks
    for idx in example_idx:
        idx = int(idx)
        w = idx // mask_n_bits
        masks[w] |= 1 << (mask_n_bits - (idx - mask_n_bits * w) - 1)
    return np.array(masks, dtype="u%d" % (mask_n_bits // 8))


# --------------------------------------------------------------------------------------
# KmerRuleClassifications -- learning/common/rules.py:99-267
# --------------------------------------------------------------------------------------
class OracleKmerRuleClassifications(object):
    """``dataset`` is anything with shape / dtype / chunks / numpy-style __getitem__
    (an ndarray works; ``chunks`` is read with getattr)."""

    def __init__(self, dataset, n_rows, block_size=None, fast=False):
        self.dataset = dataset
        self.dataset_initial_n_rows = n_rows
        self.dataset_n_rows = n_rows
        self.dataset_removed_rows = []
        self.dataset_removed_rows_mask = np.zeros(n_rows, dtype=bool)
        self.fast = fast            # True: one C pass instead of the numpy block loop (same result)
        chunks = getattr(dataset, "chunks", None)
        if block_size is None:                                   # rules.py:112-116
            self.block_size = (1, dataset.shape[1]) if chunks is None else chunks
        else:                                                    # rules.py:118-120
            if len(block_size) != 2 or not isinstance(block_size[0], int) or not isinstance(block_size[1], int):
                raise ValueError("The block size must be a tuple of 2 integers.")
            self.block_size = block_size
        if dataset.dtype == np.uint32:                           # rules.py:123-131
            self.dataset_pack_size = 32
            self.inplace_popcount = inplace_popcount_32
        elif dataset.dtype == np.uint64:
            self.dataset_pack_size = 64
            self.inplace_popcount = inplace_popcount_64
        else:
            raise ValueError("Unsupported data type for packed attribute classifications array. The supported data"
                             " types are np.uint32 and np.uint64.")

    @property
    def shape(self):                                             # rules.py:196-198
        return self.dataset_n_rows, self.dataset.shape[1] * 2

    def _relative_to_dataset_rows(self, rows):
        """rules.py:227-236 / 177-184: the (row_idx+1)-th row that has not been removed."""
        alive = np.flatnonzero(~self.dataset_removed_rows_mask)
        return [int(alive[int(r)]) for r in rows]

    def remove_rows(self, rows):                                 # rules.py:173-194
        newly = self._relative_to_dataset_rows(rows)
        if len(newly) > 0:
            self.dataset_removed_rows = sorted(set(self.dataset_removed_rows + newly))
            self.dataset_removed_rows_mask = np.zeros(self.dataset_initial_n_rows, dtype=bool)
            self.dataset_removed_rows_mask[self.dataset_removed_rows] = True
            self.dataset_n_rows = self.dataset_initial_n_rows - len(self.dataset_removed_rows)

    def get_columns(self, columns):                              # rules.py:135-171
        is_int = False
        if hasattr(columns, "__index__"):
            columns = [columns.__index__()]
            is_int = True
        elif isinstance(columns, np.ndarray):
            columns = columns.tolist()
        elif not isinstance(columns, list):
            columns = list(columns)
        n_kmers = self.dataset.shape[1]
        invert = np.array([c >= n_kmers for c in columns])
        cols = [c if c < n_kmers else c % n_kmers for c in columns]
        keep = np.ones(self.dataset.shape[0] * self.dataset_pack_size, dtype=bool)
        keep[self.dataset_initial_n_rows:] = False
        keep[self.dataset_removed_rows] = False
        uniq, inverse = np.unique(cols, return_inverse=True)
        result = unpack_binary_bytes_from_ints(np.asarray(self.dataset[:, uniq.tolist()]))[keep]
        result = result[:, inverse]
        result[:, invert] = 1 - result[:, invert]
        return result.reshape(-1) if is_int else result

    def sum_rows(self, rows):                                    # rules.py:201-267
        rows = np.asarray(rows)
        n_kmers = self.dataset.shape[1]
        result = np.zeros(n_kmers * 2, dtype=minimum_uint_size(rows.shape[0]))
        rows = np.sort(rows)
        row_mask = build_row_mask(self._relative_to_dataset_rows(rows), self.dataset_initial_n_rows,
                                  self.dataset_pack_size)
        rows_to_load = np.where(row_mask != 0)[0]
        if self.fast and self.dataset_pack_size == 64 and isinstance(self.dataset, np.ndarray) \
                and self.dataset.flags.c_contiguous:
            acc = np.zeros(n_kmers, dtype=np.uint64)
            _clib().oracle_masked_column_counts_64(self.dataset.ctypes.data, self.dataset.shape[0], n_kmers,
                                                   n_kmers, row_mask.ctypes.data, acc.ctypes.data)
            result[:n_kmers] = acc
        else:
            br, bc = self.block_size
            n_col_blocks = int(ceil(1.0 * n_kmers / bc))
            n_row_blocks = int(ceil(1.0 * len(rows_to_load) / br))
            for rb in range(n_row_blocks):
                sel = rows_to_load[rb * br:(rb + 1) * br]
                block_mask = row_mask[sel]
                for cb in range(n_col_blocks):
                    # fancy row index => a fresh C-contiguous copy, as an h5py read gives (rules.py:253)
                    block = np.ascontiguousarray(self.dataset[sel, cb * bc:(cb + 1) * bc])
                    if block.ndim == 1:
                        block = block.reshape(1, -1)
                    self.inplace_popcount(block, block_mask)                                  # rules.py:259
                    result[cb * bc:min((cb + 1) * bc, n_kmers)] += np.sum(block, axis=0).astype(result.dtype)
        result[n_kmers:] = len(rows) - result[:n_kmers]          # rules.py:265
        return result


# --------------------------------------------------------------------------------------
# SCM utility scan -- learning/learners/scm.py:238-288
# --------------------------------------------------------------------------------------
def best_utility_rules(rule_classifications, p, positive_example_idx, negative_example_idx, rule_blacklist=[]):
    n_rules = rule_classifications.shape[1]
    blacklisted = np.zeros(n_rules, dtype=bool)
    blacklisted[rule_blacklist] = True

    neg_cover = negative_example_idx.shape[0] - rule_classifications.sum_rows(negative_example_idx)   # :245
    if positive_example_idx.shape[0] > 0:                                                             # :250-253
        pos_err = positive_example_idx.shape[0] - rule_classifications.sum_rows(positive_example_idx)
    else:
        pos_err = np.zeros(n_rules, dtype=neg_cover.dtype)
    return merge_utility_blocks(neg_cover, pos_err, float(p), blacklisted)


def merge_utility_blocks(neg_cover, pos_err, p, blacklisted, block_size=UTIL_BLOCK_SIZE):
    """scm.py:258-288 -- per-block utilities, block max, isclose ties, sequential merge."""
    n_rules = neg_cover.shape[0]
    best = -np.inf
    best_idx = np.array([])
    best_pos = np.array([])
    best_neg = np.array([])
    for b in range(int(ceil(1.0 * n_rules / block_size))):
        sl = slice(b * block_size, (b + 1) * block_size)
        util = neg_cover[sl] - p * pos_err[sl]            # float64: one rounded multiply, one rounded subtract
        util[blacklisted[sl]] = -np.inf
        m = np.max(util)
        if m > best or np.allclose(best, m):
            ties = np.where(np.isclose(util, m))[0] + b * block_size
            if np.allclose(m, best):
                best_idx = np.hstack((best_idx, ties))
                best_pos = np.hstack((best_pos, pos_err[ties]))
                best_neg = np.hstack((best_neg, neg_cover[ties]))
            else:
                best = m
                best_idx = ties
                best_pos = pos_err[ties]
                best_neg = neg_cover[ties]
    return best, best_idx, best_pos, best_neg


def scm_tiebreaker(best_utility_idx, rule_risks, model_type):
    """experiments/experiment_scm.py:122-130 (== 293-301, 476-484)"""
    tie = rule_risks[best_utility_idx]
    if model_type == "conjunction":
        return best_utility_idx[np.isclose(tie, tie.min())]
    return best_utility_idx[np.isclose(tie, tie.max())]


def compute_rule_importances(rule_classifications, model_rules_idx, training_example_idx):
    """learners/scm.py:32-36"""
    cls = rule_classifications.get_columns(model_rules_idx)[training_example_idx]
    neg_pred = np.where(np.prod(cls, axis=1) == 0)[0]
    return (float(len(neg_pred)) - cls[neg_pred].sum(axis=0)) / len(neg_pred)


def scm_fit(rule_classifications, model_type, p, max_rules, positive_example_idx, negative_example_idx,
            rule_blacklist=[], tiebreaker=None, iteration_rule_importances=False):
    """learners/scm.py:54-159.  Returns (model_rule_idx, per-iteration infos, rule_importances).
    ``selected_rule`` in each info is (kmer_column, 'presence'|'absence') as the model stores it
    (disjunction appends the inverse rule, scm.py:180-184)."""
    if model_type not in ("conjunction", "disjunction"):
        raise ValueError("Unsupported model type.")
    pos = np.asarray(positive_example_idx)
    neg = np.asarray(negative_example_idx)
    if len(pos) == 0 or len(neg) == 0:
        raise ValueError("There must be positive and negative examples to train the SCM.")
    if model_type == "disjunction":                                   # :69-73
        pos, neg = neg, pos
    n_rules = rule_classifications.shape[1]
    n_kmers = n_rules // 2
    if len(rule_blacklist) > 0:                                       # :80-83
        rule_blacklist = np.unique(rule_blacklist)
        if len(rule_blacklist) == n_rules:
            raise ValueError("The blacklist cannot include all the rules.")
    training_example_idx = np.hstack((pos, neg))
    model_rules_idx, infos, importances = [], [], []
    while len(neg) > 0 and len(model_rules_idx) < max_rules:          # :88
        info = {"iteration_number": len(model_rules_idx) + 1}
        best_u, best_idx, best_pos, best_neg = best_utility_rules(rule_classifications, p, pos, neg, rule_blacklist)
        info["utility_max"] = best_u
        info["utility_argmax"] = best_idx
        cand = best_idx[np.logical_or(best_neg != 0, best_pos != 0)]  # :109
        if len(cand) == 0:
            break
        if len(cand) == 1:                                            # :117-124
            chosen = cand[0]
            info["equivalent_rules_idx"] = np.array([chosen])
        else:
            equiv = tiebreaker(cand)
            info["equivalent_rules_idx"] = equiv
            chosen = equiv[0]
        chosen = int(chosen)
        kind = "presence" if chosen < n_kmers else "absence"
        if model_type == "disjunction":
            kind = "absence" if kind == "presence" else "presence"
        info["selected_rule"] = (chosen % n_kmers, kind)
        info["selected_rule_idx"] = chosen
        model_rules_idx.append(chosen)
        cls = rule_classifications.get_columns(chosen)                # :133
        neg = neg[cls[neg] != 0]                                      # :137
        pos = pos[cls[pos] != 0]                                      # :139
        if iteration_rule_importances:                                # :144-147
            importances = compute_rule_importances(rule_classifications, model_rules_idx, training_example_idx)
            info["rule_importances"] = importances
        infos.append(info)
    if len(model_rules_idx) > 0:                                      # :153-159
        if not iteration_rule_importances:
            importances = compute_rule_importances(rule_classifications, model_rules_idx, training_example_idx)
    else:
        importances = []
    return model_rules_idx, infos, importances


# --------------------------------------------------------------------------------------
# CART split scan -- learning/learners/cart.py:54-346, common/tree.py:28-46
# --------------------------------------------------------------------------------------
def cart_altered_priors(example_idx, class_importance):
    """cart.py:71-77"""
    n_total = {c: float(len(idx)) for c, idx in example_idx.items()}
    priors = {c: 1.0 * n / sum(n_total.values()) for c, n in n_total.items()}
    denum = sum([class_importance[c] * priors[c] for c in priors.keys()])
    altered = {c: 1.0 * class_importance[c] * priors[c] / denum for c in priors.keys()}
    return n_total, altered


def gini_impurity(n_by_class, altered_priors, n_total, multiply_by_node_proba=False):
    """cart.py:85-110"""
    p_j_t = {c: 1.0 * altered_priors[c] * n_by_class[c] / n_total[c] for c in n_by_class.keys()}
    p_t = sum(p_j_t.values())
    with np.errstate(divide="ignore", invalid="ignore"):
        p_j_given_t = {c: np.divide(p_j_t[c], p_t) for c in p_j_t.keys()}
    div = sum([p_j_given_t[i] * p_j_given_t[j] for i in p_j_given_t.keys() for j in p_j_given_t.keys() if i != j])
    return div * (p_t if multiply_by_node_proba else 1.0)


def gini_rule_score(rule_classifications, example_idx, altered_priors, n_total):
    """cart.py:112-161"""
    n_kmers = rule_classifications.shape[1] // 2
    left = {c: np.asarray(rule_classifications.sum_rows(example_idx[c])[:n_kmers], dtype=float)
            for c in example_idx.keys()}
    right = {c: np.asarray(len(example_idx[c]) - left[c], dtype=float) for c in left.keys()}
    gini = np.zeros(n_kmers)
    n_blocks = int(ceil(1.0 * n_kmers / GINI_BLOCK_SIZE))
    for i in range(n_blocks):
        sl = slice(i * GINI_BLOCK_SIZE, (i + 1) * GINI_BLOCK_SIZE)
        gini[sl] = gini_impurity({c: v[sl] for c, v in left.items()}, altered_priors, n_total, True)
    for i in range(n_blocks):
        sl = slice(i * GINI_BLOCK_SIZE, (i + 1) * GINI_BLOCK_SIZE)
        gini[sl] += gini_impurity({c: v[sl] for c, v in right.items()}, altered_priors, n_total, True)
    gini[sum(left.values()) == 0] = np.inf
    gini[sum(right.values()) == 0] = np.inf
    return gini


def cross_entropy(n_by_class, altered_priors, n_total, multiply_by_node_proba=False):
    """cart.py:167-178"""
    p_j_t = {c: 1.0 * altered_priors[c] * n_by_class[c] / n_total[c] for c in n_by_class.keys()}
    p_t = sum(p_j_t.values())
    with np.errstate(divide="ignore", invalid="ignore"):
        p_j_given_t = {c: np.divide(p_j_t[c], p_t) for c in p_j_t.keys()}
        div = -1.0 * sum([np.nan_to_num(p_j_given_t[c] * np.log(p_j_given_t[c])) for c in p_j_given_t.keys()])
    return div * (p_t if multiply_by_node_proba else 1.0)


def cross_entropy_rule_score(rule_classifications, example_idx, altered_priors, n_total):
    """cart.py:180-207 (classes without examples in the node are left out, :197-198; no 100k blocking here)"""
    n_kmers = rule_classifications.shape[1] // 2
    left = {c: np.asarray(rule_classifications.sum_rows(example_idx[c])[:n_kmers], dtype=float)
            for c in example_idx.keys() if example_idx[c].size}
    right = {c: np.asarray(example_idx[c].shape[0] - left[c], dtype=float) for c in left.keys()}
    score = cross_entropy(left, altered_priors, n_total, True)
    score += cross_entropy(right, altered_priors, n_total, True)
    score[sum(left.values()) == 0] = np.inf
    score[sum(right.values()) == 0] = np.inf
    return score


def cart_tiebreaker(best_score_idx, rule_kmer_occurrences):
    """experiments/experiment_cart.py:82-94"""
    occ = rule_kmer_occurrences[best_score_idx]
    return best_score_idx[np.isclose(occ, occ.max())]


class OracleTreeNode(object):
    """The fields of common/tree.py:46-75 that the split scan reads or writes."""

    def __init__(self, depth, class_examples_idx, criterion_value, parent=None):
        self.depth = depth
        self.class_examples_idx = class_examples_idx
        self.criterion_value = criterion_value
        self.parent = parent
        self.rule_idx = None
        self.equivalent_rules_idx = None
        self.left_child = None
        self.right_child = None

    @property
    def n_examples(self):                                             # tree.py:84-90
        return sum(len(v) for v in self.class_examples_idx.values())

    @property
    def class_proportions(self):                                      # tree.py:92-99
        n = self.n_examples
        return {c: float(len(v)) / n for c, v in self.class_examples_idx.items()}

    @property
    def is_leaf(self):
        return self.rule_idx is None

    def to_dict(self):
        d = {"depth": self.depth, "criterion_value": float(self.criterion_value),
             "n_by_class": {int(c): int(len(v)) for c, v in self.class_examples_idx.items()}}
        if not self.is_leaf:
            d["rule_idx"] = int(self.rule_idx)
            d["equivalent_rules_idx"] = [int(x) for x in self.equivalent_rules_idx]
            d["left"] = self.left_child.to_dict()
            d["right"] = self.right_child.to_dict()
        return d


def cart_fit(rule_classifications, example_idx, criterion="gini", max_depth=10, min_samples_split=2,
             class_importance=None, rule_blacklist=None, tiebreaker=None):
    """cart.py:54-346, both criteria ("cross-entropy" is unreachable from the reference CLI, SURVEY section 5, and
    its scores go through np.log, whose last bit depends on the libm / SIMD dispatch of the machine).
    Returns the root OracleTreeNode."""
    if criterion not in ("gini", "cross-entropy"):
        raise ValueError("The supporting splitting criteria are: ['gini', 'cross-entropy'].")
    rule_score = gini_rule_score if criterion == "gini" else cross_entropy_rule_score
    node_impurity = gini_impurity if criterion == "gini" else cross_entropy
    if max_depth < 1:
        raise ValueError("The maximum tree depth must be greater than 1.")
    if min_samples_split < 2.0:
        raise ValueError("The minimum number of examples used to split a node must be 2 or greater.")
    min_samples_split = int(min_samples_split)
    if tiebreaker is None:
        tiebreaker = lambda x: x
    if rule_blacklist is None:
        rule_blacklist = []
    if class_importance is None:
        class_importance = {c: 1.0 for c in example_idx.keys()}
    n_total, altered = cart_altered_priors(example_idx, class_importance)

    def find_best_split(node):                                        # cart.py:219-250
        ex = node.class_examples_idx
        score = rule_score(rule_classifications, ex, altered, n_total)
        score[rule_blacklist] = np.inf                                # :231
        smin = min(score)                                             # builtin min, :234/:238
        if smin == np.inf:
            return None
        cand = np.where(score == smin)[0]
        best = tiebreaker(cand)
        sel = best[0]
        preds = rule_classifications.get_columns(sel)
        left = {c: ex[c][preds[ex[c]] == 1] for c in ex.keys()}       # :247
        right = {c: ex[c][preds[ex[c]] == 0] for c in ex.keys()}      # :248
        return sel, best, left, right

    root = OracleTreeNode(0, example_idx, node_impurity(n_total, altered, n_total))
    queue = deque([root])
    current_depth = -1
    while len(queue) > 0:                                             # cart.py:267-339
        node = queue.popleft()
        if node.depth != current_depth:
            current_depth = node.depth
            if current_depth == max_depth:
                break
        if 1.0 in node.class_proportions.values():
            continue
        if node.n_examples < min_samples_split:
            continue
        res = find_best_split(node)
        if res is None:
            continue
        sel, equiv, left, right = res
        node.rule_idx = sel
        node.equivalent_rules_idx = equiv
        node.left_child = OracleTreeNode(node.depth + 1, left,
                                         node_impurity({c: len(v) for c, v in left.items()}, altered, n_total), node)
        node.right_child = OracleTreeNode(node.depth + 1, right,
                                          node_impurity({c: len(v) for c, v in right.items()}, altered, n_total), node)
        queue.append(node.left_child)
        queue.append(node.right_child)
    return root


# --------------------------------------------------------------------------------------
# dataset split: individual k-mer risk tables -- dataset/split.py:171-188 (== 213-228 per fold)
# --------------------------------------------------------------------------------------
def kmer_risk_tables(rule_classifications, train_pos_idx, train_neg_idx):
    """-> (unique_risks, unique_risk_by_kmer, unique_risk_by_anti_kmer) exactly as split.py stores them."""
    n_kmers = rule_classifications.shape[1] // 2
    n_train = len(train_pos_idx) + len(train_neg_idx)
    kmer_risks = (len(train_pos_idx) - rule_classifications.sum_rows(train_pos_idx)[:n_kmers]).astype(float)   # :178
    kmer_risks += rule_classifications.sum_rows(train_neg_idx)[:n_kmers]                                       # :179
    kmer_risks /= n_train                                                                                      # :180
    np.round(kmer_risks, 5, out=kmer_risks)
    anti_kmer_risks = 1.0 - kmer_risks
    np.round(anti_kmer_risks, 5, out=anti_kmer_risks)
    unique_risks, inverse = np.unique(np.hstack((kmer_risks, anti_kmer_risks)), return_inverse=True)           # :184
    t = minimum_uint_size(len(unique_risks))
    return unique_risks, inverse[:n_kmers].astype(t), inverse[n_kmers:].astype(t)
