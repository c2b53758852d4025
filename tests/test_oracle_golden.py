"""CPU: the oracle (oracle/kover_oracle.py + popcount_oracle.c) against the golden vectors minted
from the REAL reference (tests/golden/make_golden.py).  Bit-exact everywhere; float64 outputs are
compared with == (the oracle must reproduce the reference's operation order)."""
import json
import os

import numpy as np
import pytest

import cases
from oracle import kover_oracle as ko


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


class Chunked(np.ndarray):
    chunks = None


def as_dataset(P, chunks):
    d = P.view(Chunked)
    d.chunks = chunks
    return d


def test_popcount_kernel(golden_dir):
    g = load(golden_dir, "popcount")
    a = g["a64"].copy()
    ko.inplace_popcount_64(a, g["m64"])
    assert np.array_equal(a, g["o64"])
    b = g["a32"].copy()
    ko.inplace_popcount_32(b, g["m32"])
    assert np.array_equal(b, g["o32"])


def test_pack_unpack(golden_dir):
    g = load(golden_dir, "pack")
    assert np.array_equal(ko.pack_binary_bytes_to_ints(g["X"], 64), g["p64"])
    assert np.array_equal(ko.pack_binary_bytes_to_ints(g["X"], 32), g["p32"])
    assert np.array_equal(cases.pack64(g["X"]), g["p64"])
    assert np.array_equal(cases.pack32_from64(g["p64"]), np.vstack([g["p32"], np.zeros((1, 9), np.uint32)])[:6])
    assert np.array_equal(ko.unpack_binary_bytes_from_ints(g["p64"]), g["u64"])
    assert np.array_equal(ko.unpack_binary_bytes_from_ints(g["p32"]), g["u32"])
    sizes = [np.dtype(ko.minimum_uint_size(v)).itemsize for v in (0, 255, 256, 65535, 65536, 2 ** 32)]
    assert sizes == g["min_uint_itemsize"].tolist()


@pytest.mark.parametrize("mname", ["M1", "M2", "M3", "M4"])
@pytest.mark.parametrize("fast", [False, True])
def test_sum_rows_get_columns(golden_dir, mname, fast):
    g = load(golden_dir, "rules_" + mname)
    X, P, n, chunks = cases.matrix(mname)
    k = P.shape[1]
    rc = ko.OracleKmerRuleClassifications(as_dataset(P, chunks), n, fast=fast)
    assert list(rc.shape) == g["shape"].tolist()
    for rname, rows in cases.row_sets(n).items():
        got = rc.sum_rows(rows)
        want = g["sum_" + rname]
        assert got.dtype == want.dtype, rname
        assert np.array_equal(got, want), rname
    for cname, cols in cases.column_sets(k).items():
        got = rc.get_columns(cols)
        want = g["col_" + cname]
        assert got.shape == want.shape and got.dtype == want.dtype and np.array_equal(got, want), cname
    if mname == "M3":
        rc32 = ko.OracleKmerRuleClassifications(as_dataset(cases.pack32_from64(P), chunks), n)
        for rname, rows in cases.row_sets(n).items():
            assert np.array_equal(rc32.sum_rows(rows), g["sum32_" + rname]), rname
        assert np.array_equal(rc32.get_columns(cases.column_sets(k)["mixed"]), g["col32_mixed"])
    if mname == "M4":
        rc2 = ko.OracleKmerRuleClassifications(as_dataset(P, chunks), n, fast=fast)
        rc2.remove_rows([3, 10])
        rc2.remove_rows([0])
        assert list(rc2.shape) == g["rm_shape"].tolist()
        assert np.array_equal(rc2.sum_rows(np.array([0, 1, 2, 9, 40, 61])), g["rm_sum"])
        assert np.array_equal(rc2.get_columns([1, k + 2]), g["rm_col"])


def test_get_columns_rejects_multi_element_ndarray():
    X, P, n, chunks = cases.matrix("M4")
    rc = ko.OracleKmerRuleClassifications(as_dataset(P, chunks), n)
    with pytest.raises(TypeError):
        rc.get_columns(np.array([1, 2]))


def test_constructor_errors():
    X, P, n, chunks = cases.matrix("M4")
    with pytest.raises(ValueError):
        ko.OracleKmerRuleClassifications(P.astype(np.uint16), n)
    with pytest.raises(ValueError):
        ko.OracleKmerRuleClassifications(P, n, block_size=(1, 2, 3))
    with pytest.raises(ValueError):
        ko.OracleKmerRuleClassifications(P, n, block_size=(1.0, 2))


@pytest.mark.parametrize("case", cases.UTILITY_CASES, ids=[c[0] for c in cases.UTILITY_CASES])
def test_best_utility_rules(golden_dir, case):
    g = load(golden_dir, case[0])
    P, n, chunks, pos, neg, p, blacklist, ubs = cases.utility_inputs(case)
    rc = ko.OracleKmerRuleClassifications(as_dataset(P, chunks), n, fast=True)
    old = ko.UTIL_BLOCK_SIZE
    try:
        n_rules = rc.shape[1]
        bl = np.zeros(n_rules, dtype=bool)
        bl[blacklist] = True
        neg_cover = neg.shape[0] - rc.sum_rows(neg)
        pos_err = pos.shape[0] - rc.sum_rows(pos)
        bu, bi, bp, bn = ko.merge_utility_blocks(neg_cover, pos_err, float(p), bl, block_size=ubs)
    finally:
        ko.UTIL_BLOCK_SIZE = old
    assert bu == float(g["best_utility"])
    assert np.array_equal(np.asarray(bi, dtype=np.int64), g["best_idx"])
    assert np.array_equal(np.asarray(bp, dtype=np.int64), g["best_pos_err"])
    assert np.array_equal(np.asarray(bn, dtype=np.int64), g["best_neg_cover"])


@pytest.mark.parametrize("case", cases.SCM_CASES, ids=[c[0] for c in cases.SCM_CASES])
def test_scm_fit(golden_dir, case, monkeypatch):
    name, mname, lname, model_type, p, max_rules, tb, blacklist, ubs = case
    g = load(golden_dir, name)
    X, P, n, chunks = cases.matrix(mname)
    y = cases.labels(lname)
    k = P.shape[1]
    monkeypatch.setattr(ko, "UTIL_BLOCK_SIZE", ubs)
    monkeypatch.setattr(ko.merge_utility_blocks, "__defaults__", (ubs,))
    rc = ko.OracleKmerRuleClassifications(as_dataset(P, chunks), n, fast=True)
    risks = cases.rule_risks(2 * k)
    tiebreaker = (lambda idx: idx) if tb == "identity" else (lambda idx: ko.scm_tiebreaker(idx, risks, model_type))
    model, infos, imp = ko.scm_fit(rc, model_type, p, max_rules, np.where(y == 1)[0], np.where(y == 0)[0],
                                   rule_blacklist=blacklist, tiebreaker=tiebreaker, iteration_rule_importances=True)
    assert len(infos) == int(g["n_iterations"])
    for t, info in enumerate(infos):
        assert info["utility_max"] == float(g["it%d_utility_max" % t])
        assert np.array_equal(np.asarray(info["utility_argmax"], dtype=np.int64), g["it%d_utility_argmax" % t])
        assert np.array_equal(np.asarray(info["equivalent_rules_idx"], dtype=np.int64), g["it%d_equivalent" % t])
        col, kind = info["selected_rule"]
        assert "%s(K%d)" % (kind.capitalize(), col) == str(g["it%d_selected" % t])
        assert np.array_equal(info["rule_importances"], g["it%d_importances" % t])
    assert np.array_equal(np.asarray(imp, dtype=np.float64), g["rule_importances"])


def tree_equal(got, want, path="root", rtol=0.0):
    """rtol = 0: bit-exact criterion values (Gini).  rtol > 0 only for the cross-entropy criterion (np.log)."""
    assert got["depth"] == want["depth"], path
    assert got["n_by_class"] == {int(c): v for c, v in want["n_by_class"].items()}, path
    if rtol:
        assert abs(got["criterion_value"] - want["criterion_value"]) <= rtol * abs(want["criterion_value"]), path
    else:
        assert got["criterion_value"] == want["criterion_value"], path
    assert ("rule_idx" in got) == ("rule_idx" in want), path
    if "rule_idx" in want:
        assert got["rule_idx"] == want["rule_idx"], path
        assert got["equivalent_rules_idx"] == want["equivalent_rules_idx"], path
        tree_equal(got["left"], want["left"], path + ".L", rtol)
        tree_equal(got["right"], want["right"], path + ".R", rtol)


@pytest.mark.parametrize("case", cases.CART_CASES, ids=[c[0] for c in cases.CART_CASES])
def test_cart_fit(golden_dir, case):
    name, mname, lname, max_depth, mss, class_importance, tb, subset_seed = case
    g = load(golden_dir, name)
    want = json.loads(str(g["tree_json"]))
    X, P, n, chunks = cases.matrix(mname)
    y = cases.labels(lname)
    rc = ko.OracleKmerRuleClassifications(as_dataset(P, chunks), n, fast=True)
    train = cases.train_subset(n, subset_seed)
    example_idx = {c: train[y[train] == c] for c in sorted(class_importance.keys())}
    tiebreaker = None
    if tb == "occurrence":
        occ = rc.sum_rows(train)
        tiebreaker = lambda idx: ko.cart_tiebreaker(idx, occ)
    root = ko.cart_fit(rc, example_idx, "gini", max_depth, mss, class_importance, [], tiebreaker)
    tree_equal(root.to_dict(), want)


def ce_case_inputs(case):
    name, mname, lname, max_depth, mss, class_importance, tb, subset_seed = case
    X, P, n, chunks = cases.matrix(mname)
    y = cases.labels(lname)
    train = cases.train_subset(n, subset_seed)
    example_idx = {c: train[y[train] == c] for c in sorted(class_importance.keys())}
    return P, n, chunks, train, example_idx


def searched_nodes(tree):
    """The nodes whose split search produced a score vector, in the learner's breadth-first order (cart.py:267-339):
    every internal node, plus the leaves that were searched without finding a split -- none in these cases."""
    out, queue = [], [tree]
    while queue:
        node = queue.pop(0)
        if "rule_idx" in node:
            out.append(node)
            queue += [node["left"], node["right"]]
    return out


@pytest.mark.parametrize("case", cases.CART_CE_CASES, ids=[c[0] for c in cases.CART_CE_CASES])
def test_cart_cross_entropy_fit(golden_dir, case):
    """a10: the reference's cross-entropy criterion (cart.py:167-207).  Trees, tie sets (equivalent rules) and
    selected rules must be identical; the raw score vectors go through np.log, whose last bit depends on the
    machine's libm / SIMD dispatch, so they are compared to 4 ulp (rtol 1e-15) -- stated tolerance, floating point."""
    name, mname, lname, max_depth, mss, class_importance, tb, subset_seed = case
    g = load(golden_dir, name)
    want = json.loads(str(g["tree_json"]))
    P, n, chunks, train, example_idx = ce_case_inputs(case)
    rc = ko.OracleKmerRuleClassifications(as_dataset(P, chunks), n, fast=True)
    tiebreaker = None
    if tb == "occurrence":
        occ = rc.sum_rows(train)
        tiebreaker = lambda idx: ko.cart_tiebreaker(idx, occ)
    root = ko.cart_fit(rc, example_idx, "cross-entropy", max_depth, mss, class_importance, [], tiebreaker)
    tree_equal(root.to_dict(), want, rtol=1e-15)
    # the score vector of the root and of every other searched node against the reference's own
    n_total, altered = ko.cart_altered_priors(example_idx, class_importance)
    score = ko.cross_entropy_rule_score(rc, example_idx, altered, n_total)
    ref_scores = g["node_scores"]
    assert ref_scores.shape[0] >= len(searched_nodes(want))
    np.testing.assert_allclose(score, ref_scores[0], rtol=1e-15, atol=0)
    assert np.array_equal(np.where(score == score.min())[0], np.where(ref_scores[0] == ref_scores[0].min())[0])


@pytest.mark.parametrize("case", cases.RISK_CASES, ids=[c[0] for c in cases.RISK_CASES])
def test_kmer_risk_tables(golden_dir, case):
    name, mname, lname, seed = case
    g = load(golden_dir, name)
    X, P, n, chunks = cases.matrix(mname)
    y = cases.labels(lname)
    train = cases.train_subset(n, seed)
    rc = ko.OracleKmerRuleClassifications(as_dataset(P, chunks), n, fast=True)
    ur, bk, ba = ko.kmer_risk_tables(rc, train[y[train] == 1], train[y[train] == 0])
    assert np.array_equal(ur, g["unique_risks"])
    assert bk.dtype == g["by_kmer"].dtype and np.array_equal(bk, g["by_kmer"])
    assert ba.dtype == g["by_anti_kmer"].dtype and np.array_equal(ba, g["by_anti_kmer"])
