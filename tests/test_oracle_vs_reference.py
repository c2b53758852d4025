"""CPU, build container only (skipped where /root/reference is absent): the checker (oracle/) against the REAL
reference, loaded in memory by oracle/ref_loader.py, on fresh random inputs -- beyond the committed golden
vectors.  This is what pins the oracle: same counts, same utilities (==), same tie sets, same fits, same trees."""
import numpy as np
import pytest

import cases
from oracle import kover_oracle as ko
from oracle import ref_loader

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_loader.reference_available(), reason="needs /root/reference")]


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load_reference()


def random_case(seed):
    rs = np.random.RandomState(5000 + seed)
    n = int(rs.choice([20, 64, 65, 130, 257, 300]))
    k = int(rs.choice([50, 333, 1000, 2500]))
    dens = rs.choice([0.05, 0.3, 0.5, 0.8, 1.0], size=k)
    X = (rs.rand(n, k) < dens[None, :]).astype(np.uint8)
    if rs.rand() < 0.5:                         # duplicate columns -> equivalent rules
        src = rs.randint(0, k, size=k // 3)
        X[:, rs.randint(0, k, size=k // 3)] = X[:, src]
    P = cases.pack64(X)
    y = (rs.rand(n) < 0.45).astype(np.uint8)
    chunks = (1, int(rs.choice([17, 100, 1000])))
    return rs, X, P, n, k, y, chunks


@pytest.mark.parametrize("seed", range(12))
def test_rules_and_utility_scan(ref, seed):
    rs, X, P, n, k, y, chunks = random_case(seed)
    rrc = ref.rules.KmerRuleClassifications(ref_loader.ChunkedArray(P, chunks), n)
    orc = ko.OracleKmerRuleClassifications(ref_loader.ChunkedArray(P, chunks), n)
    rows = np.sort(rs.choice(n, max(1, n // 2), replace=False))
    a, b = rrc.sum_rows(rows), orc.sum_rows(rows)
    assert a.dtype == b.dtype and np.array_equal(a, b)
    cols = sorted(set(int(c) for c in rs.randint(0, 2 * k, size=4)))
    assert np.array_equal(rrc.get_columns(cols), orc.get_columns(cols))
    pos, neg = np.where(y == 1)[0], np.where(y == 0)[0]
    p = float(rs.choice([0.1, 0.316, 1.0, 1.778, 10.0, 999999.0]))
    ubs = int(rs.choice([37, 500, 1000000]))
    blacklist = np.unique(rs.randint(0, 2 * k, size=rs.randint(0, 30))).tolist()
    ref.scm.UTIL_BLOCK_SIZE = ubs
    try:
        m = ref.scm.SetCoveringMachine(model_type=ref.models.conjunction, p=p, max_rules=5)
        ru, ri, rp, rn = m._get_best_utility_rules(rrc, pos, neg, rule_blacklist=blacklist)
    finally:
        ref.scm.UTIL_BLOCK_SIZE = 1000000
    bl = np.zeros(2 * k, dtype=bool)
    bl[blacklist] = True
    ou, oi, op, on = ko.merge_utility_blocks(len(neg) - orc.sum_rows(neg), len(pos) - orc.sum_rows(pos), p, bl,
                                             block_size=ubs)
    assert ru == ou and np.array_equal(np.asarray(ri, dtype=np.int64), np.asarray(oi, dtype=np.int64))
    assert np.array_equal(np.asarray(rp, dtype=np.int64), np.asarray(op, dtype=np.int64))
    assert np.array_equal(np.asarray(rn, dtype=np.int64), np.asarray(on, dtype=np.int64))


@pytest.mark.parametrize("seed", range(8))
def test_scm_fit(ref, seed):
    rs, X, P, n, k, y, chunks = random_case(100 + seed)
    model_type = "conjunction" if seed % 2 == 0 else "disjunction"
    p = float(rs.choice([0.5, 1.0, 2.0, 999999.0]))
    risks = cases.rule_risks(2 * k, seed=seed)
    tb = lambda idx: ko.scm_tiebreaker(idx, risks, model_type)
    rules = ref.rules.LazyKmerRuleList(["K%d" % i for i in range(k)], np.arange(k))
    infos = []
    m = ref.scm.SetCoveringMachine(model_type=model_type, p=p, max_rules=4)
    m.fit(rules, ref.rules.KmerRuleClassifications(ref_loader.ChunkedArray(P, chunks), n), np.where(y == 1)[0],
          np.where(y == 0)[0], tiebreaker=tb, iteration_callback=lambda i: infos.append(dict(i)))
    model, oinfos, imp = ko.scm_fit(ko.OracleKmerRuleClassifications(ref_loader.ChunkedArray(P, chunks), n),
                                    model_type, p, 4, np.where(y == 1)[0], np.where(y == 0)[0], tiebreaker=tb)
    assert len(infos) == len(oinfos)
    for a, b in zip(infos, oinfos):
        assert a["utility_max"] == b["utility_max"]
        assert np.array_equal(np.asarray(a["utility_argmax"], dtype=np.int64), np.asarray(b["utility_argmax"], dtype=np.int64))
        assert np.array_equal(np.asarray(a["equivalent_rules_idx"]), np.asarray(b["equivalent_rules_idx"]))
    assert np.array_equal(np.asarray(m.rule_importances, dtype=float), np.asarray(imp, dtype=float))


@pytest.mark.parametrize("seed", range(6))
def test_cart_fit(ref, seed):
    rs, X, P, n, k, y, chunks = random_case(200 + seed)
    ex = {0: np.where(y == 0)[0], 1: np.where(y == 1)[0]}
    imp = {0: 1.0, 1: float(rs.choice([1.0, 2.5]))}
    rules = ref.rules.LazyKmerRuleList(["K%d" % i for i in range(k)], np.arange(k))
    equiv = {}
    t = ref.cart.DecisionTreeClassifier("gini", 4, 2, imp)
    t.fit(rules, ref.rules.KmerRuleClassifications(ref_loader.ChunkedArray(P, chunks), n), ex, rule_blacklist=[],
          split_callback=lambda node, eq: equiv.__setitem__(id(node), eq))
    root = ko.cart_fit(ko.OracleKmerRuleClassifications(ref_loader.ChunkedArray(P, chunks), n), ex, "gini", 4, 2, imp)

    def walk(a, b):
        assert float(a.criterion_value) == float(b.criterion_value)
        assert (a.rule is None) == b.is_leaf
        if a.rule is not None:
            assert int(a.rule.kmer_index) == int(b.rule_idx)
            assert np.array_equal(np.asarray(equiv[id(a)]), np.asarray(b.equivalent_rules_idx))
            walk(a.left_child, b.left_child)
            walk(a.right_child, b.right_child)
    walk(t.decision_tree, root)
