"""CPU: the model-selection host logic (bound values, predictions, CV risks, best-HP rules, lock-step fit_many)
of kover_b200.learning.experiments over checker-backed shards, against golden vectors replayed from the real
reference's _bound / _predictions / metrics / fit."""
import numpy as np
import pytest

import cases
from experiment_checks import check_cart_cv_case, check_cart_experiment_case, check_driver_case, check_experiment_case
from kover_b200.learning.common.rules import KmerRuleClassifications
from kover_b200.sharding import ShardGroup, partition_columns
from oracle_shard import OracleShard


def make_rc(P, n, chunks):
    k = P.shape[1]
    shards = [OracleShard(P, n, c0, nc) for c0, nc in partition_columns(k, 2) if nc > 0]
    return KmerRuleClassifications(P, n, shard_group=ShardGroup(shards, n, k))


@pytest.mark.parametrize("case", cases.EXPERIMENT_CASES, ids=[c[0] for c in cases.EXPERIMENT_CASES])
def test_bound_and_cv_selection(golden_dir, case):
    check_experiment_case(golden_dir, case, make_rc)


@pytest.mark.parametrize("case", cases.CART_EXPERIMENT_CASES, ids=[c[0] for c in cases.CART_EXPERIMENT_CASES])
def test_pruned_tree_bound_selection(golden_dir, case):
    check_cart_experiment_case(golden_dir, case, make_rc)


@pytest.mark.parametrize("case", cases.CART_CV_CASES, ids=[c[0] for c in cases.CART_CV_CASES])
def test_pruned_tree_cv_selection(golden_dir, case):
    check_cart_cv_case(golden_dir, case, make_rc)


@pytest.mark.parametrize("case", cases.DRIVER_CASES, ids=[c[0] for c in cases.DRIVER_CASES])
def test_drivers_against_the_reference(golden_dir, case, tmp_path):
    """learn_SCM / learn_CART reading a Kover-layout HDF5 file (hdf5_lite + the dataset adapter), with the matrix
    served by checker-backed shards: the host side of the drivers against the real reference's drivers."""
    check_driver_case(golden_dir, case, str(tmp_path), make_rc)


def _tree_signature(node):
    if node.rule is None:
        return ("leaf", node.depth, tuple(sorted((int(c), len(v)) for c, v in node.class_examples_idx.items())))
    return (int(node.rule.kmer_index), float(node.criterion_value), float(node.rule.importance),
            tuple(int(x) for x in node.equivalent_rules_idx), _tree_signature(node.left_child),
            _tree_signature(node.right_child))


@pytest.mark.parametrize("mname,lname", [("M1", "M1_noisy"), ("M2", "M2_planted")])
def test_truncated_tree_equals_directly_grown_tree(mname, lname):
    """A (max_depth, min_samples_split) grid needs one growth: the tree grown at the largest depth / smallest
    min_samples_split, cut at max_depth and below the nodes too small to split, IS the tree grown with those values."""
    from kover_b200.learning.common.rules import LazyKmerRuleList
    from kover_b200.learning.experiments.experiment_cart import _truncate_tree
    from kover_b200.learning.learners.cart import DecisionTreeClassifier
    X, P, n, chunks = cases.matrix(mname)
    y = cases.labels(lname)
    rc = make_rc(P, n, chunks)
    rules = LazyKmerRuleList(["K%d" % i for i in range(P.shape[1])], np.arange(P.shape[1]))
    train = cases.train_subset(n, 7)
    ex = {0: train[y[train] == 0], 1: train[y[train] == 1]}

    def grow(depth, mss):
        t = DecisionTreeClassifier("gini", depth, mss, {0: 1.0, 1: 1.5})
        t.fit(rules, rc, ex)
        return t.decision_tree

    big = grow(8, 2)
    before = _tree_signature(big)
    for depth, mss in ((1, 2), (3, 2), (3, 9), (5, 4), (8, 30), (8, 2)):
        assert _tree_signature(_truncate_tree(big, depth, mss)) == _tree_signature(grow(depth, mss)), (depth, mss)
    assert _tree_signature(big) == before           # the grown tree itself is left untouched


def test_first_hp_scoring_exactly_one_selects_nothing_like_python2():
    """the reference compares ``hp[2] < best_hp["max_rules"]`` while best_hp is still all None; Python 2
    evaluates ``int < None`` as False and carries on (experiment_scm.py:233-246, 612-627), Python 3 raises.  A first
    HP whose score equals the initial 1.0 must select nothing, and a later HP with a smaller score must win."""
    from kover_b200.learning.experiments.experiment_scm import _bound_hp_is_better, _cv_hp_is_better
    empty = {"model_type": None, "p": None, "max_rules": None}
    for better in (_bound_hp_is_better, _cv_hp_is_better):
        assert better(1.0, ("conjunction", 0.1, 3), 1.0, dict(empty)) is False
        assert better(0.99, ("conjunction", 0.1, 3), 1.0, dict(empty)) is True
        chosen = {"model_type": "conjunction", "p": 0.1, "max_rules": 3}
        assert better(0.5, ("disjunction", 1.0, 2), 0.5, chosen) is True          # same score, fewer rules
        assert better(0.5, ("disjunction", 1.0, 3), 0.5, chosen) is True          # same score and size, p closer to 1
        assert better(0.5, ("disjunction", 10.0, 3), 0.5, chosen) is False
        assert better(0.6, ("disjunction", 1.0, 1), 0.5, chosen) is False


def _tree_dict(node):
    d = {"depth": node.depth, "criterion_value": float(node.criterion_value),
         "n_by_class": {int(c): int(len(v)) for c, v in node.class_examples_idx.items()}}
    if node.rule is not None:
        d["rule_idx"] = int(node.rule.kmer_index)
        d["equivalent_rules_idx"] = [int(x) for x in node.equivalent_rules_idx]
        d["importance"] = float(node.rule.importance)
        d["left"], d["right"] = _tree_dict(node.left_child), _tree_dict(node.right_child)
    return d


@pytest.mark.parametrize("case", list(cases.CART_CASES) + list(cases.CART_CE_CASES),
                         ids=[c[0] for c in list(cases.CART_CASES) + list(cases.CART_CE_CASES)])
def test_tree_learner_host_logic(golden_dir, case):
    """kover_b200's DecisionTreeClassifier itself -- the breadth-first growth, the level batching with node / parent
    keys, the cross-entropy formula evaluated on the host (cart.py:167-207) -- over checker-backed shards (no GPU)
    against the REAL reference's trees: Gini bit-exact, cross-entropy to 4 ulp (np.log)."""
    import json
    from kover_b200.learning.common.rules import LazyKmerRuleList
    from kover_b200.learning.learners.cart import DecisionTreeClassifier
    from oracle import kover_oracle as ko
    from test_oracle_golden import load, tree_equal
    name, mname, lname, max_depth, mss, class_importance, tb, subset_seed = case
    want = json.loads(str(load(golden_dir, name)["tree_json"]))
    X, P, n, chunks = cases.matrix(mname)
    y = cases.labels(lname)
    k = P.shape[1]
    rc = make_rc(P, n, chunks)
    train = cases.train_subset(n, subset_seed)
    example_idx = {c: train[y[train] == c] for c in sorted(class_importance.keys())}
    tiebreaker = None
    if tb == "occurrence":
        occ = rc.sum_rows(train)
        tiebreaker = lambda idx: ko.cart_tiebreaker(idx, occ)
    criterion = "cross-entropy" if case in cases.CART_CE_CASES else "gini"
    t = DecisionTreeClassifier(criterion, max_depth, mss, class_importance)
    t.fit(LazyKmerRuleList(["K%d" % i for i in range(k)], np.arange(k)), rc, example_idx, rule_blacklist=[],
          tiebreaker=tiebreaker)
    tree_equal(_tree_dict(t.decision_tree), want, rtol=1e-15 if criterion == "cross-entropy" else 0.0)
