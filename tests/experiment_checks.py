"""Shared by the CPU (checker-backed shards) and GPU tests of kover_b200.learning.experiments.experiment_scm."""
import json
import os

import numpy as np

import cases
from kover_b200.learning.common.rules import LazyKmerRuleList
from kover_b200.learning.experiments import experiment_scm as E


def check_experiment_case(golden_dir, case, make_rc):
    name, mname, lname, seed, p_values, max_rules, max_equiv, n_folds = case
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    X, P, n, chunks = cases.matrix(mname)
    labels = cases.labels(lname)
    k = P.shape[1]
    train = cases.train_subset(n, seed)
    rules = LazyKmerRuleList(["K%d" % i for i in range(k)], np.arange(k))
    risks = cases.rule_risks(2 * k)
    rc = make_rc(P, n, chunks)
    model_types = ["conjunction", "disjunction"]

    score, hp, model, importances, equiv, per_hp = E.bound_selection(
        rc, rules, labels, train, risks, model_types, p_values, max_rules, max_equiv, [], 0.05, 10000,
        np.random.RandomState(7))
    for (mt, p), by_length in per_hp.items():
        assert np.array_equal(by_length, g["bound_%s_%g" % (mt, p)]), (mt, p)        # bound values, ==
    assert score == float(g["bound_best_score"])
    assert hp == json.loads(str(g["bound_best_hp"]))
    assert str(model) == str(g["bound_best_model"])
    for t, eq in enumerate(equiv):
        assert np.array_equal(np.asarray(eq, dtype=np.int64), g["equiv_%s_%g_%d" % (hp["model_type"], hp["p"], t)])
    _, test_pred = E._predictions(model, rc, [], np.setdiff1d(np.arange(n), train))
    assert np.array_equal(test_pred, g["bound_best_model_test_predictions"])         # predictions

    folds = [(a, b, risks) for a, b in cases.experiment_folds(train, n_folds)]
    score, hp, per_hp = E.cross_validation(rc, rules, labels, folds, model_types, p_values, max_rules, [])
    for (mt, p), mean in per_hp.items():
        assert np.array_equal(mean, g["cv_%s_%g" % (mt, p)]), (mt, p)
    assert score == float(g["cv_best_score"])
    assert hp == json.loads(str(g["cv_best_hp"]))


def check_cart_experiment_case(golden_dir, case, make_rc):
    from kover_b200.learning.experiments import experiment_cart as C
    name, mname, lname, seed, max_depth, mss, imp = case
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    X, P, n, chunks = cases.matrix(mname)
    labels = cases.labels(lname)
    k = P.shape[1]
    train = cases.train_subset(n, seed)
    test = np.setdiff1d(np.arange(n), train)
    rules = LazyKmerRuleList(["K%d" % i for i in range(k)], np.arange(k))
    rc = make_rc(P, n, chunks)
    hps = {"criterion": "gini", "max_depth": max_depth, "min_samples_split": mss, "class_importance": imp}
    hps, min_score, tree, sequence = C.learn_pruned_tree_bound(rc, rules, labels, train, 2, hps, 0.05, 10000, [])
    assert np.array_equal(np.array([a for a, _ in sequence], dtype=np.float64), g["alphas"])      # pruning alphas, ==
    assert np.array_equal(np.array([b for _, b in sequence]), g["bounds"])                         # bound values, ==
    assert min_score == float(g["min_score"]) and hps["pruning_alpha"] == float(g["best_alpha"])
    assert sorted(int(r.kmer_index) for r in tree.rules) == g["best_rules"].tolist()
    assert np.array_equal(np.asarray(C._predictions(tree, rc, [], test)[1]), g["best_test_predictions"])


def check_cart_cv_case(golden_dir, case, make_rc):
    """learn_pruned_tree_cv (fold trees + master tree grown in lock-step) against the REAL reference's
    _learn_pruned_tree_cv run unmodified (tests/golden/make_golden.py::gen_cart_cv)."""
    from kover_b200.learning.experiments import experiment_cart as C
    name, mname, lname, seed, n_folds, max_depth, mss, imp = case
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    X, P, n, chunks = cases.matrix(mname)
    labels = cases.labels(lname)
    k = P.shape[1]
    train = cases.train_subset(n, seed)
    test = np.setdiff1d(np.arange(n), train)
    folds = cases.cv_folds(train, n_folds, seed)
    rules = LazyKmerRuleList(["K%d" % i for i in range(k)], np.arange(k))
    rc = make_rc(P, n, chunks)
    hps = {"criterion": "gini", "max_depth": max_depth, "min_samples_split": mss, "class_importance": imp}
    hps, score, tree = C.learn_pruned_tree_cv(rc, rules, labels, train, folds, 2, hps, [])
    assert score == float(g["cv_score"]) and hps["pruning_alpha"] == float(g["pruning_alpha"])
    assert sorted(int(r.kmer_index) for r in tree.rules) == g["best_rules"].tolist() and len(tree) == int(g["tree_size"])
    assert np.array_equal(np.asarray(C._predictions(tree, rc, [], test)[1]), g["best_test_predictions"])
