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


def check_driver_case(golden_dir, case, tmp_dir, make_rc):
    """learn_SCM / learn_CART on a Kover dataset FILE against the REAL reference's learn_SCM / learn_CART run
    unmodified on the same content (make_golden.py::gen_drivers).  make_rc(P, n, chunks) -> a rule_classifications
    to adopt (CPU tier: checker-backed shards), or None to let the driver upload the file's matrix to the GPU."""
    import json

    from kover_b200.learning.experiments import experiment_cart as cexp
    from kover_b200.learning.experiments import experiment_scm as kexp
    from kover_file import write_kover_file
    name, _, _, seed, _, learner, selection, max_equiv = case
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    d = cases.driver_dataset(case)
    path = os.path.join(tmp_dir, name + ".kover")
    write_kover_file(path, d["P"], d["n"], d["y"], d["train"], d["test"], d["folds"], chunks=d["chunks"] or (1, d["k"]),
                     fixed_strings=(seed % 2 == 0),
                     names=(d["kmer_sequences"], d["kmer_by_matrix_column"], d["genome_identifiers"]))
    rc = make_rc(d["P"], d["n"], d["chunks"]) if make_rc is not None else None
    want_cls = json.loads(str(g["classifications"]))
    if learner == "scm":
        a = cases.DRIVER_SCM_ARGS
        best_hp, score, trm, tem, model, imp, equiv, cls = kexp.learn_SCM(
            path, "split1", a["model_type"], a["p"], None, a["max_rules"], max_equiv, selection, 1, a["random_seed"], None,
            bound_delta=a["bound_delta"], bound_max_genome_size=a["bound_max_genome_size"], rule_classifications=rc)
        want_hp = json.loads(str(g["best_hp"]))
        assert best_hp["model_type"] == want_hp["model_type"] and best_hp["p"] == want_hp["p"]
        assert int(best_hp["max_rules"]) == want_hp["max_rules"]
        assert (score is None and np.isnan(g["score"])) or score == float(g["score"])
        assert trm["bound"] == float(g["train_bound"]) and trm["f1_score"][0] == float(g["train_f1"])
        assert [str(r) for r in model] == json.loads(str(g["model"]))
        assert np.array_equal(np.asarray(imp, dtype=np.float64), g["importances"])
        assert [sorted(str(r) for r in e) for e in equiv] == json.loads(str(g["equiv"]))
    else:
        a = cases.DRIVER_CART_ARGS
        best_hps, score, trm, tem, model, imp, equiv, cls = cexp.learn_CART(
            path, "split1", a["criterion"], a["max_depth"], a["min_samples_split"], a["class_importance"], a["bound_delta"],
            a["bound_max_genome_size"], None, selection, 1, None, rule_classifications=rc)
        want_hps = json.loads(str(g["best_hps"]))
        assert best_hps["criterion"] == want_hps["criterion"] and int(best_hps["max_depth"]) == want_hps["max_depth"]
        assert int(best_hps["min_samples_split"]) == want_hps["min_samples_split"]
        assert float(best_hps["pruning_alpha"]) == want_hps["pruning_alpha"]
        assert {str(c): w for c, w in best_hps["class_importance"].items()} == want_hps["class_importance"]
        assert score == float(g["score"])
        tree = model.decision_tree
        assert sorted(str(r) for r in tree.rules) == json.loads(str(g["rules"])) and len(tree) == int(g["tree_size"])
        assert sorted((str(r), float(v)) for r, v in imp.items()) == [tuple(x) for x in json.loads(str(g["importances"]))]
        assert sorted((str(r), sorted(str(e) for e in v)) for r, v in equiv.items()) == \
            [(x[0], x[1]) for x in json.loads(str(g["equiv"]))]
    assert trm["risk"][0] == float(g["train_risk"]) and tem["risk"][0] == float(g["test_risk"])
    assert {k_: list(v) for k_, v in cls.items()} == want_cls
