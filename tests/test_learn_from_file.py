"""GPU: ``learn_SCM`` end to end on a Kover-layout HDF5 file (the `kover learn scm` path: HDF5 -> HBM -> lock-step
selection -> model), against the same selection run on the in-memory matrix and the planted model."""
import numpy as np
import pytest

from kover_b200.learning.common.rules import KmerRuleClassifications, LazyKmerRuleList
from kover_b200.learning.experiments import experiment_scm as kexp
from kover_b200.utils import _pack_binary_bytes_to_ints
from kover_file import write_kover_file

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kover_file(tmp_path_factory):
    rs = np.random.RandomState(17)
    n, k = 400, 6000
    X = (rs.rand(n, k) < 0.35).astype(np.uint8)
    y = ((X[:, 70] == 1) & (X[:, 1100] == 0)).astype(np.uint8)            # planted: Presence(70) and Absence(1100)
    flip = rs.rand(n) < 0.02
    y[flip] = 1 - y[flip]
    X[:, 4000] = X[:, 70]                                                  # an equivalent k-mer
    order = np.argsort(y, kind="stable")                                   # genomes sorted by label (create.py:190-194)
    X, y = X[order], y[order]
    packed = _pack_binary_bytes_to_ints(X, 64)
    perm = rs.permutation(n)
    train, test = np.sort(perm[:270]), np.sort(perm[270:])
    fold_of = rs.randint(0, 3, size=len(train))
    folds = [(train[fold_of != f], train[fold_of == f]) for f in range(3)]
    path = str(tmp_path_factory.mktemp("kover") / "planted.kover")
    seqs, perm_k = write_kover_file(path, packed, n, y, train, test, folds, chunks=(1, 1000))
    return dict(path=path, X=X, y=y, packed=packed, train=train, test=test, folds=folds, seqs=seqs, n=n, k=k)


def _model_text(model):
    return " | ".join(sorted(str(r) for r in model))


@pytest.mark.parametrize("selection", ["bound", "cv", "none"])
def test_learn_scm_from_file(kover_file, selection, tmp_path):
    f = kover_file
    blacklist = tmp_path / "blacklist.txt"
    blacklist.write_text(f["seqs"][123] + "\n" + "A" * 12 + "\n")
    warnings = []
    out = kexp.learn_SCM(f["path"], "split1", ["conjunction", "disjunction"], [0.5, 1.0, 2.0], str(blacklist), 4, 10,
                         selection, 1, 42, None, bound_delta=0.05, bound_max_genome_size=12,
                         warning_callback=warnings.append)
    best_hp, best_score, train_metrics, test_metrics, model, importances, equiv, classifications = out
    if "A" * 12 != f["seqs"][0]:
        assert warnings and "AAAAAAAAAAAA" in warnings[0]

    # the same selection on the in-memory matrix, through the functions the golden vectors pin
    from kover_b200.dataset.ds import KoverDataset
    d = KoverDataset(f["path"])
    s = d.get_split("split1")
    rc = KmerRuleClassifications(f["packed"], f["n"], devices=[0])
    rules = LazyKmerRuleList(d.kmer_sequences[...], d.kmer_by_matrix_column[...])
    risks = np.hstack((s.unique_risk_by_kmer[...], s.unique_risk_by_anti_kmer[...]))
    bl = [123, 123 + f["k"]]
    if selection == "bound":
        w_score, w_hp, w_model, w_imp, w_equiv, _ = kexp.bound_selection(
            rc, rules, f["y"], f["train"], risks, ["conjunction", "disjunction"], [0.5, 1.0, 2.0], 4, 10, bl, 0.05, 12,
            np.random.RandomState(42))
        assert best_score == w_score and best_hp == w_hp and _model_text(model) == _model_text(w_model)
        assert np.array_equal(importances, w_imp) and train_metrics["bound"] == w_score
    elif selection == "cv":
        folds = [(np.sort(tr), np.sort(te), np.hstack((fo.unique_risk_by_kmer[...], fo.unique_risk_by_anti_kmer[...])))
                 for (tr, te), fo in zip(f["folds"], s.folds)]
        w_score, w_hp, _ = kexp.cross_validation(rc, rules, f["y"], folds, ["conjunction", "disjunction"], [0.5, 1.0, 2.0],
                                                 4, bl)
        assert best_score == w_score and best_hp == w_hp
    else:
        assert best_score is None and best_hp == {"model_type": "conjunction", "p": 0.5, "max_rules": 4}
    rc.close()
    d.close()

    # the planted concept is recovered (through its k-mer or the equivalent one) and generalises
    text = _model_text(model)
    assert ("Presence(%s)" % f["seqs"][70] in text) or ("Presence(%s)" % f["seqs"][4000] in text), text
    assert "Absence(%s)" % f["seqs"][1100] in text, text
    assert train_metrics["risk"][0] < 0.06 and test_metrics["risk"][0] < 0.1
    assert len(equiv) == len(model) and all(len(e) >= 1 for e in equiv)
    first = sorted(str(r) for r in equiv[[str(r) for r in model].index([t for t in (str(r) for r in model) if "Presence" in t][0])])
    assert "Presence(%s)" % f["seqs"][70] in first and "Presence(%s)" % f["seqs"][4000] in first     # the equivalent pair
    n_train = len(f["train"])
    assert len(classifications["train_correct"]) + len(classifications["train_errors"]) == n_train
    assert all(g.startswith("genome_") for g in classifications["test_correct"])


@pytest.mark.parametrize("selection", ["bound", "cv"])
def test_learn_cart_from_file(kover_file, selection):
    from kover_b200.learning.experiments import experiment_cart as cexp
    f = kover_file
    importances = [{0: 1.0, 1: 1.0}, {0: 1.0, 1: 2.0}]
    out = cexp.learn_CART(f["path"], "split1", ["gini"], [3, 6], [2], importances, 0.05, 12, None, selection, 1)
    best_hps, best_score, train_metrics, test_metrics, model, rule_imp, equiv, classifications = out

    # the same search on the in-memory matrix through the functions the golden vectors pin
    rc = KmerRuleClassifications(f["packed"], f["n"], devices=[0])
    rules = LazyKmerRuleList(["s%d" % i for i in range(f["k"])], np.arange(f["k"]))
    results = []
    for imp in importances:
        for depth in (3, 6):
            hps = {"criterion": "gini", "class_importance": imp, "max_depth": depth, "min_samples_split": 2}
            if selection == "bound":
                results.append(cexp.learn_pruned_tree_bound(rc, rules, f["y"], f["train"], 2, hps, 0.05, 12, [])[:3])
            else:
                results.append(cexp.learn_pruned_tree_cv(rc, rules, f["y"], f["train"], [(np.sort(a), np.sort(b)) for a, b in f["folds"]],
                                                         2, hps, []))
    rc.close()
    assert best_score == min(r[1] for r in results)
    assert best_hps["criterion"] == "gini" and best_hps["max_depth"] in (3, 6) and "pruning_alpha" in best_hps
    tree = model.decision_tree
    used = sorted(int(r.kmer_index) for r in tree.rules)
    assert (70 in used or 4000 in used) and 1100 in used, used
    assert train_metrics["risk"][0] < 0.06 and test_metrics["risk"][0] < 0.1
    assert abs(sum(rule_imp.values()) - 1.0) < 1e-12 and set(equiv.keys()) == set(tree.rules)
    root_equiv = sorted(str(r) for r in equiv[tree.rule])
    assert len(classifications["train_correct"]) + len(classifications["train_errors"]) == len(f["train"])
    assert any(f["seqs"][70] in t for t in root_equiv) or any(f["seqs"][1100] in t for t in root_equiv)


# ---------------------------------------------------------------------------------------------- against the reference
@pytest.mark.parametrize("case", [c for c in __import__("cases").DRIVER_CASES], ids=[c[0] for c in __import__("cases").DRIVER_CASES])
def test_drivers_against_the_reference(case, tmp_path, golden_dir):
    """learn_SCM / learn_CART on a Kover dataset FILE (HDF5 -> dataset adapter -> HBM -> lock-step selection) against
    the REAL reference's learn_SCM / learn_CART run unmodified on the same content (make_golden.py::gen_drivers):
    selected hyper-parameters, score, metrics, model, importances, equivalent rules and the classified genomes."""
    from experiment_checks import check_driver_case
    check_driver_case(golden_dir, case, str(tmp_path), None)
