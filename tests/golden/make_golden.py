"""Mints the committed golden vectors by running the REAL Kover reference (loaded from
/root/reference by oracle/ref_loader.py) on the seeded inputs of cases.py.

Run in the build container only:   python tests/golden/make_golden.py
Outputs: tests/golden/*.npz (small, committed).  The GPU box has no reference tree; tests
there compare against these files.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import cases  # noqa: E402
from oracle.ref_loader import ChunkedArray, load_reference  # noqa: E402

ref = load_reference()


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote %-32s %8d bytes" % (name + ".npz", os.path.getsize(path)))


def ref_rc(P, n, chunks):
    return ref.rules.KmerRuleClassifications(ChunkedArray(P, chunks), n)


# ------------------------------------------------------------------ popcount + pack (a1, a12)
def gen_popcount():
    rs = np.random.RandomState(100)
    a64 = rs.randint(0, 2 ** 63, size=(4, 257), dtype=np.int64).astype(np.uint64)
    a64 ^= rs.randint(0, 2, size=a64.shape, dtype=np.int64).astype(np.uint64) << np.uint64(63)
    a64[rs.rand(*a64.shape) < 0.1] = 0
    m64 = np.array([0xFFFFFFFFFFFFFFFF, 0, 0x8000000000000001, 0x00FF00FF00FF00FF], dtype=np.uint64)
    o64 = a64.copy()
    ref.popcount.inplace_popcount_64(o64, m64)
    a32 = rs.randint(0, 2 ** 32, size=(3, 131), dtype=np.int64).astype(np.uint32)
    a32[rs.rand(*a32.shape) < 0.1] = 0
    m32 = np.array([0xFFFFFFFF, 0x80000001, 0x0F0F0F0F], dtype=np.uint32)
    o32 = a32.copy()
    ref.popcount.inplace_popcount_32(o32, m32)
    save("popcount", a64=a64, m64=m64, o64=o64, a32=a32, m32=m32, o32=o32)


def gen_pack():
    X = cases.dense_bits(101, 130, 9, 0.5)
    p64 = ref.utils._pack_binary_bytes_to_ints(X, 64)
    p32 = ref.utils._pack_binary_bytes_to_ints(X, 32)
    u64 = ref.utils._unpack_binary_bytes_from_ints(p64)
    u32 = ref.utils._unpack_binary_bytes_from_ints(p32)
    mins = np.array([np.dtype(ref.utils._minimum_uint_size(v)).itemsize for v in (0, 255, 256, 65535, 65536, 2 ** 32)])
    save("pack", X=X, p64=p64, p32=p32, u64=u64, u32=u32, min_uint_itemsize=mins)


# ------------------------------------------------------------------ sum_rows / get_columns (a3, a4, a5)
def gen_rules():
    for mname in ("M1", "M2", "M3", "M4"):
        X, P, n, chunks = cases.matrix(mname)
        k = P.shape[1]
        out = {}
        rc = ref_rc(P, n, chunks)
        out["shape"] = np.array(rc.shape)
        for rname, rows in cases.row_sets(n).items():
            r = rc.sum_rows(rows)
            out["sum_" + rname] = r
        for cname, cols in cases.column_sets(k).items():
            out["col_" + cname] = rc.get_columns(cols)
        if mname == "M3":        # 32-bit packing (reader supports it: rules.py:123-125)
            rc32 = ref_rc(cases.pack32_from64(P), n, chunks)
            for rname, rows in cases.row_sets(n).items():
                out["sum32_" + rname] = rc32.sum_rows(rows)
            out["col32_mixed"] = rc32.get_columns(cases.column_sets(k)["mixed"])
        if mname == "M4":        # remove_rows bookkeeping (rules.py:173-194)
            rc2 = ref_rc(P, n, chunks)
            rc2.remove_rows([3, 10])
            rc2.remove_rows([0])
            out["rm_shape"] = np.array(rc2.shape)
            out["rm_sum"] = rc2.sum_rows(np.array([0, 1, 2, 9, 40, 61]))
            out["rm_col"] = rc2.get_columns([1, k + 2])
        save("rules_" + mname, **out)


# ------------------------------------------------------------------ SCM utility scan (a6)
def gen_utility():
    for case in cases.UTILITY_CASES:
        P, n, chunks, pos, neg, p, blacklist, ubs = cases.utility_inputs(case)
        ref.scm.UTIL_BLOCK_SIZE = ubs
        m = ref.scm.SetCoveringMachine(model_type=ref.models.conjunction, p=p, max_rules=10)
        bu, bi, bp, bn = m._get_best_utility_rules(ref_rc(P, n, chunks), pos, neg, rule_blacklist=blacklist)
        save(case[0], best_utility=np.float64(bu), best_idx=np.asarray(bi, dtype=np.int64),
             best_pos_err=np.asarray(bp, dtype=np.int64), best_neg_cover=np.asarray(bn, dtype=np.int64))
    ref.scm.UTIL_BLOCK_SIZE = 1000000


# ------------------------------------------------------------------ SCM fit (a7, a8)
def gen_scm():
    for case in cases.SCM_CASES:
        name, mname, lname, model_type, p, max_rules, tb, blacklist, ubs = case
        X, P, n, chunks = cases.matrix(mname)
        y = cases.labels(lname)
        k = P.shape[1]
        ref.scm.UTIL_BLOCK_SIZE = ubs
        rc = ref_rc(P, n, chunks)
        rules = ref.rules.LazyKmerRuleList(["K%d" % i for i in range(k)], np.arange(k))
        risks = cases.rule_risks(2 * k)

        def tiebreaker(idx):                      # experiment_scm.py:122-130
            if tb == "identity":
                return idx
            tie = risks[idx]
            if model_type == "conjunction":
                return idx[np.isclose(tie, tie.min())]
            return idx[np.isclose(tie, tie.max())]

        infos = []
        m = ref.scm.SetCoveringMachine(model_type=model_type, p=p, max_rules=max_rules)
        m.fit(rules, rc, np.where(y == 1)[0], np.where(y == 0)[0], rule_blacklist=blacklist,
              tiebreaker=tiebreaker, iteration_callback=lambda i: infos.append(dict(i)),
              iteration_rule_importances=True)
        out = {"n_iterations": np.array(len(infos)),
               "model_str": np.array(str(m.model)),
               "rule_importances": np.asarray(m.rule_importances, dtype=np.float64)}
        for t, info in enumerate(infos):
            out["it%d_utility_max" % t] = np.float64(info["utility_max"])
            out["it%d_utility_argmax" % t] = np.asarray(info["utility_argmax"], dtype=np.int64)
            out["it%d_equivalent" % t] = np.asarray(info["equivalent_rules_idx"], dtype=np.int64)
            out["it%d_selected" % t] = np.array(str(info["selected_rule"]))
            out["it%d_importances" % t] = np.asarray(info["rule_importances"], dtype=np.float64)
        save(name, **out)
    ref.scm.UTIL_BLOCK_SIZE = 1000000


# ------------------------------------------------------------------ CART fit (a9, a11)
def tree_to_dict(node):
    d = {"depth": node.depth, "criterion_value": float(node.criterion_value),
         "n_by_class": {int(c): int(len(v)) for c, v in node.class_examples_idx.items()}}
    if node.rule is not None:
        d["rule_idx"] = int(node.rule.kmer_index)
        d["equivalent_rules_idx"] = [int(x) for x in node.rule.equivalent_rules_idx]
        d["importance"] = float(node.rule.importance)
        d["left"] = tree_to_dict(node.left_child)
        d["right"] = tree_to_dict(node.right_child)
    return d


def gen_cart():
    for case in cases.CART_CASES:
        name, mname, lname, max_depth, mss, class_importance, tb, subset_seed = case
        X, P, n, chunks = cases.matrix(mname)
        y = cases.labels(lname)
        k = P.shape[1]
        rc = ref_rc(P, n, chunks)
        rules = ref.rules.LazyKmerRuleList(["K%d" % i for i in range(k)], np.arange(k))
        train = cases.train_subset(n, subset_seed)
        example_idx = {c: train[y[train] == c] for c in sorted(class_importance.keys())}
        if tb == "occurrence":                    # experiment_cart.py:82-94, 264
            occ = rc.sum_rows(train)
            tiebreaker = lambda idx: idx[np.isclose(occ[idx], occ[idx].max())]
        else:
            tiebreaker = None

        def split_cb(node, equiv):                # experiment_cart.py:97-106
            node.rule.equivalent_rules_idx = equiv

        t = ref.cart.DecisionTreeClassifier("gini", max_depth, mss, class_importance)
        t.fit(rules, rc, example_idx, rule_blacklist=[], tiebreaker=tiebreaker, split_callback=split_cb)
        d = tree_to_dict(t.decision_tree)
        # also the raw root scores, to pin the float64 operation order of _gini_rule_score
        save(name, tree_json=np.array(json.dumps(d)))
    # raw root gini scores for two cases (cart.py:112-161 is a closure: re-run it via a tiny fit that
    # captures rules_criterion through the tiebreaker's candidate set is not possible; instead pin the
    # scores through `candidates == where(score == min)` at every node, which tree_json already holds)


def gen_cart_ce():
    """The cross-entropy criterion (cart.py:167-207): whole trees (criterion values, selected rules, the equivalent
    rules = the tie set that survived the tiebreaker at every node) plus the raw score vector of every searched
    node, captured where _find_best_split hands it to the builtin ``min`` (cart.py:234)."""
    import builtins
    for case in cases.CART_CE_CASES:
        name, mname, lname, max_depth, mss, class_importance, tb, subset_seed = case
        X, P, n, chunks = cases.matrix(mname)
        y = cases.labels(lname)
        k = P.shape[1]
        rc = ref_rc(P, n, chunks)
        rules = ref.rules.LazyKmerRuleList(["K%d" % i for i in range(k)], np.arange(k))
        train = cases.train_subset(n, subset_seed)
        example_idx = {c: train[y[train] == c] for c in sorted(class_importance.keys())}
        tiebreaker = None
        if tb == "occurrence":
            occ = rc.sum_rows(train)
            tiebreaker = lambda idx: idx[np.isclose(occ[idx], occ[idx].max())]
        captured = []

        def spy_min(*args, **kwargs):                 # module-level name: shadows the builtin inside cart.py only
            if len(args) == 1 and isinstance(args[0], np.ndarray) and args[0].shape == (k,):
                if not captured or captured[-1] is not args[0]:
                    captured.append(args[0])
            return builtins.min(*args, **kwargs)

        def split_cb(node, equiv):
            node.rule.equivalent_rules_idx = equiv

        ref.cart.min = spy_min
        try:
            t = ref.cart.DecisionTreeClassifier("cross-entropy", max_depth, mss, class_importance)
            t.fit(rules, rc, example_idx, rule_blacklist=[], tiebreaker=tiebreaker, split_callback=split_cb)
        finally:
            del ref.cart.min
        save(name, tree_json=np.array(json.dumps(tree_to_dict(t.decision_tree))),
             node_scores=np.array([np.array(c) for c in captured]))


# ------------------------------------------------------------------ bound / CV model selection (north-star item 4)
def gen_experiments():
    """experiment_scm.py's per-HP scoring (_bound_score_hp :401-566, _cv_score_hp :102-193) needs a KoverDataset
    (HDF5); its bodies are replayed here on the REAL reference's fit / _bound / _predictions / metrics, with the
    best-HP rules of _bound_selection (:612-627) and _cross_validation (:233-246)."""
    from copy import deepcopy
    from itertools import product
    E = ref.experiment_scm
    for name, mname, lname, seed, p_values, max_rules, max_equiv, n_folds in cases.EXPERIMENT_CASES:
        X, P, n, chunks = cases.matrix(mname)
        labels = cases.labels(lname)
        k = P.shape[1]
        train = cases.train_subset(n, seed)
        rules = ref.rules.LazyKmerRuleList(["K%d" % i for i in range(k)], np.arange(k))
        risks = cases.rule_risks(2 * k)
        out = {}

        def tiebreaker(idx, model_type):
            tie = risks[idx]
            return idx[np.isclose(tie, tie.min())] if model_type == "conjunction" else idx[np.isclose(tie, tie.max())]

        # ---- bound selection
        best_score, best_hp, best_model = 1.0, {"model_type": None, "p": None, "max_rules": None}, None
        for model_type, p in product(["conjunction", "disjunction"], p_values):
            rc = ref_rc(P, n, chunks)
            rng = np.random.RandomState(7)                # every pool task unpickles its own copy
            tmp_model = ref.models.ConjunctionModel() if model_type == "conjunction" else ref.models.DisjunctionModel()
            score_by_length, model_by_length, equivalent_rules = np.ones(max_rules), [], []
            train_answers = labels[train]

            def cb(info):
                tmp_model.add(info["selected_rule"])
                model_by_length.append(deepcopy(tmp_model))
                eq = info["equivalent_rules_idx"]
                if len(eq) > max_equiv:
                    ridx = rng.choice(len(eq), max_equiv, replace=False)
                    ridx.sort()
                    eq = eq[ridx]
                if model_type == "disjunction":
                    eq = (eq + k) % (2 * k)
                equivalent_rules.append(eq)
                _, tp = E._predictions(tmp_model, P, [], train)
                score_by_length[info["iteration_number"] - 1] = E._bound(
                    train_predictions=tp, train_answers=train_answers, train_example_idx=train, model=tmp_model,
                    delta=0.05, max_genome_size=10000, rule_classifications=rc)

            m = ref.scm.SetCoveringMachine(model_type=model_type, p=p, max_rules=max_rules)
            m.fit(rules, rc, train[labels[train] == 1], train[labels[train] == 0], rule_blacklist=[],
                  tiebreaker=lambda idx: tiebreaker(idx, model_type), iteration_callback=cb,
                  iteration_rule_importances=True)
            key = "%s_%g" % (model_type, p)
            out["bound_" + key] = score_by_length.copy()
            for t, eq in enumerate(equivalent_rules):
                out["equiv_%s_%d" % (key, t)] = np.asarray(eq, dtype=np.int64)
            bi = int(np.argmin(score_by_length))
            score, hp = score_by_length[bi], (model_type, p, bi + 1)
            if (score < best_score) or (score == best_score and hp[2] < best_hp["max_rules"]) or \
                    (score == best_score and hp[2] == best_hp["max_rules"] and abs(1.0 - hp[1]) < abs(1.0 - best_hp["p"])):
                best_hp = {"model_type": hp[0], "p": hp[1], "max_rules": hp[2]}
                best_score, best_model = score, model_by_length[bi]
        out["bound_best_score"] = np.float64(best_score)
        out["bound_best_hp"] = np.array(json.dumps(best_hp))
        out["bound_best_model"] = np.array(str(best_model))
        _, test_pred = E._predictions(best_model, P, [], np.setdiff1d(np.arange(n), train))
        out["bound_best_model_test_predictions"] = test_pred

        # ---- cross-validation
        folds = cases.experiment_folds(train, n_folds)
        best_score, best_hp = 1.0, {"model_type": None, "p": None, "max_rules": None}
        for model_type, p in product(["conjunction", "disjunction"], p_values):
            fold_scores = np.ones((n_folds, max_rules + 1)) * np.inf
            for i, (ftrain, ftest) in enumerate(folds):
                rc = ref_rc(P, n, chunks)
                tmp_model = ref.models.ConjunctionModel() if model_type == "conjunction" else ref.models.DisjunctionModel()
                preds = [E._predictions(tmp_model, P, [], ftest)[1]]

                def cb(info):
                    tmp_model.add(info["selected_rule"])
                    preds.append(E._predictions(tmp_model, P, [], ftest)[1])

                m = ref.scm.SetCoveringMachine(model_type=model_type, p=p, max_rules=max_rules)
                m.fit(rules, rc, ftrain[labels[ftrain] == 1], ftrain[labels[ftrain] == 0], rule_blacklist=[],
                      tiebreaker=lambda idx: tiebreaker(idx, model_type), iteration_callback=cb)
                by_length = np.array(E._duplicate_last_element(preds, max_rules + 1))
                fold_scores[i] = E._get_binary_metrics(predictions=by_length, answers=labels[ftest])["risk"]
            mean = np.mean(fold_scores, axis=0)
            out["cv_%s_%g" % (model_type, p)] = mean
            bi = int(np.argmin(mean))
            score, hp = mean[bi], (model_type, p, bi)
            if (not np.allclose(score, best_score) and score < best_score) or \
                    (np.allclose(score, best_score) and hp[2] < best_hp["max_rules"]) or \
                    (np.allclose(score, best_score) and hp[2] == best_hp["max_rules"]
                     and not np.allclose(hp[1], best_hp["p"]) and abs(1.0 - hp[1]) < abs(1.0 - best_hp["p"])):
                best_hp = {"model_type": hp[0], "p": hp[1], "max_rules": hp[2]}
                best_score = score
        out["cv_best_score"] = np.float64(best_score)
        out["cv_best_hp"] = np.array(json.dumps(best_hp))
        save(name, **out)


def gen_cart_experiments():
    """_learn_pruned_tree_bound (experiment_cart.py:208-294) replayed on the REAL reference's fit / _prune_tree /
    _predictions / _bound (it needs a KoverDataset, i.e. HDF5, as a whole)."""
    from functools import partial
    E = ref.experiment_cart
    for name, mname, lname, seed, max_depth, mss, imp in cases.CART_EXPERIMENT_CASES:
        X, P, n, chunks = cases.matrix(mname)
        labels = cases.labels(lname)
        k = P.shape[1]
        train = cases.train_subset(n, seed)
        test = np.setdiff1d(np.arange(n), train)
        rules = ref.rules.LazyKmerRuleList(["K%d" % i for i in range(k)], np.arange(k))
        rc = ref_rc(P, n, chunks)
        t = ref.cart.DecisionTreeClassifier("gini", max_depth, mss, imp)
        t.fit(rules=rules, rule_classifications=rc, example_idx={c: train[labels[train] == c] for c in range(2)},
              rule_blacklist=[], tiebreaker=partial(E._tiebreaker, rule_kmer_occurrences=rc.sum_rows(train)),
              level_callback=None, split_callback=E._split_callback)
        alphas, trees = E._prune_tree(t.decision_tree)
        bounds, min_score, best, best_alpha = [], np.inf, None, None
        for alpha, tree in zip(alphas, trees):
            tp = E._predictions(tree, P, train, [])[0]
            b = E._bound(train_predictions=tp, train_answers=labels[train], train_example_idx=train, model=tree,
                         delta=0.05, max_genome_size=10000, rule_classifications=ref_rc(P, n, chunks), n_classes=2)
            bounds.append(b)
            if b <= min_score:
                min_score, best, best_alpha = b, tree, alpha
        save(name, alphas=np.array(alphas, dtype=np.float64), tree_sizes=np.array([len(x) for x in trees]),
             bounds=np.array(bounds), best_alpha=np.float64(best_alpha), min_score=np.float64(min_score),
             best_rules=np.array(sorted(int(r.kmer_index) for r in best.rules), dtype=np.int64),
             best_test_predictions=np.asarray(E._predictions(best, P, [], test)[1]),
             master_size=np.array(len(t.decision_tree)))


def gen_cart_cv():
    """The REAL reference's _learn_pruned_tree_cv (experiment_cart.py:297-436), run UNMODIFIED: only the HDF5-backed
    KoverDataset it opens is replaced by an in-memory stand-in with the same attributes."""
    import types
    E = ref.experiment_cart
    for name, mname, lname, seed, n_folds, max_depth, mss, imp in cases.CART_CV_CASES:
        X, P, n, chunks = cases.matrix(mname)
        labels = cases.labels(lname)
        k = P.shape[1]
        train = cases.train_subset(n, seed)
        folds = cases.cv_folds(train, n_folds, seed)

        def stub_dataset(dataset_file):
            split = types.SimpleNamespace(train_genome_idx=train, folds=[
                types.SimpleNamespace(train_genome_idx=tr, test_genome_idx=te) for tr, te in folds])
            return types.SimpleNamespace(
                get_split=lambda split_name: split,
                phenotype=types.SimpleNamespace(metadata=labels, tags=np.array(["0", "1"])),
                kmer_sequences=["K%d" % i for i in range(k)], kmer_by_matrix_column=np.arange(k),
                kmer_matrix=ChunkedArray(P, chunks), genome_count=n)

        E.KoverDataset = stub_dataset
        hps = {"criterion": "gini", "max_depth": max_depth, "min_samples_split": mss, "class_importance": imp}
        hps, score, tree = E._learn_pruned_tree_cv(hps, "unused.kover", "unused", [])
        test = np.setdiff1d(np.arange(n), train)
        save(name, cv_score=np.float64(score), pruning_alpha=np.float64(hps["pruning_alpha"]),
             best_rules=np.array(sorted(int(r.kmer_index) for r in tree.rules), dtype=np.int64),
             tree_size=np.array(len(tree)),
             best_test_predictions=np.asarray(E._predictions(tree, P, [], test)[1]))


class _SerialPool(object):
    """multiprocessing.Pool stand-in: tasks run in order in this process, each with ITS OWN copy of the callable,
    like a task unpickled in a pool worker (experiment_scm.py:594-603 hands the random generator over that way)."""

    def __init__(self, processes=None):
        pass

    def imap_unordered(self, func, iterable):
        from copy import deepcopy
        for item in iterable:
            yield deepcopy(func)(item)


def _stub_dataset(d):
    """In-memory KoverDataset with the attributes the reference's drivers read (dataset/ds.py:27-196)."""
    import types
    from oracle import kover_oracle as ko
    orc = ko.OracleKmerRuleClassifications(d["P"], d["n"], fast=True)
    y = d["y"]

    def tables(tr):
        ur, bk, ba = ko.kmer_risk_tables(orc, tr[y[tr] == 1], tr[y[tr] == 0])
        return dict(unique_risks=ur, unique_risk_by_kmer=bk, unique_risk_by_anti_kmer=ba)

    folds = [types.SimpleNamespace(name="fold_%d" % (i + 1), train_genome_idx=np.sort(tr), test_genome_idx=np.sort(te),
                                   **tables(tr)) for i, (tr, te) in enumerate(d["folds"])]
    split = types.SimpleNamespace(name="split1", train_genome_idx=d["train"], test_genome_idx=d["test"], folds=folds,
                                  **tables(d["train"]))
    ds = types.SimpleNamespace(
        get_split=lambda split_name: split, genome_count=d["n"], kmer_count=d["k"], kmer_length=12,
        classification_type="binary",
        phenotype=types.SimpleNamespace(metadata=y, tags=np.array(["0", "1"])),
        kmer_sequences=np.array(d["kmer_sequences"]), kmer_by_matrix_column=d["kmer_by_matrix_column"],
        kmer_matrix=ChunkedArray(d["P"], d["chunks"]), genome_identifiers=np.array(d["genome_identifiers"]))
    return lambda dataset_file: ds


def gen_drivers():
    """The REAL reference's learn_SCM (experiment_scm.py:675-888) and learn_CART (experiment_cart.py:521-649), run
    UNMODIFIED end to end: only the HDF5-backed KoverDataset and the fork pool are replaced by in-process
    stand-ins."""
    S, C = ref.experiment_scm, ref.experiment_cart
    for case in cases.DRIVER_CASES:
        name, _, _, _, _, learner, selection, max_equiv = case
        d = cases.driver_dataset(case)
        if learner == "scm":
            S.KoverDataset, S.Pool = _stub_dataset(d), _SerialPool
            a = cases.DRIVER_SCM_ARGS
            out = S.learn_SCM("unused.kover", "split1", a["model_type"], a["p"], None, a["max_rules"], max_equiv, selection,
                              1, a["random_seed"], None, bound_delta=a["bound_delta"],
                              bound_max_genome_size=a["bound_max_genome_size"])
            best_hp, score, trm, tem, model, imp, equiv, cls = out
            save(name, best_hp=json.dumps({k_: (v.item() if hasattr(v, "item") else v) for k_, v in best_hp.items()}),
                 score=np.float64(np.nan if score is None else score),
                 train_risk=np.float64(trm["risk"][0]), test_risk=np.float64(tem["risk"][0]),
                 train_bound=np.float64(trm["bound"]), train_f1=np.float64(trm["f1_score"][0]),
                 model=json.dumps([str(r) for r in model]), importances=np.asarray(imp, dtype=np.float64),
                 equiv=json.dumps([sorted(str(r) for r in e) for e in equiv]),
                 classifications=json.dumps({k_: list(v) for k_, v in cls.items()}))
        else:
            C.KoverDataset, C.Pool = _stub_dataset(d), _SerialPool
            a = cases.DRIVER_CART_ARGS
            out = C.learn_CART("unused.kover", "split1", a["criterion"], a["max_depth"], a["min_samples_split"],
                               a["class_importance"], a["bound_delta"], a["bound_max_genome_size"], None, selection, 1, None)
            best_hps, score, trm, tem, model, imp, equiv, cls = out
            tree = model.decision_tree
            save(name, best_hps=json.dumps({k_: ({str(c): w for c, w in v.items()} if isinstance(v, dict) else
                                                  (v.item() if hasattr(v, "item") else v)) for k_, v in best_hps.items()}),
                 score=np.float64(score), train_risk=np.float64(trm["risk"][0]), test_risk=np.float64(tem["risk"][0]),
                 rules=json.dumps(sorted(str(r) for r in tree.rules)), tree_size=np.array(len(tree)),
                 importances=json.dumps(sorted((str(r), float(v)) for r, v in imp.items())),
                 equiv=json.dumps(sorted((str(r), sorted(str(e) for e in v)) for r, v in equiv.items())),
                 classifications=json.dumps({k_: list(v) for k_, v in cls.items()}))


# ------------------------------------------------------------------ dataset split risk tables (f-4)
def gen_risk():
    """dataset/split.py:171-188 cannot be called as a function (it is inlined in _split, which needs an HDF5
    file and h5py); the lines are replayed here verbatim on the REAL reference's sum_rows."""
    for name, mname, lname, seed in cases.RISK_CASES:
        X, P, n, chunks = cases.matrix(mname)
        labels = cases.labels(lname)
        train_idx = cases.train_subset(n, seed)
        k = P.shape[1]
        kmer_matrix = ref_rc(P, n, chunks)
        train_pos_idx = train_idx[labels[train_idx] == 1]
        train_neg_idx = train_idx[labels[train_idx] == 0]
        kmer_risks = (len(train_pos_idx) - kmer_matrix.sum_rows(train_pos_idx)[: k]).astype(float)
        kmer_risks += kmer_matrix.sum_rows(train_neg_idx)[: k]
        kmer_risks /= len(train_idx)
        np.round(kmer_risks, 5, out=kmer_risks)
        anti_kmer_risks = 1.0 - kmer_risks
        np.round(anti_kmer_risks, 5, out=anti_kmer_risks)
        unique_risks, inv = np.unique(np.hstack((kmer_risks, anti_kmer_risks)), return_inverse=True)
        t = ref.utils._minimum_uint_size(len(unique_risks))
        save(name, unique_risks=unique_risks, by_kmer=inv[:k].astype(t), by_anti_kmer=inv[k:].astype(t))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "risk":
        gen_risk()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "experiments":
        gen_experiments()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "cart_experiments":
        gen_cart_experiments()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "cart_cv":
        gen_cart_cv()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "cart_ce":
        gen_cart_ce()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "drivers":
        gen_drivers()
        sys.exit(0)
    gen_popcount()
    gen_pack()
    gen_rules()
    gen_utility()
    gen_scm()
    gen_cart()
    gen_cart_ce()
    gen_risk()
    gen_experiments()
    gen_cart_experiments()
    gen_cart_cv()
    gen_drivers()
