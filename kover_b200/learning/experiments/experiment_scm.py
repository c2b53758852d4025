"""Model selection for the SCM over a device-resident matrix: sample-compression bound and cross-validation,
with all hyper-parameter combinations (and folds) fitted in LOCK-STEP so that their utility scans share
matrix passes (north-star item 4).

Mirrors the pure parts of reference learning/experiments/experiment_scm.py -- ``_predictions`` (:43-100),
the per-HP scoring of ``_cv_score_hp`` (:102-193) and ``_bound_score_hp`` (:401-566), ``_bound`` (:349-398,
including its conditional-expression precedence: with a non-empty model the two ln C(.,.) terms are NOT
added) and the best-HP rules of ``_cross_validation`` (:233-246) and ``_bound_selection`` (:612-627) --
without the HDF5 dataset object and the fork pool around them: callers pass the labels, index arrays and risk
tables they read from the dataset.  Hyper-parameter combinations are visited in the reference's
``--n-cpu 1`` order, ``product(model_types, p_values)``.
"""
import logging
from copy import deepcopy
from functools import partial
from itertools import product
from math import exp, log as ln, pi

import numpy as np
from scipy.special import comb

from ..common.models import ConjunctionModel, DisjunctionModel
from ..learners.scm import SetCoveringMachine, fit_many
from .metrics import _get_binary_metrics


def _duplicate_last_element(l, length):
    """Pad a list in place with copies of its last element (utils.py:49-54)."""
    l.extend([l[-1]] * (length - len(l)))
    return l


def _predictions(model, rule_classifications, train_example_idx, test_example_idx, progress_callback=None):
    """Predictions of an SCM model for two sets of genomes (experiment_scm.py:43-100).  Only the k-mers the model
    uses are fetched -- from the HBM-resident matrix here (``get_columns`` on presence-rule indices), from the
    HDF5 file in the reference -- and a copy of the model is re-addressed to that compact matrix."""
    report = progress_callback or (lambda task, fraction: None)
    report("Testing", 0.0)
    if len(model) == 0:                        # empty model: the input only fixes the number of predictions
        train_predictions = model.predict(np.zeros(len(train_example_idx)))
        test_predictions = model.predict(np.zeros(len(test_example_idx)))
    else:
        compact_model = deepcopy(model)
        order = np.argsort([r.kmer_index for r in model.rules])
        columns = []
        for new_column, rule_pos in enumerate(order):
            columns.append(int(compact_model.rules[rule_pos].kmer_index))
            compact_model.rules[rule_pos].kmer_index = new_column
        X = rule_classifications.get_columns(columns)
        train_predictions = compact_model.predict(X[train_example_idx])
        test_predictions = compact_model.predict(X[test_example_idx])
    report("Testing", 1.0)
    return train_predictions, test_predictions


def _greedy_compression_set(presence_by_example):
    """Chvatal's greedy set cover: genomes (rows) chosen until every k-mer (column) is present in one of them."""
    chosen = []
    while presence_by_example.shape[1] != 0:
        best = np.argmax(presence_by_example.sum(axis=1))
        chosen.append(best)
        presence_by_example = presence_by_example[:, presence_by_example[best] == 0]
    return chosen


def _bound(train_predictions, train_answers, train_example_idx, model, delta, max_genome_size, rule_classifications):
    """Sample-compression bound of an SCM model (Marchand & Sokolova 2005), experiment_scm.py:349-398."""
    compression_set = []
    if len(model) > 0:
        kmers = [r.kmer_index for r in model]
        compression_set = _greedy_compression_set(rule_classifications.get_columns(kmers)[train_example_idx])
    h_card = float(len(model))
    Z_card = float(len(compression_set) * max_genome_size)
    m = float(len(train_answers))
    mz = float(len(compression_set))
    wrong = train_predictions != train_answers
    r = float(wrong.sum() - wrong[compression_set].sum())       # errors outside the compression set
    # The reference writes  A + B + 0 if h_card == 0 else C + D : the conditional expression binds looser than
    # "+", so a non-empty model's bound holds ONLY the else branch (no ln C(m, mz), no ln C(m - mz, r)).
    if h_card == 0:
        exponent = ln(comb(m, mz, exact=True)) + ln(comb(m - mz, r, exact=True)) + 0
    else:
        exponent = (h_card * ln(2 * Z_card)) + ln(pi ** 6 * (h_card + 1) ** 2 * (r + 1) ** 2 * (mz + 1) ** 2 /
                                                  (216 * delta))
    return 1.0 - exp((-1.0 / (m - mz - r)) * exponent)


def _tiebreaker(best_utility_idx, rule_risks, model_type):
    """experiment_scm.py:122-130 (== 293-301, 476-484)"""
    tie_rule_risks = rule_risks[best_utility_idx]
    if model_type == "conjunction":
        return best_utility_idx[np.isclose(tie_rule_risks, tie_rule_risks.min())]
    # max instead of min: in the disjunction case the risks are 1.0 - conjunction risks (inverted ys)
    return best_utility_idx[np.isclose(tie_rule_risks, tie_rule_risks.max())]


def _split_labels(labels, example_idx):
    example_idx = np.asarray(example_idx)
    return example_idx[labels[example_idx] == 1].reshape(-1), example_idx[labels[example_idx] == 0].reshape(-1)


def bound_selection(rule_classifications, rules, labels, train_example_idx, rule_risks, model_types, p_values,
                    max_rules, max_equiv_rules, rule_blacklist, bound_delta, bound_max_genome_size,
                    random_generator):
    """``_bound_selection`` (:568-629) over ``_bound_score_hp`` (:401-566): every (model_type, p) is fitted on the
    training set, the bound is evaluated after every added rule, the best length per HP and then the best HP are
    kept.  -> (best_hp_score, best_hp dict, best_model, best_rule_importances, best_equiv_rules, per_hp) where
    per_hp[(model_type, p)] = score_by_length."""
    labels = np.asarray(labels)
    train_example_idx = np.asarray(train_example_idx)
    positive_example_idx, negative_example_idx = _split_labels(labels, train_example_idx)
    train_answers = labels[train_example_idx]
    n_kmers = rule_classifications.shape[1] // 2

    states, learners, fit_args = [], [], []
    for model_type, p in product(model_types, p_values):
        # the reference evaluates every HP combination in a pool task that unpickles ITS OWN copy of the
        # generator (partial argument, :594-603), so each combination starts from the same generator state
        hp_random_generator = deepcopy(random_generator)
        st = {"hp": (model_type, p), "tmp_model": ConjunctionModel() if model_type == "conjunction" else DisjunctionModel(),
              "score_by_length": np.ones(max_rules), "model_by_length": [], "equivalent_rules": [],
              "rule_importances": []}

        def _iteration_callback(iteration_infos, st=st, model_type=model_type, random_generator=hp_random_generator):
            tmp_model = st["tmp_model"]
            tmp_model.add(iteration_infos["selected_rule"])
            st["model_by_length"].append(deepcopy(tmp_model))
            st["rule_importances"].append(iteration_infos["rule_importances"])
            equiv = iteration_infos["equivalent_rules_idx"]
            if len(equiv) > max_equiv_rules:
                random_idx = random_generator.choice(len(equiv), max_equiv_rules, replace=False)
                random_idx.sort()
                equiv = equiv[random_idx]
            if model_type == "disjunction":
                equiv = (equiv + n_kmers) % (2 * n_kmers)
            iteration_infos["equivalent_rules_idx"] = equiv
            st["equivalent_rules"].append(equiv)
            _, train_predictions = _predictions(tmp_model, rule_classifications, [], train_example_idx)
            st["score_by_length"][iteration_infos["iteration_number"] - 1] = _bound(
                train_predictions=train_predictions, train_answers=train_answers, train_example_idx=train_example_idx,
                model=tmp_model, delta=bound_delta, max_genome_size=bound_max_genome_size,
                rule_classifications=rule_classifications)

        states.append(st)
        learners.append(SetCoveringMachine(model_type=model_type, p=p, max_rules=max_rules))
        fit_args.append(dict(rules=rules, positive_example_idx=positive_example_idx,
                             negative_example_idx=negative_example_idx, rule_blacklist=rule_blacklist,
                             tiebreaker=partial(_tiebreaker, rule_risks=rule_risks, model_type=model_type),
                             iteration_callback=_iteration_callback, iteration_rule_importances=True))
    fit_many(learners, fit_args, rule_classifications)

    best_hp_score = 1.0
    best_hp = {"model_type": None, "p": None, "max_rules": None}
    best_model = best_equiv_rules = best_rule_importances = None
    per_hp = {}
    for st in states:
        model_type, p = st["hp"]
        tmp_model = st["tmp_model"]
        per_hp[st["hp"]] = st["score_by_length"]
        if len(tmp_model) == 0:           # no rule was added (:529-547)
            _, train_predictions = _predictions(tmp_model, rule_classifications, [], train_example_idx)
            score = _bound(train_predictions=train_predictions, train_answers=train_answers,
                           train_example_idx=train_example_idx, model=tmp_model, delta=bound_delta,
                           max_genome_size=bound_max_genome_size, rule_classifications=rule_classifications)
            hp, model, importances, equiv = (model_type, p, 0), tmp_model, np.array([]), np.array([])
        else:
            best_score_idx = np.argmin(st["score_by_length"])
            score = st["score_by_length"][best_score_idx]
            hp = (model_type, p, best_score_idx + 1)
            model = st["model_by_length"][best_score_idx]
            importances = st["rule_importances"][best_score_idx]
            equiv = st["equivalent_rules"][: best_score_idx + 1]
        if _bound_hp_is_better(score, hp, best_hp_score, best_hp):
            best_hp["model_type"], best_hp["p"], best_hp["max_rules"] = hp
            best_hp_score, best_model, best_equiv_rules, best_rule_importances = score, model, equiv, importances
    return best_hp_score, best_hp, best_model, best_rule_importances, best_equiv_rules, per_hp


def _bound_hp_is_better(score, hp, best_hp_score, best_hp):
    """The HP replacement rule of ``_bound_selection`` (:612-627): smaller bound, then fewer rules, then p closer to 1.
    The reference runs on Python 2, where ``int < None`` is False and ``int == None`` is False: while no HP has been
    selected yet (best_hp all None) a score EQUAL to the initial 1.0 selects nothing, and a later HP with a smaller
    score is still picked up.  The ``seen`` guard reproduces that instead of Python 3's TypeError."""
    seen = best_hp["max_rules"] is not None
    return bool((score < best_hp_score) or (seen and score == best_hp_score and hp[2] < best_hp["max_rules"]) or
                (seen and score == best_hp_score and hp[2] == best_hp["max_rules"]
                 and abs(1.0 - hp[1]) < abs(1.0 - best_hp["p"])))


def _cv_hp_is_better(score, hp, best_hp_score, best_hp):
    """The HP replacement rule of ``_cross_validation`` (:233-246), with the same Python 2 ``None`` semantics."""
    seen = best_hp["max_rules"] is not None
    close = bool(np.allclose(score, best_hp_score))
    return bool((not close and score < best_hp_score) or (seen and close and hp[2] < best_hp["max_rules"]) or
                (seen and close and hp[2] == best_hp["max_rules"] and not np.allclose(hp[1], best_hp["p"])
                 and abs(1.0 - hp[1]) < abs(1.0 - best_hp["p"])))


def cross_validation(rule_classifications, rules, labels, folds, model_types, p_values, max_rules, rule_blacklist):
    """``_cross_validation`` (:196-248) over ``_cv_score_hp`` (:102-193).  folds: list of (train_example_idx,
    test_example_idx, rule_risks) in the reference's fold order (h5py group-name order).  All folds of all HP
    combinations are fitted in lock-step.  -> (best_hp_score, best_hp dict, per_hp) with
    per_hp[(model_type, p)] = mean risk by model length (0 .. max_rules)."""
    labels = np.asarray(labels)
    jobs, learners, fit_args = [], [], []
    for model_type, p in product(model_types, p_values):
        for i, (train_example_idx, test_example_idx, rule_risks) in enumerate(folds):
            train_example_idx, test_example_idx = np.asarray(train_example_idx), np.asarray(test_example_idx)
            positive_example_idx, negative_example_idx = _split_labels(labels, train_example_idx)
            tmp_model = ConjunctionModel() if model_type == "conjunction" else DisjunctionModel()
            # empty-model predictions (length = 0), before fitting (:165-169)
            preds = [_predictions(tmp_model, rule_classifications, [], test_example_idx)[1]]

            def _iteration_callback(iteration_infos, tmp_model=tmp_model, preds=preds, test_example_idx=test_example_idx):
                tmp_model.add(iteration_infos["selected_rule"])
                preds.append(_predictions(tmp_model, rule_classifications, [], test_example_idx)[1])

            jobs.append(((model_type, p), i, preds, test_example_idx))
            learners.append(SetCoveringMachine(model_type=model_type, p=p, max_rules=max_rules))
            fit_args.append(dict(rules=rules, positive_example_idx=positive_example_idx,
                                 negative_example_idx=negative_example_idx, rule_blacklist=rule_blacklist,
                                 tiebreaker=partial(_tiebreaker, rule_risks=rule_risks, model_type=model_type),
                                 iteration_callback=_iteration_callback))
    fit_many(learners, fit_args, rule_classifications)

    fold_scores = {hp: np.ones((len(folds), max_rules + 1)) * np.inf for hp in product(model_types, p_values)}
    for hp, i, preds, test_example_idx in jobs:
        by_length = np.array(_duplicate_last_element(preds, max_rules + 1))
        fold_scores[hp][i] = _get_binary_metrics(predictions=by_length, answers=labels[test_example_idx])["risk"]

    best_hp_score = 1.0
    best_hp = {"model_type": None, "p": None, "max_rules": None}
    per_hp = {}
    for hp_mt_p in product(model_types, p_values):
        score_by_model_length = np.mean(fold_scores[hp_mt_p], axis=0)
        per_hp[hp_mt_p] = score_by_model_length
        best_score_idx = np.argmin(score_by_model_length)
        score, hp = score_by_model_length[best_score_idx], (hp_mt_p[0], hp_mt_p[1], best_score_idx)
        if _cv_hp_is_better(score, hp, best_hp_score, best_hp):
            best_hp["model_type"], best_hp["p"], best_hp["max_rules"] = hp
            best_hp_score = score
    return best_hp_score, best_hp, per_hp


# ------------------------------------------------------------------------------------------------ dataset-bound driver
def _fasta_to_sequences(path):
    """All the sequences of a FASTA file, in order (utils.py:57-74)."""
    sequences, current = [], []
    with open(path, "r") as fh:
        for line in fh:
            line = line.strip()
            if line.startswith(">"):
                if current:
                    sequences.append("".join(current))
                current = []
            elif line:
                current.append(line)
    if current:
        sequences.append("".join(current))
    return sequences


def _parse_kmer_blacklist(blacklist_path, expected_kmer_len):
    """k-mers to blacklist: a FASTA file or one k-mer per line (utils.py:189-213)."""
    if any(blacklist_path.endswith(ext) for ext in (".fasta", ".fa", ".fas", ".fna")):
        kmers = _fasta_to_sequences(blacklist_path)
    else:
        with open(blacklist_path, "r") as fh:
            kmers = [line.rstrip("\n") for line in fh]
        kmers = [k for k in kmers if k]
    for kmer in kmers:
        if set(kmer).difference("ACGTacgt"):
            raise ValueError("{} is not a valid DNA sequence".format(kmer))
    if not all(len(kmer) == expected_kmer_len for kmer in kmers):
        raise ValueError("Extracted k-mers to blacklist do not have all the same length as the dataset k-mers")
    return kmers


def _as_text(x):
    return x.decode("ascii") if isinstance(x, bytes) else str(x)


def _find_rule_blacklist(dataset, kmer_blacklist_file, warning_callback):
    """Indices of the presence AND absence rules of the k-mers listed in ``kmer_blacklist_file``
    (experiment_scm.py:632-672; a dictionary lookup replaces the reference's two ``list.index`` scans per k-mer)."""
    if kmer_blacklist_file is None:
        return []
    to_blacklist = _parse_kmer_blacklist(kmer_blacklist_file, dataset.kmer_length)
    if not to_blacklist:
        return []
    sequences = [_as_text(s) for s in dataset.kmer_sequences[...].tolist()]   # (assumed upper-cased in the dataset)
    column_of_sequence_idx = dict((int(seq_idx), col) for col, seq_idx in
                                  reversed(list(enumerate(dataset.kmer_by_matrix_column[...].tolist()))))
    first_idx = {}
    for i, s in enumerate(sequences):
        first_idx.setdefault(s, i)
    n_kmers = len(sequences)
    rule_blacklist, not_found = [], []
    for kmer in to_blacklist:
        kmer = kmer.upper()
        column = column_of_sequence_idx.get(first_idx.get(kmer, -1))
        if column is None:
            not_found.append(kmer)
        else:
            rule_blacklist += [column, column + n_kmers]
    if not_found:
        warning_callback("The following kmers could not be found in the dataset: " + ", ".join(not_found))
    return rule_blacklist


def _full_train(rule_classifications, rules, labels, split_train_idx, rule_risks, model_type, p, max_rules,
                max_equiv_rules, rule_blacklist, random_generator, progress_callback):
    """One SCM on the whole training set with the selected hyper-parameters (experiment_scm.py:250-347)."""
    predictor = SetCoveringMachine(model_type=model_type, p=p, max_rules=max_rules)
    if max_rules == 0:                 # cross-validation may prefer the empty model: nothing to train
        return predictor.model, np.array([]), np.array([])
    n_kmers = rule_classifications.shape[1] // 2
    equivalent_rules, done = [], {"n_rules": 0.0}

    def _iteration_callback(iteration_infos):
        done["n_rules"] += 1
        progress_callback("Training", done["n_rules"] / max_rules)
        equiv = iteration_infos["equivalent_rules_idx"]
        if len(equiv) > max_equiv_rules:
            random_idx = random_generator.choice(len(equiv), max_equiv_rules, replace=False)
            random_idx.sort()
            equiv = equiv[random_idx]
        if model_type == "disjunction":
            equiv = (equiv + n_kmers) % (2 * n_kmers)
        iteration_infos["equivalent_rules_idx"] = equiv
        equivalent_rules.append(equiv)

    positive_example_idx, negative_example_idx = _split_labels(labels, split_train_idx)
    progress_callback("Training", 0)
    predictor.fit(rules=rules, rule_classifications=rule_classifications, positive_example_idx=positive_example_idx,
                  negative_example_idx=negative_example_idx, rule_blacklist=rule_blacklist,
                  tiebreaker=partial(_tiebreaker, rule_risks=rule_risks, model_type=model_type),
                  iteration_callback=_iteration_callback)
    return predictor.model, predictor.rule_importances, equivalent_rules


def learn_SCM(dataset_file, split_name, model_type, p, kmer_blacklist_file, max_rules, max_equiv_rules,
              parameter_selection, n_cpu, random_seed, authorized_rules=None, bound_delta=None,
              bound_max_genome_size=None, progress_callback=None, warning_callback=None, error_callback=None,
              devices=None, comm=None, rule_classifications=None):
    """``kover learn scm`` on a Kover dataset file: same arguments and return value as the reference's
    ``learn_SCM`` (experiment_scm.py:675-888).  parameter_selection: "bound", "cv" or "none" (first value of each
    hyper-parameter).  What differs is the execution: the packed ``kmer_matrix`` is streamed into HBM ONCE and every
    fit of the selection (all hyper-parameter combinations, all folds) runs in lock-step over that resident image,
    instead of one forked worker per combination, each re-reading the HDF5 chunks on every scan; ``n_cpu`` is
    accepted and ignored.  ``rule_classifications``: an already resident KmerRuleClassifications of this dataset to
    adopt instead of uploading again (several experiments on one dataset); it is then left open.
    -> (best_hp, best_hp_score, train_metrics, test_metrics, model, rule_importances, model_equivalent_rules,
    classifications)."""
    from collections import defaultdict

    from ...dataset.ds import KoverDataset
    from ..common.rules import KmerRuleClassifications, LazyKmerRuleList
    warning_callback = warning_callback or (lambda w: logging.warning(w))
    progress_callback = progress_callback or (lambda task, fraction: None)
    if error_callback is None:
        def error_callback(exception):
            raise exception
    random_generator = np.random.RandomState(random_seed)
    model_type, p = np.unique(model_type), np.unique(p)

    dataset = KoverDataset(dataset_file)
    rule_blacklist = _find_rule_blacklist(dataset, kmer_blacklist_file, warning_callback)
    split = dataset.get_split(split_name)
    labels = np.asarray(dataset.phenotype.metadata[...])
    train_example_idx = np.asarray(split.train_genome_idx[...])
    test_example_idx = np.asarray(split.test_genome_idx[...])
    rules = LazyKmerRuleList(dataset.kmer_sequences[...], dataset.kmer_by_matrix_column[...])
    owns_matrix = rule_classifications is None
    if owns_matrix:
        rule_classifications = KmerRuleClassifications(dataset.kmer_matrix, dataset.genome_count, devices=devices,
                                                       comm=comm)
    split_risks = np.hstack((split.unique_risk_by_kmer[...], split.unique_risk_by_anti_kmer[...]))
    model_types, p_values = [str(m) for m in model_type], [float(x) for x in p]

    best_model = best_importances = best_equiv = None
    if parameter_selection == "bound":
        if bound_delta is None or bound_max_genome_size is None:
            error_callback(Exception("Bound selection cannot be performed without delta and the maximum genome length."))
        best_hp_score, best_hp, best_model, best_importances, best_equiv, _ = bound_selection(
            rule_classifications, rules, labels, train_example_idx, split_risks, model_types, p_values, max_rules,
            max_equiv_rules, rule_blacklist, bound_delta, bound_max_genome_size, random_generator)
    elif parameter_selection == "cv":
        if len(split.folds) < 1:
            error_callback(Exception("Cross-validation cannot be performed on a split with no folds."))
        folds = [(np.asarray(f.train_genome_idx[...]), np.asarray(f.test_genome_idx[...]),
                  np.hstack((f.unique_risk_by_kmer[...], f.unique_risk_by_anti_kmer[...]))) for f in split.folds]
        best_hp_score, best_hp, _ = cross_validation(rule_classifications, rules, labels, folds, model_types, p_values,
                                                     max_rules, rule_blacklist)
    else:
        best_hp = {"model_type": model_types[0], "p": p_values[0], "max_rules": max_rules}
        best_hp_score = None

    if parameter_selection == "bound":
        model, rule_importances, equivalent_rules = best_model, best_importances, best_equiv
    else:
        model, rule_importances, equivalent_rules = _full_train(
            rule_classifications, rules, labels, train_example_idx, split_risks, best_hp["model_type"], best_hp["p"],
            best_hp["max_rules"], max_equiv_rules, rule_blacklist, random_generator, progress_callback)

    train_predictions, test_predictions = _predictions(model, rule_classifications, train_example_idx,
                                                       test_example_idx, progress_callback)
    train_answers = labels[train_example_idx]
    train_metrics = _get_binary_metrics(train_predictions, train_answers)
    if parameter_selection == "bound":
        train_metrics["bound"] = best_hp_score
    else:
        train_metrics["bound"] = _bound(train_predictions=train_predictions, train_answers=train_answers,
                                        train_example_idx=train_example_idx, model=model, delta=bound_delta,
                                        max_genome_size=bound_max_genome_size,
                                        rule_classifications=rule_classifications)
    test_metrics = None
    if len(test_example_idx) > 0:
        test_answers = labels[test_example_idx]
        test_metrics = _get_binary_metrics(test_predictions, test_answers)

    genome_ids = np.array([_as_text(g) for g in dataset.genome_identifiers[...].tolist()], dtype=object)
    classifications = defaultdict(list)
    classifications["train_correct"] = genome_ids[train_example_idx[train_predictions == train_answers]].tolist() \
        if train_metrics["risk"][0] < 1.0 else []
    classifications["train_errors"] = genome_ids[train_example_idx[train_predictions != train_answers]].tolist() \
        if train_metrics["risk"][0] > 0 else []
    if len(test_example_idx) > 0:
        classifications["test_correct"] = genome_ids[test_example_idx[test_predictions == test_answers]].tolist() \
            if test_metrics["risk"][0] < 1.0 else []
        classifications["test_errors"] = genome_ids[test_example_idx[test_predictions != test_answers]].tolist() \
            if test_metrics["risk"][0] > 0 else []

    model_equivalent_rules = [[rules[i] for i in equiv_idx] for equiv_idx in equivalent_rules]
    if owns_matrix:
        rule_classifications.close()
    dataset.close()
    return (best_hp, best_hp_score, train_metrics, test_metrics, model, rule_importances, model_equivalent_rules,
            classifications)
