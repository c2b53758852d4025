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
        if (score < best_hp_score) or (score == best_hp_score and hp[2] < best_hp["max_rules"]) or \
                (score == best_hp_score and hp[2] == best_hp["max_rules"]
                 and abs(1.0 - hp[1]) < abs(1.0 - best_hp["p"])):              # :612-627
            best_hp["model_type"], best_hp["p"], best_hp["max_rules"] = hp
            best_hp_score, best_model, best_equiv_rules, best_rule_importances = score, model, equiv, importances
    return best_hp_score, best_hp, best_model, best_rule_importances, best_equiv_rules, per_hp


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
        if (not np.allclose(score, best_hp_score) and score < best_hp_score) or \
                (np.allclose(score, best_hp_score) and hp[2] < best_hp["max_rules"]) or \
                (np.allclose(score, best_hp_score) and hp[2] == best_hp["max_rules"]
                 and not np.allclose(hp[1], best_hp["p"]) and abs(1.0 - hp[1]) < abs(1.0 - best_hp["p"])):   # :233-246
            best_hp["model_type"], best_hp["p"], best_hp["max_rules"] = hp
            best_hp_score = score
    return best_hp_score, best_hp, per_hp
