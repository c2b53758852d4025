"""Bound-based selection of a pruned CART tree over a device-resident matrix.

Mirrors the pure parts of reference learning/experiments/experiment_cart.py: ``_tiebreaker`` (:82-94),
``_split_callback`` (:97-106), ``_predictions`` (:120-152), ``_bound`` (:155-205) and the body of
``_learn_pruned_tree_bound`` (:208-294) without the HDF5 dataset object around it.
"""
import logging
from copy import deepcopy
from functools import partial
from math import exp, log as ln, pi

import numpy as np
from scipy.special import comb

from ..learners.cart import DecisionTreeClassifier, _prune_tree


def _tiebreaker(best_score_idx, rule_kmer_occurrences):
    """Prefer the k-mers with the most occurrences in the training set (:82-94)."""
    tie_rule_kmer_occurrences = rule_kmer_occurrences[best_score_idx]
    return best_score_idx[np.isclose(tie_rule_kmer_occurrences, tie_rule_kmer_occurrences.max())]


def _split_callback(node, equivalent_rules_idx):
    node.rule.equivalent_rules_idx = equivalent_rules_idx


def _readdress_tree(tree, rule_new_idx_by_kmer_seq):
    def _readdress(node, kmer_idx):
        if node.rule is not None:
            node.rule.kmer_index = kmer_idx[node.rule.kmer_sequence]
            _readdress(node.left_child, kmer_idx)
            _readdress(node.right_child, kmer_idx)
    new_tree = deepcopy(tree)
    _readdress(new_tree, rule_new_idx_by_kmer_seq)
    return new_tree


def _predictions(decision_tree, rule_classifications, train_example_idx, test_example_idx, progress_callback=None):
    """experiment_cart.py:120-152; the model's k-mer columns come from the HBM-resident matrix."""
    if progress_callback is None:
        progress_callback = lambda t, p: None
    progress_callback("Testing", 0.0)
    if len(decision_tree.rules) > 0:
        model_rules = decision_tree.rules
        kmer_idx_by_rule = np.array([r.kmer_index for r in model_rules])
        kmer_sequence_by_rule = np.array([r.kmer_sequence for r in model_rules])
        sort_by_idx = np.argsort(kmer_idx_by_rule)
        kmer_idx_by_rule = kmer_idx_by_rule[sort_by_idx]
        kmer_sequence_by_rule = kmer_sequence_by_rule[sort_by_idx]
        readdressed = _readdress_tree(decision_tree, dict((s, i) for i, s in enumerate(kmer_sequence_by_rule)))
        X = rule_classifications.get_columns([int(c) for c in kmer_idx_by_rule])
        train_predictions = readdressed.predict(X[train_example_idx])
        test_predictions = readdressed.predict(X[test_example_idx])
    else:                                      # the model is just a leaf
        train_predictions = decision_tree.predict(np.empty((len(train_example_idx), 1)))
        test_predictions = decision_tree.predict(np.empty((len(test_example_idx), 1)))
    progress_callback("Testing", 1.0)
    return train_predictions, test_predictions


def _bound(train_predictions, train_answers, train_example_idx, model, delta, max_genome_size, rule_classifications,
           n_classes):
    """Sample-compression bound of a decision tree (Drouin et al. 2017), experiment_cart.py:155-205."""
    compression_set = []
    if len(model.rules) > 0:
        presence_by_example = rule_classifications.get_columns([r.kmer_index for r in model.rules])[train_example_idx]
        while presence_by_example.shape[1] != 0:
            score = presence_by_example.sum(axis=1)
            best_example_relative_idx = np.argmax(score)
            compression_set.append(best_example_relative_idx)
            presence_by_example = presence_by_example[:, presence_by_example[best_example_relative_idx] == 0]
    m = float(len(train_answers))
    Z_card = float(len(compression_set))
    N_Z = Z_card * max_genome_size
    r = float((train_predictions != train_answers).sum()
              - (train_predictions[compression_set] != train_answers[compression_set]).sum())
    n = float(len(model.rules))
    return 1.0 - exp((-1.0 / (m - Z_card - r)) * (ln(comb(m, Z_card, exact=True)) +
                                                  ln(comb(m - Z_card, r, exact=True)) +
                                                  (n * ln(N_Z) if n > 0 else 0.) +
                                                  (n + 1) * ln(n_classes) +
                                                  ln(comb(2 * n + 1, n, exact=True)) +
                                                  ln(pi ** 6 * (n + 1) ** 2 * (r + 1) ** 2 * (Z_card + 1) ** 2 /
                                                     (216 * delta))))


def learn_pruned_tree_bound(rule_classifications, rules, example_labels, train_genome_idx, n_classes, hps, delta,
                            max_genome_size, rule_blacklist):
    """``_learn_pruned_tree_bound`` (:208-294): grow the master tree on the training set (split scans on the device),
    prune it by minimal cost-complexity and keep the most pruned tree with the smallest bound.
    -> (hps with "pruning_alpha", min bound, tree, [(alpha, bound)] for the whole pruning sequence)."""
    example_labels = np.asarray(example_labels)
    train_genome_idx = np.asarray(train_genome_idx)
    hps = dict(hps)
    master_predictor = DecisionTreeClassifier(criterion=hps["criterion"], max_depth=hps["max_depth"],
                                              min_samples_split=hps["min_samples_split"],
                                              class_importance=hps["class_importance"])
    master_predictor.fit(rules=rules, rule_classifications=rule_classifications,
                         example_idx={c: train_genome_idx[example_labels[train_genome_idx] == c]
                                      for c in range(n_classes)},
                         rule_blacklist=rule_blacklist,
                         tiebreaker=partial(_tiebreaker,
                                            rule_kmer_occurrences=rule_classifications.sum_rows(train_genome_idx)),
                         level_callback=None, split_callback=_split_callback)
    min_score, min_score_tree = np.inf, None
    train_answers = example_labels[train_genome_idx]
    sequence = []
    for alpha, tree in zip(*_prune_tree(master_predictor.decision_tree)):
        train_predictions = _predictions(tree, rule_classifications, train_genome_idx, [])[0]
        bound_value = _bound(train_predictions=train_predictions, train_answers=train_answers,
                             train_example_idx=train_genome_idx, model=tree, delta=delta,
                             max_genome_size=max_genome_size, rule_classifications=rule_classifications,
                             n_classes=n_classes)
        sequence.append((alpha, bound_value))
        if bound_value <= min_score:       # alphas increase: ties prefer the more pruned tree
            min_score, min_score_tree = bound_value, tree
            hps["pruning_alpha"] = alpha
    return hps, min_score, min_score_tree, sequence
