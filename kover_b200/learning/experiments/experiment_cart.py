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
    """Among equally good splits keep the k-mers that occur in the most training genomes (:82-94): they are less
    likely to isolate a phylogenetic clade and give smaller compression sets."""
    occurrences = rule_kmer_occurrences[best_score_idx]
    return best_score_idx[np.isclose(occurrences, occurrences.max())]


def _split_callback(node, equivalent_rules_idx):
    """Remember the equivalent rules of a split on its rule object (:97-106)."""
    node.rule.equivalent_rules_idx = equivalent_rules_idx


def _readdress_tree(tree, rule_new_idx_by_kmer_seq):
    """Copy of the tree whose rules point at columns of a compact matrix (k-mer sequence -> new column)."""
    clone = deepcopy(tree)
    for rule in clone.rules:
        rule.kmer_index = rule_new_idx_by_kmer_seq[rule.kmer_sequence]
    return clone


def _predictions(decision_tree, rule_classifications, train_example_idx, test_example_idx, progress_callback=None):
    """Class predictions of a tree for two sets of genomes (:120-152).  Only the k-mers the tree uses are fetched
    -- from the HBM-resident matrix here, from the HDF5 file in the reference -- and the tree is re-addressed to
    that compact matrix."""
    report = progress_callback or (lambda task, fraction: None)
    report("Testing", 0.0)
    used = decision_tree.rules
    if len(used) == 0:                         # a single leaf: the input only fixes the number of predictions
        train_predictions = decision_tree.predict(np.empty((len(train_example_idx), 1)))
        test_predictions = decision_tree.predict(np.empty((len(test_example_idx), 1)))
    else:
        order = np.argsort(np.array([r.kmer_index for r in used]))
        columns = [int(used[i].kmer_index) for i in order]
        compact_tree = _readdress_tree(decision_tree, dict((used[i].kmer_sequence, j) for j, i in enumerate(order)))
        X = rule_classifications.get_columns(columns)
        train_predictions = compact_tree.predict(X[train_example_idx])
        test_predictions = compact_tree.predict(X[test_example_idx])
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


def _bound(train_predictions, train_answers, train_example_idx, model, delta, max_genome_size, rule_classifications,
           n_classes):
    """Sample-compression bound of a decision tree (Drouin et al. 2017), experiment_cart.py:155-205."""
    compression_set = []
    if len(model.rules) > 0:
        kmers = [r.kmer_index for r in model.rules]
        compression_set = _greedy_compression_set(rule_classifications.get_columns(kmers)[train_example_idx])
    m = float(len(train_answers))
    Z_card = float(len(compression_set))
    N_Z = Z_card * max_genome_size                     # nucleotides in the compression set
    wrong = train_predictions != train_answers
    r = float(wrong.sum() - wrong[compression_set].sum())       # errors outside the compression set
    n = float(len(model.rules))
    # the terms are added in the reference's order (floating point)
    exponent = (ln(comb(m, Z_card, exact=True)) +
                ln(comb(m - Z_card, r, exact=True)) +
                (n * ln(N_Z) if n > 0 else 0.) +
                (n + 1) * ln(n_classes) +
                ln(comb(2 * n + 1, n, exact=True)) +
                ln(pi ** 6 * (n + 1) ** 2 * (r + 1) ** 2 * (Z_card + 1) ** 2 / (216 * delta)))
    return 1.0 - exp((-1.0 / (m - Z_card - r)) * exponent)


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
