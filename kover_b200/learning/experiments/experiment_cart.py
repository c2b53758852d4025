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

from ..learners.cart import DecisionTreeClassifier, _prune_tree, _to_leaf


def _tiebreaker(best_score_idx, rule_kmer_occurrences):
    """Among equally good splits keep the k-mers that occur in the most training genomes (:82-94): they are less
    likely to isolate a phylogenetic clade and give smaller compression sets."""
    occurrences = rule_kmer_occurrences[best_score_idx]
    return best_score_idx[np.isclose(occurrences, occurrences.max())]


class OccurrenceTiebreaker(object):
    """``_tiebreaker`` (:82-94) bound to the k-mer occurrences of a training set (:264 computes them with
    ``sum_rows(train)``).  When the rule_classifications can keep that table on the device, the tree learner has
    the split search apply the tiebreaker there (``device_occurrences``): the millions of ties of a small node
    never reach the host.  Calling the object applies it on the host (same result)."""

    def __init__(self, rule_classifications, train_genome_idx):
        self._rc, self._rows = rule_classifications, np.asarray(train_genome_idx)
        make = getattr(rule_classifications, "occurrence_table", None)
        table = make(self._rows) if make is not None else None
        self.device_occurrences = table if table is not None and table.table is not None else None
        self._host = None

    def _occurrences(self):
        if self._host is None:
            self._host = self.device_occurrences.host() if self.device_occurrences is not None \
                else self._rc.sum_rows(self._rows)
        return self._host

    def __call__(self, best_score_idx):
        return _tiebreaker(best_score_idx, self._occurrences())

    def release(self):
        if self.device_occurrences is not None:
            self.device_occurrences.release()
            self.device_occurrences = None


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


def _truncate_tree(tree, max_depth, min_samples_split):
    """The tree ``DecisionTreeClassifier(max_depth, min_samples_split)`` grows, derived from one grown deeper and / or
    with a smaller ``min_samples_split`` on the same examples: the breadth-first growth (cart.py:260-339) splits a
    node from its own examples only, so the shallower tree IS the deeper one cut at ``max_depth`` and below the nodes
    that are too small to split.  A hyper-parameter grid over (max_depth, min_samples_split) therefore needs ONE
    growth per (criterion, class importances, training set) instead of one per combination."""
    clone = deepcopy(tree)
    min_split = max(int(min_samples_split), 2)
    pending = [clone]
    while pending:
        node = pending.pop()
        if node.is_leaf:
            continue
        if node.depth >= max_depth or node.n_examples < min_split:
            _to_leaf(node)
        else:
            pending.extend((node.left_child, node.right_child))
    return clone


def _grown_trees(grown, kind, hps, grow):
    """Trees of ``grow(max_depth, min_samples_split)`` for ``hps``.  ``grown`` (or None): {"max_depth": the largest
    depth of the grid, "min_samples_split": its smallest value, "trees": {}} -- the trees are then grown once at
    those values per (kind, criterion, class importances) and cut down to ``hps`` (_truncate_tree)."""
    if grown is None:
        return grow(hps["max_depth"], hps["min_samples_split"])
    key = (kind, hps["criterion"], tuple(sorted(hps["class_importance"].items())))
    if key not in grown["trees"]:
        grown["trees"][key] = grow(grown["max_depth"], grown["min_samples_split"])
    return [_truncate_tree(t, hps["max_depth"], hps["min_samples_split"]) for t in grown["trees"][key]]


def learn_pruned_tree_bound(rule_classifications, rules, example_labels, train_genome_idx, n_classes, hps, delta,
                            max_genome_size, rule_blacklist, grown=None):
    """``_learn_pruned_tree_bound`` (:208-294): grow the master tree on the training set (split scans on the device),
    prune it by minimal cost-complexity and keep the most pruned tree with the smallest bound.
    grown: see _grown_trees (one growth for all the (max_depth, min_samples_split) of a grid).
    -> (hps with "pruning_alpha", min bound, tree, [(alpha, bound)] for the whole pruning sequence)."""
    example_labels = np.asarray(example_labels)
    train_genome_idx = np.asarray(train_genome_idx)
    hps = dict(hps)

    def grow(max_depth, min_samples_split):
        master_predictor = DecisionTreeClassifier(criterion=hps["criterion"], max_depth=max_depth,
                                                  min_samples_split=min_samples_split,
                                                  class_importance=hps["class_importance"])
        tiebreaker = OccurrenceTiebreaker(rule_classifications, train_genome_idx)
        try:                                 # the device occurrence table goes back even when the fit raises
            master_predictor.fit(rules=rules, rule_classifications=rule_classifications,
                                 example_idx={c: train_genome_idx[example_labels[train_genome_idx] == c]
                                              for c in range(n_classes)},
                                 rule_blacklist=rule_blacklist, tiebreaker=tiebreaker,
                                 level_callback=None, split_callback=_split_callback)
        finally:
            tiebreaker.release()
        return [master_predictor.decision_tree]

    master_tree = _grown_trees(grown, "bound", hps, grow)[0]
    min_score, min_score_tree = np.inf, None
    train_answers = example_labels[train_genome_idx]
    sequence = []
    for alpha, tree in zip(*_prune_tree(master_tree)):
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


class _AlphaIntervals(object):
    """Risk of a fold's pruned trees as a function of the pruning parameter: tree j is the fold's choice for every
    alpha in [alpha_j, alpha_j+1), the last one up to infinity.  Lookup semantics of the reference's ``BetweenDict``
    (experiment_cart.py:43-80): the FIRST interval, in insertion order, that contains the key; an interval whose
    bounds are not strictly increasing is an error."""

    def __init__(self):
        self._intervals = []

    def add(self, low, high, value):
        if not low < high:
            raise RuntimeError("First element of a BetweenDict key must be strictly less than the second element. "
                               "Got [%.6f, %.6f]" % (low, high))
        for i, (lo, hi, _) in enumerate(self._intervals):
            if lo == low and hi == high:                      # same key: the value is replaced, the position kept
                self._intervals[i] = (low, high, value)
                return
        self._intervals.append((low, high, value))

    def __getitem__(self, key):
        for low, high, value in self._intervals:
            if low <= key < high or (low <= key and high == np.inf) or (low == -np.inf and key < high):
                return value
        raise KeyError("Key '%s' is not between any values in the BetweenDict" % key)


def learn_pruned_tree_cv(rule_classifications, rules, example_labels, train_genome_idx, folds, n_classes, hps,
                         rule_blacklist, grown=None):
    """``_learn_pruned_tree_cv`` (:297-436): cost-complexity pruning with the pruning parameter chosen by k-fold
    cross-validation (Breiman et al. 1984, section 3.4).  folds: [(train_genome_idx, test_genome_idx)] in the
    reference's fold order.  The fold trees and the master tree are grown in LOCK-STEP over the resident matrix
    (learners.cart.fit_many: at every depth the nodes of all the trees share matrix passes); their k-mer
    occurrence tables (the tiebreaker's input) stay on the device, where the tiebreaker is applied.
    grown: see _grown_trees (one growth for all the (max_depth, min_samples_split) of a grid).
    -> (hps with "pruning_alpha", cross-validation score, pruned master tree)."""
    from math import sqrt

    from ..learners.cart import fit_many
    from .metrics import _get_binary_metrics
    example_labels = np.asarray(example_labels)
    train_genome_idx = np.asarray(train_genome_idx)
    folds = [(np.asarray(tr), np.asarray(te)) for tr, te in folds]
    hps = dict(hps)

    training_sets = [tr for tr, _ in folds] + [train_genome_idx]

    def grow(max_depth, min_samples_split):
        learners, fit_args, tiebreakers = [], [], []
        try:                                 # the device occurrence tables go back even when a fit raises
            for i, tr in enumerate(training_sets):
                # the tiebreaker's table is sum_rows(train) (:345, :364): kept on the device when the matrix lives there
                tiebreakers.append(OccurrenceTiebreaker(rule_classifications, tr))
                learners.append(DecisionTreeClassifier(criterion=hps["criterion"], max_depth=max_depth,
                                                       min_samples_split=min_samples_split,
                                                       class_importance=hps["class_importance"]))
                fit_args.append(dict(rules=rules,
                                     example_idx={c: tr[example_labels[tr] == c] for c in range(n_classes)},
                                     rule_blacklist=rule_blacklist, tiebreaker=tiebreakers[-1], level_callback=None,
                                     split_callback=_split_callback if i == len(folds) else None))
            fit_many(learners, fit_args, rule_classifications)
        finally:
            for tb in tiebreakers:
                tb.release()
        return [learner.decision_tree for learner in learners]

    trees_grown = _grown_trees(grown, "cv", hps, grow)
    fold_trees, master_tree = trees_grown[:-1], trees_grown[-1]

    master_alphas, master_pruned_trees = _prune_tree(master_tree)
    fold_scores_by_alpha = []
    for (_, fold_test_idx), fold_tree in zip(folds, fold_trees):
        alphas, trees = _prune_tree(fold_tree)
        answers = example_labels[fold_test_idx]
        by_alpha = _AlphaIntervals()
        for j, tree in enumerate(trees):
            risk = _get_binary_metrics(predictions=_predictions(tree, rule_classifications, [], fold_test_idx)[1],
                                       answers=answers)["risk"][0]
            by_alpha.add(alphas[j], alphas[j + 1] if j < len(alphas) - 1 else np.inf, risk)
        fold_scores_by_alpha.append(by_alpha)

    min_score, min_score_tree = np.inf, None
    for i, tree in enumerate(master_pruned_trees):
        geo_mean_alpha = sqrt(master_alphas[i] * master_alphas[i + 1]) if i < len(master_alphas) - 1 else np.inf
        cv_score = np.mean([scores[geo_mean_alpha] for scores in fold_scores_by_alpha])
        if cv_score <= min_score:          # alphas increase: ties prefer the more pruned tree
            min_score, min_score_tree = cv_score, tree
            hps["pruning_alpha"] = geo_mean_alpha
    return hps, min_score, min_score_tree


def train_tree(hp_search, criterion, class_importance, max_depth, min_samples_split, progress_callback,
               hp_search_type):
    """Best hyper-parameter combination according to ``hp_search(hps) -> (hps, score, pruned master tree)``, visited
    in the reference's ``--n-cpu 1`` order ``product(criterion, class_importance, max_depth, min_samples_split)``
    (:437-487).  Ties (``isclose``) go to the smaller tree, then to the less spread class importances -- and, like
    the reference (:480-483), a tie replaces the hyper-parameters and the score but KEEPS the earlier tree."""
    from itertools import product
    combinations = list(product(criterion, class_importance, max_depth, min_samples_split))
    best_hps, best_score, best_master_tree = None, np.inf, None
    progress_callback(hp_search_type.title(), 0.0)
    for done, (crit, importance, depth, mss) in enumerate(combinations):
        hps, score, master_tree = hp_search({"criterion": crit, "class_importance": importance, "max_depth": depth,
                                             "min_samples_split": mss})
        progress_callback(hp_search_type.title(), (done + 1.0) / len(combinations))
        if score < best_score:
            best_hps, best_score, best_master_tree = hps, score, master_tree
        elif np.isclose(score, best_score):
            n_new, n_best = len(master_tree), len(best_master_tree)
            if n_new < n_best or (n_new == n_best and np.var(list(hps["class_importance"].values())) <
                                  np.var(list(best_hps["class_importance"].values()))):
                best_hps, best_score = hps, score
    return best_score, best_hps, best_master_tree


def _find_rule_blacklist(dataset, kmer_blacklist_file, warning_callback):
    """Presence-rule indices of the k-mers to blacklist (:490-518: a tree only uses presence rules)."""
    from .experiment_scm import _find_rule_blacklist as _scm_blacklist
    n_kmers = dataset.kmer_count
    return [r for r in _scm_blacklist(dataset, kmer_blacklist_file, warning_callback) if r < n_kmers]


def learn_CART(dataset_file, split_name, criterion, max_depth, min_samples_split, class_importance, bound_delta,
               bound_max_genome_size, kmer_blacklist_file, parameter_selection, n_cpu, authorized_rules=None,
               progress_callback=None, warning_callback=None, error_callback=None, devices=None, comm=None,
               rule_classifications=None):
    """``kover learn tree`` on a Kover dataset file: same arguments and return value as the reference's
    ``learn_CART`` (:521-649).  The packed matrix is streamed into HBM once; every hyper-parameter combination is
    evaluated over that resident image (``n_cpu`` is accepted and ignored).  ``rule_classifications``: an already
    resident KmerRuleClassifications of this dataset to adopt instead of uploading again; it is then left open.
    -> (best_hps, best_hp_score, train_metrics, test_metrics, model, rule_importances, model_equivalent_rules,
    classifications)."""
    from collections import defaultdict

    from ...dataset.ds import KoverDataset
    from ..common.models import CARTModel
    from ..common.rules import KmerRuleClassifications, LazyKmerRuleList
    from .experiment_scm import _as_text
    from .metrics import _get_binary_metrics, _get_multiclass_metrics
    warning_callback = warning_callback or (lambda w: logging.warning(w))
    progress_callback = progress_callback or (lambda task, fraction: None)
    if error_callback is None:
        def error_callback(exception):
            raise exception

    dataset = KoverDataset(dataset_file)
    rule_blacklist = _find_rule_blacklist(dataset, kmer_blacklist_file, warning_callback)
    criterion, max_depth = np.unique(criterion), np.unique(max_depth)
    min_samples_split = np.unique(min_samples_split)
    # np.unique(class_importance) (:545) sorts the distinct dictionaries with Python 2's dict ordering: by size, then
    # by the value at the smallest key where they differ -- for dictionaries over the same classes that is the order
    # of their sorted items.  The order decides which HP combination wins train_tree's first-wins ties.
    importances = []
    for ci in class_importance:
        if ci not in importances:
            importances.append(ci)
    importances.sort(key=lambda ci: (len(ci), sorted(ci.items())))

    split = dataset.get_split(split_name)
    train_idx, test_idx = np.asarray(split.train_genome_idx[...]), np.asarray(split.test_genome_idx[...])
    example_labels = np.asarray(dataset.phenotype.metadata[...])
    phenotype_tags = dataset.phenotype.tags[...]
    n_classes = len(phenotype_tags)
    rules = LazyKmerRuleList(dataset.kmer_sequences, dataset.kmer_by_matrix_column)
    owns_matrix = rule_classifications is None
    if owns_matrix:
        rule_classifications = KmerRuleClassifications(dataset.kmer_matrix, dataset.genome_count, devices=devices,
                                                       comm=comm)

    # the trees of a (max_depth, min_samples_split) grid are cuts of ONE tree grown at the largest depth and the
    # smallest min_samples_split (the reference grows every combination separately, :437-464)
    grown = {"max_depth": int(max(max_depth)), "min_samples_split": int(min(min_samples_split)), "trees": {}}
    if parameter_selection == "bound":
        def hp_search(hps):
            return learn_pruned_tree_bound(rule_classifications, rules, example_labels, train_idx, n_classes, hps,
                                           bound_delta, bound_max_genome_size, rule_blacklist, grown=grown)[:3]
        search_type = "bound selection"
    elif parameter_selection == "cv":
        if len(split.folds) < 1:
            error_callback(Exception("Cross-validation cannot be performed on a split with no folds."))
        folds = [(np.asarray(f.train_genome_idx[...]), np.asarray(f.test_genome_idx[...])) for f in split.folds]

        def hp_search(hps):
            return learn_pruned_tree_cv(rule_classifications, rules, example_labels, train_idx, folds, n_classes, hps,
                                        rule_blacklist, grown=grown)
        search_type = "cross-validation"
    else:
        error_callback(ValueError("Unknown hyperparameter selection strategy specified."))
    best_hp_score, best_hps, best_master_tree = train_tree(
        hp_search, [str(c) for c in criterion], importances, [int(d) for d in max_depth],
        [int(m) for m in min_samples_split], progress_callback, search_type)

    train_predictions, test_predictions = _predictions(best_master_tree, rule_classifications, train_idx, test_idx,
                                                       progress_callback)
    train_answers, test_answers = example_labels[train_idx], example_labels[test_idx]
    binary = _as_text(dataset.classification_type) == "binary"
    metrics_of = _get_binary_metrics if binary else (lambda p_, a_: _get_multiclass_metrics(p_, a_, n_classes))
    train_metrics = metrics_of(train_predictions, train_answers)
    test_metrics = metrics_of(test_predictions, test_answers) if len(test_idx) > 0 else None

    genome_ids = np.array([_as_text(g) for g in dataset.genome_identifiers[...].tolist()], dtype=object)
    classifications = defaultdict(list)
    classifications["train_correct"] = genome_ids[train_idx[train_predictions == train_answers]].tolist() \
        if train_metrics["risk"][0] < 1.0 else []
    classifications["train_errors"] = genome_ids[train_idx[train_predictions != train_answers]].tolist() \
        if train_metrics["risk"][0] > 0 else []
    if len(test_idx) > 0:
        classifications["test_correct"] = genome_ids[test_idx[test_predictions == test_answers]].tolist() \
            if test_metrics["risk"][0] < 1.0 else []
        classifications["test_errors"] = genome_ids[test_idx[test_predictions != test_answers]].tolist() \
            if test_metrics["risk"][0] > 0 else []

    best_model = CARTModel(class_tags=phenotype_tags)
    best_model.decision_tree = best_master_tree
    model_equivalent_rules = dict((r, [rules[i] for i in r.equivalent_rules_idx]) for r in best_master_tree.rules)
    importance_sum = float(sum(r.importance for r in best_master_tree.rules))
    rule_importances = dict((r, r.importance / importance_sum) for r in best_master_tree.rules)
    if owns_matrix:
        rule_classifications.close()
    dataset.close()
    return (best_hps, best_hp_score, train_metrics, test_metrics, best_model, rule_importances, model_equivalent_rules,
            classifications)
