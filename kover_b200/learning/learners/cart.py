"""CART decision-tree learner driven by the device split scan.

Mirrors reference kover/learning/learners/cart.py:31-346: ``DecisionTreeClassifier`` with the same
constructor / ``fit`` signature, stopping rules, callbacks and exceptions.  The BFS over nodes stays
on the host (cart.py:260-339); per node, the n_classes ``sum_rows`` passes and the ~30 numpy passes
of ``_gini_rule_score`` + the Python-level ``min`` over all k-mers (cart.py:112-161, 238) become
one call to ``rule_classifications.gini_best_split`` (k_scan2 / k_count_multi + k_gini).

``_prune_tree`` (minimal cost-complexity pruning, cart.py:362-470) is host code on a handful of nodes; it is
restated at the end of this module because the bound-based model selection needs it.
"""
import logging
from collections import deque
from copy import deepcopy

import numpy as np

from ..common.tree import ProbabilisticTreeNode

UTIL_BLOCK_SIZE = 1000000


class DecisionTreeClassifier(object):
    def __init__(self, criterion, max_depth, min_samples_split, class_importance):
        supported_criteria = ["gini", "cross-entropy"]
        if criterion not in supported_criteria:
            raise ValueError("The supporting splitting criteria are: %s." % str(supported_criteria))
        self.criterion = criterion
        if max_depth < 1:
            raise ValueError("The maximum tree depth must be greater than 1.")
        self.max_depth = max_depth
        if min_samples_split < 2.0:
            raise ValueError("The minimum number of examples used to split a node must be 2 or greater.")
        self.min_samples_split = int(min_samples_split)
        self.class_importance = class_importance
        self.decision_tree = None

    def fit(self, rules, rule_classifications, example_idx, rule_blacklist=None, tiebreaker=None,
            level_callback=None, split_callback=None):
        if level_callback is None:
            level_callback = lambda x: None
        if split_callback is None:
            split_callback = lambda x, y: None
        if tiebreaker is None:
            tiebreaker = lambda x: x
        if rule_blacklist is None:
            rule_blacklist = []
        if not hasattr(rule_classifications, "gini_best_split"):
            raise TypeError("kover_b200's DecisionTreeClassifier scans a device-resident matrix: pass a "
                            "kover_b200.learning.common.rules.KmerRuleClassifications (there is no CPU path).")

        # Class priors altered by the class importances (cart.py:67-77; Breiman et al. 1984, section 4.4)
        n_total_class_examples = {c: float(len(idx)) for c, idx in example_idx.items()}
        priors = {c: 1.0 * n_examples / sum(n_total_class_examples.values())
                  for c, n_examples in n_total_class_examples.items()}
        denum = sum([self.class_importance[c] * priors[c] for c in priors.keys()])
        altered_priors = {c: 1.0 * self.class_importance[c] * priors[c] / denum for c in priors.keys()}

        def _gini_impurity(n_examples_by_class, multiply_by_node_proba=False):
            """Node impurity for scalar counts (cart.py:85-110); the per-rule vector form runs in k_gini."""
            p_j_t = {c: 1.0 * altered_priors[c] * n_examples_by_class[c] / n_total_class_examples[c]
                     for c in n_examples_by_class.keys()}
            p_t = sum(p_j_t.values())
            with np.errstate(divide='ignore', invalid='ignore'):
                p_j_given_t = {c: np.divide(p_j_t[c], p_t) for c in p_j_t.keys()}
            gini_diversity_index = sum([p_j_given_t[i] * p_j_given_t[j]
                                        for i in p_j_given_t.keys() for j in p_j_given_t.keys() if i != j])
            return gini_diversity_index * (p_t if multiply_by_node_proba else 1.0)

        def _cross_entropy(n_class_examples, multiply_by_node_proba=False):
            """cart.py:167-178"""
            p_class_node = {c: 1.0 * altered_priors[c] * n_class_examples[c] / n_total_class_examples[c]
                            for c in n_class_examples.keys()}
            node_resubstitution_estimate = sum(p_class_node.values())
            with np.errstate(divide='ignore', invalid='ignore'):
                p_class_given_node = {c: np.divide(p_class_node[c], node_resubstitution_estimate)
                                      for c in p_class_node.keys()}
                diversity_index = -1.0 * sum([np.nan_to_num(p_class_given_node[c] * np.log(p_class_given_node[c]))
                                              for c in p_class_given_node.keys()])
            return diversity_index * (node_resubstitution_estimate if multiply_by_node_proba else 1.0)

        def _cross_entropy_best_split(node_example_idx):
            """cart.py:180-207 + 231-238.  np.log is not bit-reproducible on a GPU, so the per-class counts
            come from ONE device pass (sum_rows_multi) and the score is evaluated by numpy on the host
            (SURVEY section 7).  Unreachable from the reference CLI (command_line.py:706 vs cart.py:36)."""
            classes = [c for c in node_example_idx.keys() if node_example_idx[c].size]
            counts = rule_classifications.sum_rows_multi([node_example_idx[c] for c in classes])
            left_count = {c: np.asarray(counts[i], dtype=float) for i, c in enumerate(classes)}
            right_count = {c: np.asarray(node_example_idx[c].shape[0] - left_count[c], dtype=float) for c in classes}
            score = _cross_entropy(left_count, multiply_by_node_proba=True)
            score += _cross_entropy(right_count, multiply_by_node_proba=True)
            score[sum(left_count.values()) == 0] = np.inf
            score[sum(right_count.values()) == 0] = np.inf
            score[rule_blacklist] = np.inf
            m = score.min()
            return m, np.where(score == m)[0]

        if self.criterion == "gini":
            get_criterion = _gini_impurity
            best_split = lambda ex: rule_classifications.gini_best_split(ex, altered_priors, n_total_class_examples,
                                                                         rule_blacklist)
        else:
            get_criterion = _cross_entropy
            best_split = _cross_entropy_best_split
        node_type = ProbabilisticTreeNode

        def _find_best_split(node):
            """cart.py:219-250"""
            node_example_idx = node.class_examples_idx
            min_score, candidate_rules_idx = best_split(node_example_idx)
            if min_score == np.inf:
                return None, None, None, None
            logging.debug("There are %d equivalent rules before the tiebreaker." % len(candidate_rules_idx))
            best_rules_idx = tiebreaker(candidate_rules_idx)
            logging.debug("The are %d equivalent rules after the tiebreaker." % len(best_rules_idx))
            selected_rule_idx = best_rules_idx[0]
            rule_preds = rule_classifications.get_columns(selected_rule_idx)
            left = {c: node_example_idx[c][rule_preds[node_example_idx[c]] == 1] for c in node_example_idx.keys()}
            right = {c: node_example_idx[c][rule_preds[node_example_idx[c]] == 0] for c in node_example_idx.keys()}
            return selected_rule_idx, best_rules_idx, left, right

        example_idx = {c: np.asarray(idx) for c, idx in example_idx.items()}
        root = node_type(class_examples_idx=example_idx, depth=0,
                         criterion_value=get_criterion(n_total_class_examples),
                         class_priors=altered_priors, total_n_examples_by_class=n_total_class_examples)

        nodes_to_split = deque([root])
        runtime_infos = {}
        current_depth = -1
        min_samples_split = max(self.min_samples_split, 2)

        while len(nodes_to_split) > 0:
            node = nodes_to_split.popleft()

            # Check if we have reached a new depth
            if node.depth != current_depth:
                current_depth = node.depth
                runtime_infos["depth"] = current_depth
                if current_depth > 0:
                    level_callback(runtime_infos)
                if current_depth == self.max_depth:
                    break  # the nodes of the last level must remain leaves

            if 1.0 in node.class_proportions.values():
                continue                               # pure leaf
            if node.n_examples < min_samples_split:
                continue

            selected_rule_idx, equivalent_rule_idx, left_idx, right_idx = _find_best_split(node)
            if selected_rule_idx is None:
                continue                               # no rule splits the node into two non-empty leaves

            node.rule = rules[selected_rule_idx]
            node.rule_idx = selected_rule_idx
            node.equivalent_rules_idx = equivalent_rule_idx
            node.left_child = node_type(parent=node, class_examples_idx=left_idx, depth=node.depth + 1,
                                        criterion_value=get_criterion({c: len(i) for c, i in left_idx.items()}),
                                        class_priors=altered_priors,
                                        total_n_examples_by_class=n_total_class_examples)
            node.right_child = node_type(parent=node, class_examples_idx=right_idx, depth=node.depth + 1,
                                         criterion_value=get_criterion({c: len(i) for c, i in right_idx.items()}),
                                         class_priors=altered_priors,
                                         total_n_examples_by_class=n_total_class_examples)

            # Unnormalised rule importance (cart.py:321-325)
            node.rule.importance = \
                node.breiman_info.p_t * node.criterion_value - \
                node.left_child.breiman_info.p_t * node.left_child.criterion_value - \
                node.right_child.breiman_info.p_t * node.right_child.criterion_value

            split_callback(node, equivalent_rule_idx)
            nodes_to_split.append(node.left_child)
            nodes_to_split.append(node.right_child)
            runtime_infos["model"] = root

        self.decision_tree = root

    def predict(self, X):
        if not self._is_fitted():
            raise RuntimeError("The classifier must be fitted before predicting.")
        return self.decision_tree.predict(X)

    def predict_proba(self, X):
        if not self._is_fitted():
            raise RuntimeError("The classifier must be fitted before predicting.")
        return self.decision_tree.predict_proba(X)

    def _is_fitted(self):
        return self.decision_tree is not None


def _make_leaf(node):
    node.rule = None
    node.left_child = None
    node.right_child = None


def _prune_tree(tree):
    """Minimal cost-complexity pruning (Breiman et al. 1984, ch. 3), cart.py:362-470.
    -> (alphas, trees): the nested sequence T1 > T2 > ... > {root} with the alpha at which each appears."""
    def leaf_parents(node):
        if node.is_leaf:
            return []
        if node.left_child.is_leaf and node.right_child.is_leaf:
            return [node]
        return leaf_parents(node.left_child) + leaf_parents(node.right_child)

    def initial_pruning(root):                    # Tmax -> T1: collapse splits that do not reduce R (cart.py:384-404)
        parents = leaf_parents(root)
        while len(parents) > 0:
            node = parents.pop()
            if np.allclose(node.breiman_info.R_t, node.left_child.breiman_info.R_t + node.right_child.breiman_info.R_t):
                _make_leaf(node)
                if not node.is_root and node.parent.left_child.is_leaf and node.parent.right_child.is_leaf:
                    parents.append(node.parent)

    def find_weakest_links(node):                 # cart.py:406-432
        if node.is_leaf:
            return np.inf, [node]
        RTt = sum(l.breiman_info.R_t for l in node.leaves)
        current_gt = float(node.breiman_info.R_t - RTt) / (len(node.leaves) - 1)
        left_min_gt, left_links = find_weakest_links(node.left_child)
        right_min_gt, right_links = find_weakest_links(node.right_child)
        if np.allclose(current_gt, min(left_min_gt, right_min_gt)):
            if np.allclose(left_min_gt, right_min_gt):
                return current_gt, [node] + left_links + right_links
            return current_gt, [node] + (left_links if left_min_gt < right_min_gt else right_links)
        elif current_gt < min(left_min_gt, right_min_gt):
            return current_gt, [node]
        elif np.allclose(left_min_gt, right_min_gt):
            return left_min_gt, left_links + right_links
        elif left_min_gt > right_min_gt:
            return right_min_gt, right_links
        elif left_min_gt < right_min_gt:
            return left_min_gt, left_links
        raise Exception("Unhandled case detected!")

    T1 = deepcopy(deepcopy(tree))
    initial_pruning(T1)
    sequence = [(0, T1)]
    current = T1
    while not current.is_leaf:                    # cart.py:441-463
        current = deepcopy(current)
        min_gt, weakest_links = find_weakest_links(current)
        for n in weakest_links:
            _make_leaf(n)
        sequence.append((min_gt, current))
    alphas, trees = zip(*sequence)
    return alphas, trees
