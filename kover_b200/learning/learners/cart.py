"""CART decision-tree learner driven by the device split scan.

Interface parity with reference kover/learning/learners/cart.py:31-470: ``DecisionTreeClassifier(criterion,
max_depth, min_samples_split, class_importance)``, ``fit(rules, rule_classifications, example_idx,
rule_blacklist, tiebreaker, level_callback, split_callback)``, ``decision_tree``, ``predict`` /
``predict_proba`` and ``_prune_tree``.  The breadth-first growth of the tree stays on the host; per node, what
the reference does with n_classes full-matrix ``sum_rows`` passes, ~30 numpy passes over all k-mers and a
Python-level ``min`` over 2K floats (cart.py:112-161, 219-250) is ONE call to
``rule_classifications.gini_best_split``: k_scan_tma<EPI_GINI2> (counts + Gini score + minimum fused in one pass
over the HBM-resident matrix) and k_gini2_ties_seg (the exact-tie set).
"""
import logging
from copy import deepcopy

import numpy as np

from ..common.rules import NodeExamples
from ..common.tree import ProbabilisticTreeNode

UTIL_BLOCK_SIZE = 1000000
_CRITERIA = ("gini", "cross-entropy")


class _NodeImpurity(object):
    """Impurity of a single node from scalar class counts, with the class priors altered by the class
    importances (Breiman et al. 1984, section 4.4; cart.py:67-110, 167-178).  The per-rule vector form of the Gini
    score lives in the CUDA kernels; this scalar form only labels the nodes (``criterion_value``)."""

    def __init__(self, example_idx, class_importance):
        self.n_total = dict((c, float(len(idx))) for c, idx in example_idx.items())
        n_all = sum(self.n_total.values())
        priors = dict((c, 1.0 * n / n_all) for c, n in self.n_total.items())
        norm = sum([class_importance[c] * priors[c] for c in priors.keys()])
        self.priors = dict((c, 1.0 * class_importance[c] * priors[c] / norm) for c in priors.keys())

    def _posterior(self, n_by_class):
        joint = dict((c, 1.0 * self.priors[c] * n_by_class[c] / self.n_total[c]) for c in n_by_class.keys())
        p_node = sum(joint.values())
        with np.errstate(divide="ignore", invalid="ignore"):
            return dict((c, np.divide(joint[c], p_node)) for c in joint.keys()), p_node

    def gini(self, n_by_class, multiply_by_node_proba=False):
        post, p_node = self._posterior(n_by_class)
        diversity = sum([post[a] * post[b] for a in post.keys() for b in post.keys() if a != b])
        return diversity * (p_node if multiply_by_node_proba else 1.0)

    def cross_entropy(self, n_by_class, multiply_by_node_proba=False):
        post, p_node = self._posterior(n_by_class)
        with np.errstate(divide="ignore", invalid="ignore"):
            diversity = -1.0 * sum([np.nan_to_num(post[c] * np.log(post[c])) for c in post.keys()])
        return diversity * (p_node if multiply_by_node_proba else 1.0)


class DecisionTreeClassifier(object):
    def __init__(self, criterion, max_depth, min_samples_split, class_importance):
        if criterion not in _CRITERIA:
            raise ValueError("The supporting splitting criteria are: %s." % str(list(_CRITERIA)))
        if max_depth < 1:
            raise ValueError("The maximum tree depth must be greater than 1.")
        if min_samples_split < 2.0:
            raise ValueError("The minimum number of examples used to split a node must be 2 or greater.")
        self.criterion = criterion
        self.max_depth = max_depth
        self.min_samples_split = int(min_samples_split)
        self.class_importance = class_importance      # class -> weight of its errors in the loss
        self.decision_tree = None

    # ------------------------------------------------------------------ split search
    def _cross_entropy_split(self, rule_classifications, impurity, node_idx, rule_blacklist):
        """cart.py:180-207 + 231-238.  np.log is not bit-reproducible on a GPU: the per-class counts come from ONE
        device pass (sum_rows_multi) and the reference's numpy formula is evaluated on the host.  Unreachable from
        the reference CLI (command_line.py:706 offers "crossentropy", cart.py:36 accepts "cross-entropy")."""
        classes = [c for c in node_idx.keys() if node_idx[c].size]
        counts = rule_classifications.sum_rows_multi([node_idx[c] for c in classes])
        with_kmer = dict((c, np.asarray(counts[i], dtype=float)) for i, c in enumerate(classes))
        without = dict((c, np.asarray(node_idx[c].shape[0] - with_kmer[c], dtype=float)) for c in classes)
        score = impurity.cross_entropy(with_kmer, multiply_by_node_proba=True)
        score += impurity.cross_entropy(without, multiply_by_node_proba=True)
        score[sum(with_kmer.values()) == 0] = np.inf
        score[sum(without.values()) == 0] = np.inf
        score[rule_blacklist] = np.inf
        best = score.min()
        return best, np.where(score == best)[0]

    def fit(self, rules, rule_classifications, example_idx, rule_blacklist=None, tiebreaker=None,
            level_callback=None, split_callback=None):
        levels = self._fit_levels(rules, rule_classifications, example_idx, rule_blacklist, tiebreaker,
                                  level_callback, split_callback)
        answer = None
        while True:
            try:
                nodes, impurity, blacklist, tie_table = levels.send(answer)
            except StopIteration:
                return
            answer = self._search_level(rule_classifications, nodes, impurity, blacklist, tie_table)

    def _search_level(self, rule_classifications, nodes, impurity, rule_blacklist, tie_table=None):
        """Best split of every node of ``nodes`` (class -> example indices dicts) -> [(score, candidate rules)];
        with a tie_table (see _fit_levels) the candidates are already reduced by the occurrence tiebreaker."""
        if self.criterion == "gini":
            many = getattr(rule_classifications, "gini_best_splits", None)
            if many is not None:
                if tie_table is not None:
                    return many(nodes, impurity.priors, impurity.n_total, rule_blacklist,
                                tie_tables=[tie_table] * len(nodes))
                return many(nodes, impurity.priors, impurity.n_total, rule_blacklist)
            return [rule_classifications.gini_best_split(idx, impurity.priors, impurity.n_total, rule_blacklist)
                    for idx in nodes]
        return [self._cross_entropy_split(rule_classifications, impurity, idx, rule_blacklist) for idx in nodes]

    def _fit_levels(self, rules, rule_classifications, example_idx, rule_blacklist=None, tiebreaker=None,
                    level_callback=None, split_callback=None):
        """Generator: yields (node example dicts, impurity, blacklist, tie table) once per tree level and is sent the
        list of (best score, candidate rules) of those nodes.  A tiebreaker that carries ``device_occurrences`` (a
        rules.DeviceOccurrences: the reference's occurrence tiebreaker, experiment_cart.py:82-94) is applied by the
        split search itself -- on the device -- and not called again on the host."""
        if not hasattr(rule_classifications, "gini_best_split"):
            raise TypeError("kover_b200's DecisionTreeClassifier scans a device-resident matrix: pass a "
                            "kover_b200.learning.common.rules.KmerRuleClassifications (there is no CPU path).")
        level_callback = level_callback or (lambda infos: None)
        split_callback = split_callback or (lambda node, equivalent_rules_idx: None)
        tiebreaker = tiebreaker or (lambda candidates: candidates)
        rule_blacklist = [] if rule_blacklist is None else rule_blacklist
        tie_table = getattr(tiebreaker, "device_occurrences", None)
        if self.criterion != "gini" or not hasattr(rule_classifications, "gini_best_splits"):
            tie_table = None

        example_idx = dict((c, np.asarray(idx)) for c, idx in example_idx.items())
        impurity = _NodeImpurity(example_idx, self.class_importance)
        node_criterion = impurity.gini if self.criterion == "gini" else impurity.cross_entropy

        def new_node(idx_by_class, depth, parent=None, counts=None):
            counts = counts if counts is not None else dict((c, len(i)) for c, i in idx_by_class.items())
            return ProbabilisticTreeNode(class_examples_idx=idx_by_class, depth=depth, parent=parent,
                                         criterion_value=node_criterion(counts), class_priors=impurity.priors,
                                         total_n_examples_by_class=impurity.n_total)

        # (NodeExamples: the dicts carry node / parent keys so that the level search can derive the larger of two
        #  siblings from their parent's resident counts instead of counting it)
        root = new_node(NodeExamples(example_idx), 0, counts=impurity.n_total)
        # Breadth first, one level at a time (the reference pops a FIFO queue, cart.py:260-339: same visiting
        # order).  The nodes of a level are independent, so their split searches are issued together and share
        # matrix passes; everything else -- callbacks, tiebreaker calls, child creation -- runs in node order.
        level_nodes, depth = [root], 0
        infos = {}
        min_split = max(self.min_samples_split, 2)
        while level_nodes:
            infos["depth"] = depth
            if depth > 0:
                level_callback(infos)                  # all the nodes of the previous level exist now
            if depth == self.max_depth:
                break                                  # the deepest level stays leaves
            # pure nodes and nodes too small to split stay leaves
            splittable = [node for node in level_nodes
                          if not (1.0 in node.class_proportions.values() or node.n_examples < min_split)]
            found = []
            if splittable:
                found = yield ([node.class_examples_idx for node in splittable], impurity, rule_blacklist, tie_table)
            next_level = []
            # with the tiebreaker already applied by the search, the chosen rules of the whole level are known here:
            # their columns come back in ONE get_columns call (one gather kernel / one exchange instead of one per node)
            fetched = {}
            if tie_table is not None:
                picks = [(j, cand[0]) for j, (score, cand) in enumerate(found) if score != np.inf]
                if len(picks) > 1:
                    columns = rule_classifications.get_columns([c for _, c in picks])
                    fetched = dict((j, columns[:, t]) for t, (j, _) in enumerate(picks))
            for j, (node, (best_score, candidates)) in enumerate(zip(splittable, found)):
                if best_score == np.inf:
                    continue                           # no rule separates the node into two non-empty children
                members = node.class_examples_idx
                equivalent = candidates if tie_table is not None else tiebreaker(candidates)
                if np.may_share_memory(equivalent, candidates):
                    equivalent = np.array(equivalent)  # candidates may be a view of a buffer the next level reuses
                chosen = equivalent[0]
                has_kmer = fetched[j] if j in fetched else rule_classifications.get_columns(chosen)
                left = NodeExamples(((c, members[c][has_kmer[members[c]] == 1]) for c in members.keys()), members)   # rule true
                right = NodeExamples(((c, members[c][has_kmer[members[c]] == 0]) for c in members.keys()), members)  # rule false

                node.rule = rules[chosen]
                node.rule_idx = chosen
                node.equivalent_rules_idx = equivalent
                node.left_child = new_node(left, node.depth + 1, parent=node)
                node.right_child = new_node(right, node.depth + 1, parent=node)
                # unnormalised importance: decrease of p(t) * impurity (cart.py:321-325)
                node.rule.importance = \
                    node.breiman_info.p_t * node.criterion_value - \
                    node.left_child.breiman_info.p_t * node.left_child.criterion_value - \
                    node.right_child.breiman_info.p_t * node.right_child.criterion_value
                split_callback(node, equivalent)
                next_level.extend((node.left_child, node.right_child))
                infos["model"] = root
            level_nodes, depth = next_level, depth + 1
        self.decision_tree = root

    # ------------------------------------------------------------------ predictions
    def _is_fitted(self):
        return self.decision_tree is not None

    def _fitted_tree(self):
        if not self._is_fitted():
            raise RuntimeError("The classifier must be fitted before predicting.")
        return self.decision_tree

    def predict(self, X):
        return self._fitted_tree().predict(X)

    def predict_proba(self, X):
        return self._fitted_tree().predict_proba(X)


def fit_many(learners, fit_args, rule_classifications):
    """Grow several Gini trees over the SAME rule_classifications in lock-step -- the fold trees and the master tree
    of a cross-validation (experiment_cart.py:345-378 grows them one after the other): at every depth the
    splittable nodes of ALL the trees are searched together, their class masks counted 8 nodes per matrix pass.
    fit_args: one dict of ``fit`` keyword arguments per learner.  Each learner ends in exactly the state its own
    ``fit`` would have produced; tie lists handed to a tiebreaker are consumed before the next search starts."""
    running = {}
    for i, (learner, kw) in enumerate(zip(learners, fit_args)):
        gen = learner._fit_levels(rule_classifications=rule_classifications, **kw)
        try:
            running[i] = (gen, next(gen))
        except StopIteration:
            pass
    many = getattr(rule_classifications, "gini_best_splits", None)
    while running:
        batchable = [i for i in running if learners[i].criterion == "gini" and many is not None]
        answers = {}
        # one call per distinct blacklist (the experiments use a single blacklist for all the trees)
        by_blacklist = {}
        for i in batchable:
            by_blacklist.setdefault(np.asarray(running[i][1][2], dtype=np.int64).tobytes(), []).append(i)
        for members in by_blacklist.values():
            nodes, priors, totals, tables = [], [], [], []
            for i in members:
                level_nodes, impurity, _, tie_table = running[i][1]
                nodes += level_nodes
                priors += [impurity.priors] * len(level_nodes)
                totals += [impurity.n_total] * len(level_nodes)
                tables += [tie_table] * len(level_nodes)
            found = many(nodes, priors, totals, running[members[0]][1][2],
                         tie_tables=tables if any(t is not None for t in tables) else None)
            # the tie lists may be views of one result buffer: every learner of the batch consumes its share before
            # the next search is issued (below), so they are handed over as they are
            at = 0
            for i in members:
                n = len(running[i][1][0])
                answers[i] = found[at:at + n]
                at += n
        for i in running:
            if i not in answers:
                level_nodes, impurity, blacklist, tie_table = running[i][1]
                answers[i] = learners[i]._search_level(rule_classifications, level_nodes, impurity, blacklist, tie_table)
                # (a per-learner search invalidates earlier views: materialise what the batch produced)
                answers = dict((j, [(sc, np.array(c)) for sc, c in a]) if j != i else (j, a) for j, a in answers.items())
        still_running = {}
        for i, answer in answers.items():
            gen = running[i][0]
            try:
                still_running[i] = (gen, gen.send(answer))
            except StopIteration:
                pass
        running = still_running


def _to_leaf(node):
    node.rule = node.left_child = node.right_child = None


def _prune_tree(tree):
    """Minimal cost-complexity pruning (Breiman et al. 1984, chapter 3; cart.py:362-470).
    -> (alphas, trees): the nested sequence T1 > T2 > ... > {root} and the alpha at which each tree appears."""
    def collapsible(node):
        # internal nodes whose two children are leaves
        if node.is_leaf:
            return []
        if node.left_child.is_leaf and node.right_child.is_leaf:
            return [node]
        return collapsible(node.left_child) + collapsible(node.right_child)

    def drop_useless_splits(root):
        # Tmax -> T1: a split that does not lower the resubstitution risk R is undone (cart.py:384-404)
        todo = collapsible(root)
        while todo:
            node = todo.pop()
            children_risk = node.left_child.breiman_info.R_t + node.right_child.breiman_info.R_t
            if np.allclose(node.breiman_info.R_t, children_risk):
                _to_leaf(node)
                up = node.parent
                if up is not None and up.left_child.is_leaf and up.right_child.is_leaf:
                    todo.append(up)

    def weakest_links(node):
        # g(t) = (R(t) - R(T_t)) / (|leaves(T_t)| - 1); ties are resolved with allclose like cart.py:406-432
        if node.is_leaf:
            return np.inf, [node]
        subtree_leaves = node.leaves
        g_here = float(node.breiman_info.R_t - sum(l.breiman_info.R_t for l in subtree_leaves)) / \
            (len(subtree_leaves) - 1)
        g_left, links_left = weakest_links(node.left_child)
        g_right, links_right = weakest_links(node.right_child)
        g_below = min(g_left, g_right)
        children_tie = np.allclose(g_left, g_right)
        if np.allclose(g_here, g_below):
            below = links_left + links_right if children_tie else (links_left if g_left < g_right else links_right)
            return g_here, [node] + below
        if g_here < g_below:
            return g_here, [node]
        if children_tie:
            return g_left, links_left + links_right
        if g_left > g_right:
            return g_right, links_right
        if g_left < g_right:
            return g_left, links_left
        raise Exception("Unhandled case detected!")

    current = deepcopy(tree)
    drop_useless_splits(current)
    alphas, trees = [0], [current]
    while not current.is_leaf:                      # prune the weakest links until only the root is left
        current = deepcopy(current)
        alpha, links = weakest_links(current)
        for node in links:
            _to_leaf(node)
        alphas.append(alpha)
        trees.append(current)
    return tuple(alphas), tuple(trees)
