"""Decision-tree nodes for the CART learner of this package (small host objects).

Interface parity with reference kover/learning/common/tree.py:28-264 (``BreimanInfo``, ``TreeNode``,
``ProbabilisticTreeNode`` and their attributes), because the learner, the pruning and the bound code read them.
The node statistics follow Breiman et al. (1984): eq. 2.2-2.4 and def. 2.10; their floating-point operation order
is kept (prior * n / N, sum in class order) since rule importances and pruning thresholds are compared exactly.
"""
from copy import deepcopy

import numpy as np

from .models import BaseModel


class BreimanInfo(object):
    """p(j, t), p(t), p(j | t), r(t) and R(t) of a node t."""

    def __init__(self, node_n_examples_by_class, class_priors, total_n_examples_by_class):
        classes = list(class_priors.keys())
        joint = {}
        for c in classes:
            joint[c] = class_priors[c] * node_n_examples_by_class[c] / total_n_examples_by_class[c]
        self.p_j_t = joint
        self.p_t = sum(joint.values())
        self.p_j_given_t = dict((c, joint[c] / self.p_t) for c in classes)
        self.r_t = 1.0 - max(self.p_j_given_t.values())
        self.R_t = self.r_t * self.p_t


class TreeNode(BaseModel):
    def __init__(self, depth, class_examples_idx, total_n_examples_by_class, class_priors, rule=None, parent=None,
                 left_child=None, right_child=None, criterion_value=-1):
        for arg in (class_examples_idx, total_n_examples_by_class, class_priors):
            assert isinstance(arg, dict)
        self.depth = depth
        self.parent = parent
        self.rule = rule
        self.left_child, self.right_child = left_child, right_child
        self.class_examples_idx = class_examples_idx
        self.criterion_value = criterion_value
        counts = dict((c, len(idx)) for c, idx in class_examples_idx.items())
        self.breiman_info = BreimanInfo(counts, class_priors, total_n_examples_by_class)

    def __deepcopy__(self, memo):
        """Independent copy of the (sub)tree: nodes and rule objects are copied (pruning turns nodes into leaves,
        predictions re-address rule indices), the per-class example index arrays and the node statistics are shared --
        nothing modifies them in place.  The pruning sequence copies the tree once per pruned tree (cart.py:434-470),
        which made the array copies the dominant cost of the experiment drivers."""
        clone = object.__new__(type(self))
        memo[id(self)] = clone
        for name, value in self.__dict__.items():
            if name == "class_examples_idx":
                clone.__dict__[name] = dict(value)
            elif name == "breiman_info":
                clone.__dict__[name] = value
            else:
                clone.__dict__[name] = deepcopy(value, memo)
        return clone

    # -- structure
    @property
    def is_leaf(self):
        return self.rule is None and self.left_child is None and self.right_child is None

    @property
    def is_root(self):
        return self.parent is None

    def _walk(self):
        """Pre-order traversal (node, left subtree, right subtree), iterative."""
        stack, order = [self], []
        while stack:
            node = stack.pop()
            order.append(node)
            if not node.is_leaf:
                stack.append(node.right_child)
                stack.append(node.left_child)
        return order

    @property
    def rules(self):
        return [node.rule for node in self._walk() if not node.is_leaf]

    @property
    def leaves(self):
        return [node for node in self._walk() if node.is_leaf]

    @property
    def tree_depth(self):
        return max(leaf.depth for leaf in self.leaves)

    def __iter__(self):
        return iter(enumerate(self._walk()))

    def __len__(self):
        return len(self._walk())

    # -- content
    @property
    def n_examples(self):
        return sum(len(idx) for idx in self.class_examples_idx.values())

    @property
    def class_proportions(self):
        total = self.n_examples
        return dict((c, float(len(idx)) / total) for c, idx in self.class_examples_idx.items())

    @property
    def class_prediction(self):
        posterior = self.breiman_info.p_j_given_t
        classes = list(posterior.keys())
        return classes[int(np.argmax([posterior[c] for c in classes]))]

    def __str__(self, depth=0):
        pad = "    " * depth
        if self.is_leaf:
            return "\n" + pad + str(self.class_prediction)
        return "".join([self.right_child.__str__(depth=depth + 1),
                        "\n" + pad + "   /",
                        "\n" + pad + str(self.rule),
                        "\n" + pad + "   \\",
                        self.left_child.__str__(depth=depth + 1)])


class ProbabilisticTreeNode(TreeNode):
    def _leaf_of(self, x):
        """The leaf an example (one row of the 0/1 k-mer matrix) falls into: rule true -> left child."""
        node = self
        while not node.is_leaf:
            node = node.left_child if node.rule.classify(x) else node.right_child
        return node

    def predict_proba(self, X):
        """[n_classes, n_examples]: class posterior of the leaf every example reaches.  The examples are routed down
        the tree node by node (one vectorised rule evaluation per internal node) instead of one by one."""
        X = np.ascontiguousarray(X)
        classes = list(self.class_examples_idx.keys())
        out = np.zeros((len(classes), X.shape[0]))
        pending = [(self, np.arange(X.shape[0]))]
        while pending:
            node, rows = pending.pop()
            if rows.size == 0:
                continue
            if node.is_leaf:
                posterior = node.breiman_info.p_j_given_t
                for c in classes:
                    out[c][rows] = posterior[c]
                continue
            goes_left = node.rule.classify(X[rows]) != 0          # rule true -> left child
            pending.append((node.left_child, rows[goes_left]))
            pending.append((node.right_child, rows[~goes_left]))
        return out

    def predict(self, X):
        return np.argmax(self.predict_proba(X), axis=0)
