"""Tree nodes for the CART learner (host objects).

Mirrors reference kover/learning/common/tree.py:28-264: ``BreimanInfo`` (Breiman et al. 1984,
eqs. 2.2-2.4, def. 2.10), ``TreeNode`` and ``ProbabilisticTreeNode``.  Out of the hot path; the
arithmetic is restated with the reference's operation order because node probabilities feed the
rule importances (cart.py:321-325).
"""
import numpy as np

from .models import BaseModel


class BreimanInfo(object):
    def __init__(self, node_n_examples_by_class, class_priors, total_n_examples_by_class):
        self.p_j_t = {c: class_priors[c] * node_n_examples_by_class[c] / total_n_examples_by_class[c]
                      for c in class_priors.keys()}                                    # eq. 2.2
        self.p_t = sum(self.p_j_t.values())                                            # eq. 2.3
        self.p_j_given_t = {c: self.p_j_t[c] / self.p_t for c in class_priors.keys()}  # eq. 2.4
        self.r_t = 1.0 - max(self.p_j_given_t.values())                                # def. 2.10
        self.R_t = self.r_t * self.p_t


class TreeNode(BaseModel):
    def __init__(self, depth, class_examples_idx, total_n_examples_by_class, class_priors, rule=None, parent=None,
                 left_child=None, right_child=None, criterion_value=-1):
        assert isinstance(class_examples_idx, dict)
        assert isinstance(total_n_examples_by_class, dict)
        assert isinstance(class_priors, dict)
        self.rule = rule
        self.parent = parent
        self.left_child = left_child
        self.right_child = right_child
        self.class_examples_idx = class_examples_idx
        self.depth = depth
        self.criterion_value = criterion_value
        self.breiman_info = BreimanInfo(
            node_n_examples_by_class={c: len(idx) for c, idx in class_examples_idx.items()},
            class_priors=class_priors, total_n_examples_by_class=total_n_examples_by_class)

    @property
    def is_leaf(self):
        return self.rule is None and self.left_child is None and self.right_child is None

    @property
    def is_root(self):
        return self.parent is None

    @property
    def n_examples(self):
        return sum(len(idx) for idx in self.class_examples_idx.values())

    @property
    def class_proportions(self):
        n_examples = self.n_examples
        return {c: float(len(idx)) / n_examples for c, idx in self.class_examples_idx.items()}

    @property
    def class_prediction(self):
        p = self.breiman_info.p_j_given_t
        return list(p.keys())[int(np.argmax(list(p.values())))]

    def _preorder(self):
        nodes = [self]
        if not self.is_leaf:
            nodes += self.left_child._preorder()
            nodes += self.right_child._preorder()
        return nodes

    @property
    def rules(self):
        return [n.rule for n in self._preorder() if not n.is_leaf]

    @property
    def leaves(self):
        return [n for n in self._preorder() if n.is_leaf]

    @property
    def tree_depth(self):
        return max(n.depth for n in self.leaves)

    def __iter__(self):
        for node_id, node in enumerate(self._preorder()):
            yield node_id, node

    def __len__(self):
        return len(self._preorder())

    def __str__(self, depth=0):
        if self.is_leaf:
            return "\n" + ("    " * depth) + str(self.class_prediction)
        s = self.right_child.__str__(depth=depth + 1)
        s += "\n" + ("    " * depth + "   ") + "/"
        s += "\n" + ("    " * depth) + str(self.rule)
        s += "\n" + ("    " * depth + "   ") + "\\"
        s += self.left_child.__str__(depth=depth + 1)
        return s


class ProbabilisticTreeNode(TreeNode):
    def predict(self, X):
        return np.argmax(self.predict_proba(np.ascontiguousarray(X)), axis=0)

    def predict_proba(self, X):
        X = np.ascontiguousarray(X)
        class_probabilities = np.zeros((len(self.class_examples_idx.keys()), X.shape[0]))
        for i, x in enumerate(X):
            x = x.reshape(1, -1)
            node = self
            while not node.is_leaf:
                node = node.left_child if node.rule.classify(x) else node.right_child   # rule TRUE -> left
            for c in self.class_examples_idx.keys():
                class_probabilities[c][i] = node.breiman_info.p_j_given_t[c]
        return class_probabilities
