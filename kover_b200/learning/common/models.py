"""Model containers filled by the learners of this package (small host objects).

Interface parity with reference kover/learning/common/models.py:20-182 -- class names, ``add`` / ``remove`` /
``predict`` / ``predict_proba`` / ``learner`` / ``type`` / ``len`` / ``str`` -- so that code written against the
reference (report writers, the experiment drivers) keeps working.  The arithmetic is restated, not copied: a
conjunction is the AND of its rules' 0/1 outputs, a disjunction the OR, both returned as float32 "probabilities"
like the reference and thresholded at 0.5 by ``predict``.
"""
import numpy as np

conjunction, disjunction = "conjunction", "disjunction"
scm, cart = "scm", "cart"


class BaseModel(object):
    learner = None

    def predict(self, X):
        raise NotImplementedError()

    def predict_proba(self, X):
        raise NotImplementedError()

    def _to_string(self):
        raise NotImplementedError()

    def __str__(self):
        return self._to_string()


class SCMModel(BaseModel):
    """An ordered set of rules combined by a logical operator (given by the subclass)."""
    learner = scm
    type = None
    _joiner = " "

    def __init__(self):
        self.rules = []

    # -- container protocol
    def add(self, rule):
        self.rules.append(rule)

    def remove(self, index):
        self.rules.pop(index)

    def __len__(self):
        return len(self.rules)

    def __iter__(self):
        return iter(list(self.rules))

    def __eq__(self, other):
        return self.__dict__ == other.__dict__

    @property
    def example_dependencies(self):
        deps = []
        for rule in self.rules:
            deps.extend(rule.example_dependencies)
        return deps

    # -- predictions
    def _votes(self, X):
        """float32 [n_rules, n_examples] of the rules' 0/1 outputs (empty models give a [0, n] array)."""
        n = X.shape[0]
        if not self.rules:
            return np.zeros((0, n), dtype=np.float32)
        return np.stack([np.asarray(rule.classify(X), dtype=np.float32).reshape(n) for rule in self.rules])

    def predict(self, X):
        return (self.predict_proba(X) > 0.5).astype(np.uint8)

    def _to_string(self, separator=None):
        return (self._joiner if separator is None else separator).join(str(rule) for rule in self.rules)


class ConjunctionModel(SCMModel):
    type = conjunction
    _joiner = " and "

    def predict_proba(self, X):
        # 1 where every rule outputs 1 (an empty conjunction is true)
        return np.prod(self._votes(X), axis=0, dtype=np.float32)


class DisjunctionModel(SCMModel):
    type = disjunction
    _joiner = " or "

    def predict_proba(self, X):
        # 1 where at least one rule outputs 1 (an empty disjunction is false)
        none_fired = np.prod(np.float32(1.0) - self._votes(X), axis=0, dtype=np.float32)
        return np.float32(1.0) - none_fired


class CARTModel(BaseModel):
    learner = cart

    def __init__(self, class_tags=None):
        self.decision_tree = None
        self.class_tags = class_tags

    def _require_tree(self):
        if self.decision_tree is None:
            raise RuntimeError("A decision tree must be fitted prior to calling predict.")
        return self.decision_tree

    def predict(self, X):
        return np.asarray(self._require_tree().predict(X), dtype=np.uint8)

    def predict_proba(self, X):
        return self._require_tree().predict_proba(X)

    def _to_string(self):
        return str(self.decision_tree)

    def __len__(self):
        return len(self.decision_tree) if self.decision_tree is not None else 0

    @property
    def depth(self):
        return self.decision_tree.tree_depth if self.decision_tree is not None else 0
