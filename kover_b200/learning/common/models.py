"""Model containers the learners fill (host objects, a few lines each).

Mirrors reference kover/learning/common/models.py:20-182: same names and prediction arithmetic
(float32 products for the SCM models).  Out of the hot path; restated so that ``fit`` has something
to return and predictions can be compared with the reference's.
"""
import numpy as np

conjunction = "conjunction"
disjunction = "disjunction"
scm = "scm"
cart = "cart"


class BaseModel(object):
    def predict(self, X):
        raise NotImplementedError()

    def predict_proba(self, X):
        raise NotImplementedError()

    @property
    def learner(self):
        raise NotImplementedError()

    def __str__(self):
        return self._to_string()


class SCMModel(BaseModel):
    def __init__(self):
        super(SCMModel, self).__init__()
        self.rules = []

    def add(self, rule):
        self.rules.append(rule)

    def remove(self, index):
        del self.rules[index]

    def predict(self, X):
        predictions = self.predict_proba(X)
        predictions[predictions > 0.5] = 1
        predictions[predictions <= 0.5] = 0
        return np.asarray(predictions, dtype=np.uint8)

    @property
    def learner(self):
        return scm

    @property
    def type(self):
        raise NotImplementedError()

    def _to_string(self, separator=" "):
        return separator.join([str(a) for a in self.rules])

    def __eq__(self, other):
        return self.__dict__ == other.__dict__

    def __iter__(self):
        for ba in self.rules:
            yield ba

    def __len__(self):
        return len(self.rules)


class ConjunctionModel(SCMModel):
    def predict_proba(self, X):
        predictions = np.ones(X.shape[0], np.float32)
        for a in self.rules:
            predictions *= a.classify(X)
        return predictions

    @property
    def type(self):
        return conjunction

    def __str__(self):
        return self._to_string(separator=" and ")


class DisjunctionModel(SCMModel):
    def predict_proba(self, X):
        predictions = np.ones(X.shape[0], dtype=np.float32)
        for a in self.rules:
            predictions *= 1.0 - a.classify(X)
        return 1.0 - predictions

    @property
    def type(self):
        return disjunction

    def __str__(self):
        return self._to_string(separator=" or ")


class CARTModel(BaseModel):
    def __init__(self, class_tags=None):
        super(CARTModel, self).__init__()
        self.decision_tree = None
        self.class_tags = class_tags

    def predict(self, X):
        if self.decision_tree is None:
            raise RuntimeError("A decision tree must be fitted prior to calling predict.")
        return np.asarray(self.decision_tree.predict(X), dtype=np.uint8)

    def predict_proba(self, X):
        if self.decision_tree is None:
            raise RuntimeError("A decision tree must be fitted prior to calling predict.")
        return self.decision_tree.predict_proba(X)

    @property
    def learner(self):
        return cart

    def _to_string(self):
        return str(self.decision_tree)

    def __len__(self):
        return 0 if self.decision_tree is None else len(self.decision_tree)

    @property
    def depth(self):
        return 0 if self.decision_tree is None else self.decision_tree.tree_depth
