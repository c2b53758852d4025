"""Binary classification metrics (host, trivial).  Same keys and values as reference
learning/experiments/metrics.py:24-60 (``risk``, ``tp``, ``fp``, ``tn``, ``fn``, ``precision``, ``sensitivity``,
``recall``, ``specificity``, ``f1_score``; one list entry per row of ``predictions``)."""
from collections import defaultdict

import numpy as np


def _ratio(num, den):
    return 1.0 * num / den if den != 0 else -np.inf


def _get_binary_metrics(predictions, answers):
    predictions = np.atleast_2d(predictions)
    positives, negatives = answers == 1, answers == 0
    metrics = defaultdict(list)
    for row in predictions:
        tp = int(np.count_nonzero(row[positives] == 1))
        fn = int(np.count_nonzero(row[positives] == 0))
        fp = int(np.count_nonzero(row[negatives] == 1))
        tn = int(np.count_nonzero(row[negatives] == 0))
        precision, recall = _ratio(tp, tp + fp), _ratio(tp, tp + fn)
        values = {
            "risk": 1.0 * int(np.count_nonzero(row != answers)) / len(answers),
            "tp": tp, "fp": fp, "tn": tn, "fn": fn,
            "precision": precision, "sensitivity": recall, "recall": recall,
            "specificity": _ratio(tn, fp + tn),
            "f1_score": 2.0 * precision * recall / (precision + recall) if (precision + recall) > 0.0 else -np.inf,
        }
        for key, value in values.items():
            metrics[key].append(value)
    return metrics


def _get_multiclass_metrics(predictions, answers, nb_class):
    """``risk`` and the [actual][predicted] ``confusion_matrix`` per row of ``predictions`` (metrics.py:65-92)."""
    predictions = np.atleast_2d(predictions)
    metrics = defaultdict(list)
    for row in predictions:
        metrics["risk"].append(1.0 * int(np.count_nonzero(row != answers)) / len(answers))
        metrics["confusion_matrix"].append([[int(np.count_nonzero(row[answers == actual] == predicted))
                                             for predicted in range(nb_class)] for actual in range(nb_class)])
    return metrics
