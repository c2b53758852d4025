"""Binary classification metrics (host, trivial).  Mirrors reference learning/experiments/metrics.py:24-60."""
from collections import defaultdict

import numpy as np


def _get_binary_metrics(predictions, answers):
    if len(predictions.shape) == 1:
        predictions = predictions.reshape(1, -1)
    metrics = defaultdict(list)
    for i in range(predictions.shape[0]):
        p = predictions[i]
        risk = 1.0 * len(p[p != answers]) / len(answers)
        tp = len(np.where(p[answers == 1] == 1)[0])
        fp = len(np.where(p[answers == 0] == 1)[0])
        tn = len(np.where(p[answers == 0] == 0)[0])
        fn = len(np.where(p[answers == 1] == 0)[0])
        precision = 1.0 * tp / (tp + fp) if (tp + fp) != 0 else -np.inf
        sensitivity = recall = 1.0 * tp / (tp + fn) if (tp + fn) != 0 else -np.inf
        specificity = 1.0 * tn / (fp + tn) if (fp + tn) != 0 else -np.inf
        f1_score = 2.0 * precision * recall / (precision + recall) if (precision + recall) > 0.0 else -np.inf
        for key, value in (("risk", risk), ("tp", tp), ("fp", fp), ("tn", tn), ("fn", fn), ("precision", precision),
                           ("sensitivity", sensitivity), ("recall", recall), ("specificity", specificity),
                           ("f1_score", f1_score)):
            metrics[key].append(value)
    return metrics
