"""CPU restatement of the reference's evaluation metrics -- TEST INFRASTRUCTURE (only tests/, smoke() and bench.py's
baseline legs may import this package).

Follows core/utils/metric.py: ``_get_correct_score`` (132-157), ``set_metrics`` (49-107), ``get_metrics`` (109-130).
Pinned against outputs of the unmodified reference classes: tests/golden/eval_metric.npz, made by
``oracle/make_golden_eval.py``.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np


def correct_and_confusion(out: np.ndarray, target: np.ndarray, maxk: int):
    """metric.py:147-157: the ``maxk`` best classes per row in descending order (ties: lower index first),
    correct[j, i] = (pred[j, i] == target[i]), conf[target[i], pred[0, i]] += 1."""
    b, c = out.shape
    order = np.argsort(-out.astype(np.float64), axis=1, kind="stable")[:, :maxk]  # (b, maxk)
    preds = order.T  # (maxk, b)
    correct = preds == target.reshape(1, -1)
    conf = np.zeros((c, c), np.float32)
    for i1, i2 in zip(target.reshape(-1), preds[0]):
        conf[i1, i2] += 1
    return correct, preds, conf


class MetricOracle:
    """Accumulates exactly like the reference ``Metric`` (losses given as float32 scalars)."""

    def __init__(self, num_classes: "OrderedDict[str, int]", topk, no_batches: int, loss_keys):
        self.topk = list(topk)
        self.no_batches = no_batches
        self.multi = len(num_classes) > 1
        self.accuracy = OrderedDict((k, [0] * len(self.topk)) for k in num_classes)
        self.conf = OrderedDict((k, np.zeros((n, n), np.float32)) for k, n in num_classes.items())
        if self.multi:
            self.accuracy["all_class"] = [0] * len(self.topk)
        self.loss = OrderedDict((k, 0) for k in loss_keys)

    def set_metrics(self, out, target, batch_size, batch_loss):
        correct = OrderedDict()
        for key in out:
            if key == "weights":
                continue
            correct[key], _, cm = correct_and_confusion(out[key], target[key], max(self.topk))
            self.conf[key] += cm
        for key in self.loss:
            self.loss[key] += float(np.float32(batch_loss[key]))  # metric.py:79-93 (.item() of an fp32 scalar)
        scale = np.float32(100.0 / batch_size)  # metric.py:102/105: fp32 tensor.mul_(python scalar)
        for key in self.accuracy:
            for i, k in enumerate(self.topk):
                if key == "all_class":
                    c = None
                    for corr in correct.values():
                        hit = corr[:k].sum(0)
                        c = hit if c is None else c * hit
                    n = c.sum()
                else:
                    n = correct[key][:k].sum()
                self.accuracy[key][i] += float(np.float32(n) * scale)

    def get_metrics(self):
        acc = OrderedDict((k, [round(x / self.no_batches, 2) for x in v]) for k, v in self.accuracy.items())
        loss = OrderedDict((k, round(v / self.no_batches, 5)) for k, v in self.loss.items())
        return loss, acc, self.conf
