"""Metric wrappers (recman/tf/core/metric.py:5-27); sklearn's log_loss lost its eps argument, so clip here."""
import numpy as np


class LogLoss:
    def __init__(self, eps=1e-07):
        self.eps = eps

    def __call__(self, y_true, y_pred):
        from sklearn.metrics import log_loss

        p = np.clip(np.asarray(y_pred, dtype=np.float64), self.eps, 1 - self.eps)
        return log_loss(y_true, p, labels=[0, 1])

    def __str__(self):
        return "logloss"

    __repr__ = __str__


class RocAucScore:
    def __call__(self, y_true, y_pred):
        from sklearn.metrics import roc_auc_score

        return roc_auc_score(y_true, y_pred)

    def __str__(self):
        return "roc_auc"

    __repr__ = __str__
