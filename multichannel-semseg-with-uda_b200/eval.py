"""Evaluation counts of the reference's eval.py on the GPU: the confusion matrix is accumulated on the device straight
from the testers' predictions (no PNG round trip), the scores derived from it are the reference's formulas.

    fast_hist(a, b, n)            eval.py:21-23     a = ground truth, b = prediction
    per_class_iu(hist)            eval.py:26-27
    calc_fw_iu(hist)              eval.py:30-35
    calc_pixel_accuracy(hist)     eval.py:38-40
    calc_mean_accuracy(hist)      eval.py:43-46
"""
import numpy as np

from mcd_b200 import pipeline


def fast_hist(a, b, n):
    """GPU tensors (uint8 / int64) -> numpy int64 [n, n]; same counts as the reference's bincount."""
    return pipeline.hist_matrix(pipeline.fast_hist(a, b, n), n)


class ConfusionMatrix:
    """hist += fast_hist(gt, pred, n) over a whole test set without leaving the device (eval.py:128-134 loop)."""

    def __init__(self, n):
        self.n, self.counts = n, None

    def update(self, gt, pred):
        self.counts = pipeline.fast_hist(gt, pred, self.n, self.counts)

    def hist(self):
        return pipeline.hist_matrix(self.counts, self.n)


def per_class_iu(hist):
    tp = np.diag(hist)
    return tp / (hist.sum(1) + hist.sum(0) - tp)


def calc_fw_iu(hist):
    pred, gt, tp = hist.sum(0), hist.sum(1), np.diag(hist)
    return np.nansum((gt * tp) / (pred + gt - tp)) / gt.sum()


def calc_pixel_accuracy(hist):
    return np.diag(hist).sum() / hist.sum(1).sum()


def calc_mean_accuracy(hist):
    return np.nanmean(np.diag(hist) / hist.sum(1))
