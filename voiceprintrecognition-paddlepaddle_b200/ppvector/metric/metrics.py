"""EER / minDCF from a flat list of scores and 0/1 labels -- same definitions as ppvector/metric/metrics.py:4-37 of the
reference (sorted-threshold sweep; EER by linear interpolation around the FNR/FPR crossing; minDCF with p_target 0.01,
c_miss = c_fa = 1, normalised by the default cost).  Host numpy: it runs once per evaluation on ~1e6 scores."""
import numpy as np


def compute_fnr_fpr(scores, labels, weights=None):
    """reference: metrics.py:4-19.  Returns (fnr, fpr, thresholds) along the ascending-score sweep."""
    order = np.argsort(scores)
    thresholds = scores[order]
    labels = labels[order]
    w = np.ones(labels.shape, dtype=np.float64) if weights is None else weights[order]
    target_w = w * (labels == 1)
    impostor_w = w * (labels == 0)
    fnr = np.cumsum(target_w) / np.sum(target_w)
    fpr = 1.0 - np.cumsum(impostor_w) / np.sum(impostor_w)
    return fnr, fpr, thresholds


def compute_eer(fnr, fpr, scores=None):
    """reference: metrics.py:22-31"""
    diff = fnr - fpr
    x1 = np.flatnonzero(diff >= 0)[0]
    x2 = np.flatnonzero(diff < 0)[-1]
    a = (fnr[x1] - fpr[x1]) / (fpr[x2] - fpr[x1] - (fnr[x2] - fnr[x1]))
    eer = fnr[x1] + a * (fnr[x2] - fnr[x1])
    if scores is not None:
        return eer, np.sort(scores)[x1]
    return eer


def compute_dcf(fnr, fpr, p_target=0.01, c_miss=1, c_fa=1):
    """reference: metrics.py:34-37"""
    c_det = np.min(c_miss * fnr * p_target + c_fa * fpr * (1 - p_target))
    return c_det / min(c_miss * p_target, c_fa * (1 - p_target))
