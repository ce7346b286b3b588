"""Verification metrics from a flat array of scores and 0/1 labels: the FNR / FPR operating curve, the equal error rate with its
threshold, and the minimum detection cost.  Definitions are those of ppvector/metric/metrics.py:4-37 of the reference
(thresholds swept over the sorted scores; EER by linear interpolation between the two operating points around the FNR = FPR
crossing; minDCF with p_target 0.01, c_miss = c_fa = 1, normalised by the cost of the trivial system); tests/test_metrics.py
pins the outputs to the reference's own functions.  Host numpy: this runs once per evaluation on ~1e6 scores."""
import numpy as np


def compute_fnr_fpr(scores, labels, weights=None):
    """(fnr, fpr, thresholds), all aligned with the ascending-score order: fnr[i] / fpr[i] are the miss and false-alarm rates
    when everything up to and including the i-th smallest score is rejected."""
    scores = np.asarray(scores)
    perm = np.argsort(scores)
    lab = np.asarray(labels)[perm]
    mass = np.ones(lab.shape[0], dtype=np.float64) if weights is None else np.asarray(weights, dtype=np.float64)[perm]
    rejected_targets = np.cumsum(np.where(lab == 1, mass, 0.0))
    rejected_impostors = np.cumsum(np.where(lab == 0, mass, 0.0))
    miss_rate = rejected_targets / rejected_targets[-1]
    false_alarm_rate = 1.0 - rejected_impostors / rejected_impostors[-1]
    return miss_rate, false_alarm_rate, scores[perm]


def compute_eer(fnr, fpr, scores=None):
    """Equal error rate; with ``scores`` also the score at the crossing (the decision threshold)."""
    gap = np.asarray(fnr) - np.asarray(fpr)
    hi = int(np.argmax(gap >= 0))                        # first operating point with FNR >= FPR
    lo = len(gap) - 1 - int(np.argmax(gap[::-1] < 0))    # last operating point with FNR <  FPR
    # straight lines through the two points: FNR and FPR meet at fraction `frac` of the way from hi to lo
    frac = gap[hi] / ((fpr[lo] - fpr[hi]) - (fnr[lo] - fnr[hi]))
    eer = fnr[hi] + frac * (fnr[lo] - fnr[hi])
    if scores is None:
        return eer
    return eer, np.sort(scores)[hi]


def compute_dcf(fnr, fpr, p_target=0.01, c_miss=1, c_fa=1):
    """Minimum of the detection cost over the curve, divided by the cost of always accepting / always rejecting."""
    cost = c_miss * p_target * np.asarray(fnr) + c_fa * (1.0 - p_target) * np.asarray(fpr)
    trivial = min(c_miss * p_target, c_fa * (1.0 - p_target))
    return float(np.min(cost)) / trivial
