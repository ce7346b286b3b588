"""Verification metrics from a flat array of scores and 0/1 labels: the FNR / FPR operating curve, the equal error rate with its
threshold, and the minimum detection cost.  Definitions are those of ppvector/metric/metrics.py:4-37 of the reference
(thresholds swept over the sorted scores; EER by linear interpolation between the two operating points around the FNR = FPR
crossing; minDCF with p_target 0.01, c_miss = c_fa = 1, normalised by the cost of the trivial system); tests/test_metrics.py
pins the outputs to the reference's own functions.  The numpy functions keep the reference's signatures (array-returning
``compute_fnr_fpr``); ``eer_mindcf_gpu`` / ``eer_mindcf_from_matrix_gpu`` run the same definitions on the device (libppv_b200
``ppv_eer_mindcf*``: radix sort + one sweep, csrc/metrics.cu) and are what ``PPVectorTrainer.evaluate`` uses."""
import ctypes as C

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


def _eer_call(n, launch):
    import torch

    from ppvector import _lib
    lib = _lib.load()
    dev = torch.device('cuda', torch.cuda.current_device())
    nbytes = lib.ppv_eer_workspace_bytes(n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = torch.empty(4, dtype=torch.float64, device=dev)
    launch(lib, out, ws, nbytes)
    eer, thr, dcf, nt = out.cpu().tolist()
    return float(eer), float(dcf), float(thr)


def eer_mindcf_gpu(scores, labels, p_target=0.01, c_miss=1.0, c_fa=1.0):
    """scores [n] float32, labels [n] (1 = target) CUDA tensors (or arrays) -> (eer, min_dcf, threshold) like
    compute_fnr_fpr + compute_eer + compute_dcf of metrics.py:4-37, computed on the device."""
    import torch

    from ppvector import _lib
    s = torch.as_tensor(scores).to(device='cuda', dtype=torch.float32).contiguous().reshape(-1)
    lab = torch.as_tensor(labels).to(device=s.device, dtype=torch.int32).contiguous().reshape(-1)
    assert s.numel() == lab.numel()
    with torch.cuda.device(s.device):
        return _eer_call(s.numel(), lambda lib, out, ws, nb: _lib.check(
            lib.ppv_eer_mindcf(_lib.ptr(s), _lib.ptr(lab), s.numel(), float(p_target), float(c_miss), float(c_fa), _lib.ptr(out),
                               C.c_void_p(ws.data_ptr()), nb, _lib.current_stream()), 'ppv_eer_mindcf'))


def eer_mindcf_from_matrix_gpu(score_matrix, trial_labels, enroll_labels, p_target=0.01, c_miss=1.0, c_fa=1.0):
    """The evaluation loop's form (trainer.py:416-431): scores [M,N] (CUDA), label[i,j] = (trial_labels[i] == enroll_labels[j]);
    neither the flattened labels nor the scores visit the host."""
    import torch

    from ppvector import _lib
    s = torch.as_tensor(score_matrix).to(device='cuda', dtype=torch.float32).contiguous()
    M, N = s.shape
    tl = torch.as_tensor(trial_labels).to(device=s.device, dtype=torch.int32).contiguous()
    el = torch.as_tensor(enroll_labels).to(device=s.device, dtype=torch.int32).contiguous()
    assert tl.numel() == M and el.numel() == N
    with torch.cuda.device(s.device):
        return _eer_call(M * N, lambda lib, out, ws, nb: _lib.check(
            lib.ppv_eer_mindcf_matrix(_lib.ptr(s), _lib.ptr(tl), _lib.ptr(el), M, N, float(p_target), float(c_miss), float(c_fa),
                                      _lib.ptr(out), C.c_void_p(ws.data_ptr()), nb, _lib.current_stream()), 'ppv_eer_mindcf_matrix'))
