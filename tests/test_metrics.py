"""CPU: ppvector.metric.metrics against golden values produced by the REFERENCE's own metrics.py (pure numpy, so it
could be run in the authoring container: tests/golden/make_golden.py) -- the one piece of the path that IS pinned to
reference outputs."""
import numpy as np

from ppvector.metric.metrics import compute_dcf, compute_eer, compute_fnr_fpr


def test_eer_mindcf_match_reference(golden_dir):
    g = np.load(f"{golden_dir}/metrics_ref.npz")
    fnr, fpr, thr = compute_fnr_fpr(g["scores"], g["labels"])
    eer, threshold = compute_eer(fnr, fpr, g["scores"])
    assert np.array_equal(fnr[:50], g["fnr_head"]) and np.array_equal(fpr[-50:], g["fpr_tail"])
    assert float(eer) == float(g["eer"]) and float(threshold) == float(g["threshold"])
    assert float(compute_dcf(fnr, fpr)) == float(g["min_dcf"])


def test_perfect_and_chance_systems():
    s = np.concatenate([np.full(100, 0.9, np.float32), np.full(900, 0.1, np.float32)]) + np.linspace(0, 1e-3, 1000, dtype=np.float32)
    lab = np.concatenate([np.ones(100, np.int32), np.zeros(900, np.int32)])
    fnr, fpr, _ = compute_fnr_fpr(s, lab)
    assert compute_eer(fnr, fpr) < 1e-2 and compute_dcf(fnr, fpr) < 2e-2
