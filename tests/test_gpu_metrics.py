"""GPU: EER / minDCF (csrc/metrics.cu) against the reference's own metrics.py outputs (tests/golden/metrics_ref.npz was produced by
/root/reference/ppvector/metric/metrics.py) and against the host mirror on random score sets; enrol-DB retrieval arg-max."""
import numpy as np
import pytest
import torch

from ppvector.metric.cosine import cosine_matrix, retrieval
from ppvector.metric.metrics import compute_dcf, compute_eer, compute_fnr_fpr, eer_mindcf_from_matrix_gpu, eer_mindcf_gpu

pytestmark = pytest.mark.gpu


def host(scores, labels):
    fnr, fpr, _ = compute_fnr_fpr(scores, labels)
    eer, thr = compute_eer(fnr, fpr, scores)
    return float(eer), float(compute_dcf(fnr, fpr)), float(thr)


def test_against_the_reference_metrics_fixture(cuda, golden_dir):
    g = np.load(f"{golden_dir}/metrics_ref.npz")
    eer, dcf, thr = eer_mindcf_gpu(torch.from_numpy(g["scores"]).to(cuda), torch.from_numpy(g["labels"]).to(cuda))
    assert abs(eer - float(g["eer"])) < 1e-12
    assert abs(dcf - float(g["min_dcf"])) < 1e-12
    assert thr == pytest.approx(float(g["threshold"]), abs=0)


@pytest.mark.parametrize("n", [64, 257, 2048, 2049, 100003, 1000000])  # metrics.py itself raises when no FNR < FPR point exists (tiny sets)
def test_random_sets_equal_the_host_definitions(cuda, n):
    rng = np.random.default_rng(n)
    labels = (rng.random(n) < 0.2).astype(np.int32)
    labels[0], labels[-1] = 1, 0  # both classes present
    scores = (rng.standard_normal(n) * 0.3 + 0.5 * labels).astype(np.float32)
    scores[rng.integers(0, n, n // 50)] *= -1.0  # negative keys exercise the order-preserving float mapping
    want = host(scores, labels)
    got = eer_mindcf_gpu(torch.from_numpy(scores).to(cuda), torch.from_numpy(labels).to(cuda))
    # ties between equal scores may order differently (numpy's argsort is unspecified among ties): continuous scores have none here
    assert len(np.unique(scores)) == n or n > 100000
    assert abs(got[0] - want[0]) < 1e-9 and abs(got[1] - want[1]) < 1e-9 and got[2] == pytest.approx(want[2], abs=1e-7)


def test_matrix_form_matches_flat_form_and_config5_size(cuda):
    g = torch.Generator().manual_seed(5)
    M, N = 1000, 1000
    tl = torch.randint(0, 40, (M,), generator=g, dtype=torch.int32)
    el = torch.randint(0, 40, (N,), generator=g, dtype=torch.int32)
    Et = torch.randn(M, 192, generator=g) + torch.nn.functional.one_hot(tl.long(), 192).float() * 3
    Ee = torch.randn(N, 192, generator=g) + torch.nn.functional.one_hot(el.long(), 192).float() * 3
    S = cosine_matrix(Et.to(cuda), Ee.to(cuda))
    a = eer_mindcf_from_matrix_gpu(S, tl, el)
    labels = (tl[:, None] == el[None, :]).int().reshape(-1)
    b = eer_mindcf_gpu(S.reshape(-1), labels.to(cuda))
    assert a == b
    want = host(S.cpu().numpy().reshape(-1), labels.numpy())
    assert abs(a[0] - want[0]) < 1e-9 and abs(a[1] - want[1]) < 1e-9


def test_retrieval_argmax(cuda):
    g = torch.Generator().manual_seed(9)
    db = torch.randn(37, 192, generator=g)
    q = db[[5, 5, 30]] + 0.05 * torch.randn(3, 192, generator=g)
    q = torch.cat([q, torch.randn(1, 192, generator=g)])
    names = [f"user{i}" for i in range(37)]
    r = retrieval(q.to(cuda), db.to(cuda), threshold=0.6, names=names)
    assert [x[0] for x in r] == ["user5", "user5", "user30", None]
    sim = torch.nn.functional.cosine_similarity(q[0], db[5], dim=0).item()
    assert abs(r[0][1] - round(sim, 5)) < 2e-5
