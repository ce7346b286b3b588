"""GPU: CAM++ forward (FCM head, stride-2 TDNN over frame pairs, context-aware masked dense TDNN blocks, statistics pooling)
vs the fp64 oracle and golden embeddings; SURVEY.md §8 row a8.  Tolerance: cosine scores within 1e-4 of the reference path."""
import numpy as np
import pytest
import torch

from oracle import campplus as oc
from oracle import head as oh
from ppvector.models.campplus import CAMPPlus

pytestmark = pytest.mark.gpu

TAPS = ["head.layer1", "head.layer2", "tdnn", "block1", "transit1", "block2", "transit2", "block3", "out_nonlinear", "stats"]


@pytest.fixture(scope="module")
def W64():
    return oc.make_campplus_weights(seed=1000, dtype=torch.float64)


@pytest.fixture(scope="module")
def model(cuda, W64):
    m = CAMPPlus(input_size=80).eval()
    m.load_state_dict({k: v.float() for k, v in W64.items()}, strict=True)
    return m.to(cuda)


def test_param_count_and_names(W64):
    assert oc.count_params(W64) == 6859232  # README.md:72 "CAM++ 6.8 M"
    m = CAMPPlus(input_size=80)
    assert sorted(m.state_dict().keys()) == sorted(W64.keys())


# T = 64: one context segment; 298 (3 s): two segments, the second short; 451: odd length, three segments
@pytest.mark.parametrize("T", [64, 298, 451])
def test_stagewise_taps_and_embedding(cuda, model, W64, golden_dir, T):
    g = np.load(f"{golden_dir}/campplus_seed1000.npz")
    gi = torch.Generator().manual_seed(4000 + T)
    f = torch.randn(2, T, 80, generator=gi, dtype=torch.float64)
    f = f - f.mean(1, keepdim=True)
    taps = {}
    ref = oc.campplus_forward(f, W64, taps=taps)
    emb = model(f.float().to(cuda))
    torch.cuda.synchronize()
    for name in TAPS:
        got = model.read_tap(name, 2, T).double().cpu()
        want = taps[name]
        if name.startswith("head."):
            want = want.permute(0, 2, 3, 1)
        elif name != "stats":
            want = want.transpose(1, 2)
        assert got.shape == want.shape, (name, got.shape, want.shape)
        rel = (got - want).norm() / want.norm()
        assert rel < 5e-5, (name, rel.item())
    emb = emb.double().cpu()
    assert np.abs(emb.numpy() - g[f"emb_T{T}"]).max() < 1e-4
    cos = torch.nn.functional.cosine_similarity(emb, ref)
    assert (1 - cos).max() < 1e-8
    assert np.abs(oh.cosine_matrix(emb.numpy(), emb.numpy()) - oh.cosine_matrix(ref.numpy(), ref.numpy())).max() < 1e-4


@pytest.mark.parametrize("B,T", [(1, 3), (3, 33), (5, 201), (2, 1000)])
def test_shapes(cuda, model, W64, B, T):
    gi = torch.Generator().manual_seed(B * 100 + T)
    f = torch.randn(B, T, 80, generator=gi)
    ref = oc.campplus_forward(f[:2].double(), W64)
    emb = model(f.to(cuda)).double().cpu()
    assert emb.shape == (B, 192)
    rel = (emb[: ref.shape[0]] - ref).norm(dim=1) / ref.norm(dim=1)
    assert rel.max() < 1e-4, rel


def test_batch_independence(cuda, model):
    gi = torch.Generator().manual_seed(13)
    f = torch.randn(8, 298, 80, generator=gi).to(cuda)
    emb = model(f)
    assert torch.isfinite(emb).all()
    for b in (0, 7):
        assert torch.equal(model(f[b:b + 1]), emb[b:b + 1])


def test_repeat_is_bitwise_stable(cuda, model):
    gi = torch.Generator().manual_seed(14)
    f = torch.randn(4, 149, 80, generator=gi).to(cuda)
    a = model(f).clone()
    _ = model(torch.randn(4, 149, 80, generator=gi).to(cuda))
    assert torch.equal(model(f), a)
