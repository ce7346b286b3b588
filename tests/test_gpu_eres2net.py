"""GPU: ERes2Net forward (Res2Net splits, AFF fusion, bottom-up stage fusion, TSTP) vs the fp64 oracle and golden
embeddings; SURVEY.md §8 row a7.  Tolerance: cosine scores within 1e-4 of the reference path."""
import numpy as np
import pytest
import torch

from oracle import eres2net as oe
from oracle import head as oh
from ppvector.models.eres2net import ERes2Net

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def W64():
    return oe.make_eres2net_weights(seed=1000, dtype=torch.float64)


@pytest.fixture(scope="module")
def model(cuda, W64):
    m = ERes2Net(input_size=80).eval()
    m.load_state_dict({k: v.float() for k, v in W64.items()}, strict=True)
    return m.to(cuda)


def test_param_count_and_names(W64):
    assert oe.count_params(W64) == 6620128  # README.md:71 "ERes2Net 6.6 M"
    m = ERes2Net(input_size=80)
    assert sorted(m.state_dict().keys()) == sorted(W64.keys())


@pytest.mark.parametrize("T", [64, 149])
def test_stagewise_taps_and_embedding(cuda, model, W64, golden_dir, T):
    g = np.load(f"{golden_dir}/eres2net_seed1000.npz")
    gi = torch.Generator().manual_seed(3000 + T)
    f = torch.randn(2, T, 80, generator=gi, dtype=torch.float64)
    f = f - f.mean(1, keepdim=True)
    taps = {}
    ref = oe.eres2net_forward(f, W64, taps=taps)
    emb = model(f.float().to(cuda))
    torch.cuda.synchronize()
    for name in ["layer1", "layer2", "layer3", "layer4", "fuse12", "fuse123", "fuse1234", "stats"]:
        got = model.read_tap(name, 2, T).double().cpu()
        want = taps[name] if name == "stats" else taps[name].permute(0, 2, 3, 1)
        assert got.shape == want.shape, (name, got.shape, want.shape)
        rel = (got - want).norm() / want.norm()
        assert rel < 5e-5, (name, rel.item())
    emb = emb.double().cpu()
    assert np.abs(emb.numpy() - g[f"emb_T{T}"]).max() < 1e-4
    cos = torch.nn.functional.cosine_similarity(emb, ref)
    assert (1 - cos).max() < 1e-8
    assert np.abs(oh.cosine_matrix(emb.numpy(), emb.numpy()) - oh.cosine_matrix(ref.numpy(), ref.numpy())).max() < 1e-4


@pytest.mark.parametrize("B,T", [(1, 16), (3, 33), (4, 298)])
def test_shapes(cuda, model, W64, B, T):
    gi = torch.Generator().manual_seed(B * 100 + T)
    f = torch.randn(B, T, 80, generator=gi)
    ref = oe.eres2net_forward(f[:2].double(), W64)
    emb = model(f.to(cuda)).double().cpu()
    assert emb.shape == (B, 192)
    rel = (emb[: ref.shape[0]] - ref).norm(dim=1) / ref.norm(dim=1)
    assert rel.max() < 1e-4, rel


def test_batch_independence(cuda, model):
    gi = torch.Generator().manual_seed(12)
    f = torch.randn(8, 298, 80, generator=gi).to(cuda)
    emb = model(f)
    assert torch.isfinite(emb).all()
    for b in (0, 7):
        assert torch.equal(model(f[b:b + 1]), emb[b:b + 1])
