"""GPU: ERes2NetV2 forward (reference eres2net.py:266-462: chunk widths 13 / 26 / 52 / 104 zero-padded to 32 / 32 / 64 / 128 columns on the
device, AFF blocks in layers 3-4, layer3_ds + fuse34, TSTP) vs the fp64 oracle -- itself pinned to the reference's ERes2NetV2 class in
tests/test_oracle_vs_reference.py -- and vs the reference's own outputs (tests/golden/ref_models.npz); SURVEY.md §8 row f4.
Tolerance: cosine scores within 1e-4 of the reference path."""
import numpy as np
import pytest
import torch

from oracle import eres2net as oe
from oracle import head as oh
from ppvector.models.eres2net import ERes2NetV2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def W64():
    return oe.make_eres2net_weights(seed=1000, dtype=torch.float64, base_width=26, version=2)


@pytest.fixture(scope="module")
def model(cuda, W64):
    m = ERes2NetV2(input_size=80).eval()
    m.load_state_dict({k: v.float() for k, v in W64.items()}, strict=True)
    return m.to(cuda)


def test_names_equal_the_oracle_state_dict(W64):
    m = ERes2NetV2(input_size=80)
    assert sorted(m.state_dict().keys()) == sorted(W64.keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(W64[k].shape), k


@pytest.mark.parametrize("T", [64, 149])
def test_stagewise_taps_and_embedding(cuda, model, W64, T):
    gi = torch.Generator().manual_seed(3000 + T)
    f = torch.randn(2, T, 80, generator=gi, dtype=torch.float64)
    f = f - f.mean(1, keepdim=True)
    taps = {}
    ref = oe.eres2net_forward(f, W64, taps=taps, base_width=26, version=2)
    emb = model(f.float().to(cuda))
    torch.cuda.synchronize()
    for name in ["layer1", "layer2", "layer3", "layer4", "fuse34", "stats"]:
        got = model.read_tap(name, 2, T).double().cpu()
        want = taps[name] if name == "stats" else taps[name].permute(0, 2, 3, 1)
        assert got.shape == want.shape, (name, got.shape, want.shape)
        rel = (got - want).norm() / want.norm()
        assert rel < 5e-5, (name, rel.item())
    emb = emb.double().cpu()
    cos = torch.nn.functional.cosine_similarity(emb, ref)
    assert (1 - cos).max() < 1e-8
    assert np.abs(oh.cosine_matrix(emb.numpy(), emb.numpy()) - oh.cosine_matrix(ref.numpy(), ref.numpy())).max() < 1e-4


@pytest.mark.parametrize("T", [98, 298])
def test_embedding_equals_the_reference_class_output(cuda, model, golden_dir, T):
    """the inputs and embeddings the REFERENCE's ERes2NetV2 produced under the paddle shim (make_ref_fixtures.py, seed-1000 weights)"""
    g = np.load(f"{golden_dir}/ref_models.npz")
    want = g[f"eres2netv2_T{T}_emb"]
    gen = torch.Generator().manual_seed(3000 + T)  # make_ref_fixtures.feats("eres2net", T): SEEDS["eres2net"] + T, mean-subtracted
    f = torch.randn(want.shape[0], T, 80, generator=gen, dtype=torch.float64)
    f = f - f.mean(1, keepdim=True)
    emb = model(f.float().to(cuda)).double().cpu().numpy()
    rel = np.linalg.norm(emb - want, axis=1) / np.linalg.norm(want, axis=1)
    assert rel.max() < 1e-4, rel


@pytest.mark.parametrize("B,T", [(1, 16), (3, 33), (4, 298)])
def test_shapes(cuda, model, W64, B, T):
    gi = torch.Generator().manual_seed(B * 100 + T)
    f = torch.randn(B, T, 80, generator=gi)
    ref = oe.eres2net_forward(f[:2].double(), W64, base_width=26, version=2)
    emb = model(f.to(cuda)).double().cpu()
    assert emb.shape == (B, 192)
    rel = (emb[: ref.shape[0]] - ref).norm(dim=1) / ref.norm(dim=1)
    assert rel.max() < 1e-4, rel


def test_batch_independence(cuda, model):
    gi = torch.Generator().manual_seed(12)
    f = torch.randn(8, 298, 80, generator=gi).to(cuda)
    emb = model(f)
    assert torch.isfinite(emb).all()
    for b in (0, 3, 7):
        assert torch.equal(model(f[b:b + 1]), emb[b:b + 1])
