"""CPU: the ECAPA-TDNN / head oracle against its golden vectors and the reference's structural known-answers."""
import math

import numpy as np
import pytest
import torch

from oracle import ecapa, head


def test_param_count_matches_reference_readme():
    # README.md:70-76 "EcapaTdnn 6.1 M"; SURVEY.md §6: 6.19 M trainable with the current code
    W = ecapa.make_ecapa_weights()
    assert ecapa.count_params(W) == 6194048
    # paddle.summary known-answers (README.md:303-351, 64-mel config): first conv 64*5*512+512, ASP tdnn, ASP conv, fc
    S = ecapa.ecapa_param_shapes(input_size=64)
    n = lambda k: int(np.prod(S[k]))
    assert n("blocks.0.conv.conv.weight") + n("blocks.0.conv.conv.bias") == 164352
    assert n("asp.tdnn.conv.conv.weight") + n("asp.tdnn.conv.conv.bias") == 589952
    assert n("asp.conv.conv.weight") + n("asp.conv.conv.bias") == 198144
    assert n("fc.conv.weight") + n("fc.conv.bias") == 590016


@pytest.mark.parametrize("T", [98, 298])
def test_embedding_golden(golden_dir, T):
    g = np.load(f"{golden_dir}/ecapa_seed1000.npz")
    W = ecapa.make_ecapa_weights(seed=1000, dtype=torch.float32)
    gi = torch.Generator().manual_seed(1000 + T)
    f = torch.randn(3, T, 80, generator=gi, dtype=torch.float64)
    f = (f - f.mean(1, keepdim=True)).float()
    emb = ecapa.ecapa_forward(f, W).double().numpy()
    ref = g[f"emb_T{T}"]
    cos = (emb * ref).sum(1) / np.linalg.norm(emb, axis=1) / np.linalg.norm(ref, axis=1)
    assert np.all(1 - cos < 1e-10)
    assert np.abs(emb - ref).max() < 1e-5


def test_reflect_same_padding():
    # utils.py:79-93: pad (L_in - L_out)//2 both sides with reflect; k=3,d=4 -> pad 4
    x = torch.arange(10.0).view(1, 1, 10)
    w = torch.zeros(1, 1, 3)
    w[0, 0, 0] = 1.0  # picks x[t - d]
    y = ecapa.conv1d_same_reflect(x, w, None, dilation=4)
    assert y.shape == x.shape
    assert y.flatten().tolist() == [4, 3, 2, 1, 0, 1, 2, 3, 4, 5]


def test_aam_and_cosine_golden(golden_dir):
    g = np.load(f"{golden_dir}/head_seed1000.npz")
    emb, W, labels = torch.from_numpy(g["emb"]), torch.from_numpy(g["W"]), torch.from_numpy(g["labels"])
    logits = head.cosine_logits(emb, W)
    assert np.abs(logits.numpy() - g["logits"]).max() < 1e-12
    assert logits.abs().max() <= 1 + 1e-12
    for margin, ls in [(0.0, 0.0), (0.2, 0.0), (0.3, 0.1)]:
        loss = head.aam_loss(logits, labels, margin=margin, scale=32.0, label_smoothing=ls)
        assert abs(loss.item() - float(g[f"loss_m{margin}_ls{ls}"])) < 1e-9
    # margin 0 == plain scaled softmax CE
    ce = torch.nn.functional.cross_entropy(32.0 * logits, labels)
    assert abs(ce.item() - float(g["loss_m0.0_ls0.0"])) < 1e-9
    C = head.cosine_matrix(g["cos_A"], g["cos_B"])
    assert np.abs(C - g["cos_AB"]).max() < 1e-12
    assert abs(head.cosine_pair(g["cos_A"][3], g["cos_B"][5]) - C[3, 5]) < 1e-12
    idx = np.array([[0, 1], [2, 2], [16, 3]])
    assert np.allclose(head.cosine_pairlist(g["cos_A"], idx), [head.cosine_pair(g["cos_A"][i], g["cos_A"][j]) for i, j in idx])


def test_margin_scheduler():
    # scheduler.py:43-102: 0 before 30 % of epochs, exp ramp, final after 70 %
    kw = dict(epochs=60, steps_per_epoch=100, initial_margin=0.0, final_margin=0.3)
    assert head.margin_schedule(0, **kw) == 0.0
    assert head.margin_schedule(18 * 100 - 1, **kw) == 0.0
    assert head.margin_schedule(42 * 100, **kw) == 0.3
    mid = head.margin_schedule(30 * 100, **kw)
    assert 0.0 < mid < 0.3
    expect = 0.3 * (1.0 - math.exp(0.5 * math.log(1e-3 / (1.0 + 1e-6))))
    assert abs(mid - expect) < 1e-12


# ------------------------------------------------------------------------------------------------ CAM++
def test_campplus_param_count_and_golden(golden_dir):
    from oracle import campplus
    W = campplus.make_campplus_weights(seed=1000, dtype=torch.float32)
    assert campplus.count_params(W) == 6859232  # README.md:72 "CAM++ 6.8 M"
    g = np.load(f"{golden_dir}/campplus_seed1000.npz")
    gi = torch.Generator().manual_seed(4000 + 64)
    f = torch.randn(2, 64, 80, generator=gi, dtype=torch.float64)
    f = (f - f.mean(1, keepdim=True)).float()
    emb = campplus.campplus_forward(f, W).double().numpy()
    assert np.abs(emb - g["emb_T64"]).max() < 2e-5


@pytest.mark.parametrize("T", [1, 99, 100, 101, 149, 250])
def test_campplus_seg_pooling_is_ceil_mode_avg_pool(T):
    # campplus.py:95-106: avg_pool1d(kernel 100, stride 100, ceil_mode=True), expand each segment value over 100 frames, crop to T
    from oracle import campplus
    x = torch.randn(2, 5, T, dtype=torch.float64)
    seg = torch.nn.functional.avg_pool1d(x, kernel_size=100, stride=100, ceil_mode=True)
    want = seg.unsqueeze(-1).expand(*seg.shape, 100).reshape(*seg.shape[:-1], -1)[..., :T]
    assert torch.allclose(campplus.seg_pooling(x), want, atol=1e-12)


def test_campplus_tdnn_frame_pair_identity():
    # the stride-2 k5 conv over frames equals five taps over (even | odd) frame pairs at offsets -1,-1,0,0,+1 (csrc/campplus.cu)
    g = torch.Generator().manual_seed(5)
    for T in (7, 8):
        x = torch.randn(1, 3, T, generator=g, dtype=torch.float64)
        w = torch.randn(4, 3, 5, generator=g, dtype=torch.float64)
        want = torch.nn.functional.conv1d(x, w, stride=2, padding=2)
        T2 = (T - 1) // 2 + 1
        ev = torch.zeros(1, 3, T2 + 2, dtype=torch.float64)
        od = torch.zeros(1, 3, T2 + 2, dtype=torch.float64)
        ev[..., 1:1 + (T + 1) // 2] = x[..., 0::2]
        od[..., 1:1 + T // 2] = x[..., 1::2]
        got = torch.zeros_like(want)
        for k, (src, off) in enumerate([(ev, -1), (od, -1), (ev, 0), (od, 0), (ev, 1)]):
            got += torch.einsum("oc,bct->bot", w[:, :, k], src[..., 1 + off:1 + off + T2])
        assert torch.allclose(got, want, atol=1e-12)
