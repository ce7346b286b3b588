"""Oracle: CAM++ (CAMPPlus) forward on the CPU (torch functional ops, fp32 or fp64).

TEST INFRASTRUCTURE -- see oracle/__init__.py.

Weights: flat dict keyed like the reference's Paddle state_dict for ``CAMPPlus`` (default arguments of configs/cam++.yml:
growth_rate 32, bn_size 4, init_channels 128, blocks (12, 24, 16) with kernel 3 and dilations (1, 2, 2)).

Follows
  * ppvector/models/campplus.py:216-251  BasicResBlock (stride on the frequency axis only)
  * ppvector/models/campplus.py:254-289  FCM head, reshape [B, C, F/8, T] -> [B, C*F/8, T]
  * ppvector/models/campplus.py:37-64    TDNNLayer (Conv1D k5 stride 2, zero padding 2, BatchNorm, ReLU)
  * ppvector/models/campplus.py:67-106   CAMLayer: local conv * sigmoid(MLP(mean over time + 100-frame segment mean))
  * ppvector/models/campplus.py:109-141  CAMDenseTDNNLayer: BN-ReLU -> 1x1 -> BN-ReLU -> CAMLayer
  * ppvector/models/campplus.py:144-171  dense connectivity (concat on channels)
  * ppvector/models/campplus.py:174-186  TransitLayer: BN-ReLU -> 1x1
  * ppvector/models/campplus.py:24-31    statistics_pooling: mean | unbiased std (no epsilon)
  * ppvector/models/campplus.py:189-204  DenseLayer: 1x1 -> BatchNorm
  * ppvector/models/campplus.py:292-346  CAMPPlus

Paddle semantics restated here: nn.Conv1D / nn.Conv2D zero-pad and carry a bias by default (the ``bias=False`` arguments of
the reference's layers are not forwarded to the convs); F.avg_pool1d(ceil_mode=True) with the default exclusive=True
averages the last, short segment over the frames it really holds.
"""
import math
from typing import Dict

import torch
import torch.nn.functional as F

from oracle.ecapa import batchnorm_eval

BLOCKS = ((12, 3, 1), (24, 3, 2), (16, 3, 2))  # (num_layers, kernel_size, dilation), campplus.py:316


def conv2d_bn(x, W, conv, bn, stride=(1, 1), padding=0):
    return batchnorm_eval(F.conv2d(x, W[conv + ".weight"], W[conv + ".bias"], stride=stride, padding=padding), W, bn)


def basic_res_block(x, W, p, stride):
    """campplus.py:245-251"""
    out = F.relu(conv2d_bn(x, W, p + ".conv1", p + ".bn1", stride=(stride, 1), padding=1))
    out = conv2d_bn(out, W, p + ".conv2", p + ".bn2", padding=1)
    sc = x
    if p + ".shortcut.0.weight" in W:
        sc = conv2d_bn(x, W, p + ".shortcut.0", p + ".shortcut.1", stride=(stride, 1))
    return F.relu(out + sc)


def fcm(x, W, taps=None):
    """campplus.py:278-289.  x [B,F,T] -> [B, 32*F/8, T]"""
    out = F.relu(conv2d_bn(x.unsqueeze(1), W, "head.conv1", "head.bn1", padding=1))
    for li in (1, 2):
        for bi in range(2):
            out = basic_res_block(out, W, f"head.layer{li}.{bi}", 2 if bi == 0 else 1)
        if taps is not None:
            taps[f"head.layer{li}"] = out
    out = F.relu(conv2d_bn(out, W, "head.conv2", "head.bn2", stride=(2, 1), padding=1))
    return out.reshape(out.shape[0], out.shape[1] * out.shape[2], out.shape[3])


def bn_relu(x, W, p):
    return F.relu(batchnorm_eval(x, W, p + ".batchnorm"))


def seg_pooling(x, seg_len=100):
    """campplus.py:95-106 (avg): per-segment mean broadcast back over the segment's frames."""
    T = x.shape[-1]
    nseg = (T + seg_len - 1) // seg_len
    out = torch.empty_like(x)
    for s in range(nseg):
        a, b = s * seg_len, min((s + 1) * seg_len, T)
        out[..., a:b] = x[..., a:b].mean(dim=-1, keepdim=True)
    return out


def cam_layer(x, W, p, dilation):
    """campplus.py:88-93"""
    y = F.conv1d(x, W[p + ".linear_local.weight"], W[p + ".linear_local.bias"], padding=dilation, dilation=dilation)
    context = x.mean(-1, keepdim=True) + seg_pooling(x)
    context = F.relu(F.conv1d(context, W[p + ".linear1.weight"], W[p + ".linear1.bias"]))
    m = torch.sigmoid(F.conv1d(context, W[p + ".linear2.weight"], W[p + ".linear2.bias"]))
    return y * m


def campplus_forward(feats, W: Dict[str, torch.Tensor], taps=None):
    """campplus.py:342-346.  feats [B,T,F] -> [B,embd_dim]."""
    x = fcm(feats.transpose(1, 2), W, taps)
    if taps is not None:
        taps["head"] = x
    x = F.conv1d(x, W["xvector.tdnn.linear.weight"], W["xvector.tdnn.linear.bias"], stride=2, padding=2)
    x = bn_relu(x, W, "xvector.tdnn.nonlinear")
    if taps is not None:
        taps["tdnn"] = x
    for bi, (num_layers, _k, dil) in enumerate(BLOCKS, start=1):
        for li in range(1, num_layers + 1):
            p = f"xvector.block{bi}.tdnnd{li}"
            h = bn_relu(x, W, p + ".nonlinear1")
            h = F.conv1d(h, W[p + ".linear1.weight"], W[p + ".linear1.bias"])
            h = bn_relu(h, W, p + ".nonlinear2")
            x = torch.cat((x, cam_layer(h, W, p + ".cam_layer", dil)), dim=1)
        if taps is not None:
            taps[f"block{bi}"] = x
        p = f"xvector.transit{bi}"
        x = bn_relu(x, W, p + ".nonlinear")
        x = F.conv1d(x, W[p + ".linear.weight"], W[p + ".linear.bias"])
        if taps is not None:
            taps[f"transit{bi}"] = x
    x = bn_relu(x, W, "xvector.out_nonlinear")
    if taps is not None:
        taps["out_nonlinear"] = x
    stats = torch.cat((x.mean(dim=-1), x.std(dim=-1, unbiased=True)), dim=-1)
    if taps is not None:
        taps["stats"] = stats
    e = F.conv1d(stats.unsqueeze(-1), W["xvector.dense.linear.weight"], W["xvector.dense.linear.bias"]).squeeze(-1)
    return batchnorm_eval(e, W, "xvector.dense.nonlinear.batchnorm")


def campplus_param_shapes(input_size=80, embd_dim=192, growth_rate=32, bn_size=4, init_channels=128):
    S = {}

    def conv2(p, cin, cout, k):
        S[p + ".weight"] = (cout, cin, k, k)
        S[p + ".bias"] = (cout,)

    def conv1(p, cin, cout, k):
        S[p + ".weight"] = (cout, cin, k)
        S[p + ".bias"] = (cout,)

    def bn(p, c):
        for n in ("weight", "bias", "_mean", "_variance"):
            S[f"{p}.{n}"] = (c,)

    m = 32
    conv2("head.conv1", 1, m, 3)
    bn("head.bn1", m)
    for li in (1, 2):
        for bi in range(2):
            p = f"head.layer{li}.{bi}"
            conv2(p + ".conv1", m, m, 3)
            bn(p + ".bn1", m)
            conv2(p + ".conv2", m, m, 3)
            bn(p + ".bn2", m)
            if bi == 0:
                conv2(p + ".shortcut.0", m, m, 1)
                bn(p + ".shortcut.1", m)
    conv2("head.conv2", m, m, 3)
    bn("head.bn2", m)
    channels = m * math.ceil(input_size / 8)
    conv1("xvector.tdnn.linear", channels, init_channels, 5)
    bn("xvector.tdnn.nonlinear.batchnorm", init_channels)
    channels = init_channels
    bnc = bn_size * growth_rate
    for bi, (num_layers, k, _d) in enumerate(BLOCKS, start=1):
        for li in range(1, num_layers + 1):
            p = f"xvector.block{bi}.tdnnd{li}"
            cin = channels + (li - 1) * growth_rate
            bn(p + ".nonlinear1.batchnorm", cin)
            conv1(p + ".linear1", cin, bnc, 1)
            bn(p + ".nonlinear2.batchnorm", bnc)
            conv1(p + ".cam_layer.linear_local", bnc, growth_rate, k)
            conv1(p + ".cam_layer.linear1", bnc, bnc // 2, 1)
            conv1(p + ".cam_layer.linear2", bnc // 2, growth_rate, 1)
        channels += num_layers * growth_rate
        bn(f"xvector.transit{bi}.nonlinear.batchnorm", channels)
        conv1(f"xvector.transit{bi}.linear", channels, channels // 2, 1)
        channels //= 2
    bn("xvector.out_nonlinear.batchnorm", channels)
    conv1("xvector.dense.linear", channels * 2, embd_dim, 1)
    bn("xvector.dense.nonlinear.batchnorm", embd_dim)
    return S


def make_campplus_weights(seed=1000, dtype=torch.float32, **shape_args) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    W = {}
    for name, shape in campplus_param_shapes(**shape_args).items():
        is_bn = ".batchnorm." in name or ".bn" in name or ".shortcut.1." in name
        if name.endswith("_variance") or (is_bn and name.endswith(".weight")):
            t = torch.rand(shape, generator=g, dtype=torch.float64) + 0.5
        elif name.endswith("_mean") or (is_bn and name.endswith(".bias")):
            t = torch.randn(shape, generator=g, dtype=torch.float64) * 0.1
        elif name.endswith(".weight"):
            fan_in = int(torch.tensor(shape[1:]).prod())
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * math.sqrt(3.0 / fan_in)
        else:
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * 0.1
        W[name] = t.to(dtype)
    return W


def count_params(W) -> int:
    return sum(v.numel() for k, v in W.items() if not (k.endswith("_mean") or k.endswith("_variance")))
