"""Oracle: ECAPA-TDNN forward on the CPU (torch functional ops, fp32 or fp64).

TEST INFRASTRUCTURE -- see oracle/__init__.py.

Weights are a flat ``dict[str, torch.Tensor]`` keyed exactly like the reference's
Paddle ``state_dict`` for ``EcapaTdnn`` (Conv1D ``weight`` [Cout,Cin,k] / ``bias``;
BatchNorm1D ``weight`` / ``bias`` / ``_mean`` / ``_variance``), e.g.
``blocks.1.res2net_block.blocks.3.conv.conv.weight``.

Follows
  * ppvector/models/utils.py:65-93    Conv1d: reflect "same" padding then conv
  * ppvector/models/utils.py:96-119   BatchNorm1d (eps 1e-5), eval mode
  * ppvector/models/utils.py:122-148  TDNNBlock = norm(relu(conv(x)))   (conv -> ReLU -> BN)
  * ppvector/models/ecapa_tdnn.py:36-47    Res2NetBlock.forward
  * ppvector/models/ecapa_tdnn.py:69-82    SEBlock.forward
  * ppvector/models/ecapa_tdnn.py:132-142  SERes2NetBlock.forward
  * ppvector/models/ecapa_tdnn.py:245-276  EcapaTdnn.forward
  * ppvector/models/pooling.py:86-125      AttentiveStatisticsPooling.forward
"""
import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
ASP_EPS = 1e-12


def conv1d_same_reflect(x, w, b, dilation=1):
    """utils.py:65-93.  stride 1: L_out = L_in - d*(k-1); pad (L_in-L_out)//2 both sides, reflect."""
    k = w.shape[-1]
    L_in = x.shape[-1]
    L_out = (L_in - dilation * (k - 1) - 1) // 1 + 1
    pad = (L_in - L_out) // 2
    if pad > 0:
        x = F.pad(x, (pad, pad), mode="reflect")
    return F.conv1d(x, w, b, dilation=dilation)


def batchnorm_eval(x, W: Dict[str, torch.Tensor], prefix: str):
    """utils.py:96-119 in eval mode: (x - mean) / sqrt(var + eps) * gamma + beta over channel axis 1."""
    g, b = W[prefix + ".weight"], W[prefix + ".bias"]
    m, v = W[prefix + "._mean"], W[prefix + "._variance"]
    shape = [1, -1] + [1] * (x.dim() - 2)
    return (x - m.view(shape)) / torch.sqrt(v.view(shape) + BN_EPS) * g.view(shape) + b.view(shape)


def tdnn_block(x, W, prefix, dilation=1, bn=batchnorm_eval, taps=None):
    """utils.py:147  return self.norm(self.activation(self.conv(x)));  ``bn`` = batchnorm_eval or a train-mode BatchNorm
    (oracle/train.py)"""
    y = conv1d_same_reflect(x, W[prefix + ".conv.conv.weight"], W[prefix + ".conv.conv.bias"], dilation)
    out = bn(F.relu(y), W, prefix + ".norm.norm")
    if taps is not None:  # layer-level taps (gradient localisation in tests): conv output and block output
        taps[prefix + ".z"] = y
        taps[prefix] = out
    return out


def res2net_block(x, W, prefix, scale=8, dilation=1, bn=batchnorm_eval, taps=None):
    """ecapa_tdnn.py:36-47"""
    ys = []
    y_i = None
    for i, x_i in enumerate(torch.chunk(x, scale, dim=1)):
        if i == 0:
            y_i = x_i
        elif i == 1:
            y_i = tdnn_block(x_i, W, f"{prefix}.blocks.{i - 1}", dilation, bn, taps)
        else:
            y_i = tdnn_block(x_i + y_i, W, f"{prefix}.blocks.{i - 1}", dilation, bn, taps)
        ys.append(y_i)
    return torch.cat(ys, dim=1)


def length_to_mask(length, max_len):
    """utils.py:8-19"""
    return (torch.arange(max_len, dtype=length.dtype)[None, :] < length[:, None]).to(length.dtype)


def se_block(x, W, prefix, lengths=None):
    """ecapa_tdnn.py:69-82"""
    L = x.shape[-1]
    if lengths is not None:
        mask = length_to_mask(lengths * L, L).unsqueeze(1).to(x.dtype)
        total = mask.sum(dim=2, keepdim=True)
        s = (x * mask).sum(dim=2, keepdim=True) / total
    else:
        s = x.mean(dim=2, keepdim=True)
    s = F.relu(F.conv1d(s, W[prefix + ".conv1.conv.weight"], W[prefix + ".conv1.conv.bias"]))
    s = torch.sigmoid(F.conv1d(s, W[prefix + ".conv2.conv.weight"], W[prefix + ".conv2.conv.bias"]))
    return s * x


def se_res2net_block(x, W, prefix, scale=8, dilation=1, lengths=None, bn=batchnorm_eval, taps=None):
    """ecapa_tdnn.py:132-142 (in_channels == out_channels for the shipped config -> no shortcut conv)"""
    residual = x
    if prefix + ".shortcut.conv.weight" in W:
        residual = F.conv1d(x, W[prefix + ".shortcut.conv.weight"], W[prefix + ".shortcut.conv.bias"])
    x = tdnn_block(x, W, prefix + ".tdnn1", bn=bn, taps=taps)
    x = res2net_block(x, W, prefix + ".res2net_block", scale, dilation, bn, taps)
    if taps is not None:
        taps[prefix + ".res2net_block"] = x
    x = tdnn_block(x, W, prefix + ".tdnn2", bn=bn, taps=taps)
    x = se_block(x, W, prefix + ".se_block", lengths)
    return x + residual


def attentive_stats_pool(x, W, prefix="asp", lengths=None, global_context=True, bn=batchnorm_eval):
    """pooling.py:86-125"""
    N, C, L = x.shape

    def stats(x, m, eps=ASP_EPS):
        mean = (m * x).sum(2)
        std = torch.sqrt((m * (x - mean.unsqueeze(2)).pow(2)).sum(2).clamp(min=eps))
        return mean, std

    if lengths is None:
        lengths = torch.ones(N, dtype=x.dtype)
    mask = length_to_mask(lengths.to(x.dtype) * L, L).unsqueeze(1)
    if global_context:
        total = mask.sum(dim=2, keepdim=True)
        mean, std = stats(x, mask / total)
        attn = torch.cat([x, mean.unsqueeze(2).expand(-1, -1, L), std.unsqueeze(2).expand(-1, -1, L)], dim=1)
    else:
        attn = x
    attn = torch.tanh(tdnn_block(attn, W, prefix + ".tdnn", bn=bn))
    attn = F.conv1d(attn, W[prefix + ".conv.conv.weight"], W[prefix + ".conv.conv.bias"])
    attn = attn.masked_fill(mask.expand(-1, C, -1) == 0, float("-inf"))
    attn = F.softmax(attn, dim=2)
    mean, std = stats(x, attn)
    return torch.cat((mean, std), dim=1)


def ecapa_forward(feats, W: Dict[str, torch.Tensor], lengths: Optional[torch.Tensor] = None,
                  dilations=(1, 2, 3, 4, 1), res2net_scale=8, taps: Optional[dict] = None, bn=batchnorm_eval, layer_taps=False,
                  pooling_type="ASP", global_context=True):
    """ecapa_tdnn.py:245-276.  feats [B,T,F] -> embedding [B,embd_dim].

    ``taps`` (optional dict) is filled with intermediate activations for per-layer
    parity tests.
    """
    x = feats.transpose(1, 2)
    xl = []
    x = tdnn_block(x, W, "blocks.0", dilations[0], bn)
    xl.append(x)
    if taps is not None:
        taps["blocks.0"] = x
    nblk = len(dilations) - 2
    for i in range(1, nblk + 1):
        x = se_res2net_block(x, W, f"blocks.{i}", res2net_scale, dilations[i], lengths, bn, taps if layer_taps else None)
        xl.append(x)
        if taps is not None:
            taps[f"blocks.{i}"] = x
    x = torch.cat(xl[1:], dim=1)
    x = tdnn_block(x, W, "mfa", dilations[-1], bn)
    if taps is not None:
        taps["mfa"] = x
    if pooling_type == "ASP":
        x = attentive_stats_pool(x, W, "asp", lengths, global_context=global_context, bn=bn)
        if taps is not None:
            taps["asp"] = x
        x = bn(x, W, "asp_bn.norm")
    else:
        # ecapa_tdnn.py:221-241 with pooling.py:8-66; these heads use paddle.nn.BatchNorm1D directly (keys asp_bn.weight, ...)
        if pooling_type == "SAP":
            a = torch.tanh(F.conv1d(x, W["asp.linear1.weight"], W["asp.linear1.bias"]))
            a = F.softmax(F.conv1d(a, W["asp.linear2.weight"], W["asp.linear2.bias"]), dim=2)
            x = (a * x).sum(dim=2)
        elif pooling_type == "TAP":
            x = x.mean(dim=2)
        elif pooling_type == "TSP":
            x = torch.cat((x.mean(dim=2), x.var(dim=2, unbiased=True)), dim=1)
        else:
            raise Exception(f"unknown pooling_type {pooling_type}")
        if taps is not None:
            taps["asp"] = x
        x = bn(x, W, "asp_bn")
    x = F.conv1d(x.unsqueeze(2), W["fc.conv.weight"], W["fc.conv.bias"]).squeeze(-1)
    return x


# ----------------------------------------------------------------------------------------------
# Seeded weights (the reference ships none: README.md:70-102) -- SURVEY.md §8(d) config 2
# ----------------------------------------------------------------------------------------------
def ecapa_param_shapes(input_size=80, embd_dim=192, channels=(512, 512, 512, 512, 1536),
                       kernel_sizes=(5, 3, 3, 3, 1), attention_channels=128, res2net_scale=8,
                       se_channels=128, pooling_type="ASP", global_context=True):
    """Name -> shape for every tensor in the reference EcapaTdnn state_dict (ecapa_tdnn.py:145-243)."""
    S = {}

    def conv(prefix, cin, cout, k):
        S[prefix + ".weight"] = (cout, cin, k)
        S[prefix + ".bias"] = (cout,)

    def bn(prefix, c):
        for n in ("weight", "bias", "_mean", "_variance"):
            S[f"{prefix}.{n}"] = (c,)

    def tdnn(prefix, cin, cout, k):
        conv(prefix + ".conv.conv", cin, cout, k)
        bn(prefix + ".norm.norm", cout)

    tdnn("blocks.0", input_size, channels[0], kernel_sizes[0])
    for i in range(1, len(channels) - 1):
        p = f"blocks.{i}"
        cin, cout = channels[i - 1], channels[i]
        tdnn(p + ".tdnn1", cin, cout, 1)
        for j in range(res2net_scale - 1):
            tdnn(f"{p}.res2net_block.blocks.{j}", cout // res2net_scale, cout // res2net_scale, 3)
        tdnn(p + ".tdnn2", cout, cout, 1)
        conv(p + ".se_block.conv1.conv", cout, se_channels, 1)
        conv(p + ".se_block.conv2.conv", se_channels, cout, 1)
        if cin != cout:
            conv(p + ".shortcut.conv", cin, cout, 1)
    C = channels[-1]
    tdnn("mfa", C, C, kernel_sizes[-1])
    if pooling_type == "ASP":
        tdnn("asp.tdnn", 3 * C if global_context else C, attention_channels, 1)  # pooling.py:75-78
        conv("asp.conv.conv", attention_channels, C, 1)
        bn("asp_bn.norm", 2 * C)
        conv("fc.conv", 2 * C, embd_dim, 1)
        return S
    pooled = 2 * C if pooling_type == "TSP" else C
    if pooling_type == "SAP":
        conv("asp.linear1", C, 128, 1)
        conv("asp.linear2", 128, C, 1)
    bn("asp_bn", pooled)
    conv("fc.conv", pooled, embd_dim, 1)
    return S


def make_ecapa_weights(seed=1000, dtype=torch.float32, **shape_args) -> Dict[str, torch.Tensor]:
    """Seeded random weights: conv ~ U(+-1/sqrt(fan_in)) (Paddle's Conv1D default is also a
    fan-in scaled init), BN gamma ~ U(0.5,1.5), beta ~ N(0,0.1), running mean ~ N(0,0.1),
    running var ~ U(0.5,1.5) -- perturbed BN statistics so BN bugs are visible
    (SURVEY.md §8(d) config 2)."""
    g = torch.Generator().manual_seed(seed)
    W = {}
    for name, shape in ecapa_param_shapes(**shape_args).items():
        if name.endswith("_variance") or name.endswith("norm.weight") or name == "asp_bn.weight":
            t = torch.rand(shape, generator=g, dtype=torch.float64) + 0.5
        elif name.endswith("_mean") or name.endswith("norm.bias") or name == "asp_bn.bias":
            t = torch.randn(shape, generator=g, dtype=torch.float64) * 0.1
        elif name.endswith(".weight"):
            fan_in = shape[1] * shape[2]
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound
        else:  # conv bias
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * 0.1
        W[name] = t.to(dtype)
    return W


def count_params(W) -> int:
    return sum(v.numel() for k, v in W.items() if not (k.endswith("_mean") or k.endswith("_variance")))
