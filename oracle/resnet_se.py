"""Oracle: ResNetSE forward on the CPU (torch functional ops, fp32 or fp64).

TEST INFRASTRUCTURE -- see oracle/__init__.py.

Weights: flat dict keyed like the reference's Paddle state_dict for ``ResNetSE`` (Conv2D ``weight`` [Cout,Cin,kh,kw] /
``bias``; BatchNorm2D ``weight``/``bias``/``_mean``/``_variance``; Linear ``weight`` [in,out] (Paddle layout) / ``bias``).

Follows
  * ppvector/models/resnet_se.py:24-45    SEBottleneck.forward (conv-BN-ReLU x2, conv-BN, SE, + residual, ReLU)
  * ppvector/models/resnet_se.py:59-63    SELayer.forward (global average pool, Linear-ReLU-Linear-Sigmoid, scale)
  * ppvector/models/resnet_se.py:107-119  _make_layer (1x1 strided conv + BN downsample on the first block)
  * ppvector/models/resnet_se.py:121-139  ResNetSE.forward
  * ppvector/models/pooling.py:86-125     AttentiveStatisticsPooling (shared with oracle/ecapa.py)
"""
import math
from typing import Dict

import torch
import torch.nn.functional as F

from oracle.ecapa import attentive_stats_pool, batchnorm_eval

EXPANSION = 2


def se_layer(x, W, prefix):
    """resnet_se.py:59-63"""
    y = x.mean(dim=(2, 3))
    y = F.relu(y @ W[prefix + ".fc.0.weight"] + W[prefix + ".fc.0.bias"])
    y = torch.sigmoid(y @ W[prefix + ".fc.2.weight"] + W[prefix + ".fc.2.bias"])
    return x * y[:, :, None, None]


def se_bottleneck(x, W, prefix, stride):
    """resnet_se.py:24-45"""
    residual = x
    out = F.conv2d(x, W[prefix + ".conv1.weight"], W[prefix + ".conv1.bias"])
    out = F.relu(batchnorm_eval(out, W, prefix + ".bn1"))
    out = F.conv2d(out, W[prefix + ".conv2.weight"], W[prefix + ".conv2.bias"], stride=stride, padding=1)
    out = F.relu(batchnorm_eval(out, W, prefix + ".bn2"))
    out = F.conv2d(out, W[prefix + ".conv3.weight"], W[prefix + ".conv3.bias"])
    out = batchnorm_eval(out, W, prefix + ".bn3")
    out = se_layer(out, W, prefix + ".se")
    if prefix + ".downsample.0.weight" in W:
        residual = F.conv2d(x, W[prefix + ".downsample.0.weight"], W[prefix + ".downsample.0.bias"], stride=stride)
        residual = batchnorm_eval(residual, W, prefix + ".downsample.1")
    return F.relu(out + residual)


def resnet_se_forward(feats, W: Dict[str, torch.Tensor], layers=(3, 4, 6, 3), taps=None):
    """resnet_se.py:121-139.  feats [B,T,F] -> [B,embd_dim]."""
    x = feats.transpose(1, 2).unsqueeze(1)  # [B,1,F,T]
    x = F.conv2d(x, W["conv1.weight"], W["conv1.bias"], padding=1)
    x = F.relu(batchnorm_eval(x, W, "bn1"))
    if taps is not None:
        taps["conv1"] = x
    for li, nblocks in enumerate(layers, start=1):
        for bi in range(nblocks):
            stride = 2 if (li > 1 and bi == 0) else 1
            x = se_bottleneck(x, W, f"layer{li}.{bi}", stride)
        if taps is not None:
            taps[f"layer{li}"] = x
    x = x.reshape(x.shape[0], -1, x.shape[-1])  # [B, C*F', T']
    if taps is not None:
        taps["flat"] = x
    x = attentive_stats_pool(x, W, "pooling")
    if taps is not None:
        taps["asp"] = x
    x = batchnorm_eval(x, W, "bn2.norm")
    x = x @ W["linear.weight"] + W["linear.bias"]
    x = batchnorm_eval(x, W, "bn3.norm")
    return x


def resnet_se_param_shapes(input_size=80, layers=(3, 4, 6, 3), num_filters=(32, 64, 128, 256), embd_dim=192,
                           attention_channels=128, reduction=8):
    S = {}

    def conv(p, cin, cout, k):
        S[p + ".weight"] = (cout, cin, k, k)
        S[p + ".bias"] = (cout,)

    def bn(p, c):
        for n in ("weight", "bias", "_mean", "_variance"):
            S[f"{p}.{n}"] = (c,)

    conv("conv1", 1, num_filters[0], 3)
    bn("bn1", num_filters[0])
    inplanes = num_filters[0]
    for li, (nblocks, planes) in enumerate(zip(layers, num_filters), start=1):
        for bi in range(nblocks):
            p = f"layer{li}.{bi}"
            stride = 2 if (li > 1 and bi == 0) else 1
            conv(p + ".conv1", inplanes, planes, 1)
            bn(p + ".bn1", planes)
            conv(p + ".conv2", planes, planes, 3)
            bn(p + ".bn2", planes)
            conv(p + ".conv3", planes, planes * EXPANSION, 1)
            bn(p + ".bn3", planes * EXPANSION)
            c = planes * EXPANSION
            S[p + ".se.fc.0.weight"] = (c, c // reduction)
            S[p + ".se.fc.0.bias"] = (c // reduction,)
            S[p + ".se.fc.2.weight"] = (c // reduction, c)
            S[p + ".se.fc.2.bias"] = (c,)
            if bi == 0 and (stride != 1 or inplanes != planes * EXPANSION):
                conv(p + ".downsample.0", inplanes, planes * EXPANSION, 1)
                bn(p + ".downsample.1", planes * EXPANSION)
            inplanes = planes * EXPANSION
    cat = num_filters[3] * EXPANSION * (input_size // 8)
    S["pooling.tdnn.conv.conv.weight"] = (attention_channels, 3 * cat, 1)
    S["pooling.tdnn.conv.conv.bias"] = (attention_channels,)
    bn("pooling.tdnn.norm.norm", attention_channels)
    S["pooling.conv.conv.weight"] = (cat, attention_channels, 1)
    S["pooling.conv.conv.bias"] = (cat,)
    bn("bn2.norm", 2 * cat)
    S["linear.weight"] = (2 * cat, embd_dim)
    S["linear.bias"] = (embd_dim,)
    bn("bn3.norm", embd_dim)
    return S


def make_resnet_se_weights(seed=1000, dtype=torch.float32, **shape_args) -> Dict[str, torch.Tensor]:
    """Seeded random weights with perturbed BatchNorm statistics (same recipe as oracle/ecapa.py)."""
    g = torch.Generator().manual_seed(seed)
    W = {}
    for name, shape in resnet_se_param_shapes(**shape_args).items():
        is_bn_scale = name.endswith("_variance") or (name.endswith(".weight") and len(shape) == 1)
        is_bn_shift = name.endswith("_mean") or (name.endswith(".bias") and (".bn" in name or name.startswith("bn") or ".norm." in name
                                                                              or ".downsample.1." in name))
        if is_bn_scale:
            t = torch.rand(shape, generator=g, dtype=torch.float64) + 0.5
        elif is_bn_shift:
            t = torch.randn(shape, generator=g, dtype=torch.float64) * 0.1
        elif name.endswith(".weight"):
            if len(shape) == 2:      # Linear [in, out]
                fan_in = shape[0]
            else:
                fan_in = 1
                for d in shape[1:]:
                    fan_in *= d
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) / math.sqrt(fan_in)
        else:
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * 0.1
        W[name] = t.to(dtype)
    return W


def count_params(W) -> int:
    return sum(v.numel() for k, v in W.items() if not (k.endswith("_mean") or k.endswith("_variance")))
