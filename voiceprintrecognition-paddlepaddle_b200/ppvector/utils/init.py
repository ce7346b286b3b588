"""Parameter initialisation for the backbone mirrors.

``paddle_default_init_`` reproduces the defaults a freshly constructed reference model gets from Paddle
(`paddle.nn.Conv1D/Conv2D`: Normal(0, sqrt(2 / fan_in)), bias 0; `paddle.nn.Linear` / `create_parameter`: Xavier-uniform,
bias 0; BatchNorm: weight 1, bias 0, _mean 0, _variance 1), so that from-scratch training starts where the reference does.
``seeded_state_dict`` is the synthetic-weights recipe of SURVEY.md §8(d) config 2 (BatchNorm statistics perturbed so that
BatchNorm bugs are visible); bench.py uses it for both arms.
"""
import math

import torch


def _kind(name, t):
    if name.endswith("._mean"):
        return "bn_mean"
    if name.endswith("._variance"):
        return "bn_var"
    if t.dim() == 1:
        return "vec"
    return "mat"


def paddle_default_init_(module, generator=None):
    bn = {n.rsplit(".", 1)[0] for n, _ in module.named_buffers() if n.endswith("._mean")}
    with torch.no_grad():
        for name, p in module.named_parameters():
            owner = name.rsplit(".", 1)[0]
            if owner in bn:
                p.fill_(1.0 if name.endswith(".weight") else 0.0)
            elif p.dim() >= 3:      # conv weight [Cout, Cin, *k]
                fan_in = p[0].numel()
                p.normal_(0.0, math.sqrt(2.0 / fan_in), generator=generator)
            elif p.dim() == 2:      # Linear / classifier weight [in, out]: Xavier uniform
                bound = math.sqrt(6.0 / (p.shape[0] + p.shape[1]))
                p.uniform_(-bound, bound, generator=generator)
            else:
                p.zero_()
    return module


def seeded_state_dict(module, seed=1000):
    """name -> fp32 tensor for every entry of ``module.state_dict()``: conv / linear weights fan-in scaled, biases small,
    BatchNorm gamma ~ U(0.5,1.5), beta / running mean ~ N(0,0.1²), running variance ~ U(0.5,1.5)."""
    g = torch.Generator().manual_seed(seed)
    sd = module.state_dict()
    bn = {n.rsplit(".", 1)[0] for n in sd if n.endswith("._mean")}
    out = {}
    for name, t in sd.items():
        owner = name.rsplit(".", 1)[0]
        shape = tuple(t.shape)
        if owner in bn:
            if name.endswith("._variance") or name.endswith(".weight"):
                v = torch.rand(shape, generator=g) + 0.5
            else:
                v = torch.randn(shape, generator=g) * 0.1
        elif t.dim() >= 3:
            bound = 1.0 / math.sqrt(t[0].numel())
            v = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif t.dim() == 2:
            bound = 1.0 / math.sqrt(shape[0])
            v = (torch.rand(shape, generator=g) * 2 - 1) * bound
        else:
            v = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
        out[name] = v.to(torch.float32)
    return out
