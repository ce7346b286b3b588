"""paddle.nn stand-in (TEST INFRASTRUCTURE, see ../../README.md)."""
import math

import torch
import torch.nn as tnn
import torch.nn.functional as TF

from . import functional  # noqa: F401
from . import initializer  # noqa: F401


def _as_p(x):
    from .. import Tensor
    return x.as_subclass(Tensor) if isinstance(x, torch.Tensor) and not isinstance(x, Tensor) else x


class Layer(tnn.Module):
    """paddle.nn.Layer: sublayer API names on top of torch.nn.Module; inputs are promoted to the shim Tensor."""

    def __call__(self, *args, **kw):
        args = tuple(_as_p(a) for a in args)
        return super().__call__(*args, **kw)

    def add_sublayer(self, name, sublayer):
        self.add_module(str(name), sublayer)
        return sublayer

    def sublayers(self, include_self=False):
        mods = list(self.modules())
        return mods if include_self else mods[1:]

    def set_state_dict(self, sd):
        own = self.state_dict()
        missing = [k for k in own if k not in sd]
        extra = [k for k in sd if k not in own]
        assert not missing and not extra, f"state_dict mismatch: missing {missing[:5]} extra {extra[:5]}"
        with torch.no_grad():
            for k, v in own.items():
                src = torch.as_tensor(sd[k])
                assert tuple(src.shape) == tuple(v.shape), (k, tuple(src.shape), tuple(v.shape))
                v.copy_(src.to(v.dtype))

    def create_parameter(self, shape, attr=None, dtype=None, is_bias=False, default_initializer=None):
        from .. import create_parameter
        return create_parameter(shape, attr=attr, is_bias=is_bias, default_initializer=default_initializer)


class Sequential(Layer):
    """paddle.nn.Sequential: positional layers or (name, layer) tuples; sublayers named '0','1',... or by name."""

    def __init__(self, *layers):
        super().__init__()
        for i, l in enumerate(layers):
            if isinstance(l, (tuple, list)):
                self.add_sublayer(l[0], l[1])
            else:
                self.add_sublayer(str(i), l)

    def forward(self, x):
        for m in self._modules.values():
            x = m(x)
        return x

    def __iter__(self):
        return iter(self._modules.values())

    def __len__(self):
        return len(self._modules)

    def __getitem__(self, i):
        return list(self._modules.values())[i]


class LayerList(Layer):
    def __init__(self, sublayers=None):
        super().__init__()
        for l in (sublayers or []):
            self.append(l)

    def append(self, l):
        self.add_sublayer(str(len(self._modules)), l)
        return self

    def __iter__(self):
        return iter(self._modules.values())

    def __len__(self):
        return len(self._modules)

    def __getitem__(self, i):
        return list(self._modules.values())[i]


def _tup(v, n):
    return tuple(v) if isinstance(v, (tuple, list)) else (v,) * n


class _ConvNd(Layer):
    nd = 1

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 padding_mode="zeros", weight_attr=None, bias_attr=None, data_format=None):
        super().__init__()
        assert padding_mode == "zeros"
        n = self.nd
        self._k, self._s, self._p, self._d, self._g = _tup(kernel_size, n), _tup(stride, n), _tup(padding, n), _tup(dilation, n), groups
        self.weight = tnn.Parameter(torch.empty(out_channels, in_channels // groups, *self._k, dtype=torch.get_default_dtype()))
        fan_in = (in_channels // groups) * math.prod(self._k)
        tnn.init.normal_(self.weight, 0.0, math.sqrt(2.0 / fan_in))  # Paddle's conv default initializer
        if bias_attr is False:
            self.bias = None
        else:
            self.bias = tnn.Parameter(torch.zeros(out_channels, dtype=torch.get_default_dtype()))

    def forward(self, x):
        f = TF.conv1d if self.nd == 1 else TF.conv2d
        return f(x, self.weight, self.bias, stride=self._s, padding=self._p, dilation=self._d, groups=self._g)


class Conv1D(_ConvNd):
    nd = 1


class Conv2D(_ConvNd):
    nd = 2


class _BatchNorm(Layer):
    """paddle BatchNorm: state_dict keys weight, bias, _mean, _variance; momentum 0.9 == torch momentum 0.1;
    the running variance tracks the BIASED batch variance."""

    def __init__(self, num_features, momentum=0.9, epsilon=1e-5, weight_attr=None, bias_attr=None, data_format=None,
                 use_global_stats=None, name=None):
        super().__init__()
        d = torch.get_default_dtype()
        self.weight = tnn.Parameter(torch.ones(num_features, dtype=d))
        self.bias = tnn.Parameter(torch.zeros(num_features, dtype=d))
        self.register_buffer("_mean", torch.zeros(num_features, dtype=d))
        self.register_buffer("_variance", torch.ones(num_features, dtype=d))
        self._momentum, self._epsilon, self._use_global_stats = momentum, epsilon, use_global_stats

    def forward(self, x):
        shape = [1, -1] + [1] * (x.dim() - 2)
        use_global = (not self.training) if self._use_global_stats is None else self._use_global_stats
        if use_global:
            m, v = self._mean, self._variance
        else:
            dims = [i for i in range(x.dim()) if i != 1]
            m = torch.mean(x, dim=dims)
            v = torch.var(x, dim=dims, unbiased=False)
            with torch.no_grad():
                self._mean.mul_(self._momentum).add_((1 - self._momentum) * m.detach())
                self._variance.mul_(self._momentum).add_((1 - self._momentum) * v.detach())
        return (x - m.view(shape)) / torch.sqrt(v.view(shape) + self._epsilon) * self.weight.view(shape) + self.bias.view(shape)


class BatchNorm1D(_BatchNorm):
    pass


class BatchNorm2D(_BatchNorm):
    pass


class Linear(Layer):
    def __init__(self, in_features, out_features, weight_attr=None, bias_attr=None, name=None):
        super().__init__()
        d = torch.get_default_dtype()
        self.weight = tnn.Parameter(torch.empty(in_features, out_features, dtype=d))
        tnn.init.xavier_uniform_(self.weight)
        self.bias = None if bias_attr is False else tnn.Parameter(torch.zeros(out_features, dtype=d))

    def forward(self, x):
        y = x @ self.weight
        return y if self.bias is None else y + self.bias


class ReLU(Layer):
    def __init__(self, name=None):
        super().__init__()

    def forward(self, x):
        return TF.relu(x)


class Sigmoid(Layer):
    def __init__(self, name=None):
        super().__init__()

    def forward(self, x):
        return torch.sigmoid(x)


class Tanh(Layer):
    def __init__(self, name=None):
        super().__init__()

    def forward(self, x):
        return torch.tanh(x)


class Silu(Layer):
    def __init__(self, name=None):
        super().__init__()

    def forward(self, x):
        return x * torch.sigmoid(x)


class Hardtanh(Layer):
    def __init__(self, min=-1.0, max=1.0, name=None):  # noqa: A002
        super().__init__()
        self._min, self._max = min, max

    def forward(self, x):
        return torch.clamp(x, self._min, self._max)


class Identity(Layer):
    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, x):
        return x


class PReLU(Layer):
    def __init__(self, num_parameters=1, init=0.25, weight_attr=None, data_format="NCHW", name=None):
        super().__init__()
        self._weight = tnn.Parameter(torch.full((num_parameters,), init, dtype=torch.get_default_dtype()))

    def forward(self, x):
        return TF.prelu(x, self._weight)


class AdaptiveAvgPool2D(Layer):
    def __init__(self, output_size, data_format="NCHW", name=None):
        super().__init__()
        assert output_size == 1

    def forward(self, x):
        return torch.mean(x, dim=(2, 3), keepdim=True)


class MaxPool2D(Layer):
    def __init__(self, kernel_size, stride=None, padding=0, ceil_mode=False, **kw):
        super().__init__()
        self._a = (kernel_size, stride, padding, ceil_mode)

    def forward(self, x):
        k, s, p, c = self._a
        return TF.max_pool2d(x, k, s, p, ceil_mode=c)


class AvgPool2D(Layer):
    def __init__(self, kernel_size, stride=None, padding=0, ceil_mode=False, exclusive=True, **kw):
        super().__init__()
        self._a = (kernel_size, stride, padding, ceil_mode, exclusive)

    def forward(self, x):
        k, s, p, c, e = self._a
        return TF.avg_pool2d(x, k, s, p, ceil_mode=c, count_include_pad=not e)


class CrossEntropyLoss(Layer):
    def __init__(self, weight=None, ignore_index=-100, reduction="mean", soft_label=False, axis=-1, use_softmax=True,
                 label_smoothing=0.0, name=None):
        super().__init__()
        assert weight is None and reduction in ("mean", "sum") and not soft_label and use_softmax
        self._ls, self._red = label_smoothing, reduction

    def forward(self, input, label):  # noqa: A002
        return TF.cross_entropy(input, label.reshape(-1).long(), label_smoothing=self._ls, reduction=self._red)


class loss:  # noqa: N801  (paddle.nn.loss namespace, used by a loss that is out of scope)
    class MarginRankingLoss(Layer):
        def __init__(self, margin=0.0, reduction="mean", name=None):
            super().__init__()
            self._m = margin

        def forward(self, input, other, label):  # noqa: A002
            return torch.clamp(-label * (input - other) + self._m, min=0).mean()
