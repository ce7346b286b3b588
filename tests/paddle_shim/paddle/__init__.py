"""Stand-in for the `paddle` package over torch -- TEST INFRASTRUCTURE, see ../README.md.

Only what the reference's model / loss / scheduler files call is implemented; every op's assumed Paddle
semantics is listed in ../README.md.  Tensors are a torch.Tensor subclass carrying Paddle's method
signatures (``transpose(perm)``, ``axis=`` keywords, ``tile``, ``clip``, ``astype`` ...).
"""
import builtins

import torch

float32 = torch.float32
float64 = torch.float64
int64 = torch.int64
int32 = torch.int32
bool = torch.bool  # noqa: A001  (paddle.bool)

_DTYPES = {"float32": torch.float32, "float64": torch.float64, "int": torch.int32, "int32": torch.int32,
           "int64": torch.int64, "bool": torch.bool, "float": torch.float32}


def _dt(d):
    """Paddle dtype -> torch dtype.  'float32' means "the working float precision": the fixture generator runs the reference graph in
    fp64 (torch default dtype), and a float32 temporary inside it (armloss.py:24) must not round the fixture to single precision."""
    if d is None:
        return None
    if isinstance(d, str):
        d = _DTYPES[d]
    if d is torch.float32:
        return torch.get_default_dtype()
    return d


def _axis(kw):
    """paddle's axis= / keepdim= keywords -> torch's dim= / keepdim="""
    out = {}
    if "axis" in kw:
        a = kw.pop("axis")
        if a is not None:
            out["dim"] = tuple(a) if isinstance(a, (list, tuple)) else a
    if "keepdim" in kw:
        out["keepdim"] = kw.pop("keepdim")
    assert not kw, f"paddle_shim: unsupported keywords {kw}"
    return out


class Tensor(torch.Tensor):
    """torch.Tensor with Paddle's method signatures; torch's default __torch_function__ keeps the subclass."""

    # ---- shape ops
    def transpose(self, perm, *rest):
        if rest:  # torch-style call from torch internals
            return super().transpose(perm, *rest)
        return self.permute(*perm)

    def unsqueeze(self, axis=None, dim=None):
        return super().unsqueeze(axis if axis is not None else dim)

    def unsqueeze_(self, axis):
        return super().unsqueeze_(axis)

    def squeeze(self, axis=None, dim=None):
        a = axis if axis is not None else dim
        return super().squeeze() if a is None else super().squeeze(a)

    def reshape(self, shape, *rest):
        return super().reshape(*([shape] if not rest else [shape, *rest]))

    def flatten(self, start_axis=0, stop_axis=-1):
        return super().flatten(start_axis, stop_axis)

    def tile(self, repeat_times):
        return self.repeat(*repeat_times)

    def expand(self, shape, *rest):
        return super().expand(*(list(shape) if not rest else [shape, *rest]))

    def astype(self, dtype):
        return self.to(_dt(dtype))

    def numpy(self):
        return self.detach().as_subclass(torch.Tensor).cpu().numpy()

    # ---- reductions (axis as first positional or keyword)
    def _reduce(self, fn, args, kw):
        if args:
            kw = dict(kw, axis=args[0])
            if len(args) > 1:
                kw["keepdim"] = args[1]
        return fn(**_axis(kw))

    def sum(self, *args, **kw):
        return self._reduce(super().sum, args, kw)

    def mean(self, *args, **kw):
        return self._reduce(super().mean, args, kw)

    def max(self, *args, **kw):
        if not args and not kw:
            return super().max()
        r = self._reduce(super().max, args, kw)
        return r.values if hasattr(r, "values") else r

    def min(self, *args, **kw):
        if not args and not kw:
            return super().min()
        r = self._reduce(super().min, args, kw)
        return r.values if hasattr(r, "values") else r

    def std(self, axis=None, unbiased=True, keepdim=False):
        if axis is None:
            return super().std(unbiased=unbiased)
        return super().std(dim=axis, unbiased=unbiased, keepdim=keepdim)

    def var(self, axis=None, unbiased=True, keepdim=False):
        if axis is None:
            return super().var(unbiased=unbiased)
        return super().var(dim=axis, unbiased=unbiased, keepdim=keepdim)

    def clip(self, min=None, max=None):  # noqa: A002
        return self.clamp(min=min, max=max)

    def pow(self, y):
        return super().pow(y)


def _wrap(t):
    return t.as_subclass(Tensor) if isinstance(t, torch.Tensor) and not isinstance(t, Tensor) else t


def to_tensor(data, dtype=None, place=None, stop_gradient=True):
    if isinstance(data, torch.Tensor):
        t = data.detach().clone() if dtype is None else data.to(_dt(dtype))
    else:
        t = torch.as_tensor(data, dtype=_dt(dtype))
        if dtype is None and t.dtype.is_floating_point:
            t = t.to(torch.get_default_dtype())
    return _wrap(t)


def ones(shape, dtype=None):
    return _wrap(torch.ones(list(shape), dtype=_dt(dtype) or torch.get_default_dtype()))


def zeros(shape, dtype=None):
    return _wrap(torch.zeros(list(shape), dtype=_dt(dtype) or torch.get_default_dtype()))


def ones_like(x, dtype=None):
    return torch.ones_like(x, dtype=_dt(dtype))


def zeros_like(x, dtype=None):
    return torch.zeros_like(x, dtype=_dt(dtype))


def arange(start=0, end=None, step=1, dtype=None):
    if end is None:
        start, end = 0, start
    d = _dt(dtype)
    return _wrap(torch.arange(start, end, step, dtype=d))


def concat(x, axis=0):
    return torch.cat(list(x), dim=axis)


def split(x, num_or_sections, axis=0):
    """paddle.split: an int is the NUMBER of equal sections (torch.split's int is the section size)."""
    if isinstance(num_or_sections, int):
        assert x.shape[axis] % num_or_sections == 0
        return list(torch.split(x, x.shape[axis] // num_or_sections, dim=axis))
    return list(torch.split(x, list(num_or_sections), dim=axis))


def chunk(x, chunks, axis=0):
    assert x.shape[axis] % chunks == 0
    return list(torch.chunk(x, chunks, dim=axis))


def where(condition, x=None, y=None):
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x, dtype=y.dtype)
    if not isinstance(y, torch.Tensor):
        y = torch.as_tensor(y, dtype=x.dtype)
    return torch.where(condition, x, y)


def reshape(x, shape):
    return torch.reshape(x, list(shape))


def expand(x, shape):
    return x.expand(list(shape))


def squeeze(x, axis=None):
    return torch.squeeze(x) if axis is None else torch.squeeze(x, axis)


def unsqueeze(x, axis):
    return torch.unsqueeze(x, axis)


def mean(x, axis=None, keepdim=False):
    return torch.mean(x) if axis is None else torch.mean(x, dim=axis, keepdim=keepdim)


def sum(x, axis=None, dtype=None, keepdim=False):  # noqa: A001
    return torch.sum(x) if axis is None else torch.sum(x, dim=axis, keepdim=keepdim)


def var(x, axis=None, unbiased=True, keepdim=False):
    return torch.var(x, unbiased=unbiased) if axis is None else torch.var(x, dim=axis, unbiased=unbiased, keepdim=keepdim)


def max(x, axis=None, keepdim=False):  # noqa: A001
    return torch.max(x) if axis is None else torch.amax(x, dim=axis, keepdim=keepdim)


def min(x, axis=None, keepdim=False):  # noqa: A001
    return torch.min(x) if axis is None else torch.amin(x, dim=axis, keepdim=keepdim)


def norm(x, p=2, axis=None, keepdim=False):
    return torch.linalg.vector_norm(x, ord=p, dim=axis, keepdim=keepdim)


def masked_select(x, mask):
    return torch.masked_select(x, mask)


sqrt, exp, log, tanh, pow, multiply, divide, matmul = (torch.sqrt, torch.exp, torch.log, torch.tanh, torch.pow,  # noqa: A001
                                                       torch.multiply, torch.divide, torch.matmul)


class ParamAttr:
    def __init__(self, name=None, initializer=None, learning_rate=1.0, regularizer=None, trainable=True):
        self.initializer = initializer


def create_parameter(shape, dtype="float32", attr=None, is_bias=False, default_initializer=None):
    """Values are placeholders: fixtures always load seeded weights by name afterwards."""
    p = torch.nn.Parameter(torch.zeros(list(shape), dtype=torch.get_default_dtype()))
    init = getattr(attr, "initializer", None) or default_initializer
    if init is not None:
        init(p)
    return p


def seed(s):
    torch.manual_seed(s)


class _NoGrad:
    def __enter__(self):
        self._g = torch.no_grad()
        return self._g.__enter__()

    def __exit__(self, *a):
        return self._g.__exit__(*a)


def no_grad():
    return _NoGrad()


from . import nn  # noqa: E402,F401
from . import optimizer  # noqa: E402,F401

del builtins
