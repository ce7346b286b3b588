"""paddle.nn.functional stand-in (TEST INFRASTRUCTURE, see ../../README.md)."""
import torch
import torch.nn.functional as TF


def relu(x, name=None):
    return TF.relu(x)


def softmax(x, axis=-1, dtype=None, name=None):
    return TF.softmax(x, dim=axis)


def one_hot(x, num_classes, name=None):
    return TF.one_hot(x.long(), num_classes).to(torch.get_default_dtype())


def normalize(x, p=2, axis=1, epsilon=1e-12, name=None):
    return x / torch.linalg.vector_norm(x, ord=p, dim=axis, keepdim=True).clamp(min=epsilon)


def linear(x, weight, bias=None, name=None):
    """paddle: weight is [in, out]"""
    y = x @ weight
    return y if bias is None else y + bias


def pad(x, pad, mode="constant", value=0.0, data_format="NCHW", name=None):  # noqa: A002
    """Only the form the reference uses: a [left, right] pad of the last axis of an NCL tensor."""
    assert data_format == "NCL" and len(pad) == 2 and x.dim() == 3, (data_format, pad, tuple(x.shape))
    if pad[0] == 0 and pad[1] == 0:
        return x
    if mode == "constant":
        return TF.pad(x, tuple(pad), mode="constant", value=value)
    return TF.pad(x, tuple(pad), mode=mode)


def avg_pool1d(x, kernel_size, stride=None, padding=0, exclusive=True, ceil_mode=False, name=None):
    return TF.avg_pool1d(x, kernel_size, stride, padding, ceil_mode=ceil_mode, count_include_pad=not exclusive)


def max_pool1d(x, kernel_size, stride=None, padding=0, return_mask=False, ceil_mode=False, name=None):
    return TF.max_pool1d(x, kernel_size, stride, padding, ceil_mode=ceil_mode)
