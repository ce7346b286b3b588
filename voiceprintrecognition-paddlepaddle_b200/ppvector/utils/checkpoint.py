"""Checkpoint I/O for the drop-in entry points.

The reference saves ``model.pdparams`` with ``paddle.save`` (ppvector/utils/checkpoint.py:104-159): a pickle of
{name: numpy array}.  Reading real ``.pdparams`` files needs no Paddle -- it is a plain
pickle of numpy arrays for the 2.x formats -- and is supported here on a best-effort basis; ``.npz`` and torch
``.pt/.pth`` state dicts are the native formats of this build."""
import os
import pickle

import numpy as np


def load_state_dict_file(path):
    if os.path.isdir(path):
        for name in ('model.pdparams', 'model.npz', 'model.pth', 'model.pt'):
            if os.path.exists(os.path.join(path, name)):
                path = os.path.join(path, name)
                break
        else:
            raise FileNotFoundError(f'{path} 模型不存在！')
    if path.endswith('.npz'):
        with np.load(path) as z:
            return {k: z[k] for k in z.files}
    if path.endswith(('.pt', '.pth')):
        import torch
        sd = torch.load(path, map_location='cpu')
        return {k: v.numpy() for k, v in sd.items()}
    with open(path, 'rb') as f:  # .pdparams: pickle of {name: ndarray} (paddle.save protocol 2-4)
        obj = pickle.load(f, encoding='latin1')
    out = {}
    for k, v in obj.items():
        if isinstance(v, np.ndarray):
            out[k] = v
        elif isinstance(v, (tuple, list)) and len(v) == 2 and isinstance(v[1], np.ndarray):
            out[k] = v[1]  # (name, ndarray) form used by some paddle versions
    return out


def save_state_dict_npz(state_dict, path):
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    np.savez(path, **{k: (v.detach().cpu().numpy() if hasattr(v, 'detach') else np.asarray(v)) for k, v in state_dict.items()})
