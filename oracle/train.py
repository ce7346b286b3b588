"""Oracle: one ECAPA-TDNN training step on the CPU -- train-mode forward, AAM-softmax loss, autograd backward, Adam.

TEST INFRASTRUCTURE -- see oracle/__init__.py.

Follows
  * ppvector/trainer.py:206-229            forward -> loss -> backward -> optimizer.step -> clear_grad
  * ppvector/models/utils.py:96-119        BatchNorm1d -> paddle.nn.BatchNorm1D in train mode: normalise with the batch mean and
                                           the BIASED batch variance over (batch, time); running = 0.9 * running + 0.1 * batch
                                           (Paddle keeps the biased variance in the running estimate -- recalled)
  * ppvector/models/fc.py:41-53, ppvector/loss/aamloss.py:28-53   (oracle/head.py)
  * ppvector/optimizer/__init__.py:12-18   paddle.optimizer.Adam(weight_decay=wd): coupled L2, bias-corrected -- identical to
                                           torch.optim.Adam(weight_decay=wd)
  * ppvector/optimizer/scheduler.py:43-102 MarginScheduler
The backward pass is torch autograd over the functional forward of oracle/ecapa.py.
"""
import math
from typing import Dict

import torch

from oracle import ecapa, head

BN_MOMENTUM = 0.9


def make_bn_train(new_stats: Dict[str, torch.Tensor]):
    """BatchNorm in train mode; the updated running statistics are written to ``new_stats``."""

    def bn(x, W, prefix):
        dims = [0] + list(range(2, x.dim()))
        mean = x.mean(dim=dims)
        var = x.var(dim=dims, unbiased=False)
        shape = [1, -1] + [1] * (x.dim() - 2)
        y = (x - mean.view(shape)) / torch.sqrt(var.view(shape) + ecapa.BN_EPS) * W[prefix + ".weight"].view(shape) + W[prefix + ".bias"].view(shape)
        new_stats[prefix + "._mean"] = (BN_MOMENTUM * W[prefix + "._mean"] + (1 - BN_MOMENTUM) * mean).detach()
        new_stats[prefix + "._variance"] = (BN_MOMENTUM * W[prefix + "._variance"] + (1 - BN_MOMENTUM) * var).detach()
        return y

    return bn


def is_stat(name):
    return name.endswith("._mean") or name.endswith("._variance")


def train_step_grads(feats, labels, W: Dict[str, torch.Tensor], Wcls, margin=0.2, scale=32.0, easy_margin=False, label_smoothing=0.0, taps=None,
                     layer_taps=False):
    """-> (loss, grads dict incl. 'classifier.weight', new running stats, cosine logits)"""
    P = {k: v.clone().requires_grad_(not is_stat(k)) for k, v in W.items()}
    Wc = Wcls.clone().requires_grad_(True)
    new_stats = {}
    emb = ecapa.ecapa_forward(feats, P, taps=taps, bn=make_bn_train(new_stats), layer_taps=layer_taps)
    logits = head.cosine_logits(emb, Wc)
    loss = head.aam_loss(logits, labels, margin=margin, scale=scale, easy_margin=easy_margin, label_smoothing=label_smoothing)
    if taps is not None:
        for v in taps.values():  # gradients of the intermediate activations stay readable as taps[name].grad
            v.retain_grad()
    loss.backward()
    grads = {k: v.grad for k, v in P.items() if not is_stat(k)}
    grads["classifier.weight"] = Wc.grad
    if taps is not None:
        taps["emb"] = emb.detach()
    return loss.detach(), grads, new_stats, logits.detach()


def margin_at(step, increase_start_step, fix_step, initial_margin=0.0, final_margin=0.3, increase_type="exp"):
    """scheduler.py:82-102 (MarginScheduler.iter_margin)"""
    if step < increase_start_step:
        return initial_margin
    if step >= fix_step:
        return final_margin
    cur = step - increase_start_step
    inc = fix_step - increase_start_step
    if increase_type == "exp":
        ratio = 1.0 - math.exp((cur / inc) * math.log(1e-3 / (1.0 + 1e-6))) * 1.0
    else:
        ratio = 1.0 * cur / inc
    return initial_margin + (final_margin - initial_margin) * ratio


def train_loop(feat_batches, label_batches, W, Wcls, lr=1e-3, weight_decay=1e-6, margins=None, scale=32.0, label_smoothing=0.0):
    """Runs len(feat_batches) steps of Adam; returns the loss curve and the final state."""
    W = {k: v.clone() for k, v in W.items()}
    names = [k for k in W if not is_stat(k)]
    params = [W[k].requires_grad_(True) for k in names]
    Wc = Wcls.clone().requires_grad_(True)
    opt = torch.optim.Adam(params + [Wc], lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=weight_decay)
    losses = []
    for i, (f, y) in enumerate(zip(feat_batches, label_batches)):
        new_stats = {}
        emb = ecapa.ecapa_forward(f, W, bn=make_bn_train(new_stats))
        logits = head.cosine_logits(emb, Wc)
        loss = head.aam_loss(logits, y, margin=0.0 if margins is None else margins[i], scale=scale, label_smoothing=label_smoothing)
        opt.zero_grad()
        loss.backward()
        opt.step()
        for k, v in new_stats.items():
            W[k] = v
        losses.append(loss.item())
    return losses, {k: v.detach() for k, v in W.items()}, Wc.detach()
