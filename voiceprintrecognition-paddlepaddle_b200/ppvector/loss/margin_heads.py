"""AMLoss, ARMLoss, CELoss -- drop-ins for ppvector/loss/amloss.py:5-28, armloss.py:5-35, celoss.py:5-22: the other softmax heads over
the cosine logits of ``SpeakerIdentification``.  They run on the same fused CUDA head as AAMLoss (``ppv_aam_forward`` /
``ppv_aam_backward`` with a head selector, csrc/aam.cu): cosines from ``features`` and the classifier weight, the head's margin rule on
the fly, online-softmax cross-entropy (``reduction='sum'`` / batch size = mean), label smoothing; backward in one pass.

  AMLoss   z = scale * (cos - margin * onehot)
  ARMLoss  z as AMLoss, then every entry whose z is below its row's target z is replaced by 0
  CELoss   z = the logits as they are (no scale)
"""
from torch import nn

from ppvector import _lib
from ppvector.loss.aamloss import _AAMFunction


class _MarginHead(nn.Module):
    head = None

    def __init__(self, margin=0.2, scale=30, label_smoothing=0.0):
        super().__init__()
        self.margin, self.scale, self.label_smoothing = margin, scale, label_smoothing
        self.easy_margin = self.head  # what the training loop forwards to the C ABI as the head selector

    def forward(self, inputs, labels):
        weight = inputs.get('_weight')
        if weight is None:
            raise _lib.PPVError(f"{type(self).__name__} on B200 needs the classifier weight: pass SpeakerIdentification's output dict")
        return _AAMFunction.apply(inputs['features'], weight, labels, self.margin, self.scale, self.head, self.label_smoothing)

    def update(self, margin=0.2):
        self.margin = margin


class AMLoss(_MarginHead):
    head = _lib.PPV_HEAD_AM


class ARMLoss(_MarginHead):
    head = _lib.PPV_HEAD_ARM


class CELoss(_MarginHead):
    head = _lib.PPV_HEAD_CE

    def __init__(self, label_smoothing=0.0):
        super().__init__(margin=0.0, scale=1.0, label_smoothing=label_smoothing)

    def update(self, margin=0.2):  # celoss.py:21-22
        pass
