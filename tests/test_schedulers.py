"""CPU: learning-rate and margin schedules of the training loop (ppvector/optimizer/scheduler.py of the reference)."""
import math

import pytest

from oracle import train as ot
from ppvector.optimizer.scheduler import MarginScheduler, cosine_decay_with_warmup


class _Crit:
    def update(self, margin):
        self.margin = margin


def test_margin_scheduler_matches_the_reference_formula():
    crit = _Crit()
    ms = MarginScheduler(crit, increase_start_epoch=2, fix_epoch=5, step_per_epoch=10, initial_margin=0.0, final_margin=0.3)
    assert crit.margin == 0.0
    seen = []
    for step in range(60):
        ms.step()
        seen.append(crit.margin)
        assert abs(crit.margin - ot.margin_at(step, 20, 50, 0.0, 0.3)) < 1e-12
    assert seen[19] == 0.0 and 0.0 <= seen[20] < seen[35] < seen[49] < 0.3 and seen[50] == 0.3 == seen[-1]
    # exponential ramp: ratio = 1 - exp(x * ln(1e-3 / (1 + 1e-6)))
    x = (35 - 20) / 30
    assert abs(seen[35] - 0.3 * (1 - math.exp(x * math.log(1e-3 / (1 + 1e-6))))) < 1e-12


def test_warmup_cosine_table():
    sch = cosine_decay_with_warmup(learning_rate=1e-3, step_per_epoch=10, fix_epoch=6, warmup_epoch=2, min_lr=1e-5)
    lrs = []
    for _ in range(70):
        lrs.append(sch.get_lr())
        sch.step()
    assert lrs[0] == 0.0 and abs(lrs[10] - 5e-4) < 1e-12 and abs(max(lrs) - 1e-3) < 1e-9
    peak = lrs.index(max(lrs))
    assert 19 <= peak <= 21 and all(a <= b + 1e-15 for a, b in zip(lrs[:peak], lrs[1:peak + 1]))
    assert all(a >= b - 1e-15 for a, b in zip(lrs[peak:59], lrs[peak + 1:60]))
    assert lrs[-1] >= 1e-5 - 1e-12 and lrs[-1] < 2e-5
    assert sch.state_dict() == {'last_epoch': 70}


def test_unknown_scheduler_raises():
    from ppvector.optimizer import build_lr_scheduler
    from ppvector.utils.utils import dict_to_object
    cfg = dict_to_object({'optimizer_conf': {'scheduler': 'CosineAnnealingDecay', 'scheduler_args': {}}, 'train_conf': {'max_epoch': 3}})
    with pytest.raises(NotImplementedError):
        build_lr_scheduler(10, cfg)
