"""Learning-rate and margin schedules of the reference (ppvector/optimizer/scheduler.py:6-102), host side.

``cosine_decay_with_warmup`` builds the same piecewise table as the reference (linear warm-up over ``warmup_epoch`` epochs, then
half-cosine to ``min_lr`` at ``fix_epoch``) and returns an object with paddle's ``LRScheduler`` surface (``get_lr`` / ``step`` /
``state_dict``).  ``MarginScheduler`` is the reference class with the same constructor and methods."""
import math


class PiecewiseLR:
    """paddle.optimizer.lr.PiecewiseDecay(boundaries, values): values[i] while step < boundaries[i], values[-1] afterwards."""

    def __init__(self, boundaries, values):
        assert len(values) == len(boundaries) + 1 or len(values) == len(boundaries)
        self.boundaries, self.values = list(boundaries), list(values)
        self.last_epoch = 0

    def get_lr(self):
        for b, v in zip(self.boundaries, self.values):
            if self.last_epoch < b:
                return v
        return self.values[-1]

    def step(self, epoch=None):
        self.last_epoch = self.last_epoch + 1 if epoch is None else epoch

    def state_dict(self):
        return {'last_epoch': self.last_epoch}

    def set_state_dict(self, sd):
        self.last_epoch = int(sd.get('last_epoch', 0))


def cosine_decay_with_warmup(learning_rate, step_per_epoch, fix_epoch=1000, warmup_epoch=5, min_lr=0.0):
    """reference: scheduler.py:6-40 (same table, including its off-by-one conventions)"""
    boundary, value = [], []
    warmup_steps = warmup_epoch * step_per_epoch
    for i in range(warmup_steps + 1):
        if warmup_steps > 0:
            value.append(learning_rate * (i / warmup_steps))
        if i > 0:
            boundary.append(i)
    max_iters = fix_epoch * int(step_per_epoch)
    warmup_iters = len(boundary)
    for i in range(int(boundary[-1]) if boundary else 0, max_iters):
        boundary.append(i)
        value.append(min_lr + (learning_rate - min_lr) * 0.5 * (math.cos((i - warmup_iters) * math.pi / (max_iters - warmup_iters)) + 1))
    return PiecewiseLR(boundary, value)


class MarginScheduler:
    """reference: scheduler.py:43-102"""

    def __init__(self, criterion, increase_start_epoch, fix_epoch, step_per_epoch, initial_margin=0.0, final_margin=0.3, increase_type='exp'):
        assert hasattr(criterion, 'update'), "Loss function not has 'update()' attributes."
        self.criterion = criterion
        self.increase_start_step = increase_start_epoch * step_per_epoch
        self.fix_step = fix_epoch * step_per_epoch
        self.initial_margin, self.final_margin, self.increase_type = initial_margin, final_margin, increase_type
        self.margin = initial_margin
        self.current_step = 0
        self.increase_step = self.fix_step - self.increase_start_step
        self.init_margin()

    def init_margin(self):
        self.criterion.update(margin=self.initial_margin)

    def step(self, current_step=None):
        if current_step is not None:
            self.current_step = current_step
        self.margin = self.iter_margin()
        self.criterion.update(margin=self.margin)
        self.current_step += 1

    def iter_margin(self):
        if self.current_step < self.increase_start_step:
            return self.initial_margin
        if self.current_step >= self.fix_step:
            return self.final_margin
        a, b = 1.0, 1e-3
        current_step = self.current_step - self.increase_start_step
        if self.increase_type == 'exp':
            ratio = 1.0 - math.exp((current_step / self.increase_step) * math.log(b / (a + 1e-6))) * a
        else:
            ratio = 1.0 * current_step / self.increase_step
        return self.initial_margin + (self.final_margin - self.initial_margin) * ratio

    def get_margin(self):
        return self.margin
