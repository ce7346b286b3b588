"""Learning-rate and margin schedules of the training loop, host side.

``cosine_decay_with_warmup`` reproduces the learning rates of the reference's function of the same name
(ppvector/optimizer/scheduler.py:6-40, a ``PiecewiseDecay`` table) in closed form: step s < W (W = warmup_epoch * step_per_epoch)
gives ``lr * s / W``; later steps give ``min_lr + (lr - min_lr) * (1 + cos(pi * (s - W) / (M - W))) / 2`` with M = fix_epoch *
step_per_epoch, held at its last value from step M - 1 on.  The object has paddle's scheduler surface (``get_lr`` / ``step`` /
``state_dict`` / ``set_state_dict``).  ``MarginScheduler`` has the reference class's constructor and methods
(scheduler.py:43-102): the AAM margin stays at ``initial_margin`` until ``increase_start_epoch``, then ramps (exponentially by
default) to ``final_margin`` at ``fix_epoch``."""
import math


class WarmupCosineLR:
    def __init__(self, learning_rate, step_per_epoch, fix_epoch=1000, warmup_epoch=5, min_lr=0.0):
        self.peak, self.floor = float(learning_rate), float(min_lr)
        self.warmup_steps = int(warmup_epoch) * int(step_per_epoch)
        self.total_steps = int(fix_epoch) * int(step_per_epoch)
        self.last_epoch = 0  # paddle's name for "steps taken so far"

    def lr_at(self, step):
        W, M = self.warmup_steps, self.total_steps
        if step < W:
            return self.peak * step / W
        s = min(step, max(M - 1, W))
        span = max(M - W, 1)
        return self.floor + (self.peak - self.floor) * 0.5 * (1.0 + math.cos(math.pi * (s - W) / span))

    def get_lr(self):
        return self.lr_at(self.last_epoch)

    def step(self, epoch=None):
        self.last_epoch = self.last_epoch + 1 if epoch is None else int(epoch)

    def state_dict(self):
        return {'last_epoch': self.last_epoch}

    def set_state_dict(self, state):
        self.last_epoch = int(state.get('last_epoch', 0))


def cosine_decay_with_warmup(learning_rate, step_per_epoch, fix_epoch=1000, warmup_epoch=5, min_lr=0.0):
    return WarmupCosineLR(learning_rate, step_per_epoch, fix_epoch=fix_epoch, warmup_epoch=warmup_epoch, min_lr=min_lr)


def margin_at_step(step, start_step, fix_step, initial_margin, final_margin, increase_type='exp'):
    """Margin in force at training step ``step`` (0-based)."""
    if step < start_step:
        return initial_margin
    if step >= fix_step:
        return final_margin
    progress = (step - start_step) / (fix_step - start_step)
    if increase_type == 'exp':
        # 1 - (1e-3 / (1 + 1e-6)) ** progress: 0 at the start, 0.999 just before fix_step
        ramp = 1.0 - math.exp(progress * math.log(1e-3 / (1.0 + 1e-6)))
    else:
        ramp = progress
    return initial_margin + (final_margin - initial_margin) * ramp


class MarginScheduler:
    def __init__(self, criterion, increase_start_epoch, fix_epoch, step_per_epoch, initial_margin=0.0, final_margin=0.3,
                 increase_type='exp'):
        if not hasattr(criterion, 'update'):
            raise AssertionError("Loss function not has 'update()' attributes.")
        self.criterion = criterion
        self.increase_start_step = increase_start_epoch * step_per_epoch
        self.fix_step = fix_epoch * step_per_epoch
        self.increase_step = self.fix_step - self.increase_start_step
        self.initial_margin, self.final_margin, self.increase_type = initial_margin, final_margin, increase_type
        self.current_step = 0
        self.margin = initial_margin
        self.init_margin()

    def init_margin(self):
        self.criterion.update(margin=self.initial_margin)

    def iter_margin(self):
        return margin_at_step(self.current_step, self.increase_start_step, self.fix_step, self.initial_margin, self.final_margin, self.increase_type)

    def step(self, current_step=None):
        if current_step is not None:
            self.current_step = current_step
        self.margin = self.iter_margin()
        self.criterion.update(margin=self.margin)
        self.current_step += 1

    def get_margin(self):
        return self.margin
