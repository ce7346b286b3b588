"""paddle.optimizer.lr stand-in (TEST INFRASTRUCTURE)."""


class LRScheduler:
    def __init__(self, learning_rate=0.1, last_epoch=-1, verbose=False):
        self.base_lr = float(learning_rate)
        self.last_epoch = last_epoch
        self.last_lr = self.base_lr
        self.step()

    def __call__(self):
        return self.last_lr

    def step(self, epoch=None):
        self.last_epoch = self.last_epoch + 1 if epoch is None else epoch
        self.last_lr = self.get_lr()

    def get_lr(self):
        raise NotImplementedError


class PiecewiseDecay(LRScheduler):
    """values[i] for the first i with last_epoch < boundaries[i], else values[-1] (Paddle docs)."""

    def __init__(self, boundaries, values, last_epoch=-1, verbose=False):
        self.boundaries, self.values = list(boundaries), list(values)
        super().__init__(last_epoch=last_epoch, verbose=verbose)

    def get_lr(self):
        for i, b in enumerate(self.boundaries):
            if self.last_epoch < b:
                return self.values[i]
        return self.values[len(self.values) - 1]
