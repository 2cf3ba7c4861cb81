"""Training-loop callbacks with Horovod's Keras semantics.

Parity: `horovod/_keras/callbacks.py`, `horovod/keras/callbacks.py`
(`BroadcastGlobalVariablesCallback`, `MetricAverageCallback`,
`LearningRateScheduleCallback`, `LearningRateWarmupCallback`).  Keras drives
callbacks itself; here `CallbackList` is the small driver a torch loop calls
(`on_train_begin`, `on_epoch_begin/end`, `on_batch_begin/end`).  Callbacks act
on a ``torch.optim.Optimizer`` (or a `collectives.DistributedOptimizer`) through
its `param_groups`.
"""
import torch

from . import collectives as hvd


class Callback(object):
    def set_context(self, model=None, optimizer=None):
        self.model, self.optimizer = model, optimizer

    def on_train_begin(self, logs=None):
        pass

    def on_epoch_begin(self, epoch, logs=None):
        pass

    def on_batch_begin(self, batch, logs=None):
        pass

    def on_batch_end(self, batch, logs=None):
        pass

    def on_epoch_end(self, epoch, logs=None):
        pass


class CallbackList(object):
    def __init__(self, callbacks, model=None, optimizer=None):
        self.callbacks = list(callbacks)
        for c in self.callbacks:
            c.set_context(model, optimizer)

    def __getattr__(self, name):
        if not name.startswith("on_"):
            raise AttributeError(name)

        def call(*a, **kw):
            for c in self.callbacks:
                getattr(c, name)(*a, **kw)
        return call


class BroadcastGlobalVariablesCallback(Callback):
    """broadcast model parameters/buffers and optimizer state from `root_rank` at
    the start of training, so every worker starts from the same point (random
    init or a checkpoint restored on one rank)"""

    def __init__(self, root_rank=0):
        self.root_rank, self.done = root_rank, False

    def on_train_begin(self, logs=None):
        if self.done:
            return
        if self.model is not None:
            hvd.broadcast_parameters(self.model.state_dict(), self.root_rank)
        if self.optimizer is not None:
            hvd.broadcast_optimizer_state(self.optimizer, self.root_rank)
        self.done = True


class MetricAverageCallback(Callback):
    """average the epoch's metrics over all workers, in place in `logs`
    (metric names are reduced in sorted order so all ranks agree)"""

    def on_epoch_end(self, epoch, logs=None):
        if not logs:
            return
        for k in sorted(logs):
            v = logs[k]
            if isinstance(v, (int, float)) or (torch.is_tensor(v) and v.numel() == 1):
                t = torch.as_tensor(float(v), dtype=torch.float64)
                logs[k] = float(hvd.allreduce(t, average=True, name="metric.%s" % k))


class LearningRateScheduleCallback(Callback):
    """lr = initial_lr × multiplier(epoch) for `start_epoch` ≤ epoch < `end_epoch`.

    `multiplier` is a constant or a function of the (possibly fractional) epoch;
    with `staircase` it is evaluated at whole epochs, otherwise at every batch
    (needs `steps_per_epoch`).  `momentum_correction` rescales the momentum
    buffers by new_lr/old_lr when the rate changes (Goyal et al., as in Horovod)."""

    def __init__(self, multiplier, start_epoch=0, end_epoch=None, staircase=True,
                 momentum_correction=True, steps_per_epoch=None, initial_lr=None):
        self.multiplier = multiplier if callable(multiplier) else (lambda epoch: multiplier)
        self.start_epoch, self.end_epoch = start_epoch, end_epoch
        self.staircase, self.momentum_correction = staircase, momentum_correction
        self.steps_per_epoch, self.initial_lr = steps_per_epoch, initial_lr
        self.current_epoch = None
        if not staircase and not steps_per_epoch:
            raise ValueError("a smooth schedule (staircase=False) needs steps_per_epoch")

    def _opt(self):
        o = self.optimizer
        return o.optimizer if isinstance(o, hvd.DistributedOptimizer) else o

    def on_train_begin(self, logs=None):
        if self.initial_lr is None:
            self.initial_lr = [g["lr"] for g in self._opt().param_groups]
        elif not isinstance(self.initial_lr, (list, tuple)):
            self.initial_lr = [self.initial_lr] * len(self._opt().param_groups)

    def _in_range(self, epoch):
        return epoch >= self.start_epoch and (self.end_epoch is None or epoch < self.end_epoch)

    def _adjust(self, epoch):
        opt = self._opt()
        for g, base in zip(opt.param_groups, self.initial_lr):
            old, new = g["lr"], base * self.multiplier(epoch)
            g["lr"] = new
            if self.momentum_correction and old > 0 and new != old:
                for p in g["params"]:
                    buf = opt.state.get(p, {}).get("momentum_buffer")
                    if buf is not None:
                        buf.mul_(new / old)

    def on_epoch_begin(self, epoch, logs=None):
        self.current_epoch = epoch
        if self.staircase and self._in_range(epoch):
            self._adjust(epoch)

    def on_batch_begin(self, batch, logs=None):
        if self.staircase or self.current_epoch is None:
            return
        epoch = self.current_epoch + float(batch) / self.steps_per_epoch
        if self._in_range(epoch):
            self._adjust(epoch)

    def on_epoch_end(self, epoch, logs=None):
        if logs is not None:
            logs["lr"] = self._opt().param_groups[0]["lr"]


class LearningRateWarmupCallback(LearningRateScheduleCallback):
    """gradual warm-up from lr/size to lr over `warmup_epochs`
    (lr here is the already size-scaled rate, as in Horovod's recipe):
    multiplier(epoch) = 1/size · (epoch·(size−1)/warmup_epochs + 1)"""

    def __init__(self, warmup_epochs=5, momentum_correction=True, steps_per_epoch=None,
                 verbose=0, initial_lr=None):
        size = hvd.size()

        def multiplier(epoch):
            return 1.0 / size * (epoch * (size - 1) / float(warmup_epochs) + 1)
        super().__init__(multiplier, start_epoch=0, end_epoch=warmup_epochs, staircase=False,
                         momentum_correction=momentum_correction,
                         steps_per_epoch=steps_per_epoch, initial_lr=initial_lr)
        self.warmup_epochs, self.verbose = warmup_epochs, verbose

    def on_epoch_end(self, epoch, logs=None):
        super().on_epoch_end(epoch, logs)
        if epoch == self.warmup_epochs - 1:
            # land exactly on the target rate
            for g, base in zip(self._opt().param_groups, self.initial_lr):
                g["lr"] = base
            if self.verbose and hvd.rank() == 0:
                print("Epoch %d: finished gradual learning rate warmup to %g." %
                      (epoch + 1, self.initial_lr[0]))
