"""Checkpoint save / restore-on-start.

Parity: reference `common/lib.py:38-56` (`build_ckpt_hooks`: a Saver over all
global variables + `CheckpointSaverHook(save_steps | save_secs)` installed as
a chief-only hook) and the runners passing ``checkpoint_dir`` only on worker 0
so an existing checkpoint is restored automatically at start
(`mpi/runner.py:178-193`, `hybrid/runner.py:243-257`); partitioned variables
are saved as slices of one logical variable.

Format: ``<ckpt_dir>/model.ckpt-<global_step>.pt`` holding full *logical*
tensors keyed by single-device variable names (weights, optimizer slots, EMA
shadows, global_step) plus a ``checkpoint`` index file naming the latest one —
so it can be resumed under a different world size, run option or partition
count.  Gathering of sharded state is a collective (all ranks participate);
only the chief writes.
"""
import os
import time

import torch

from .log import parallax_log

INDEX = "checkpoint"
PREFIX = "model.ckpt-"


def latest_checkpoint(ckpt_dir):
    if not ckpt_dir or not os.path.isdir(ckpt_dir):
        return None
    idx = os.path.join(ckpt_dir, INDEX)
    if os.path.exists(idx):
        with open(idx) as f:
            name = f.read().strip()
        path = os.path.join(ckpt_dir, name)
        if os.path.exists(path):
            return path
    cands = [f for f in os.listdir(ckpt_dir)
             if f.startswith(PREFIX) and f.endswith(".pt")]
    if not cands:
        return None
    cands.sort(key=lambda f: int(f[len(PREFIX):-3]))
    return os.path.join(ckpt_dir, cands[-1])


class CheckpointSaver(object):
    """`CheckpointSaverHook` analogue: `after_step` is called by the session
    after every training step on *all* workers."""

    def __init__(self, engine, ckpt_config, is_chief):
        self.engine = engine
        self.dir = ckpt_config.ckpt_dir
        self.save_steps = ckpt_config.save_ckpt_steps
        self.save_secs = ckpt_config.save_ckpt_secs
        self.is_chief = is_chief
        self._last_time = time.time()
        self._last_step = None
        self.enabled = bool(self.dir) and (self.save_steps or self.save_secs)

    def restore_if_present(self):
        """Restore-on-start; every rank loads the same logical state and keeps
        its own shard (the reference instead restores on worker 0 and
        broadcasts, `mpi/runner.py:134-139,197`)."""
        if not self.dir:
            return None
        path = latest_checkpoint(self.dir)
        # all ranks must agree on whether a checkpoint exists
        path = self.engine.comm.broadcast_object(path, 0)
        if path is None:
            return None
        sd = torch.load(path, map_location="cpu", weights_only=False)
        self.engine.load_state_dict(sd)
        parallax_log.info("restored checkpoint %s (global_step=%d)",
                          path, sd["global_step"])
        return path

    def _due(self, step):
        if self.save_steps:
            return step % int(self.save_steps) == 0 and step != self._last_step
        if self.save_secs:
            # decision must be identical on all ranks: chief decides
            due = (time.time() - self._last_time) >= float(self.save_secs)
            return bool(self.engine.comm.broadcast_object(due, 0))
        return False

    def after_step(self, step):
        if not self.enabled or not self._due(step):
            return None
        return self.save(step)

    def save(self, step=None):
        step = self.engine.global_step if step is None else step
        sd = self.engine.state_dict()          # collective
        path = None
        if self.is_chief:
            os.makedirs(self.dir, exist_ok=True)
            name = "%s%d.pt" % (PREFIX, step)
            path = os.path.join(self.dir, name)
            tmp = path + ".tmp"
            torch.save(sd, tmp)
            os.replace(tmp, path)
            with open(os.path.join(self.dir, INDEX), "w") as f:
                f.write(name)
            parallax_log.info("saved checkpoint %s", path)
        self._last_time = time.time()
        self._last_step = step
        return path
