"""Stall watchdog (Python face of `ops/csrc/runtime/watchdog.cpp`).

Knobs (parity with HOROVOD_STALL_CHECK_TIME_SECONDS /
HOROVOD_STALL_SHUTDOWN_TIME_SECONDS, `horovod/common/operations.cc:1023-1027`):
``PARALLAX_STALL_CHECK_TIME_SECONDS`` (default 60) and
``PARALLAX_STALL_SHUTDOWN_TIME_SECONDS`` (default 0 = never).
"""
import ctypes
import os

from .. import ops
from ..consts import (PARALLAX_STALL_CHECK_TIME_SECONDS,
                      PARALLAX_STALL_SHUTDOWN_TIME_SECONDS)

_i, _d, _vp, _ll = ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_longlong
ops.register_signatures({
    "px_watchdog_start": (_i, [_i, _i, _d, _d, _vp, ctypes.c_size_t]),
    "px_watchdog_beat": (None, [_ll]),
    "px_watchdog_arm": (None, [_vp, _ll]),
    "px_watchdog_disarm": (None, []),
    "px_watchdog_stalls": (_i, []),
    "px_watchdog_stop": (_i, []),
})


class Watchdog(object):
    def __init__(self, rank, world, heap=None, warn_s=None, shutdown_s=None):
        self.L = ops.lib()
        warn_s = float(os.environ.get(PARALLAX_STALL_CHECK_TIME_SECONDS, 60)) \
            if warn_s is None else warn_s
        shutdown_s = float(os.environ.get(PARALLAX_STALL_SHUTDOWN_TIME_SECONDS, 0)) \
            if shutdown_s is None else shutdown_s
        pad, words = 0, 0
        if heap is not None:
            pad, words = heap.pad.local_ptr, heap.pad.nbytes // 4
        self.L.px_watchdog_start(rank, world, warn_s, shutdown_s, _vp(pad), words)
        self._ev = None
        self.device = heap is not None

    def step_enqueued(self, step):
        """Call after a step's work was enqueued: progress = its completion."""
        if self.device:
            import torch
            self._ev = torch.cuda.Event()
            self._ev.record()
            self.L.px_watchdog_arm(_vp(self._ev.cuda_event), int(step))
        else:
            self.L.px_watchdog_beat(int(step))

    def beat(self, step=0):
        self.L.px_watchdog_beat(int(step))

    def stalls(self):
        return int(self.L.px_watchdog_stalls())

    def stop(self):
        self.L.px_watchdog_disarm()
        self.L.px_watchdog_stop()
