"""Autotuner (Python face of `ops/csrc/runtime/autotune.cpp`).

``PARALLAX_AUTOTUNE=1`` makes the engine tune its communication knobs during
the first steps (Horovod: ``HOROVOD_AUTOTUNE``, CSV log via
``HOROVOD_AUTOTUNE_LOG`` — `horovod/common/parameter_manager.cc:96-101`).
Knobs must be identical on all ranks (the per-CTA barrier slots depend on the
grid size), so rank 0 decides and the values are broadcast through the
control plane each time they change.
"""
import ctypes
import os
import time

from .. import ops
from ..consts import PARALLAX_AUTOTUNE, PARALLAX_AUTOTUNE_LOG

_i, _d = ctypes.c_int, ctypes.c_double
_pd, _pi = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
ops.register_signatures({
    "px_autotune_create": (_i, [_i, _pd, _pd, _i, _pi, _i, _i, _i, ctypes.c_uint]),
    "px_autotune_current": (_i, [_i, _pd, _pi]),
    "px_autotune_report": (_i, [_i, _d]),
    "px_autotune_num_points": (_i, [_i]),
    "px_autotune_best_score": (_d, [_i]),
    "px_autotune_destroy": (_i, [_i]),
})


class BayesianTuner(object):
    """Continuous knobs ``{name: (lo, hi)}`` searched jointly by GP/EI, then
    categorical knobs ``{name: [values]}`` swept one after the other."""

    def __init__(self, continuous, categorical=None, samples_per_point=5,
                 max_points=20, warmups=3, seed=1):
        self.L = ops.lib()
        self.cnames = list(continuous)
        self.knames = list(categorical or {})
        self.kvalues = [list(categorical[k]) for k in self.knames]
        nd, nk = len(self.cnames), len(self.knames)
        lo = (ctypes.c_double * max(nd, 1))(*[continuous[n][0] for n in self.cnames])
        hi = (ctypes.c_double * max(nd, 1))(*[continuous[n][1] for n in self.cnames])
        ks = (ctypes.c_int * max(nk, 1))(*[len(v) for v in self.kvalues])
        self.h = self.L.px_autotune_create(nd, lo, hi, nk, ks, samples_per_point,
                                           max_points, warmups, seed)
        self._x = (ctypes.c_double * max(nd, 1))()
        self._k = (ctypes.c_int * max(nk, 1))()
        self.done = False

    def current(self):
        rc = self.L.px_autotune_current(self.h, self._x, self._k)
        self.done = rc == 1
        out = {n: self._x[i] for i, n in enumerate(self.cnames)}
        out.update({n: self.kvalues[i][self._k[i]] for i, n in enumerate(self.knames)})
        return out

    def report(self, score):
        rc = self.L.px_autotune_report(self.h, float(score))
        if rc == 2:
            self.done = True
        return rc

    def num_points(self):
        return self.L.px_autotune_num_points(self.h)

    def best_score(self):
        return self.L.px_autotune_best_score(self.h)

    def close(self):
        self.L.px_autotune_destroy(self.h)


class EngineAutotuner(object):
    """Tunes `comm_blocks` (CTAs of the dense comm kernels) and `sparse_blocks`
    (CTA cap of the sparse kernels) from per-step throughput."""

    def __init__(self, engine):
        self.engine = engine
        self.tuner = BayesianTuner({"comm_blocks": (4.0, 64.0),
                                    "sparse_blocks": (32.0, 592.0)},
                                   max_points=12) if engine.comm.rank == 0 else None
        self.log = os.environ.get(PARALLAX_AUTOTUNE_LOG)
        self._t = None
        self.done = False
        self._apply(self._decide())

    @staticmethod
    def wanted():
        return os.environ.get(PARALLAX_AUTOTUNE, "0") not in ("0", "", "false", "False")

    def _decide(self):
        vals = None
        if self.tuner is not None:
            cur = self.tuner.current()
            vals = (int(round(cur["comm_blocks"])), int(round(cur["sparse_blocks"])),
                    self.tuner.done)
        return self.engine.comm.broadcast_object(vals, 0)

    def _apply(self, vals):
        cb, sb, done = vals
        eng = self.engine
        eng.fabric.max_blocks = max(1, min(128, cb))
        for grp in getattr(eng, "sparse_groups", ()):
            grp.max_blocks = max(1, sb)
        self.done = done
        if self.log and eng.comm.rank == 0:
            with open(self.log, "a") as f:
                f.write("%d,%d,%d,%s\n" % (eng.global_step, cb, sb, done))

    def step_begin(self):
        import torch
        torch.cuda.synchronize(self.engine.comm.device)
        self._t = time.perf_counter()

    def step_end(self):
        import torch
        if self.done:
            return
        torch.cuda.synchronize(self.engine.comm.device)
        dt = time.perf_counter() - self._t
        changed = 0
        if self.tuner is not None:
            changed = self.tuner.report(1.0 / max(dt, 1e-9))
        changed = self.engine.comm.broadcast_object(changed, 0)
        if changed:
            self._apply(self._decide())
