"""`ParallaxSession` — the session-like handle returned by `parallel_run`.

Parity: reference `common/session_context.py:35-92` (`_parallax_run`: fetch and
feed names of the single-device graph are remapped to replicas, results come
back as a *list per replica*, feeds must be lists with one value per
replica — `doc/parallax_api.md:23-38`), `:54-71` (window timing for the
partition search), `:74-88` (profile steps) and the MonitoredTrainingSession
hooks (checkpoint saver).  Unlike the reference there is no extra blocking
``run(global_step)`` per step.
"""
import os
import time

import numpy as np
import torch

from . import consts
from .graph import GLOBAL_STEP, TRAIN_OP
from .log import parallax_log


def _flatten_fetches(fetches):
    """-> (list of leaf names, rebuild(fn leaf->value))"""
    if isinstance(fetches, str):
        return [fetches], lambda vals: vals[fetches]
    if isinstance(fetches, dict):
        subs = {k: _flatten_fetches(v) for k, v in fetches.items()}
        names = [n for s in subs.values() for n in s[0]]
        return names, lambda vals: {k: s[1](vals) for k, s in subs.items()}
    if isinstance(fetches, (list, tuple)):
        subs = [_flatten_fetches(v) for v in fetches]
        names = [n for s in subs for n in s[0]]
        typ = type(fetches)
        return names, lambda vals: typ(s[1](vals) for s in subs)
    name = getattr(fetches, "name", None)
    if name is not None:
        return [name], lambda vals: vals[name]
    raise TypeError("cannot fetch %r" % (fetches,))


class ParallaxSession(object):
    def __init__(self, engine, num_workers, worker_id,
                 num_replicas_per_worker=1, saver=None, profiler=None):
        self.engine = engine
        self.graph = engine.graph
        self.num_workers = num_workers
        self.worker_id = worker_id
        self.num_replicas_per_worker = num_replicas_per_worker
        self.saver = saver
        self.profiler = profiler
        self.fetch_mode = "numpy"
        self._search = os.environ.get(consts.PARALLAX_SEARCH, "False") == "True"
        self._search_addr = os.environ.get(consts.PARALLAX_SEARCH_ADDR)
        self._window = []
        self._reported = False
        self._closed = False

    # ------------------------------------------------------------------ feeds
    def _convert_feed(self, feed_dict):
        feeds = {}
        n = self.num_replicas_per_worker
        for k, v in (feed_dict or {}).items():
            name = k if isinstance(k, str) else getattr(k, "name")
            if name not in self.graph.placeholders:
                raise KeyError("%r is not a placeholder of the graph (%s)"
                               % (name, self.graph.placeholders))
            if isinstance(v, (list, tuple)):
                if len(v) != n:
                    raise ValueError(
                        "feed for %r must be a list with one value per "
                        "replica (%d), got %d" % (name, n, len(v)))
                v = v[0]
            feeds[name] = self._to_device(v)
        return feeds

    def _to_device(self, v):
        """Feeds stay where the user put them (typically pinned host memory);
        the engine issues the H2D copy itself — straight into the captured
        graph's static input buffers when the step is graph-replayed."""
        if isinstance(v, np.ndarray):
            v = torch.from_numpy(v)
        elif not torch.is_tensor(v):
            v = torch.as_tensor(v)
        return v

    # -------------------------------------------------------------------- run
    def run(self, fetches, feed_dict=None):
        assert not self._closed, "session is closed"
        names, rebuild = _flatten_fetches(fetches)
        feeds = self._convert_feed(feed_dict)
        training = TRAIN_OP in names
        eng = self.engine
        step_no = eng.global_step
        profiling = self.profiler is not None and \
            self.profiler.is_profile_step(step_no)
        if profiling:
            self.profiler.start()
        t0 = time.time()
        if training:
            out = eng.train_step(feeds)
        else:
            need_fwd = any(n not in (GLOBAL_STEP,) for n in names)
            out = eng.eval_step(feeds) if need_fwd else {}
        vals = {}
        for n in names:
            if n == GLOBAL_STEP:
                v = eng.global_step
            elif n == TRAIN_OP:
                v = None
            else:
                if n not in out:
                    raise KeyError("%r is not fetchable (have %s)"
                                   % (n, sorted(out) + [GLOBAL_STEP, TRAIN_OP]))
                v = self._convert_value(out[n])
            vals[n] = [v] * 1 if self.num_replicas_per_worker == 1 else \
                [v] * self.num_replicas_per_worker
        if profiling:
            if eng.comm.is_cuda:
                torch.cuda.synchronize()
            self.profiler.stop(step_no)
        if training:
            self._after_train_step(time.time() - t0)
        return rebuild(vals)

    def _convert_value(self, v):
        if not torch.is_tensor(v):
            return v
        if self.fetch_mode == "torch":
            return v
        v = v.detach()
        if v.dtype == torch.bfloat16:
            v = v.float()
        a = v.cpu().numpy()
        return a.item() if a.ndim == 0 else a

    def _after_train_step(self, dt):
        eng = self.engine
        if self.saver is not None:
            self.saver.after_step(eng.global_step)
        if self._search and not self._reported:
            s = eng.global_step
            if consts.NUM_ITERATIONS_FOR_WARMUP < s <= consts.NUM_ITERATIONS_FOR_TEST:
                self._window.append(dt)
            if s >= consts.NUM_ITERATIONS_FOR_TEST and self._window:
                self._reported = True
                mean = float(np.mean(self._window))
                parallax_log.info("partition search: mean step %.3f ms", mean * 1e3)
                if self._search_addr:
                    from .partitions import send_exec_time
                    try:
                        send_exec_time(self._search_addr, mean)
                    except Exception as e:  # pragma: no cover
                        parallax_log.warning("could not report exec time: %s", e)

    # ------------------------------------------------------------------ misc
    def should_stop(self):
        return self._closed

    def save_checkpoint(self):
        if self.saver is None:
            raise RuntimeError("no CheckPointConfig.ckpt_dir configured")
        return self.saver.save()

    def close(self):
        if self._closed:
            return
        self._closed = True
        self.engine.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
