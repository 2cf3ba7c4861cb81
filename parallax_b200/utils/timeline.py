"""Chrome-trace timeline (Python face of `ops/csrc/runtime/timeline.cpp`).

Enable with ``PARALLAX_TIMELINE=<file>`` (rank 0 by default, like
``HOROVOD_TIMELINE`` — `horovod/common/operations.cc:1012-1021`) or
programmatically::

    from parallax_b200.utils import timeline
    timeline.start("trace.json")
    with timeline.activity("emb", "SPARSE_PUSH", gpu=True): ...
    timeline.stop()

Rows (tid) are named tensors / buckets / tables; activity names follow
Horovod's (`horovod/common/common.h:31-55`) where they still apply:
WAIT_FOR_DATA, ALLREDUCE, ALLGATHER, BROADCAST, plus DENSE_STEP, SPARSE_PUSH,
SPARSE_CLAIM, SPARSE_APPLY, LOOKUP and STEP / CYCLE_START markers.
"""
import contextlib
import ctypes
import os

from .. import ops
from ..consts import PARALLAX_TIMELINE

_c, _i, _ll, _vp = ctypes.c_char_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p
ops.register_signatures({
    "px_timeline_start": (_i, [_c, _i, _i]),
    "px_timeline_enabled": (_i, []),
    "px_timeline_event": (_i, [_c, _c, ctypes.c_char, _i, _ll, _c]),
    "px_timeline_gpu_begin": (_i, [_c, _c, _i, _c, _vp]),
    "px_timeline_gpu_end": (_i, [_i, _vp]),
    "px_timeline_written": (ctypes.c_long, []),
    "px_timeline_dropped": (ctypes.c_long, []),
    "px_timeline_stop": (_i, []),
})

_rows = {}
_on = False


def _tid(row):
    if row not in _rows:
        _rows[row] = len(_rows) + 1
        ops.lib().px_timeline_event(b"thread_name", b"__metadata", b"M", _rows[row], 0,
                                    str(row).encode())
    return _rows[row]


def start(path, rank=0, use_cuda=None):
    global _on
    if use_cuda is None:
        try:
            import torch
            use_cuda = torch.cuda.is_available()
        except Exception:
            use_cuda = False
    rc = ops.lib().px_timeline_start(str(path).encode(), int(rank), 1 if use_cuda else 0)
    if rc < 0:
        raise OSError("cannot open timeline file %r" % path)
    _rows.clear()
    _on = True


def start_from_env(rank):
    path = os.environ.get(PARALLAX_TIMELINE)
    if path and rank == 0 and not enabled():
        start(path, rank)
        return True
    return False


def enabled():
    return _on


def stop():
    global _on
    if _on:
        ops.lib().px_timeline_stop()
        _on = False


def instant(name, row="global", args=""):
    if _on:
        ops.lib().px_timeline_event(name.encode(), b"marker", b"i", _tid(row), 0,
                                    args.encode())


def begin(row, name, args=""):
    if _on:
        ops.lib().px_timeline_event(name.encode(), b"op", b"B", _tid(row), 0, args.encode())


def end(row, name):
    if _on:
        ops.lib().px_timeline_event(name.encode(), b"op", b"E", _tid(row), 0, b"")


@contextlib.contextmanager
def activity(row, name, gpu=False, stream=None, args=""):
    """Host ('B'/'E') or GPU (CUDA-event timed 'X') activity on `row`."""
    if not _on:
        yield
        return
    L = ops.lib()
    if gpu:
        import torch
        s = stream if stream is not None else torch.cuda.current_stream()
        h = L.px_timeline_gpu_begin(name.encode(), b"gpu", _tid(row), args.encode(),
                                    _vp(s.cuda_stream))
        try:
            yield
        finally:
            L.px_timeline_gpu_end(h, _vp(s.cuda_stream))
    else:
        begin(row, name, args)
        try:
            yield
        finally:
            end(row, name)


def stats():
    L = ops.lib()
    return {"written": L.px_timeline_written(), "dropped": L.px_timeline_dropped()}
