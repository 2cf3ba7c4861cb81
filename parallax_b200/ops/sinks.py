"""Gradient sinks and the weight-gradient side stream.

The reference packs every dense gradient into Horovod's fusion buffer with one
`cudaMemcpyAsync` per tensor and unpacks it afterwards
(`horovod/common/ops/cuda_operations.cc:105-121`).  Here the bucket IS the
gradient storage: a fused op that produces a parameter's gradient asks for the
parameter's *sink* — the slice of the symmetric gradient bucket registered by the
dense group — and lets its GEMM write there (`out=`); the dense group then sees a
gradient that already lives in the bucket and skips the pack copy.  Ordinary
autograd-produced gradients still take one multi-tensor copy.

Weight-gradient GEMMs do not feed the rest of the backward pass, so fused ops may
issue them on a side stream (`side_stream()`), overlapping the latency-bound
recurrent chain and the sparse push; the event recorded by `mark_async` is what
the bucket's fused reduce/update kernel waits for.
"""
import torch

_SINKS = {}          # param storage ptr -> (grad view shaped like the param, deliver callback)
_SIDE = {}


def register(param, grad_view, deliver):
    _SINKS[param.data_ptr()] = (grad_view, deliver)


def unregister_all(params=None):
    if params is None:
        _SINKS.clear()
        return
    for p in params:
        _SINKS.pop(p.data_ptr(), None)


def get(param):
    """The bucket slice that receives `param`'s gradient (shaped like `param`), or None
    when the parameter is not managed by a dense group (host fabric, plain torch)."""
    ent = _SINKS.get(param.data_ptr())
    return None if ent is None else ent[0]


def deliver(param, event=None):
    """Tell the dense group that `param`'s gradient is complete in its sink (`event`:
    recorded on the side stream that produced it, or None if the current stream did).
    The producer then returns None to autograd for this input: the gradient never
    travels through AccumulateGrad, so it cannot be copied or read too early."""
    _SINKS[param.data_ptr()][1](event)


def side_stream(device, which=0):
    """The device's side stream number `which` (0: weight-gradient GEMMs / prefetches,
    1: bandwidth-bound reductions that overlap those GEMMs)."""
    key = (torch.device(device).index or 0, int(which))
    s = _SIDE.get(key)
    if s is None:
        s = _SIDE[key] = torch.cuda.Stream(device)
    return s
