"""Thin, allocation-free Python wrappers around the native kernels.

Each function launches exactly one kernel (or a fixed short sequence) on the
given/current stream; all are CUDA-graph capturable.
"""
import ctypes

import torch

from .. import ops
from ..optim import KIND_ID

_vp = ctypes.c_void_p
DT = {torch.float32: 0, torch.bfloat16: 1}

# launch counter — bench.py reports `gpu_launches` from it
launches = {"n": 0}


def _count(k=1):
    launches["n"] += k


def _s(stream):
    if stream is None:
        stream = torch.cuda.current_stream()
    return _vp(stream.cuda_stream)


def _p(t):
    return _vp(t.data_ptr()) if t is not None else _vp(0)


def allreduce_twoshot(heap, buf_cptrs, n, dtype, scale, channels, sumsq=None,
                      max_blocks=32, stream=None):
    L = ops.lib()
    _count()
    ops.check(L.px_allreduce_twoshot(
        buf_cptrs, _p(heap.pads_dev()), _p(heap.epoch), channels[0], channels[1],
        n, DT[dtype], scale, _p(sumsq), heap.rank, heap.world, max_blocks,
        _s(stream)), "allreduce_twoshot")


def allreduce_oneshot(heap, src, dst, stage_buf, n, dtype, scale, channel,
                      sumsq=None, max_blocks=8, stream=None):
    L = ops.lib()
    _count()
    ops.check(L.px_allreduce_oneshot(
        _p(src), _p(dst), stage_buf.c_ptrs(), stage_buf.nbytes // 2,
        _p(heap.pads_dev()), _p(heap.epoch), channel, n, DT[dtype], scale,
        _p(sumsq), heap.rank, heap.world, max_blocks, _s(stream)),
        "allreduce_oneshot")


def broadcast(heap, buf_cptrs, nbytes, root, channels, max_blocks=32, stream=None):
    L = ops.lib()
    _count()
    ops.check(L.px_broadcast(buf_cptrs, _p(heap.pads_dev()), _p(heap.epoch),
                             channels[0], channels[1], nbytes, root, heap.rank,
                             heap.world, max_blocks, _s(stream)), "broadcast")


def allgather(heap, buf_cptrs, slice_bytes, channels, max_blocks=32, stream=None):
    L = ops.lib()
    _count()
    ops.check(L.px_allgather(buf_cptrs, _p(heap.pads_dev()), _p(heap.epoch),
                             channels[0], channels[1], slice_bytes, heap.rank,
                             heap.world, max_blocks, _s(stream)), "allgather")


def dense_step(heap, grads_cptrs, params_cptrs, master, slot0, slot1, ema, red,
               hp, clip, sumsq, n, avg, ema_decay, kind, mode, dtype, channels,
               rank=None, world=None, max_blocks=32, stream=None, use_mc=False):
    L = ops.lib()
    _count()
    ops.check(L.px_dense_step(
        grads_cptrs, params_cptrs, _p(master), _p(slot0), _p(slot1), _p(ema),
        _p(red), _p(hp), _p(clip), _p(sumsq), n, avg, ema_decay, KIND_ID[kind],
        mode, DT[dtype], _p(heap.pads_dev()) if heap is not None else _vp(0),
        _p(heap.epoch) if heap is not None else _vp(0), channels[0], channels[1],
        heap.rank if rank is None else rank,
        heap.world if world is None else world, max_blocks, 1 if use_mc else 0,
        _s(stream)), "dense_step")


def clip_scale(sumsq_total, max_norm, scale_out, norm_out, zero_after, stream=None):
    L = ops.lib()
    _count()
    ops.check(L.px_clip_scale(_p(sumsq_total), max_norm, _p(scale_out),
                              _p(norm_out), _p(zero_after), _s(stream)),
              "clip_scale")


def dense_async(my_grads, my_params, master_c, slot0_c, slot1_c, hp, clip, n,
                kind, dtype, rank, world, max_blocks=64, stream=None):
    L = ops.lib()
    _count()
    ops.check(L.px_dense_async(_p(my_grads), _p(my_params), master_c, slot0_c,
                               slot1_c, _p(hp), _p(clip), n, KIND_ID[kind],
                               DT[dtype], rank, world, max_blocks, _s(stream)),
              "dense_async")


def sumsq(x, n, dtype, mul, out, stream=None):
    L = ops.lib()
    _count()
    ops.check(L.px_sumsq(_p(x), n, DT[dtype], mul, _p(out), _s(stream)), "sumsq")


def make_geom(layout, D4):
    g = ops.PxTableGeom()
    g.V, g.P, g.W = layout.V, layout.P, layout.world
    g.rows_per_part, g.D4 = layout.rows_per_part, D4
    g.strategy = 0 if layout.strategy == "mod" else 1
    g.replicated = 1 if layout.replicated else 0
    g.extras = getattr(layout, "_extras", 0)
    g.base = getattr(layout, "_base", 0)
    return g


def sparse_lookup(ids, n, tables_dev, out, pend_ids, geom, hdr_ptr, ctl, wait,
                  stream=None):
    L = ops.lib()
    _count()
    ops.check(L.px_sparse_lookup(
        _p(ids), 1 if ids.dtype == torch.int64 else 0, n, _p(tables_dev), _p(out),
        DT[out.dtype], _p(pend_ids), ctypes.byref(geom), _vp(hdr_ptr), _p(ctl),
        1 if wait else 0, _s(stream)), "sparse_lookup")


def sparse_dedup(pend_ids, n, hbits, keys, slot_u, uniq_id, uniq_k, uniq_cnt,
                 pos2u, ctl, geom, dedup, use_smem, stream=None):
    L = ops.lib()
    _count(1 if use_smem else (3 if dedup else 1))
    ops.check(L.px_sparse_dedup(_p(pend_ids), n, hbits, _p(keys), _p(slot_u),
                                _p(uniq_id), _p(uniq_k), _p(uniq_cnt), _p(pos2u),
                                _p(ctl), ctypes.byref(geom), 1 if dedup else 0,
                                1 if use_smem else 0, _s(stream)), "sparse_dedup")


def sparse_push(pend_grads, n, pos2u, uniq_id, uniq_k, uniq_cnt, staging, ctl,
                rings_dev, hdrs_dev, ring_ids_off, cap, geom, scale, rank,
                max_blocks=592, stream=None):
    L = ops.lib()
    _count(2)
    ops.check(L.px_sparse_push(_p(pend_grads), DT[pend_grads.dtype], n, _p(pos2u),
                               _p(uniq_id), _p(uniq_k), _p(uniq_cnt), _p(staging),
                               _p(ctl), _p(rings_dev), _p(hdrs_dev), ring_ids_off, cap,
                               ctypes.byref(geom), scale, rank, max_blocks,
                               _s(stream)), "sparse_push")


def sparse_claim(ring_ptr, hdr_ptr, ring_ids_off, cap, slotmap, ctl, geom,
                 max_blocks=64, stream=None):
    L = ops.lib()
    _count()
    ops.check(L.px_sparse_claim(_vp(ring_ptr), _vp(hdr_ptr), ring_ids_off, cap,
                                _p(slotmap), _p(ctl), ctypes.byref(geom),
                                max_blocks, _s(stream)), "sparse_claim")


def sparse_apply(ring_ptr, hdr_ptr, ring_ids_off, cap, slotmap, table, slot0,
                 slot1, hp, avg, kind, ctl, hdrs_dev, geom, rank, use_slotmap,
                 max_blocks=64, stream=None):
    L = ops.lib()
    _count()
    ops.check(L.px_sparse_apply(_vp(ring_ptr), _vp(hdr_ptr), ring_ids_off, cap,
                                _p(slotmap), _p(table), _p(slot0), _p(slot1),
                                _p(hp), avg, KIND_ID[kind], _p(ctl), _p(hdrs_dev),
                                ctypes.byref(geom), rank, 1 if use_slotmap else 0,
                                max_blocks, _s(stream)), "sparse_apply")


def sparse_async_apply(pend_grads, n, pos2u, uniq_id, uniq_k, uniq_cnt, staging, ctl,
                       tables_dev, slot0s_dev, slot1s_dev, hp, scale, kind, geom,
                       max_blocks=592, stream=None):
    L = ops.lib()
    _count(2)
    ops.check(L.px_sparse_async_apply(
        _p(pend_grads), DT[pend_grads.dtype], n, _p(pos2u), _p(uniq_id), _p(uniq_k),
        _p(uniq_cnt), _p(staging), _p(ctl), _p(tables_dev), _p(slot0s_dev),
        _p(slot1s_dev), _p(hp), scale, KIND_ID[kind], ctypes.byref(geom), max_blocks,
        _s(stream)), "sparse_async_apply")
