"""NVLS multicast buffers (Python face of `ops/csrc/runtime/mcast.cpp`).

`MulticastBuffer(fabric, nbytes)` is a collective constructor: every rank
allocates a VMM segment, rank 0 creates the NVSwitch multicast object and ships
its file descriptor to the peers over abstract Unix sockets (SCM_RIGHTS), all
ranks add their device, bind their memory and map the multicast view.
`mc_ptr` is the address `multimem.ld_reduce` / `multimem.st` operate on;
`tensor()` is the ordinary (unicast) view of this rank's copy.
"""
import ctypes
import itertools
import os

import torch

from .. import ops
from .symmetric import _CAI

_vp, _i = ctypes.c_void_p, ctypes.c_int
_PP = ctypes.POINTER(ctypes.c_void_p)
ops.register_signatures({
    "px_mc_last_error": (ctypes.c_char_p, []),
    "px_mc_supported": (_i, []),
    "px_fd_listen": (_i, [ctypes.c_char_p, _i]),
    "px_fd_send": (_i, [ctypes.c_char_p, _i, _i, _i]),
    "px_fd_recv": (_i, [ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "px_fd_close_listener": (None, []),
    "px_fd_close": (None, [_i]),
    "px_mc_round_size": (ctypes.c_longlong, [ctypes.c_size_t, _i]),
    "px_mc_seg_create": (_i, [ctypes.c_size_t, _PP, _PP]),
    "px_mc_create_export": (_i, [_vp, _i, ctypes.POINTER(_i)]),
    "px_mc_import": (_i, [_vp, _i]),
    "px_mc_add_device": (_i, [_vp]),
    "px_mc_bind_map": (_i, [_vp, _PP]),
    "px_mc_seg_destroy": (_i, [_vp]),
    "px_allreduce_nvls": (_i, [_vp, _vp, _vp, _i, _i, ctypes.c_size_t, _i, ctypes.c_float,
                               _i, _i, _i, _vp]),
})
_seq = itertools.count()


class MulticastError(RuntimeError):
    pass


def _ck(rc, what):
    if rc != 0:
        raise MulticastError("%s: %s" % (what, ops.lib().px_mc_last_error().decode()))


def supported(comm):
    """Collective: True iff every rank's device supports NVSwitch multicast."""
    if not comm.is_cuda:
        return False
    try:
        mine = bool(ops.lib().px_mc_supported())
    except Exception:
        mine = False
    return all(comm.all_gather_object(mine))


class MulticastBuffer(object):
    def __init__(self, fabric, nbytes):
        L = ops.lib()
        comm = fabric.comm
        self.fabric, self.L = fabric, L
        W, rank = comm.world, comm.rank
        self.device = fabric.device
        job = comm.broadcast_object("%d_%d" % (os.getpid(), next(_seq)), 0)
        size = L.px_mc_round_size(int(nbytes), W)
        if size < 0:
            raise MulticastError(L.px_mc_last_error().decode())
        size = max(comm.all_gather_object(int(size)))
        self.nbytes = size
        seg, uc = _vp(), _vp()
        _ck(L.px_mc_seg_create(size, ctypes.byref(seg), ctypes.byref(uc)), "seg_create")
        self.seg, self.uc_ptr = seg, uc.value
        _ck(L.px_fd_listen(job.encode(), rank), "fd_listen")
        comm.barrier()
        ok, err = True, ""
        # rank 0 creates the multicast object first and tells everybody whether it
        # worked — peers must never block in recv on an fd that will not come
        fd0 = _i(-1)
        if rank == 0:
            try:
                _ck(L.px_mc_create_export(seg, W, ctypes.byref(fd0)), "mc_create")
            except MulticastError as e:
                ok, err = False, str(e)
        ok0, err0 = comm.broadcast_object((ok, err), 0)
        if not ok0:
            L.px_fd_close_listener()
            L.px_mc_seg_destroy(seg)
            raise MulticastError("multicast object creation failed on rank 0: %s" % err0)
        try:
            if rank == 0:
                for r in range(1, W):
                    _ck(L.px_fd_send(job.encode(), r, fd0.value, 0), "fd_send")
                L.px_fd_close(fd0.value)
            else:
                fd, tag = _i(), _i()
                _ck(L.px_fd_recv(ctypes.byref(fd), ctypes.byref(tag)), "fd_recv")
                _ck(L.px_mc_import(seg, fd.value), "mc_import")
                L.px_fd_close(fd.value)
            _ck(L.px_mc_add_device(seg), "add_device")
        except MulticastError as e:
            ok, err = False, str(e)
        L.px_fd_close_listener()
        oks = comm.all_gather_object((ok, err))
        if not all(o for o, _ in oks):
            raise MulticastError("multicast setup failed: %s" % [e for o, e in oks if not o])
        mc = _vp()
        ok, err = True, ""
        try:
            _ck(L.px_mc_bind_map(seg, ctypes.byref(mc)), "bind_map")
        except MulticastError as e:
            ok, err = False, str(e)
        oks = comm.all_gather_object((ok, err))
        if not all(o for o, _ in oks):
            raise MulticastError("multicast bind/map failed: %s" % [e for o, e in oks if not o])
        self.mc_ptr = mc.value
        torch.cuda.synchronize(self.device)
        comm.barrier()
        self._bytes = None

    # -- SymmBuffer-compatible face (used by the dense engine) -----------------
    @property
    def local_ptr(self):
        return self.uc_ptr

    def mc_c_ptrs(self):
        """ctypes array whose entry 0 is the multicast address."""
        arr = getattr(self, "_mc_arr", None)
        if arr is None:
            arr = (ctypes.c_void_p * 1)(self.mc_ptr)
            self._mc_arr = arr
        return arr

    def tensor(self, dtype, numel=None):
        if self._bytes is None:
            self._bytes = torch.as_tensor(_CAI(self.uc_ptr, self.nbytes, self),
                                          device=self.device)
        es = torch.empty((), dtype=dtype).element_size()
        if numel is None:
            numel = self.nbytes // es
        return self._bytes[:numel * es].view(dtype)

    def allreduce_(self, n, dtype, scale, channels, max_blocks=32, stream=None):
        """In-place NVLS all-reduce of the first `n` elements (n % (W·16B) == 0)."""
        from . import nvops
        heap = self.fabric.heap
        nvops._count()
        ops.check(self.L.px_allreduce_nvls(
            _vp(self.mc_ptr), nvops._p(heap.pads_dev()), nvops._p(heap.epoch), channels[0],
            channels[1], n, nvops.DT[dtype], scale, heap.rank, heap.world, max_blocks,
            nvops._s(stream)), "allreduce_nvls")

    def close(self):
        self._bytes = None
        self.L.px_mc_seg_destroy(self.seg)
