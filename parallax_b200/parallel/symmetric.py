"""Symmetric heap + signal pads: Python face of `ops/csrc/runtime/symm.cpp`.

Every rank allocates the same sequence of segments; segment i of rank r is
addressable from every rank (CUDA IPC over NVLink).  The reference's
equivalents are NCCL communicators bootstrapped over MPI
(`horovod/common/ops/nccl_operations.cc:111-153`) and TF's gRPC/verbs/GDR
rendezvous (`tensorflow/core/distributed_runtime/rpc/grpc_worker_service.cc:427-500`,
`tensorflow/contrib/{verbs,gdr}`).

Two exchange strategies:
* `IpcExchange`  — one process per GPU; 64-byte IPC handles travel through
  the control plane (`Comm.all_gather_object`).
* `LocalExchange` — a world simulated inside one process (all "ranks" on one
  GPU, or one rank per visible GPU with direct peer access).  Used by the
  single-GPU kernel tests: the kernels cannot tell the difference.
"""
import ctypes

import torch

from .. import ops

PX_MAX_RANKS = 16
PX_MAX_BLOCKS = 128
PX_NUM_CHANNELS = 8
PAD_BYTES = PX_NUM_CHANNELS * PX_MAX_BLOCKS * PX_MAX_RANKS * 4

# barrier channel plan (kernels on one stream may share a pair)
CH_COMM = (0, 1)       # comm stream: dense buckets
CH_MAIN = (2, 3)       # main stream: broadcast / generic collectives
CH_SMALL = 4           # one-shot all-reduce (norms, small tensors)
CH_USER = (5, 6)       # public ops API


class _CAI(object):
    """Minimal __cuda_array_interface__ holder so torch can alias raw memory."""

    def __init__(self, ptr, nbytes, owner):
        self.__cuda_array_interface__ = {
            "shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
            "version": 3, "strides": None}
        self._owner = owner


class SymmBuffer(object):
    def __init__(self, heap, local_ptr, nbytes, name):
        self.heap = heap
        self.local_ptr = local_ptr
        self.nbytes = nbytes
        self.name = name
        self.peer_ptrs = None       # list[int], natural rank order
        self._bytes = None
        self._dev_ptrs = None
        self._c_arrays = {}

    def bytes_tensor(self):
        if self._bytes is None:
            self._bytes = torch.as_tensor(
                _CAI(self.local_ptr, self.nbytes, self), device=self.heap.device)
        return self._bytes

    def tensor(self, dtype, numel=None, byte_offset=0):
        es = torch.empty((), dtype=dtype).element_size()
        b = self.bytes_tensor()
        if numel is None:
            numel = (self.nbytes - byte_offset) // es
        return b[byte_offset:byte_offset + numel * es].view(dtype)

    def c_ptrs(self, byte_offset=0):
        """ctypes array of `world` void* (host) — peers' pointers + offset."""
        key = byte_offset
        arr = self._c_arrays.get(key)
        if arr is None:
            arr = (ctypes.c_void_p * len(self.peer_ptrs))(
                *[p + byte_offset for p in self.peer_ptrs])
            self._c_arrays[key] = arr
        return arr

    def dev_ptrs(self, byte_offset=0):
        """Device int64 tensor of `world` pointers (for dynamic indexing)."""
        t = torch.tensor([p + byte_offset for p in self.peer_ptrs],
                         dtype=torch.int64, device=self.heap.device)
        return t


class IpcExchange(object):
    def __init__(self, comm):
        self.comm = comm
        self.rank, self.world = comm.rank, comm.world

    def exchange(self, local_ptr, nbytes):
        L = ops.lib()
        if self.world == 1:
            return [local_ptr]
        h = ctypes.create_string_buffer(64)
        ops.check(L.px_ipc_export(ctypes.c_void_p(local_ptr), h), "ipc_export")
        allh = self.comm.all_gather_object((bytes(h.raw), nbytes))
        ptrs = []
        for r, (hb, nb) in enumerate(allh):
            if nb != nbytes:
                raise RuntimeError(
                    "symmetric allocation size mismatch: rank %d has %d bytes, "
                    "rank %d has %d" % (self.rank, nbytes, r, nb))
            if r == self.rank:
                ptrs.append(local_ptr)
            else:
                out = ctypes.c_void_p()
                ops.check(L.px_ipc_import(hb, ctypes.byref(out)), "ipc_import")
                ptrs.append(out.value)
        return ptrs


class LocalWorld(object):
    """Shared registry for ranks simulated inside one process."""

    def __init__(self, world):
        self.world = world
        self.table = {}     # seq -> [ptr per rank]

    def exchange_for(self, rank):
        return LocalExchange(self, rank)


class LocalExchange(object):
    def __init__(self, lw, rank):
        self.lw, self.rank, self.world = lw, rank, lw.world
        self.seq = 0
        self.pending = []

    def exchange(self, local_ptr, nbytes):
        slot = self.lw.table.setdefault(self.seq, [None] * self.world)
        slot[self.rank] = local_ptr
        self.seq += 1
        return slot         # filled in as the other ranks allocate (same list)


class SymmetricHeap(object):
    def __init__(self, device, exchange):
        self.device = torch.device(device)
        self.ex = exchange
        self.rank, self.world = exchange.rank, exchange.world
        self.buffers = []
        self.L = ops.lib()
        torch.cuda.set_device(self.device)
        torch.cuda.current_stream()     # make sure the context exists
        ops.check(self.L.px_set_device(self.device.index or 0), "set_device")
        # signal pad + epoch counters
        self.pad = self.alloc(PAD_BYTES, "signal_pad")
        self.epoch = torch.zeros(PX_NUM_CHANNELS * PX_MAX_BLOCKS,
                                 dtype=torch.int32, device=self.device)
        self._pads_dev = None

    def alloc(self, nbytes, name="buf"):
        nbytes = (int(nbytes) + 255) // 256 * 256
        out = ctypes.c_void_p()
        ops.check(self.L.px_symm_alloc(nbytes, ctypes.byref(out)), "symm_alloc")
        buf = SymmBuffer(self, out.value, nbytes, name)
        buf.peer_ptrs = self.ex.exchange(out.value, nbytes)
        self.buffers.append(buf)
        return buf

    def pads_dev(self):
        if self._pads_dev is None:
            assert all(p is not None for p in self.pad.peer_ptrs), \
                "simulated ranks must all be constructed before first use"
            self._pads_dev = self.pad.dev_ptrs()
        return self._pads_dev

    def barrier(self, channel=CH_MAIN[0], stream=None):
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        ops.check(self.L.px_barrier(
            ctypes.c_void_p(self.pads_dev().data_ptr()),
            ctypes.c_void_p(self.epoch.data_ptr()), channel, self.rank,
            self.world, 1, ctypes.c_void_p(s.cuda_stream)), "barrier")

    def free(self, buf):
        """Release one segment (collective: every rank frees the same segment)."""
        if buf not in self.buffers:
            return
        self.buffers.remove(buf)
        buf._bytes = None
        if isinstance(self.ex, IpcExchange) and buf.peer_ptrs:
            for r, p in enumerate(buf.peer_ptrs):
                if r != self.rank and p:
                    self.L.px_ipc_close(ctypes.c_void_p(p))
        self.L.px_symm_free(ctypes.c_void_p(buf.local_ptr))

    def close(self):
        for b in self.buffers:
            b._bytes = None
            if isinstance(self.ex, IpcExchange) and b.peer_ptrs:
                for r, p in enumerate(b.peer_ptrs):
                    if r != self.rank and p:
                        self.L.px_ipc_close(ctypes.c_void_p(p))
            self.L.px_symm_free(ctypes.c_void_p(b.local_ptr))
        self.buffers = []
