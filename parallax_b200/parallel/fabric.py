"""Process-group plumbing shared by both backends.

`Comm` is the bootstrap/control plane: rank/world discovery, object
exchange, barriers and the *library* collectives.  It replaces the
reference's ssh+env+MPI_Init+gRPC bootstrap (`hybrid/runner.py:195-214`,
`horovod/common/operations.cc:907-1007`).  Library collectives here are used
only (a) on the host fabric (CPU tests, gloo), (b) to exchange IPC handles
for the symmetric heap and (c) by the NCCL baseline — never on the NVLink
hot path, which talks through peer memory in our own kernels.
"""
import os
import datetime

import torch
import torch.distributed as dist

from ..log import parallax_log


_GLOO_GROUPS = {}      # one gloo side group per process, shared by all Comms


class Comm(object):
    def __init__(self, rank=0, world=1, local_rank=0, device=None, group=None,
                 owns_pg=False):
        self.rank = int(rank)
        self.world = int(world)
        self.local_rank = int(local_rank)
        self.device = device if device is not None else torch.device("cpu")
        self.group = group
        self.owns_pg = owns_pg
        self._host_group = None

    # -- construction --------------------------------------------------------
    @classmethod
    def from_env(cls, device=None, timeout_s=600):
        """Build from torchrun-style env (RANK/WORLD_SIZE/LOCAL_RANK/
        MASTER_ADDR/MASTER_PORT); world 1 if absent."""
        world = int(os.environ.get("WORLD_SIZE", "1"))
        rank = int(os.environ.get("RANK", "0"))
        local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
        # `run --start-timeout`: how long ranks wait for each other at start-up
        timeout_s = int(os.environ.get("PARALLAX_START_TIMEOUT", timeout_s))
        if device is None:
            if torch.cuda.is_available():
                device = torch.device("cuda", local_rank % torch.cuda.device_count())
            else:
                device = torch.device("cpu")
        if device.type == "cuda":
            torch.cuda.set_device(device)
        owns = False
        if world > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29511")
            backend = "nccl" if device.type == "cuda" else "gloo"
            kw = {}
            if backend == "nccl":
                kw["device_id"] = device
            dist.init_process_group(
                backend=backend, rank=rank, world_size=world,
                timeout=datetime.timedelta(seconds=timeout_s), **kw)
            owns = True
        return cls(rank, world, local_rank, device,
                   dist.group.WORLD if world > 1 else None, owns)

    @property
    def is_cuda(self):
        return self.device.type == "cuda"

    @property
    def distributed(self):
        return self.world > 1

    def host_group(self):
        """A gloo group for CPU-side object exchange when the main group is
        NCCL (keeps control traffic off the GPU streams)."""
        if not self.distributed:
            return None
        if self._host_group is None:
            if dist.get_backend(self.group) == "gloo":
                self._host_group = self.group
            else:
                key = id(self.group)
                if key not in _GLOO_GROUPS:
                    _GLOO_GROUPS[key] = dist.new_group(backend="gloo")
                self._host_group = _GLOO_GROUPS[key]
        return self._host_group

    # -- control plane -------------------------------------------------------
    def barrier(self):
        if self.distributed:
            dist.barrier(group=self.host_group())

    def all_gather_object(self, obj):
        if not self.distributed:
            return [obj]
        out = [None] * self.world
        dist.all_gather_object(out, obj, group=self.host_group())
        return out

    def all_reduce_max_int(self, value):
        """max over ranks of a small host integer (gloo, off the GPU streams)."""
        if not self.distributed:
            return int(value)
        t = torch.tensor([int(value)], dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.host_group())
        return int(t.item())

    def broadcast_object(self, obj, src=0):
        if not self.distributed:
            return obj
        box = [obj]
        dist.broadcast_object_list(box, src=src, group=self.host_group())
        return box[0]

    # -- library collectives (host fabric / baseline only) ---------------------
    def all_reduce_sum_(self, t):
        if self.distributed:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self._grp_for(t))
        return t

    def broadcast_(self, t, src=0):
        if self.distributed:
            dist.broadcast(t, src=src, group=self._grp_for(t))
        return t

    def all_gather_tensors(self, t):
        """All-gather equally-shaped tensors -> list of `world` tensors."""
        if not self.distributed:
            return [t]
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t.contiguous(), group=self._grp_for(t))
        return out

    def all_gather_varlen(self, t):
        """All-gather tensors whose dim 0 differs per rank."""
        if not self.distributed:
            return [t]
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        sizes = [int(x.item()) for x in self.all_gather_tensors(n)]
        m = max(sizes) if sizes else 0
        pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype,
                          device=t.device)
        if t.shape[0]:
            pad[:t.shape[0]] = t
        gathered = self.all_gather_tensors(pad)
        return [g[:s] for g, s in zip(gathered, sizes)]

    def all_to_all_varlen(self, t, send_counts):
        """Exchange row segments: the first ``send_counts[0]`` rows of `t` go to
        rank 0, the next ``send_counts[1]`` to rank 1, …  Returns
        ``(received rows concatenated in source-rank order, recv_counts)`` — the
        PS-style all-to-all of (index, row) pairs (BASELINE.md "PS-style as
        all-to-all") used by the library fabric."""
        if not self.distributed:
            return t, list(send_counts)
        W = self.world
        sc = torch.tensor(list(send_counts), dtype=torch.int64, device=t.device)
        rc = torch.empty(W, dtype=torch.int64, device=t.device)
        dist.all_to_all_single(rc, sc, group=self._grp_for(t))
        recv_counts = [int(x) for x in rc.tolist()]
        out = torch.empty((sum(recv_counts),) + tuple(t.shape[1:]), dtype=t.dtype,
                          device=t.device)
        dist.all_to_all_single(out, t.contiguous(), recv_counts, list(send_counts),
                               group=self._grp_for(t))
        return out, recv_counts

    def _grp_for(self, t):
        if t.device.type == "cpu":
            return self.host_group()
        return self.group

    def shutdown(self):
        if self.owns_pg and dist.is_initialized():
            try:
                dist.destroy_process_group()
            except Exception as e:  # pragma: no cover
                parallax_log.debug("destroy_process_group: %s", e)
