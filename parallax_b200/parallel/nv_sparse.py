"""Sparse variables on the NVLink fabric.

`NVSparseTable` is the storage of one row-partitioned variable (fp32 master rows,
optimizer slots and — for bf16 models — a bf16 *shadow* copy that lookups read, all
in symmetric memory).  `NVSparseGroup` is the machinery of one *group* of tables
that are looked up with the same ids (LM1B: ``softmax_w`` + ``softmax_b``; every
other table is a group of one): per step ONE remote-gather lookup kernel per
lookup call, ONE push kernel (SMEM local aggregation + P2P stores + flag) and ONE
owner kernel (cross-source merge + sparse optimizer + flag) — see
`ops/csrc/kernels/sparse.cu`.

Reference semantics kept (file:line in /root/reference/parallax/parallax):
* sparse sync across workers — `core/python/common/graph_transform_lib.py:1558-1946`
  (accumulate every worker's IndexedSlices on the variable's server, apply once,
  workers read the result);
* local aggregation — `:1372-1556` (`PSConfig.local_aggregation`);
* average_option — SUM vs ÷num_workers (`:101-102,385-387`);
* boundary between workers and servers — `:1315-1370`: size-increasing casts run
  on the consumer (owner) side, so bf16 gradients cross the wire as bf16 and are
  widened/accumulated in fp32 by the owner (`boundary_between_workers_and_servers
  =False` widens on the sender instead, fp32 on the wire);
* variable-size receive negotiation — `horovod/common/ops/collective_operations.cc
  :80-90`: ring capacity follows the largest per-rank row count; in eager mode the
  counts are agreed on every step (one small host all-reduce), under CUDA-graph
  replay shapes are static.
"""
import ctypes
import math

import torch

from .. import optim as _optim
from ..log import parallax_log
from . import modes
from .layout import TableLayout

_vp = ctypes.c_void_p
_ES = {torch.float32: 4, torch.bfloat16: 2}
_DT = {torch.float32: 0, torch.bfloat16: 1}


def _lib():
    from .. import ops
    return ops.lib()


def _count(k=1):
    from . import nvops
    nvops.launches["n"] += k


def _sp(stream):
    return _vp(stream.cuda_stream)


class HPStage(object):
    """Device copy of an optimizer's hyper-parameter vector, refreshed once per step
    through a ring of pinned buffers (the host may run several CUDA-graph replays ahead
    of the device, so one re-used pinned buffer could be overwritten before its H2D copy
    has executed)."""
    DEPTH = 8

    def __init__(self, optimizer, device):
        self.optimizer = optimizer
        self.dev = torch.zeros(_optim.HP_SIZE, dtype=torch.float32, device=device)
        self.host = [torch.zeros(_optim.HP_SIZE, dtype=torch.float32).pin_memory()
                     for _ in range(self.DEPTH)]
        self.events = [None] * self.DEPTH
        self.i = 0
        self.step = None

    def upload(self, step):
        if self.step == step:
            return self.dev
        i = self.i
        self.i = (i + 1) % self.DEPTH
        if self.events[i] is not None:
            self.events[i].synchronize()
        h = self.host[i]
        for j, v in enumerate(self.optimizer.hyper(step)):
            h[j] = v
        self.dev.copy_(h, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[i] = ev
        self.step = step
        return self.dev


def hp_stage(fabric, optimizer):
    cache = fabric.__dict__.setdefault("_hp_stages", {})
    st = cache.get(id(optimizer))
    if st is None:
        st = cache[id(optimizer)] = HPStage(optimizer, fabric.device)
    return st


class NVSparseTable(object):
    """Storage + per-table facade.  Kernels run through `self.group`."""

    def __init__(self, name, weight, num_partitions, strategy, optimizer, fabric,
                 route, graph, config, init=None, out_dtype=None, options=None,
                 owners=None, auto_group=True):
        self.name = name
        self.fabric, self.heap = fabric, fabric.heap
        self.comm = fabric.comm
        self.rank, self.world, self.device = fabric.rank, fabric.world, fabric.device
        self.route, self.optimizer, self.graph, self.config = route, optimizer, graph, config
        self.kind = optimizer.kind
        self.nslots = _optim.NUM_SLOTS[self.kind]
        self.V, self.D = int(weight.shape[0]), int(weight.shape[1])
        self.Dp = (self.D + 3) // 4 * 4
        self.D4 = self.Dp // 4
        self.replicated = route.sparse == modes.SPARSE_ALLGATHER
        self.layout = TableLayout(self.V, num_partitions, self.world, strategy,
                                  replicated=self.replicated, owners=owners)
        self.average = bool(config.average_sparse)
        ps = config.communication_config.ps_config
        self.local_aggregation = bool(ps.local_aggregation)
        self.scale = graph.scale_for(name)
        # PSConfig.boundary_between_workers_and_servers (graph_transform_lib.py:1315-1370):
        # True  — post-processing that does not grow the data (ScaleGradients) runs on the
        #         sender, the widening bf16→fp32 cast on the owner: bf16 on the wire;
        # False — nothing is moved: rows are widened to fp32 by the sender and the scale
        #         runs on the owner.
        self.boundary = bool(ps.boundary_between_workers_and_servers)
        opts = options or {}
        self.options = opts
        self.out_dtype = out_dtype or torch.float32
        self.anchor_device = self.device
        self.capacity_hint = (opts.get("sparse_capacity") or {}).get(name)
        L = self.layout
        rows = L.rows_local
        heap = self.heap
        self.tab_buf = heap.alloc(rows * self.Dp * 4, "table:" + name)
        self.table = self.tab_buf.tensor(torch.float32, rows * self.Dp).view(rows, self.Dp)
        self.slot_bufs, self.slots = [], []
        for v in optimizer.slot_init():
            sb = heap.alloc(rows * self.Dp * 4, "slot:" + name)
            t = sb.tensor(torch.float32, rows * self.Dp).view(rows, self.Dp)
            t.fill_(v)
            self.slot_bufs.append(sb)
            self.slots.append(t)
        # bf16 shadow rows for bf16 models: lookups move half the bytes over NVLink / HBM;
        # the owner kernel refreshes the shadow row together with the fp32 master row
        self.Dps = (self.D4 + 1) // 2 * 8          # shadow row length (16-byte vectors)
        want = opts.get("sparse_shadow", "auto")
        self.use_shadow = (self.out_dtype == torch.bfloat16 and bool(want) and
                           (want is True or rows * self.Dps * 2 <= (16 << 30)))
        self.shadow_buf = self.shadow = None
        if self.use_shadow:
            self.shadow_buf = heap.alloc(rows * self.Dps * 2, "shadow:" + name)
            self.shadow = self.shadow_buf.tensor(torch.bfloat16, rows * self.Dps) \
                .view(rows, self.Dps)
        self._init_weights(weight, init)
        self._ptrs = {}
        self.ring_buf = None
        self.staging = None
        self.group = None
        self.stats = {"pushed_rows": 0, "steps": 0}
        if auto_group:
            NVSparseGroup([self])

    # ------------------------------------------------------------- storage
    def _init_weights(self, weight, init):
        L = self.layout
        if weight.device.type == "meta":
            # lazy: initialise only this owner's rows on the device
            gen = torch.Generator(device=self.device)
            gen.manual_seed(int(init["seed"]) * 1000003 + (0 if self.replicated
                                                           else self.rank))
            self.table[:, :self.D].uniform_(-init["scale"], init["scale"], generator=gen)
            if self.Dp != self.D:
                self.table[:, self.D:].zero_()
        else:
            w = weight.detach().to(torch.float32)
            for g, l in L.owner_chunks(0 if self.replicated else self.rank):
                self.table[l.to(self.device), :self.D] = w[g].to(self.device)
        self.refresh_shadow()

    def refresh_shadow(self):
        if self.shadow is not None:
            self.shadow.zero_()
            self.shadow[:, :self.Dp].copy_(self.table)

    def dev_ptrs(self, what):
        """Device array of every rank's pointer for `what` (built lazily: in a simulated
        world the peers allocate after us)."""
        t = self._ptrs.get(what)
        if t is None:
            if what == "table":
                buf = self.tab_buf
            elif what == "shadow":
                buf = self.shadow_buf
            elif what == "ring":
                buf = self.ring_buf
            else:
                buf = self.slot_bufs[int(what[4:])]
            if self.replicated and what != "ring":
                # every replica reads and updates its own full copy
                t = torch.tensor([buf.local_ptr] * self.world, dtype=torch.int64,
                                 device=self.device)
            else:
                t = buf.dev_ptrs()
            self._ptrs[what] = t
        return t

    # ------------------------------------------------- per-table facade (group of 1..)
    def lookup(self, flat_ids, record=True):
        if record and len(self.group.tables) > 1:
            raise RuntimeError(
                "table %r belongs to the co-lookup group %s: training lookups must go "
                "through parallax.nn.lookup_many()" % (self.name, self.group.name))
        outs, pend = self.group.lookup(flat_ids, record=record,
                                       members=[self] if not record else None)
        return outs[0], pend

    def add_pending(self, token, grad_rows):
        self.group.add_pending(token, [grad_rows])

    def begin_step(self, step):
        self.group.begin_step(step)

    def finish_step(self, step, stream=None):
        self.group.finish_step(step, stream)

    def stage_push(self, step, stream=None):
        self.group.stage_push(step, stream)

    def stage_apply(self, step, stream=None):
        self.group.stage_apply(step, stream)

    def warm(self, n):
        self.group.warm(n)

    def _ensure_capacity(self, n):
        self.group._ensure_capacity(n)

    @property
    def cap(self):
        return self.group.cap

    # -------------------------------------------------------------- checkpoint
    def local_rows(self, what="weight"):
        """(global ids, rows [n, D]) of the real rows this rank owns — the unit of a
        sharded checkpoint (no cross-rank traffic)."""
        src = self.table if what == "weight" else self.slots[int(what)]
        gs, rows = [], []
        for g, l in self.layout.owner_chunks(0 if self.replicated else self.rank):
            gs.append(g)
            rows.append(src[l.to(self.device), :self.D].cpu())
        if not gs:
            return torch.zeros(0, dtype=torch.int64), torch.zeros(0, self.D)
        return torch.cat(gs), torch.cat(rows)

    def _gather_full(self, local):
        L, W = self.layout, self.world
        local = local[:, :self.D].contiguous()
        out = torch.zeros(self.V, self.D)
        if self.replicated or W == 1:
            g, l = L.global_ids_of_owner(0 if self.replicated else self.rank)
            out[g] = local.cpu()[l]
            return out
        shards = self.comm.all_gather_tensors(local)
        for o in range(W):
            g, l = L.global_ids_of_owner(o)
            out[g] = shards[o].cpu()[l]
        return out

    def full_weight(self):
        torch.cuda.synchronize(self.device)
        return self._gather_full(self.table)

    def full_slots(self):
        torch.cuda.synchronize(self.device)
        return [self._gather_full(s) for s in self.slots]

    def load_full(self, weight, slots=None):
        for g, l in self.layout.owner_chunks(0 if self.replicated else self.rank):
            l = l.to(self.device)
            self.table[l, :self.D] = weight.float()[g].to(self.device)
            if slots is not None:
                for s, full in zip(self.slots, slots):
                    s[l, :self.D] = full.float()[g].to(self.device)
        self.refresh_shadow()
        torch.cuda.synchronize(self.device)

    def load_rows(self, ids, rows, what="weight"):
        """Scatter (global id, row) pairs into this rank's shard; ids owned by other
        ranks are ignored (sharded-checkpoint restore, any source layout)."""
        L = self.layout
        ids = ids.to(torch.int64)
        own = torch.ones_like(ids, dtype=torch.bool) if self.replicated else \
            (L.owner_of(ids) == self.rank)
        if own.any():
            l = L.local_row_of(ids[own]).to(self.device)
            dst = self.table if what == "weight" else self.slots[int(what)]
            dst[l, :self.D] = rows[own].float().to(self.device)

    def release(self):
        """Free this table's symmetric segments (collective)."""
        torch.cuda.synchronize(self.device)
        if self.comm.distributed:
            self.comm.barrier()
        if self.group is not None:
            self.group.release_shared()
        for b in [self.tab_buf, self.shadow_buf, self.ring_buf] + list(self.slot_bufs):
            if b is not None:
                self.heap.free(b)
        self.table = self.shadow = None
        self.slots = []


class NVSparseGroup(object):
    """Tables with one placement that are looked up with the same ids."""
    _seq = 0

    def __init__(self, tables, name=None):
        from .. import ops
        L = _lib()
        gmax = int(L.px_sparse_group_max())
        if not 1 <= len(tables) <= gmax:
            raise ValueError("a co-lookup group holds 1..%d tables" % gmax)
        t0 = tables[0]
        for t in tables[1:]:
            if not t.layout.same_placement(t0.layout):
                raise ValueError(
                    "co-lookup group: %r and %r differ in rows / partitions / strategy / "
                    "owner placement" % (t0.name, t.name))
            if (t.kind in _optim.KINDS) != (t0.kind in _optim.KINDS):
                raise ValueError("co-lookup group mixes optimizer families")
        self.tables = list(tables)
        self.name = name or "+".join(t.name for t in tables)
        for t in tables:
            t.group = self
        self.fabric, self.heap, self.comm = t0.fabric, t0.heap, t0.comm
        self.rank, self.world, self.device = t0.rank, t0.world, t0.device
        self.route, self.layout = t0.route, t0.layout
        self.replicated = t0.replicated
        self.local_aggregation = t0.local_aggregation
        self.boundary = t0.boundary
        opts = t0.options
        # PSConfig.protocol == "nccl": the in-engine library baseline — same engine, same
        # kernels for the optimizer, but every byte that crosses GPUs goes through NCCL
        # collectives (Horovod's IndexedSlices path: all-gather of ids and rows)
        self.protocol = opts.get("_protocol", "nvlink")
        self.max_blocks = int(opts.get("sparse_blocks", 148 * 2))
        self.early_push = bool(opts.get("sparse_early_push", True))
        hints = [t.capacity_hint for t in tables if t.capacity_hint]
        self.capacity_hint = max(hints) if hints else None
        self.hp = hp_stage(self.fabric, t0.optimizer)
        lay = self.layout
        self._owners_dev = torch.tensor(lay.owners, dtype=torch.int32, device=self.device)
        self._slots_dev = torch.tensor(lay.slots, dtype=torch.int32, device=self.device)
        g = ops.PxGroupGeom()
        g.V, g.P, g.W, g.rows_per_part = lay.V, lay.P, lay.world, lay.rows_per_part
        g.strategy = 0 if lay.strategy == "mod" else 1
        g.replicated = 1 if lay.replicated else 0
        g.extras, g.base = getattr(lay, "_extras", 0), getattr(lay, "_base", 0)
        g.part_owner = self._owners_dev.data_ptr()
        g.part_slot = self._slots_dev.data_ptr()
        self.geom = g
        self.ctl = torch.zeros(int(L.px_sparse_ctl_bytes()) // 4, dtype=torch.int32,
                               device=self.device)
        self._t_off = int(L.px_sparse_ctl_time_offset())
        self._ovf_off = int(L.px_sparse_ctl_overflow_offset())
        self.hdr_buf = self.heap.alloc(int(L.px_sparse_hdr_words()) * 4, "hdr:" + self.name)
        self.ids_buf = None
        self.slotmap = torch.full((lay.rows_local,), -1, dtype=torch.int32,
                                  device=self.device)
        self.next = None
        self.cap = 0
        self.scratch_n = 0
        self.wire_dtype = None
        self._hdrs_dev = self._ids_dev = None
        self.calls = []            # (pend ids, [grad rows per table]) per lookup this step
        self._fwd_calls = self._bwd_calls = 0
        self._cur_step = 0
        self._done_step = -1
        self._last_n = 1
        NVSparseGroup._seq += 1

    # ---------------------------------------------------------------- forward
    # ------------------------------------------------- library-collective (NCCL) arm
    def _nccl(self):
        return (self.protocol == "nccl" and self.world > 1 and self.comm.distributed and
                self.route.sync)

    def _place(self, ids):
        """(valid, owner, local row) of global ids — device tensor arithmetic."""
        lay = self.layout
        valid = (ids >= 0) & (ids < lay.V)
        idc = ids.clamp(0, lay.V - 1).to(torch.int64)
        if lay.replicated:
            return valid, torch.zeros_like(idc), idc
        if lay.strategy == "mod":
            p, idx = idc % lay.P, idc // lay.P
        else:
            thr = lay._extras * (lay._base + 1)
            p = torch.where(idc < thr, idc // (lay._base + 1),
                            (idc - lay._extras) // max(lay._base, 1))
            start = torch.where(p < lay._extras, p * (lay._base + 1),
                                p * lay._base + lay._extras)
            idx = idc - start
        owner = self._owners_dev[p].to(torch.int64)
        local = self._slots_dev[p].to(torch.int64) * lay.rows_per_part + idx
        return valid, owner, local

    def _lookup_nccl(self, members, ids, n, record):
        import torch.distributed as dist
        W, grp = self.world, self.comm.group
        ids64 = ids.to(torch.int64)
        all_ids = torch.empty(W * n, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(all_ids, ids64, group=grp)
        valid, owner, local = self._place(all_ids)
        mine = valid & (owner == self.rank)
        rows = torch.where(mine, local, torch.zeros_like(local))
        outs = []
        for t in members:
            src = t.shadow[:, :t.Dp] if t.use_shadow else t.table
            part = src.index_select(0, rows).to(t.out_dtype) * mine[:, None].to(t.out_dtype)
            out = torch.empty(n, t.Dp, dtype=t.out_dtype, device=self.device)
            dist.reduce_scatter_tensor(out, part, group=grp)
            outs.append(out if t.Dp == t.D else out[:, :t.D])
        pend = None
        if record:
            v0, _, _ = self._place(ids64)
            pend = torch.where(v0, ids64, torch.full_like(ids64, -1)).to(torch.int32)
            self._fwd_calls += 1
        _count()
        return outs, pend

    def _push_apply_nccl(self, pend_ids, grads, n, cs):
        """all-gather (ids, rows) from every rank, then the owner kernel applies the rows
        this rank owns (entries of other owners are marked -1)."""
        import torch.distributed as dist
        from .. import ops
        L = _lib()
        W, grp = self.world, self.comm.group
        nt = len(self.tables)
        with torch.cuda.stream(cs):
            all_ids = torch.empty(W * n, dtype=torch.int32, device=self.device)
            dist.all_gather_into_tensor(all_ids, pend_ids, group=grp)
            valid, owner, local = self._place(all_ids.to(torch.int64))
            mine = valid if self.replicated else (valid & (owner == self.rank))
            ring_ids = torch.where(mine, local, torch.full_like(local, -1)).to(torch.int32)
            nxt = torch.empty(W * n, dtype=torch.int32, device=self.device)
            descs = (ops.PxOwnerTable * nt)()
            keep = [all_ids, ring_ids, nxt]
            for d, t, g in zip(descs, self.tables, grads):
                all_g = torch.empty(W * n, t.Dp, dtype=g.dtype, device=self.device)
                dist.all_gather_into_tensor(all_g, g, group=grp)
                keep.append(all_g)
                d.ring, d.table = all_g.data_ptr(), t.table.data_ptr()
                d.slot0 = t.slots[0].data_ptr() if t.nslots > 0 else 0
                d.slot1 = t.slots[1].data_ptr() if t.nslots > 1 else 0
                d.slot2 = t.slots[2].data_ptr() if t.nslots > 2 else 0
                d.shadow = t.shadow.data_ptr() if t.use_shadow else 0
                d.hp, d.D4, d.kind = self.hp.dev.data_ptr(), t.D4, _optim.KIND_ID[t.kind]
                d.avg = ((1.0 / W) if t.average else 1.0) * t.scale
            if not torch.cuda.is_current_stream_capturing():
                for k_ in keep:
                    k_.record_stream(cs)
            self._keep_n = (keep, descs)
            blocks = max(1, min(self.max_blocks, (n * W + 7) // 8))
            _count()
            ops.check(L.px_sparse_owner(
                descs, nt, _DT[grads[0].dtype], _vp(ring_ids.data_ptr()),
                _vp(self.hdr_buf.local_ptr), _vp(self.hdrs_dev.data_ptr()),
                _vp(self.slotmap.data_ptr()), _vp(nxt.data_ptr()), n,
                ctypes.byref(self.geom), _vp(self.ctl.data_ptr()), self.rank, 1, blocks, n,
                _sp(cs)), "sparse_owner(nccl)")

    def lookup(self, flat_ids, record=True, members=None):
        from .. import ops
        L = _lib()
        members = self.tables if members is None else members
        n = int(flat_ids.numel())
        ids = flat_ids if flat_ids.is_cuda else flat_ids.to(self.device, non_blocking=True)
        if ids.dtype not in (torch.int64, torch.int32):
            ids = ids.to(torch.int64)
        ids = ids.contiguous()
        if self._nccl() and not self.replicated and n > 0:
            return self._lookup_nccl(members, ids, n, record)
        descs = (ops.PxLookupTable * len(members))()
        outs = []
        for d, t in zip(descs, members):
            if t.use_shadow:
                out = torch.empty((n, t.Dps), dtype=torch.bfloat16, device=self.device)
                d.srcs, d.src_bf16, d.out_bf16 = t.dev_ptrs("shadow").data_ptr(), 1, 1
            else:
                out = torch.empty((n, t.Dp), dtype=t.out_dtype, device=self.device)
                d.srcs, d.src_bf16 = t.dev_ptrs("table").data_ptr(), 0
                d.out_bf16 = 1 if t.out_dtype == torch.bfloat16 else 0
            d.out, d.D4 = out.data_ptr(), t.D4
            outs.append(out if out.shape[1] == t.D else out[:, :t.D])
        pend = torch.empty(n, dtype=torch.int32, device=self.device) if record else None
        if record:
            self._fwd_calls += 1
        if n > 0:
            _count()
            ops.check(L.px_sparse_lookup(
                _vp(ids.data_ptr()), 1 if ids.dtype == torch.int64 else 0, n, descs,
                len(members), _vp(pend.data_ptr()) if pend is not None else _vp(0),
                ctypes.byref(self.geom), _vp(self.hdr_buf.local_ptr),
                _vp(self.ctl.data_ptr()),
                1 if (self.route.sync and self.world > 1 and not self._nccl()) else 0,
                _sp(torch.cuda.current_stream(self.device))), "sparse_lookup")
        return outs, pend

    def add_pending(self, token, grads):
        gs = []
        for t, g in zip(self.tables, grads):
            g = g.reshape(-1, t.D)
            if t.Dp != t.D:
                g = torch.nn.functional.pad(g, (0, t.Dp - t.D))
            if g.dtype not in _DT:
                g = g.float()
            gs.append(g.contiguous())
        dts = {g.dtype for g in gs}
        if len(dts) > 1:
            gs = [g.float() for g in gs]
        self.calls.append((token, gs))
        self._bwd_calls += 1
        if self.early_push and self._bwd_calls == self._fwd_calls and self._cur_step > 0 \
                and self.ring_ready():
            self._run_step(self._cur_step)

    def ring_ready(self):
        """Early push needs every lazy allocation done (first step runs at the end)."""
        if self._nccl():
            return True
        return self.scratch_n > 0 and (not self.route.sync or self.ids_buf is not None)

    # ----------------------------------------------------------------- capacity
    def _negotiated(self, n):
        """Largest per-rank row count of this step (Horovod negotiates allgather sizes on
        every cycle, `collective_operations.cc:80-90`).  With a `sparse_capacity` hint, in
        a simulated world or under stream capture nothing is exchanged."""
        if (not self.route.sync or self.world == 1 or not self.comm.distributed or
                self.capacity_hint or torch.cuda.is_current_stream_capturing()):
            return n
        return self.comm.all_reduce_max_int(n)

    def _ensure_capacity(self, n):
        dev = self.device
        if n > self.scratch_n:
            cap_n = max(int(n * 1.25) + 16, 64)
            for t in self.tables:
                # fp32 rows for ids carried by several positions (kept zero between steps
                # by the push kernel's flush pass)
                t.staging = torch.zeros(cap_n // 2 + 2, t.Dp, dtype=torch.float32, device=dev)
            self.scratch_n = cap_n
        if not self.route.sync:
            return
        m = self._negotiated(n)
        if self.ids_buf is not None and m <= self.cap:
            return
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError(
                "sparse group %s: %d gradient rows exceed the ring capacity %d inside a "
                "captured CUDA graph (shapes must be static under cuda_graph)" %
                (self.name, n, self.cap))
        want = int(self.capacity_hint or max(int(m * 1.5) + 16, 64))
        if self.capacity_hint and m > want:
            raise RuntimeError(
                "sparse group %s: %d gradient rows in one step exceed sess_config"
                "['sparse_capacity'] = %d" % (self.name, m, want))
        if self.ids_buf is not None:
            parallax_log.info("sparse group %s: growing receive rings %d -> %d rows/source",
                              self.name, self.cap, want)
            torch.cuda.synchronize(dev)
            if self.comm.distributed:
                self.comm.barrier()
            self.heap.free(self.ids_buf)
            for t in self.tables:
                self.heap.free(t.ring_buf)
                t._ptrs.pop("ring", None)
        self.cap = want
        W = self.world
        for t in self.tables:
            t.ring_buf = self.heap.alloc(W * self.cap * t.Dp * 4, "ring:" + t.name)
        self.ids_buf = self.heap.alloc(W * self.cap * 4, "ring_ids:" + self.name)
        self.next = torch.empty(W * self.cap, dtype=torch.int32, device=dev)
        self._ids_dev = None
        if self.comm.distributed:
            torch.cuda.synchronize(dev)
            self.comm.barrier()

    @property
    def hdrs_dev(self):
        if self._hdrs_dev is None:
            self._hdrs_dev = self.hdr_buf.dev_ptrs()
        return self._hdrs_dev

    @property
    def ids_dev(self):
        if self._ids_dev is None:
            self._ids_dev = self.ids_buf.dev_ptrs()
        return self._ids_dev

    def warm(self, n):
        """Allocate everything a step of `n` gradient rows needs (no lazy allocation /
        pointer upload will happen inside the step)."""
        self._ensure_capacity(n)
        for t in self.tables:
            t.dev_ptrs("table")
            for i in range(t.nslots):
                t.dev_ptrs("slot%d" % i)
            if t.use_shadow:
                t.dev_ptrs("shadow")
            if self.route.sync:
                t.dev_ptrs("ring")
        if self.route.sync:
            self.ids_dev, self.hdrs_dev

    # --------------------------------------------------------------------- step
    def begin_step(self, step):
        self.hp.upload(step)
        self._cur_step = step
        self._fwd_calls = self._bwd_calls = 0

    def finish_step(self, step, stream=None):
        if self._done_step == step and not self.calls:
            return                           # already pushed from the backward pass
        self._run_step(step, stream)

    def _run_step(self, step, stream=None):
        from ..utils import timeline
        self._done_step = step
        if self._nccl():
            cs = stream if stream is not None else self.fabric.comm_stream
            calls, self.calls = self.calls, []
            if not calls:
                raise RuntimeError("protocol='nccl': every rank must look the group %s up in "
                                   "every step (collective)" % self.name)
            nt = len(self.tables)
            pend_ids = calls[0][0] if len(calls) == 1 else torch.cat([c[0] for c in calls])
            grads = [calls[0][1][k] if len(calls) == 1 else
                     torch.cat([c[1][k] for c in calls]) for k in range(nt)]
            cur = torch.cuda.current_stream(self.device)
            if cs is not cur:
                cs.wait_stream(cur)
            self._push_apply_nccl(pend_ids, grads, int(pend_ids.numel()), cs)
            return
        if timeline.enabled():
            cs = stream if stream is not None else self.fabric.comm_stream
            with timeline.activity(self.name, "SPARSE_PUSH_APPLY", gpu=True, stream=cs,
                                   args="rows=%d" % sum(c[0].numel() for c in self.calls)):
                self.stage_push(step, stream)
                if self.route.sync:
                    self.stage_apply(step, stream)
        else:
            self.stage_push(step, stream)
            if self.route.sync:
                self.stage_apply(step, stream)

    def stage_push(self, step, stream=None):
        """Sender side, one kernel: local aggregation + push (or remote apply in async
        mode).  Separate from `stage_apply` so that a world simulated on one GPU can
        enqueue every rank's push before any rank's (spinning) owner kernel."""
        from .. import ops
        L = _lib()
        cs = stream if stream is not None else self.fabric.comm_stream
        calls, self.calls = self.calls, []
        nt = len(self.tables)
        if calls:
            pend_ids = calls[0][0] if len(calls) == 1 else torch.cat([c[0] for c in calls])
            grads = [calls[0][1][k] if len(calls) == 1 else
                     torch.cat([c[1][k] for c in calls]) for k in range(nt)]
        else:
            pend_ids = torch.empty(0, dtype=torch.int32, device=self.device)
            grads = [torch.empty((0, t.Dp), dtype=torch.float32, device=self.device)
                     for t in self.tables]
        n = int(pend_ids.numel())
        self._ensure_capacity(max(n, 1))
        self._last_n = n
        for t in self.tables:
            t.stats["pushed_rows"] += n
            t.stats["steps"] += 1
        gdt = grads[0].dtype
        sync = self.route.sync
        if sync:
            if self.wire_dtype is None:
                # fixed for the lifetime of the rings: bf16 gradients stay bf16 on the wire
                # when the boundary optimisation is on, everything else travels as fp32
                self.wire_dtype = torch.bfloat16 if (gdt == torch.bfloat16 and self.boundary) \
                    else torch.float32
            if self.wire_dtype == torch.bfloat16 and gdt != torch.bfloat16:
                grads = [g.to(torch.bfloat16) for g in grads]   # dtype changed mid-run
                gdt = torch.bfloat16
        cur = torch.cuda.current_stream(self.device)
        if cs is not cur:
            cs.wait_stream(cur)
            if not torch.cuda.is_current_stream_capturing():
                pend_ids.record_stream(cs)
                for g in grads:
                    g.record_stream(cs)
        descs = (ops.PxPushTable * nt)()
        for d, t, g in zip(descs, self.tables, grads):
            d.grads, d.staging = g.data_ptr(), t.staging.data_ptr()
            d.hp, d.D4, d.kind = self.hp.dev.data_ptr(), t.D4, _optim.KIND_ID[t.kind]
            if sync:
                d.rings = t.dev_ptrs("ring").data_ptr()
                d.scale = t.scale if self.boundary else 1.0
            else:
                d.tables = t.dev_ptrs("table").data_ptr()
                d.slot0s = t.dev_ptrs("slot0").data_ptr() if t.nslots > 0 else 0
                d.slot1s = t.dev_ptrs("slot1").data_ptr() if t.nslots > 1 else 0
                d.slot2s = t.dev_ptrs("slot2").data_ptr() if t.nslots > 2 else 0
                d.shadows = t.dev_ptrs("shadow").data_ptr() if t.use_shadow else 0
                d.scale = t.scale
        self._keep = (pend_ids, grads, descs)
        _count()
        ops.check(L.px_sparse_push(
            _vp(pend_ids.data_ptr()), n, descs, nt, _DT[gdt],
            _DT[self.wire_dtype] if sync else 0, 0 if sync else 1,
            _vp(self.ids_dev.data_ptr()) if sync else _vp(0),
            _vp(self.hdrs_dev.data_ptr()) if sync else _vp(0), self.cap,
            ctypes.byref(self.geom), _vp(self.ctl.data_ptr()), self.rank,
            1 if self.local_aggregation else 0, self.max_blocks, _sp(cs)), "sparse_push")

    def stage_apply(self, step, stream=None):
        """Owner side, one kernel: merge rows from all sources, apply the optimizer."""
        from .. import ops
        L = _lib()
        cs = stream if stream is not None else self.fabric.comm_stream
        n = max(self._last_n, 1)
        nt = len(self.tables)
        # 16 half-warps per CTA, one entry per half-warp: as many CTAs as fit on the device at once
        # (4 per SM at 64 registers; the merge variant is a cooperative launch)
        blocks = max(1, min(max(self.max_blocks, 148 * 4) if self.max_blocks >= 148
                            else self.max_blocks,
                            (n * (self.world if self.replicated else 1) + 15) // 16))
        use_merge = self.world > 1 or not self.local_aggregation
        descs = (ops.PxOwnerTable * nt)()
        for d, t in zip(descs, self.tables):
            d.ring, d.table = t.ring_buf.local_ptr, t.table.data_ptr()
            d.slot0 = t.slots[0].data_ptr() if t.nslots > 0 else 0
            d.slot1 = t.slots[1].data_ptr() if t.nslots > 1 else 0
            d.slot2 = t.slots[2].data_ptr() if t.nslots > 2 else 0
            d.shadow = t.shadow.data_ptr() if t.use_shadow else 0
            d.hp, d.D4, d.kind = self.hp.dev.data_ptr(), t.D4, _optim.KIND_ID[t.kind]
            avg = (1.0 / self.world) if t.average else 1.0
            if not self.boundary:
                avg *= t.scale
            d.avg = avg
        self._keep_o = descs
        _count()
        ops.check(L.px_sparse_owner(
            descs, nt, _DT[self.wire_dtype], _vp(self.ids_buf.local_ptr),
            _vp(self.hdr_buf.local_ptr), _vp(self.hdrs_dev.data_ptr()),
            _vp(self.slotmap.data_ptr()), _vp(self.next.data_ptr()), self.cap,
            ctypes.byref(self.geom), _vp(self.ctl.data_ptr()), self.rank,
            1 if use_merge else 0, blocks, -1, _sp(cs)), "sparse_owner")

    # ---------------------------------------------------------------- inspection
    def device_times(self):
        """%globaltimer stamps (ns) written by the last push / owner kernels:
        push start, pushed flag published, owner start, all sources arrived, applied
        published.  Valid under CUDA-graph replay (the kernels write them every run)."""
        raw = self.ctl.view(torch.uint8)[self._t_off:self._t_off + 104].clone() \
            .view(torch.int64).tolist()
        d = dict(zip(("push_start", "pushed", "owner_start", "arrived", "applied"), raw))
        d["push_phases"] = raw[5:13]        # CTA 0 of the push kernel: end of each phase
        return d

    def overflow_count(self):
        return int(self.ctl.view(torch.uint8)[self._ovf_off:self._ovf_off + 4]
                   .clone().view(torch.int32).item())

    def release_shared(self):
        for b in (self.hdr_buf, self.ids_buf):
            if b is not None:
                self.heap.free(b)
        self.hdr_buf = self.ids_buf = None
