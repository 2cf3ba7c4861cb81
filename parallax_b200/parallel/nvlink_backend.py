"""NVLink fabric backend: the engine's production path on B200.

Dense variables → `NVDenseGroup` (buckets in symmetric memory, fused
reduce-scatter + optimizer + parameter all-gather kernel per bucket, launched
from autograd hooks on a dedicated comm stream so it overlaps backward).
Sparse variables → `NVSparseTable` (row-partitioned shard in symmetric
memory; remote-gather lookup; dedup + P2P push; owner-side claim/apply).

Static schedule: every kernel that synchronises with peers is issued on ONE
comm stream in an order fixed at build time (buckets in index order, then
tables in name order) — identical on all ranks, so no cross-rank wait cycle
can form and no per-step negotiation is needed (what Horovod's coordinator
does every 5 ms tick, `horovod/common/operations.cc:1274-1590`).
"""
import ctypes
import math

import torch

from .. import optim as _optim
from ..log import parallax_log
from . import modes, nvops
from .layout import TableLayout
from .symmetric import (SymmetricHeap, IpcExchange, CH_COMM, CH_MAIN, CH_SMALL)

MODE_FUSED, MODE_REDUCE, MODE_UPDATE = 0, 1, 2
_ES = {torch.float32: 4, torch.bfloat16: 2}


class NVFabric(object):
    def __init__(self, comm, exchange=None, options=None):
        self.comm = comm
        self.device = comm.device
        self.rank, self.world = comm.rank, comm.world
        self.options = options or {}
        ex = exchange if exchange is not None else IpcExchange(comm)
        if exchange is None and comm.distributed:
            import socket
            hosts = set(comm.all_gather_object(socket.gethostname()))
            if len(hosts) > 1:
                raise RuntimeError(
                    "the NVLink fabric addresses peers through CUDA IPC / NVSwitch "
                    "multicast and therefore spans ONE NVLink domain (one HGX/DGX box); "
                    "this job spans hosts %s. Run one job per box, or use "
                    "sess_config={'fabric': 'library'} (same engine on torch.distributed/"
                    "NCCL collectives) across boxes."
                    % sorted(hosts))
        self.heap = SymmetricHeap(self.device, ex)
        self.comm_stream = torch.cuda.Stream(self.device, priority=-1)
        # double-buffered staging for the one-shot all-reduce (norms, scalars)
        self.small_stage = self.heap.alloc(2 * 65536, "oneshot_stage")
        self.max_blocks = int(self.options.get("comm_blocks", 32))
        if isinstance(ex, IpcExchange):
            self.heap.pads_dev()        # eager: no lazy H2D inside a step
        if comm.distributed:
            comm.barrier()

    def close(self):
        torch.cuda.synchronize(self.device)
        if self.comm.distributed:
            self.comm.barrier()
        self.heap.close()


class _Bucket(object):
    pass


class NVDenseGroup(object):
    def __init__(self, named_params, optimizer, fabric, route, graph,
                 options=None):
        self.fabric, self.route, self.graph = fabric, route, graph
        self.optimizer = optimizer
        self.heap = fabric.heap
        self.rank, self.world = fabric.rank, fabric.world
        self.device = fabric.device
        opts = options or {}
        self.options = opts
        # bucket size: sess_config["bucket_bytes"], else PARALLAX_FUSION_THRESHOLD (Horovod's
        # HOROVOD_FUSION_THRESHOLD, `operations.cc:1030`), else 32 MiB
        import os
        from .. import consts
        self.bucket_bytes = int(opts.get("bucket_bytes") or
                                os.environ.get(consts.PARALLAX_FUSION_THRESHOLD) or (32 << 20))
        self.update = opts.get("dense_update", "sharded")   # or "replicated"
        if not route.sync:
            self.update = "async"
        # PSConfig.protocol == "nccl": library fallback for the dense reduction
        # (the in-engine baseline); every other protocol value = NVLink kernels.
        self.protocol = opts.get("_protocol", "nvlink")
        if self.protocol == "nccl" and route.sync:
            self.update = "replicated"
        # PSConfig.replicate_variables (PS run option): True = owners push the
        # updated values into every mirror right after the update; False =
        # workers pull the owners' slices at the start of the next step.
        self.pull_mirrors = (route.run_option == "PS" and route.sync and
                             not opts.get("_replicate_variables", True) and
                             self.update == "sharded" and self.world > 1)
        self.names = [n for n, _ in named_params]
        self.params = [p for _, p in named_params]
        _optim.require_fused(optimizer.kind, "NVLink fabric")
        self.kind = optimizer.kind
        self.nslots = _optim.NUM_SLOTS[self.kind]
        self.clip_rules = graph.clip_rules()
        self.ema_rule = graph.ema
        self.last_grad_norm = {}
        self.hp_host = torch.zeros(_optim.HP_SIZE, dtype=torch.float32).pin_memory()
        self.hp = torch.zeros(_optim.HP_SIZE, dtype=torch.float32, device=self.device)
        self._build_buckets()
        self._install_hooks()
        self._next = 0
        self._clip_pending = {}

    # ---------------------------------------------------------------- build
    def _clip_index(self, name):
        for i, r in enumerate(self.clip_rules):
            if r.applies_to(name):
                return i
        return -1

    def _build_buckets(self):
        W = self.world
        open_b, buckets = {}, []
        for name, p in reversed(list(zip(self.names, self.params))):
            assert p.dtype in _ES, "unsupported parameter dtype %s" % p.dtype
            key = (p.dtype, self._clip_index(name))
            b = open_b.get(key)
            if b is None:
                b = _Bucket()
                b.dtype, b.clip = key
                b.items, b.n = [], 0
                open_b[key] = b
                buckets.append(b)
            vn = 16 // _ES[p.dtype]
            off = b.n
            b.items.append((name, p, off, p.numel()))
            b.n = off + (p.numel() + vn - 1) // vn * vn
            if b.n * _ES[p.dtype] >= self.bucket_bytes:
                del open_b[key]
        self.buckets = buckets
        heap = self.heap
        # NVLS: bucket buffers bound to an NVSwitch multicast object, the fused
        # kernel then reduces with multimem.ld_reduce and broadcasts parameters
        # with multimem.st.  Measured (profiles/allreduce_sweep_8gpu.json): wins
        # from 256 KB up at 8 GPUs, loses at 2.  "auto" enables it on the
        # configuration it was validated and measured on (a full 8-GPU box; the
        # end-to-end gain there is ~1 %); `dense_nvls=True` forces it for 4..7.
        want = self.options.get("dense_nvls", "auto")
        self.nvls = False
        if W > 1 and self.update == "sharded" and not self.pull_mirrors and want:
            from . import multicast
            if want is True or (want == "auto" and W >= 8):
                try:
                    self.nvls = multicast.supported(self.fabric.comm)
                except Exception:
                    self.nvls = False
        for bi, b in enumerate(buckets):
            vn = 16 // _ES[b.dtype]
            quantum = W * vn * 32
            b.index = bi
            b.n = (b.n + quantum - 1) // quantum * quantum
            es = _ES[b.dtype]
            b.mc = False
            if self.nvls:
                from . import multicast
                try:
                    b.grad_buf = multicast.MulticastBuffer(self.fabric, b.n * es)
                    b.param_buf = multicast.MulticastBuffer(self.fabric, b.n * es)
                    b.mc = True
                except multicast.MulticastError as e:
                    # MulticastBuffer agrees on success/failure across ranks, so
                    # every rank takes this branch together
                    parallax_log.warning("NVLS unavailable (%s); using P2P kernels", e)
                    self.nvls = False
            if not b.mc:
                b.grad_buf = heap.alloc(b.n * es, "grad%d" % bi)
                b.param_buf = heap.alloc(b.n * es, "param%d" % bi)
            b.grad_flat = b.grad_buf.tensor(b.dtype, b.n)
            b.param_flat = b.param_buf.tensor(b.dtype, b.n)
            b.grad_views, b.scales = [], []
            with torch.no_grad():
                for name, p, off, numel in b.items:
                    b.param_flat[off:off + numel].copy_(p.detach().reshape(-1))
                    p.data = b.param_flat[off:off + numel].view(p.shape)
                    b.grad_views.append(b.grad_flat[off:off + numel].view(p.shape))
                    b.scales.append(self.graph.scale_for(name))
            b.need_scale = any(s != 1.0 for s in b.scales)
            b.pending = [None] * len(b.items)
            b.ready = 0
            b.launched = False
            b.event = torch.cuda.Event()
        # every replica starts from rank 0's values
        if W > 1:
            torch.cuda.synchronize(self.device)
            for b in buckets:
                if b.mc:      # bootstrap only: no peer unicast mappings on NVLS buffers
                    self.fabric.comm.broadcast_(b.param_flat, 0)
                else:
                    nvops.broadcast(heap, b.param_buf.c_ptrs(), b.n * _ES[b.dtype], 0,
                                    CH_MAIN, self.fabric.max_blocks)
            torch.cuda.synchronize(self.device)
        for b in buckets:
            self._alloc_state(b)
        # clip-rule scalars
        self.clip_state = {}
        for ci in sorted({b.clip for b in buckets if b.clip >= 0}):
            st = _Bucket()
            st.local = torch.zeros(4, dtype=torch.float32, device=self.device)
            st.total = torch.zeros(4, dtype=torch.float32, device=self.device)
            st.scale = torch.ones(1, dtype=torch.float32, device=self.device)
            st.norm = torch.zeros(1, dtype=torch.float32, device=self.device)
            st.buckets = [b for b in buckets if b.clip == ci]
            self.clip_state[ci] = st

    def _alloc_state(self, b):
        W, dev = self.world, self.device
        sharded = self.update == "sharded"
        b.slice = b.n // W if (sharded or self.update == "async") else b.n
        lo = self.rank * b.slice if b.slice != b.n else 0
        init = self.optimizer.slot_init()

        def mk(fill=None, src=None):
            if self.update == "async":
                sb = self.heap.alloc(b.slice * 4, "state")
                t = sb.tensor(torch.float32, b.slice)
                t._symm = sb
            else:
                t = torch.empty(b.slice, dtype=torch.float32, device=dev)
                t._symm = None
            if src is not None:
                t.copy_(src)
            else:
                t.fill_(fill)
            return t
        b.master = mk(src=b.param_flat[lo:lo + b.slice].float())
        b.slots = [mk(fill=v) for v in init]
        b.ema = None
        if self.ema_rule is not None and any(
                self.ema_rule.applies_to(n) for n, _, _, _ in b.items):
            b.ema = mk(src=b.master)
        b.red = torch.empty(b.slice, dtype=torch.float32, device=dev) \
            if (b.clip >= 0 and sharded) else None

    def _install_hooks(self):
        for b in self.buckets:
            for idx, (name, p, off, numel) in enumerate(b.items):
                p.register_post_accumulate_grad_hook(self._make_hook(b, idx))

    def _make_hook(self, b, idx):
        def hook(p):
            b.pending[idx] = p.grad
            p.grad = None
            b.ready += 1
            if b.ready == len(b.items):
                self._bucket_ready(b)
        return hook

    # ----------------------------------------------------------------- step
    def begin_step(self, step):
        hp = self.optimizer.hyper(step)
        for i, v in enumerate(hp):
            self.hp_host[i] = v
        self.hp.copy_(self.hp_host, non_blocking=True)
        if self.pull_mirrors and step > 1:
            # mirror refresh deferred to "first use": all-gather of the owners'
            # parameter slices before the forward pass
            cur = torch.cuda.current_stream(self.device)
            for b in self.buckets:
                nvops.allgather(self.heap, b.param_buf.c_ptrs(),
                                (b.n // self.world) * _ES[b.dtype], CH_MAIN,
                                self.fabric.max_blocks, stream=cur)

    def _bucket_ready(self, b):
        views, grads = [], []
        for v, g in zip(b.grad_views, b.pending):
            if g is None:
                v.zero_()
            else:
                views.append(v)
                grads.append(g if g.dtype == v.dtype else g.to(v.dtype))
        if grads:
            torch._foreach_copy_(views, grads)
            nvops._count(1)
        if b.need_scale:
            for v, s in zip(b.grad_views, b.scales):
                if s != 1.0:
                    v.mul_(s)
        b.pending = [None] * len(b.items)
        b.event.record(torch.cuda.current_stream(self.device))
        b.is_ready = True
        self._drain()

    def _drain(self):
        """Launch consecutive ready buckets in index order on the comm stream."""
        while self._next < len(self.buckets) and \
                getattr(self.buckets[self._next], "is_ready", False):
            b = self.buckets[self._next]
            self._launch(b)
            self._next += 1

    def _launch(self, b):
        from ..utils import timeline
        if timeline.enabled():
            with timeline.activity("bucket%d" % b.index, "DENSE_STEP", gpu=True,
                                   stream=self.fabric.comm_stream,
                                   args="%s n=%d %s" % (self.update, b.n, self.kind)):
                self._launch_impl(b)
        else:
            self._launch_impl(b)

    def _launch_impl(self, b):
        fab, heap, cs = self.fabric, self.heap, self.fabric.comm_stream
        cs.wait_event(b.event)
        W = self.world
        mb = fab.max_blocks
        ema_decay = self.ema_rule.decay if self.ema_rule is not None else 0.0
        s0 = b.slots[0] if self.nslots > 0 else None
        s1 = b.slots[1] if self.nslots > 1 else None
        st = self.clip_state.get(b.clip)
        if self.update == "sharded":
            if st is None:
                nvops.dense_step(heap, self._grad_sources(b), self._param_targets(b),
                                 b.master, s0, s1, b.ema, None, self.hp, None,
                                 None, b.n, 1.0 / W, ema_decay, self.kind,
                                 MODE_FUSED, b.dtype, CH_COMM, max_blocks=mb,
                                 stream=cs, use_mc=b.mc)
            else:
                nvops.dense_step(heap, self._grad_sources(b), self._param_targets(b),
                                 b.master, s0, s1, b.ema, b.red, self.hp, None,
                                 st.local, b.n, 1.0 / W, ema_decay, self.kind,
                                 MODE_REDUCE, b.dtype, CH_COMM, max_blocks=mb,
                                 stream=cs, use_mc=b.mc)
                if b is st.buckets[-1]:
                    self._finish_clip(st, cs)
                    for bb in st.buckets:
                        t0 = bb.slots[0] if self.nslots > 0 else None
                        t1 = bb.slots[1] if self.nslots > 1 else None
                        nvops.dense_step(heap, self._grad_sources(bb),
                                         self._param_targets(bb), bb.master, t0, t1,
                                         bb.ema, bb.red, self.hp, st.scale, None,
                                         bb.n, 1.0 / W, ema_decay, self.kind,
                                         MODE_UPDATE, bb.dtype, CH_COMM,
                                         max_blocks=mb, stream=cs, use_mc=bb.mc)
        elif self.update == "replicated":
            # classic AR: all-reduce (mean) then every replica updates itself
            if self.protocol == "nccl" and W > 1:
                import torch.distributed as dist
                with torch.cuda.stream(cs):
                    dist.all_reduce(b.grad_flat, group=self.fabric.comm.group)
                    b.grad_flat.mul_(1.0 / W)
                    if st is not None:
                        # slice-local Σg² so the cross-rank sum equals the global norm²
                        sl = b.n // W
                        st.local[0] += b.grad_flat[self.rank * sl:(self.rank + 1) * sl] \
                            .float().pow(2).sum()
            else:
                nvops.allreduce_twoshot(heap, b.grad_buf.c_ptrs(), b.n, b.dtype,
                                        1.0 / W, CH_COMM,
                                        sumsq=st.local if st is not None else None,
                                        max_blocks=mb, stream=cs)
            if st is None:
                self._local_update(b, None, cs)
            elif b is st.buckets[-1]:
                self._finish_clip(st, cs)
                for bb in st.buckets:
                    self._local_update(bb, st.scale, cs)
        else:  # async PS
            clip = None
            if st is not None:
                nvops.sumsq(b.grad_flat, b.n, b.dtype, 1.0, st.local, stream=cs)
                if b is st.buckets[-1]:
                    nvops.clip_scale(st.local, self.clip_rules[b.clip].max_norm,
                                     st.scale, st.norm, st.local, stream=cs)
                    for bb in st.buckets:
                        self._async_update(bb, st.scale, cs)
            else:
                self._async_update(b, clip, cs)
        b.launched = True

    def _grad_sources(self, b):
        return b.grad_buf.mc_c_ptrs() if b.mc else b.grad_buf.c_ptrs()

    def _param_targets(self, b):
        """Where the fused kernel stores updated parameters: every peer's mirror
        (push; one multimem.st on NVLS buffers) or only the local buffer (pull
        mode; peers fetch later)."""
        if b.mc:
            return b.param_buf.mc_c_ptrs()
        if not self.pull_mirrors:
            return b.param_buf.c_ptrs()
        arr = getattr(b, "_self_targets", None)
        if arr is None:
            arr = (ctypes.c_void_p * self.world)(*([b.param_buf.local_ptr] * self.world))
            b._self_targets = arr
        return arr

    def _finish_clip(self, st, cs):
        rule = self.clip_rules[st.buckets[0].clip]
        if self.world > 1:
            nvops.allreduce_oneshot(self.heap, st.local, st.total,
                                    self.fabric.small_stage, 4, torch.float32,
                                    1.0, CH_SMALL, stream=cs)
            nvops.clip_scale(st.total, rule.max_norm, st.scale, st.norm, st.local,
                             stream=cs)
        else:
            nvops.clip_scale(st.local, rule.max_norm, st.scale, st.norm, st.local,
                             stream=cs)

    def _local_update(self, b, clip, cs):
        ema_decay = self.ema_rule.decay if self.ema_rule is not None else 0.0
        s0 = b.slots[0] if self.nslots > 0 else None
        s1 = b.slots[1] if self.nslots > 1 else None
        g = (ctypes.c_void_p * 1)(b.grad_buf.local_ptr)
        p = (ctypes.c_void_p * 1)(b.param_buf.local_ptr)
        b._keep = (g, p)
        nvops.dense_step(self.heap, g, p, b.master, s0, s1, b.ema, None, self.hp,
                         clip, None, b.n, 1.0, ema_decay, self.kind, MODE_FUSED,
                         b.dtype, CH_COMM, rank=0, world=1, stream=cs)

    def _async_update(self, b, clip, cs):
        W = self.world
        mc = b.master._symm.c_ptrs()
        s0 = b.slots[0]._symm.c_ptrs() if self.nslots > 0 else None
        s1 = b.slots[1]._symm.c_ptrs() if self.nslots > 1 else None
        nvops.dense_async(b.grad_flat, b.param_flat, mc, s0, s1, self.hp, clip,
                          b.n, self.kind, b.dtype, self.rank, W,
                          max_blocks=self.fabric.max_blocks * 2, stream=cs)

    def finish_step(self, step):
        # buckets whose parameters received no gradient this step
        for b in self.buckets:
            if not getattr(b, "is_ready", False):
                self._bucket_ready(b)
        self._drain()
        assert self._next == len(self.buckets)
        torch.cuda.current_stream(self.device).wait_stream(self.fabric.comm_stream)
        for b in self.buckets:
            b.ready, b.is_ready, b.launched = 0, False, False
        self._next = 0

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    # ----------------------------------------------------------- checkpoint
    def _gather_full(self, attr, idx=None):
        """name -> full fp32 tensor (CPU) reconstructed from slices."""
        out = {}
        comm = self.fabric.comm
        for b in self.buckets:
            t = getattr(b, attr)
            if idx is not None:
                t = t[idx] if t is not None and idx < len(t) else None
            if t is None:
                continue
            if b.slice != b.n and self.world > 1:
                full = torch.cat(comm.all_gather_tensors(t.contiguous()))
            else:
                full = t
            full = full.detach().float().cpu()
            for name, p, off, numel in b.items:
                out[name] = full[off:off + numel].view(p.shape).clone()
        return out

    def state_dict(self):
        torch.cuda.synchronize(self.device)
        sd = {"master": self._gather_full("master"), "slots": {}, "ema": {}}
        per_slot = [self._gather_full("slots", i) for i in range(self.nslots)]
        for n in self.names:
            sd["slots"][n] = [ps[n] for ps in per_slot]
        if self.ema_rule is not None:
            full = self._gather_full("ema")
            sd["ema"] = {n: v for n, v in full.items()
                         if self.ema_rule.applies_to(n)}
        return sd

    def load_state_dict(self, sd):
        torch.cuda.synchronize(self.device)
        for b in self.buckets:
            lo = self.rank * b.slice if b.slice != b.n else 0
            hi = lo + b.slice

            def fill(dst, src_by_name):
                full = torch.zeros(b.n, dtype=torch.float32)
                have = dst.detach().float().cpu()
                full[lo:hi] = have
                for name, p, off, numel in b.items:
                    if name in src_by_name:
                        full[off:off + numel] = src_by_name[name].reshape(-1).float()
                dst.copy_(full[lo:hi].to(self.device))
                return full
            full = fill(b.master, sd["master"])
            # parameters (all of them, not just my slice)
            pf = b.param_flat.detach().float().cpu()
            for name, p, off, numel in b.items:
                if name in sd["master"]:
                    pf[off:off + numel] = sd["master"][name].reshape(-1).float()
            b.param_flat.copy_(pf.to(self.device).to(b.dtype))
            for i in range(self.nslots):
                fill(b.slots[i], {n: v[i] for n, v in sd["slots"].items()
                                  if len(v) > i})
            if b.ema is not None:
                fill(b.ema, sd.get("ema", {}))
        torch.cuda.synchronize(self.device)

    def ema_value(self, name):
        return self._gather_full("ema")[name]


# ===========================================================================
class NVSparseTable(object):
    SMEM_MAX_N = 8192

    def __init__(self, name, weight, num_partitions, strategy, optimizer, fabric,
                 route, graph, config, init=None, out_dtype=None, options=None):
        self.name = name
        self.fabric, self.heap = fabric, fabric.heap
        self.comm = fabric.comm
        self.rank, self.world, self.device = fabric.rank, fabric.world, fabric.device
        self.route, self.optimizer = route, optimizer
        _optim.require_fused(optimizer.kind, "NVLink fabric")
        self.kind = optimizer.kind
        self.nslots = _optim.NUM_SLOTS[self.kind]
        self.V, self.D = int(weight.shape[0]), int(weight.shape[1])
        self.Dp = (self.D + 3) // 4 * 4
        self.D4 = self.Dp // 4
        self.replicated = route.sparse == modes.SPARSE_ALLGATHER
        self.layout = TableLayout(self.V, num_partitions, self.world, strategy,
                                  replicated=self.replicated)
        self.geom = nvops.make_geom(self.layout, self.D4)
        self.average = bool(config.average_sparse)
        ps = config.communication_config.ps_config
        self.local_aggregation = bool(ps.local_aggregation)
        self.scale = graph.scale_for(name)
        # PSConfig.boundary_between_workers_and_servers: where gradient
        # post-processing (the ScaleGradients factor) runs — on the sender inside
        # the push kernel (True, the reference moves such ops to the side that
        # shrinks/keeps the wire traffic, graph_transform_lib.py:1315-1370) or on
        # the owner inside the apply kernel (False).
        self.scale_on_sender = bool(ps.boundary_between_workers_and_servers)
        opts = options or {}
        self.out_dtype = out_dtype or torch.float32
        self.anchor_device = self.device
        self.max_blocks = int(opts.get("sparse_blocks", 148 * 4))
        self.capacity_hint = (opts.get("sparse_capacity") or {}).get(name)
        # push/apply as soon as the table's last gradient of the step has arrived
        # (softmax tables: right after the loss backward, overlapping the LSTM
        # backward) instead of after the whole backward pass.  The comm-stream
        # order stays identical on all ranks because autograd order is.
        self.early_push = bool(opts.get("sparse_early_push", True))
        self._fwd_calls = self._bwd_calls = 0
        self._cur_step = 0
        self._done_step = -1
        L = self.layout
        rows = L.rows_local
        # table + slots in symmetric memory (async mode updates them remotely)
        self.tab_buf = self.heap.alloc(rows * self.Dp * 4, "table:" + name)
        self.table = self.tab_buf.tensor(torch.float32, rows * self.Dp).view(rows, self.Dp)
        self.slot_bufs, self.slots = [], []
        for v in optimizer.slot_init():
            sb = self.heap.alloc(rows * self.Dp * 4, "slot:" + name)
            t = sb.tensor(torch.float32, rows * self.Dp).view(rows, self.Dp)
            t.fill_(v)
            self.slot_bufs.append(sb)
            self.slots.append(t)
        self._init_weights(weight, init)
        if self.replicated:
            # every replica reads and updates its own full copy
            self.tables_dev = torch.tensor([self.tab_buf.local_ptr] * self.world,
                                           dtype=torch.int64, device=self.device)
            self.slots_dev = [torch.tensor([sb.local_ptr] * self.world,
                                           dtype=torch.int64, device=self.device)
                              for sb in self.slot_bufs]
        else:
            self._tables_dev = None
            self._slots_dev = None
        self.slotmap = torch.full((rows,), -1, dtype=torch.int32, device=self.device)
        self.ctl = torch.zeros(ops_ctl_words(), dtype=torch.int32, device=self.device)
        self.hp_host = torch.zeros(_optim.HP_SIZE, dtype=torch.float32).pin_memory()
        self.hp = torch.zeros(_optim.HP_SIZE, dtype=torch.float32, device=self.device)
        # header (flags) lives in its own small symmetric segment
        self.hdr_buf = self.heap.alloc(256, "hdr:" + name)
        self._hdrs_dev = None
        self.ring_buf = None
        self.cap = 0
        self.scratch_n = 0
        self.calls = []          # (ids32, grad) per lookup this step
        self.stats = {"pushed_rows": 0, "steps": 0}

    def _init_weights(self, weight, init):
        L = self.layout
        g, l = L.global_ids_of_owner(0 if self.replicated else self.rank)
        if weight.device.type == "meta":
            # lazy: initialise only this owner's rows, in chunks, on the device
            gen = torch.Generator(device=self.device)
            gen.manual_seed(int(init["seed"]) * 1000003 + (0 if self.replicated
                                                           else self.rank))
            self.table[:, :self.D].uniform_(-init["scale"], init["scale"],
                                            generator=gen)
            if self.Dp != self.D:
                self.table[:, self.D:].zero_()
        else:
            w = weight.detach().to(torch.float32)
            chunk = 1 << 20
            for s in range(0, g.numel(), chunk):
                gi, li = g[s:s + chunk], l[s:s + chunk].to(self.device)
                self.table[li, :self.D] = w[gi].to(self.device)

    # ---------------------------------------------------------------- forward
    def lookup(self, flat_ids, record=True):
        n = int(flat_ids.numel())
        ids = flat_ids if flat_ids.is_cuda else flat_ids.to(self.device,
                                                             non_blocking=True)
        if ids.dtype not in (torch.int64, torch.int32):
            ids = ids.to(torch.int64)
        ids = ids.contiguous()
        out = torch.empty((n, self.Dp), dtype=self.out_dtype, device=self.device)
        pend = torch.empty(n, dtype=torch.int32, device=self.device) if record else None
        if record:
            self._fwd_calls += 1
        nvops.sparse_lookup(ids, n, self._tdev(), out, pend, self.geom,
                            self.hdr_buf.local_ptr, self.ctl,
                            wait=self.route.sync and self.world > 1)
        if self.Dp != self.D:
            out = out[:, :self.D]
        return out, pend

    def _tdev(self):
        """Device array of every rank's table pointer (built lazily: in a
        simulated world the peers allocate after us)."""
        if self.replicated:
            return self.tables_dev
        if self._tables_dev is None:
            self._tables_dev = self.tab_buf.dev_ptrs()
            self._slots_dev = [sb.dev_ptrs() for sb in self.slot_bufs]
        return self._tables_dev

    def _sdev(self, i):
        self._tdev()
        sd = self.slots_dev if self.replicated else self._slots_dev
        return sd[i] if i < len(sd) else None

    def add_pending(self, token, grad_rows):
        g = grad_rows
        if self.Dp != self.D:
            g = torch.nn.functional.pad(g, (0, self.Dp - self.D))
        if g.dtype not in (torch.float32, torch.bfloat16):
            g = g.float()
        self.calls.append((token, g.contiguous()))
        self._bwd_calls += 1
        if self.early_push and self._bwd_calls == self._fwd_calls and self._cur_step > 0 \
                and self.ring_ready():
            self._run_step(self._cur_step)

    def ring_ready(self):
        """Early push needs every lazy allocation done (first step runs at the end)."""
        return self.scratch_n > 0 and (not self.route.sync or self.ring_buf is not None)

    # ----------------------------------------------------------------- update
    def _ensure_capacity(self, n):
        if n > self.scratch_n:
            cap = max(int(n * 1.25) + 16, 64)
            dev = self.device
            mk = lambda: torch.empty(cap, dtype=torch.int32, device=dev)
            self.uniq_id, self.uniq_k, self.uniq_cnt, self.pos2u = mk(), mk(), mk(), mk()
            # fp32 staging rows for ids carried by several positions (kept zero
            # between steps by the flush kernel)
            self.staging = torch.zeros(cap, self.Dp, dtype=torch.float32, device=dev)
            self.hbits = max(6, int(math.ceil(math.log2(max(2 * cap, 2)))))
            self.use_smem = cap <= self.SMEM_MAX_N
            if not self.use_smem:
                self.keys = torch.full((1 << self.hbits,), -1, dtype=torch.int32, device=dev)
                self.slot_u = torch.empty(1 << self.hbits, dtype=torch.int32, device=dev)
            else:
                self.keys = self.slot_u = None
            self.scratch_n = cap
        if self.route.sync and self.ring_buf is None:
            want = self.capacity_hint or max(int(n * 1.25) + 16, 64)
            if self.comm.distributed:
                want = max(self.comm.all_gather_object(int(want)))
            self.cap = int(want)
            rows_b = self.world * self.cap * self.Dp * 4
            ids_b = self.world * self.cap * 4
            self.ring_ids_off = (rows_b + 255) // 256 * 256
            self.ring_buf = self.heap.alloc(self.ring_ids_off + ids_b, "ring:" + self.name)
            self._rings_dev = None
            if self.comm.distributed:
                torch.cuda.synchronize(self.device)
                self.comm.barrier()
        if self.route.sync and n > self.cap:
            raise RuntimeError(
                "sparse table %r: %d gradient rows in one step exceed the ring "
                "capacity %d fixed at the first step; pass sess_config="
                "{'sparse_capacity': {%r: N}}" % (self.name, n, self.cap, self.name))

    @property
    def hdrs_dev(self):
        if self._hdrs_dev is None:
            self._hdrs_dev = self.hdr_buf.dev_ptrs()
        return self._hdrs_dev

    @property
    def rings_dev(self):
        if self._rings_dev is None:
            self._rings_dev = self.ring_buf.dev_ptrs()
        return self._rings_dev

    def warm(self, n):
        """Allocate everything a step of `n` gradient rows needs (no lazy
        allocation / pointer upload will happen inside the step)."""
        self._ensure_capacity(n)
        self._tdev()
        if self.route.sync:
            self.rings_dev, self.hdrs_dev

    def begin_step(self, step):
        hp = self.optimizer.hyper(step)
        for i, v in enumerate(hp):
            self.hp_host[i] = v
        self.hp.copy_(self.hp_host, non_blocking=True)
        self._cur_step = step
        self._fwd_calls = self._bwd_calls = 0

    def finish_step(self, step, stream=None):
        if self._done_step == step and not self.calls:
            return                           # already pushed from the backward pass
        self._run_step(step, stream)

    def _run_step(self, step, stream=None):
        from ..utils import timeline
        self._done_step = step
        if timeline.enabled():
            cs = stream if stream is not None else self.fabric.comm_stream
            with timeline.activity(self.name, "SPARSE_PUSH_APPLY", gpu=True, stream=cs,
                                   args="rows=%d" % sum(c[0].numel() for c in self.calls)):
                self._finish_step_impl(step, stream)
        else:
            self._finish_step_impl(step, stream)

    def _finish_step_impl(self, step, stream=None):
        self.stage_push(step, stream)
        if self.route.sync:
            self.stage_apply(step, stream)

    def stage_push(self, step, stream=None):
        """Sender side: local aggregation + push (or remote apply in async
        mode).  Separate from `stage_apply` so that a world simulated on one GPU
        can enqueue every rank's push before any rank's (spinning) owner kernels —
        streams of one process may share a hardware queue."""
        cs = stream if stream is not None else self.fabric.comm_stream
        calls, self.calls = self.calls, []
        if calls:
            pend_ids = calls[0][0] if len(calls) == 1 else torch.cat([c[0] for c in calls])
            grads = calls[0][1] if len(calls) == 1 else torch.cat([c[1] for c in calls])
        else:
            pend_ids = torch.empty(0, dtype=torch.int32, device=self.device)
            grads = torch.empty((0, self.Dp), dtype=torch.float32, device=self.device)
        n = int(pend_ids.numel())
        self._ensure_capacity(max(n, 1))
        self._last_n = n
        self.stats["pushed_rows"] += n
        self.stats["steps"] += 1
        cur = torch.cuda.current_stream(self.device)
        cs.wait_stream(cur)
        if not torch.cuda.is_current_stream_capturing():
            pend_ids.record_stream(cs)
            grads.record_stream(cs)
        nvops.sparse_dedup(pend_ids, n, self.hbits, self.keys, self.slot_u,
                           self.uniq_id, self.uniq_k, self.uniq_cnt, self.pos2u,
                           self.ctl, self.geom, self.local_aggregation,
                           self.use_smem, stream=cs)
        s0d, s1d = self._sdev(0), self._sdev(1)
        if not self.route.sync:
            nvops.sparse_async_apply(grads, n, self.pos2u, self.uniq_id, self.uniq_k,
                                     self.uniq_cnt, self.staging, self.ctl,
                                     self._tdev(), s0d, s1d, self.hp, self.scale,
                                     self.kind, self.geom, self.max_blocks, stream=cs)
            return
        nvops.sparse_push(grads, n, self.pos2u, self.uniq_id, self.uniq_k,
                          self.uniq_cnt, self.staging, self.ctl, self.rings_dev,
                          self.hdrs_dev, self.ring_ids_off, self.cap, self.geom,
                          self.scale if self.scale_on_sender else 1.0, self.rank,
                          self.max_blocks, stream=cs)

    def stage_apply(self, step, stream=None):
        """Owner side: merge rows from all sources, apply the sparse optimizer."""
        cs = stream if stream is not None else self.fabric.comm_stream
        n = getattr(self, "_last_n", 1)
        # 8 warps per CTA, one row per warp; bounded by the configured cap
        blk_own = max(1, min(self.max_blocks, (n * self.world + 7) // 8))
        need_claim = self.world > 1 or not self.local_aggregation
        if need_claim:
            nvops.sparse_claim(self.ring_buf.local_ptr, self.hdr_buf.local_ptr,
                               self.ring_ids_off, self.cap, self.slotmap, self.ctl,
                               self.geom, blk_own, stream=cs)
        avg = (1.0 / self.world) if self.average else 1.0
        if not self.scale_on_sender:
            avg *= self.scale
        nvops.sparse_apply(self.ring_buf.local_ptr, self.hdr_buf.local_ptr,
                           self.ring_ids_off, self.cap, self.slotmap, self.table,
                           self.slots[0] if self.nslots > 0 else None,
                           self.slots[1] if self.nslots > 1 else None, self.hp, avg,
                           self.kind, self.ctl, self.hdrs_dev, self.geom, self.rank,
                           need_claim, blk_own, stream=cs)

    # -------------------------------------------------------------- checkpoint
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

    def release(self):
        """Free this table's symmetric segments (collective)."""
        torch.cuda.synchronize(self.device)
        if self.comm.distributed:
            self.comm.barrier()
        for b in [self.tab_buf, self.hdr_buf, self.ring_buf] + list(self.slot_bufs):
            if b is not None:
                self.heap.free(b)
        self.table = None
        self.slots = []

    def load_full(self, weight, slots=None):
        g, l = self.layout.global_ids_of_owner(0 if self.replicated else self.rank)
        l = l.to(self.device)
        self.table[l, :self.D] = weight.float()[g].to(self.device)
        if slots is not None:
            for s, full in zip(self.slots, slots):
                s[l, :self.D] = full.float()[g].to(self.device)
        torch.cuda.synchronize(self.device)


def ops_ctl_words():
    from .. import ops
    return int(ops.lib().px_sparse_ctl_bytes()) // 4


# ===========================================================================
def build_nvlink(engine):
    """Module surgery for the NVLink fabric (called by `TrainEngine._build`)."""
    from .engine import ShardedEmbedding, _set_submodule
    from .. import ops
    if not torch.cuda.is_available():
        raise RuntimeError("the NVLink fabric needs a CUDA device")
    ops.lib()      # loud failure if the native library is missing
    g, comm, cfg = engine.graph, engine.comm, engine.config
    # convolution algorithms are picked by measurement during the eager warm-up
    # steps (shapes are static per graph); TF's cuDNN autotune did the same for the
    # reference (tensorflow/core/kernels/conv_ops.cc:746-783)
    torch.backends.cudnn.benchmark = bool(
        (cfg.sess_config or {}).get("cudnn_benchmark", True)
        if isinstance(cfg.sess_config, dict) else True)
    opts = dict(cfg.sess_config) if isinstance(cfg.sess_config, dict) else {}
    ps_cfg = cfg.communication_config.ps_config
    opts["_protocol"] = "nccl" if ps_cfg.protocol == "nccl" else "nvlink"
    opts["_replicate_variables"] = bool(ps_cfg.replicate_variables)
    fabric = NVFabric(comm, exchange=opts.get("_exchange"), options=opts)
    engine.fabric = fabric
    dev = comm.device
    cdt = opts.get("compute_dtype")
    cdt = {None: None, "float32": torch.float32, "fp32": torch.float32,
           "bfloat16": torch.bfloat16, "bf16": torch.bfloat16}.get(cdt, cdt)
    # sparse tables first (their weights may be meta / huge)
    sparse_items = list(engine.analysis.sparse_modules.items())
    for path, mod in sorted(sparse_items, key=lambda kv: kv[0]):
        pname = path + ".weight" if path else "weight"
        info = engine.analysis.variables[pname]
        part = getattr(mod, "partitioner", None)
        t = NVSparseTable(
            pname, mod.weight, info.partitions,
            part.strategy if part is not None else "mod", g.sparse_optimizer,
            fabric, engine.route, g, cfg,
            init={"seed": getattr(mod, "init_seed", 1234),
                  "scale": getattr(mod, "init_scale", 0.05)},
            out_dtype=cdt or torch.float32, options=opts)
        engine.tables[pname] = t
        _set_submodule(engine.model, path, ShardedEmbedding(t))
    engine.model.to(dev)
    if cdt is not None:
        for p in engine.model.parameters():
            if p.is_floating_point() and not getattr(p, "_parallax_skip", False):
                p.data = p.data.to(cdt)
        for b in engine.model.buffers():
            if b.is_floating_point() and opts.get("cast_buffers", False):
                b.data = b.data.to(cdt)
    dense_named = [(n, p) for n, p in engine.model.named_parameters()
                   if p.requires_grad and not getattr(p, "_parallax_skip", False)]
    if g.trainable() and dense_named:
        engine.dense = NVDenseGroup(dense_named, g.optimizer, fabric, engine.route,
                                    g, options=opts)
    torch.cuda.synchronize(dev)
    parallax_log.info(
        "nvlink fabric: rank %d/%d, %d dense buckets, %d sparse tables, "
        "symmetric heap %.1f MiB", comm.rank, comm.world,
        len(engine.dense.buckets) if engine.dense else 0, len(engine.tables),
        ops.lib().px_symm_live_bytes() / 2 ** 20)
