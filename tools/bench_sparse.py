"""Device time of the sparse-path kernels at the LM1B shapes on ONE GPU (world 1):
lookup / push / owner for the softmax group (793470 x 512 + 793470 x 1, 10752 ids) and the
embedding table (2560 ids), bf16 gradients.  Each kernel timed alone with CUDA events over
50 launches (fresh random ids every launch, tables >> L2).
Usage: python tools/bench_sparse.py [--ncu]   (--ncu: 3 iterations only, for a capture)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import parallax_b200 as parallax
from parallax_b200 import optim
from parallax_b200.graph import Graph
from parallax_b200.parallel import modes
from parallax_b200.parallel.nvlink_backend import NVFabric, NVSparseTable, NVSparseGroup
from parallax_b200.parallel.symmetric import LocalWorld

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from gpu_utils import FakeComm

ITERS = 3 if "--ncu" in sys.argv else 50
V = 793470
for a_ in sys.argv:
    if a_.startswith("--V="):
        V = int(a_[4:])          # a small V makes every row L2 / TLB resident (latency probe)
fab = NVFabric(FakeComm(0, 1, "cuda:0"), exchange=LocalWorld(1).exchange_for(0))
opt = optim.Adagrad(0.2, 1.0)
route = modes.route_for("HYBRID", True)
cfg = parallax.Config(run_option="HYBRID")
if "--merge" in sys.argv:       # force the owner's cross-source merge path on one GPU
    cfg.communication_config = parallax.CommunicationConfig(
        parallax.PSConfig(local_aggregation=False))
graph = Graph(torch.nn.Linear(1, 1), optimizer=opt)
meta = lambda d: torch.empty(V, d, device="meta")
init = {"seed": 1, "scale": 0.05}
o = {"sparse_early_push": False}
mk = lambda name, d: NVSparseTable(name, meta(d), 32, "mod", opt, fab, route, graph, cfg,
                                   init=init, out_dtype=torch.bfloat16, options=o,
                                   auto_group=False)
g_soft = NVSparseGroup([mk("softmax_w", 512), mk("softmax_b", 1)])
g_emb = NVSparseGroup([mk("emb", 512)])


def ev():
    return torch.cuda.Event(enable_timing=True)


def bench(grp, n, name):
    grp.warm(n)
    torch.cuda.synchronize()
    tl = tp = to = 0.0
    for it in range(ITERS + 2):
        ids = torch.randint(0, V, (n,), device="cuda")
        grads = [torch.randn(n, t.D, device="cuda").bfloat16() for t in grp.tables]
        grp.begin_step(it + 1)
        torch.cuda.synchronize()
        e = [ev() for _ in range(4)]
        cs = torch.cuda.current_stream()
        e[0].record()
        outs, tok = grp.lookup(ids)
        e[1].record()
        grp.add_pending(tok, grads)
        grp.stage_push(it + 1, stream=cs)
        e[2].record()
        grp.stage_apply(it + 1, stream=cs)
        e[3].record()
        torch.cuda.synchronize()
        if it >= 2:
            tl += e[0].elapsed_time(e[1]); tp += e[1].elapsed_time(e[2]); to += e[2].elapsed_time(e[3])
    d = grp.device_times()
    ph = d["push_phases"]
    print("   push kernel (device timer, CTA 0): total %.1f us; phases scan1 %.1f | slots %.1f | "
          "scan3a(chunk0) %.1f | ship %.1f | flush %.1f | fence+ticket %.1f ; owner %.1f us" % (
              (d["pushed"] - d["push_start"]) / 1e3, (ph[0] - d["push_start"]) / 1e3,
              (ph[1] - ph[0]) / 1e3, (ph[2] - ph[1]) / 1e3, (ph[3] - ph[2]) / 1e3,
              (ph[4] - ph[3]) / 1e3, (ph[5] - ph[4]) / 1e3,
              (d["applied"] - d["owner_start"]) / 1e3), flush=True)
    if ph[6] and ph[7] >= ph[6] > d["owner_start"]:
        print("   owner (merge): link %.1f us | grid barrier %.1f us | apply+publish %.1f us" % (
            (ph[6] - d["owner_start"]) / 1e3, (ph[7] - ph[6]) / 1e3,
            (d["applied"] - ph[7]) / 1e3), flush=True)
    rowb = sum(t.Dp for t in grp.tables)
    print("%-10s n=%-6d lookup %6.1f us  push %6.1f us  owner %6.1f us   "
          "(push moves %.1f MB bf16, owner %.1f MB)" %
          (name, n, tl / ITERS * 1e3, tp / ITERS * 1e3, to / ITERS * 1e3,
           n * rowb * 2 * 2 / 1e6, n * rowb * (2 + 16 + 2) / 1e6), flush=True)


bench(g_soft, 10752, "softmax")
bench(g_emb, 2560, "emb")
fab.close()
