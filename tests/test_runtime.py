"""Native runtime pieces that do not need a GPU: timeline writer, Bayesian
autotuner, op registry (response-cache analogue), stall watchdog, collectives
API on the host fabric (gloo), error catalogue."""
import json
import os
import subprocess
import sys
import time

import pytest
import torch

from tests.dist_utils import run_distributed

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_timeline_writes_valid_chrome_trace(tmp_path):
    from parallax_b200.utils import timeline
    p = tmp_path / "trace.json"
    timeline.start(str(p), rank=0, use_cuda=False)
    for i in range(100):
        with timeline.activity("bucket%d" % (i % 3), "DENSE_STEP", args="n=%d" % i):
            pass
        timeline.instant("CYCLE_START")
    timeline.stop()
    ev = json.load(open(p))
    names = {e["name"] for e in ev}
    assert {"DENSE_STEP", "CYCLE_START", "thread_name"} <= names
    assert sum(1 for e in ev if e["ph"] == "B") == sum(1 for e in ev if e["ph"] == "E") == 100
    # horovod/test/test_timeline.py asserts marker strings are present
    txt = open(p).read()
    assert "CYCLE_START" in txt and "DENSE_STEP" in txt


def test_engine_emits_timeline(tmp_path, monkeypatch):
    import parallax_b200 as parallax
    from parallax_b200 import optim
    from parallax_b200.models.simple import MLPWithEmbedding
    p = tmp_path / "t.json"
    monkeypatch.setenv("PARALLAX_TIMELINE", str(p))
    g = parallax.Graph(MLPWithEmbedding(64), optimizer=optim.GradientDescent(0.1))
    sess, *_ = parallax.parallel_run(g, "localhost", parallax_config=parallax.Config())
    ids = torch.randint(0, 64, (4, 3))
    for _ in range(3):
        sess.run(["loss", "train_op"], {"ids": [ids], "labels": [torch.zeros(4, dtype=torch.long)]})
    sess.close()
    sess.engine.timeline.stop()
    ev = json.load(open(p))
    assert sum(1 for e in ev if e["name"] == "STEP" and e["ph"] == "B") == 3


def test_bayesian_tuner_finds_optimum():
    from parallax_b200.utils.autotune import BayesianTuner
    t = BayesianTuner({"fusion_mb": (0, 64), "cycle_ms": (1, 100)}, {"hier": [0, 1]},
                      samples_per_point=3, max_points=14, warmups=1)
    f = lambda a, b, h: 10 - ((a - 24) / 16) ** 2 - ((b - 40) / 30) ** 2 + (1.0 if h else 0)
    n = 0
    while not t.done and n < 500:
        c = t.current()
        t.report(f(c["fusion_mb"], c["cycle_ms"], c["hier"]))
        n += 1
    best = t.current()
    assert t.done and best["hier"] == 1
    assert f(best["fusion_mb"], best["cycle_ms"], 1) > 10.3
    assert t.num_points() == 14
    t.close()


def test_registry_lru_and_invalidation():
    from parallax_b200 import collectives as hvd
    L = hvd._registry_lib()
    L.px_registry_reset(4)
    assert L.px_registry_lookup(b"a", 1) == 0             # MISS
    bits = [L.px_registry_put(n.encode(), 1) for n in "abcd"]
    assert sorted(bits) == [0, 1, 2, 3]
    assert L.px_registry_lookup(b"a", 1) == 1             # HIT (and now most recent)
    assert L.px_registry_lookup(b"b", 2) == 2             # INVALID: signature changed
    L.px_registry_put(b"e", 1)                            # evicts LRU = "b"
    assert L.px_registry_lookup(b"b", 1) == 0
    assert L.px_registry_lookup(b"a", 1) == 1
    st = hvd.registry_stats()
    assert st["evictions"] == 1 and st["size"] == 4 and st["invalid"] == 1


def _hvd_worker(rank, world):
    from parallax_b200 import collectives as hvd
    hvd.init()
    assert (hvd.rank(), hvd.size()) == (rank, world)
    out = {}
    x = torch.arange(10, dtype=torch.float32) * (rank + 1)
    out["sum"] = hvd.allreduce(x, average=False, name="t.sum")
    out["avg"] = hvd.allreduce(x, average=True, name="t.avg")
    out["again"] = hvd.allreduce(x, average=False, name="t.sum")     # registry hit
    out["gather"] = hvd.allgather(torch.full((2 + rank, 3), float(rank)), name="g")
    out["bcast"] = hvd.broadcast(torch.full((4,), float(rank)), root_rank=1, name="b")
    sp = torch.sparse_coo_tensor(torch.tensor([[rank, 5]]), torch.tensor([[1.0, 2.0]] * 2),
                                 (8, 2))
    out["sparse"] = hvd.allreduce(sp, average=False, name="sp").to_dense()
    h = hvd.allreduce_async(x, name="dup")
    try:
        hvd.allreduce_async(x, name="dup")
        out["dup_error"] = False
    except hvd.HorovodInternalError:
        out["dup_error"] = True
    hvd.synchronize(h)
    # mismatched shapes must raise on every rank (horovod test_tensorflow.py error cases)
    try:
        hvd.allreduce(torch.zeros(3 + rank), name="bad")
        out["mismatch_error"] = False
    except hvd.HorovodInternalError:
        out["mismatch_error"] = True
    # mismatched broadcast root / allgather trailing dims are errors too
    try:
        hvd.broadcast(torch.zeros(2), root_rank=rank, name="badroot")
        out["root_error"] = False
    except hvd.HorovodInternalError:
        out["root_error"] = True
    try:
        hvd.allgather(torch.zeros(2, 3 + rank), name="badgather")
        out["gather_error"] = False
    except hvd.HorovodInternalError:
        out["gather_error"] = True
    g = hvd.grouped_allreduce([torch.ones(3) * (rank + 1), torch.ones(2, 2) * rank],
                              average=False)
    out["grouped"] = [t.clone() for t in g]
    out["fp16"] = hvd.allreduce(torch.ones(4) * (rank + 1), average=True,
                                compression=hvd.Compression.fp16)
    out["stats"] = hvd.registry_stats()
    # DistributedOptimizer averages gradients
    w = torch.nn.Parameter(torch.ones(3))
    opt = hvd.DistributedOptimizer(torch.optim.SGD([w], lr=1.0), [("w", w)])
    (w * float(rank + 1)).sum().backward()
    opt.step()
    out["w"] = w.detach().clone()
    hvd.shutdown()
    return out


def test_collectives_api_host_fabric():
    res = run_distributed(_hvd_worker, 2)
    for r, o in enumerate(res):
        base = torch.arange(10, dtype=torch.float32)
        torch.testing.assert_close(o["sum"], base * 3)          # tensor × Σ(rank+1)
        torch.testing.assert_close(o["avg"], base * 1.5)
        torch.testing.assert_close(o["again"], base * 3)
        assert o["gather"].shape == (5, 3) and float(o["gather"][2:].min()) == 1.0
        assert float(o["bcast"].min()) == 1.0
        exp = torch.zeros(8, 2)
        exp[0] += torch.tensor([1.0, 2.0]); exp[1] += torch.tensor([1.0, 2.0])
        exp[5] += 2 * torch.tensor([1.0, 2.0])
        torch.testing.assert_close(o["sparse"], exp)
        assert o["dup_error"] and o["mismatch_error"] and o["root_error"] and o["gather_error"]
        torch.testing.assert_close(o["grouped"][0], torch.full((3,), 3.0))
        torch.testing.assert_close(o["grouped"][1], torch.full((2, 2), 1.0))
        torch.testing.assert_close(o["fp16"], torch.full((4,), 1.5))
        assert o["stats"]["hits"] >= 1
        torch.testing.assert_close(o["w"], torch.ones(3) - 1.5)


STALL_SCRIPT = r"""
import sys, time
sys.path.insert(0, %r)
from parallax_b200.utils.watchdog import Watchdog
w = Watchdog(rank=0, world=2, heap=None, warn_s=0.3, shutdown_s=1.0)
w.beat(1)
time.sleep(30)          # "a peer never arrives": no more heartbeats
print("NOT REACHED")
"""


def test_stall_watchdog_shuts_down():
    """horovod/test/test_stall.py: a stalled rank must warn, then terminate."""
    t0 = time.time()
    r = subprocess.run([sys.executable, "-c", STALL_SCRIPT % ROOT], capture_output=True,
                       text=True, timeout=60)
    assert r.returncode == 17, (r.returncode, r.stderr[-500:])
    assert "no progress" in r.stderr and "terminating" in r.stderr
    assert "NOT REACHED" not in r.stdout
    assert time.time() - t0 < 20
