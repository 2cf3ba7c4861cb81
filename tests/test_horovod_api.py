"""More of Horovod's user API on the host fabric (gloo, world_size 2):
`DistributedOptimizer(compression, backward_passes_per_step)`,
`broadcast_optimizer_state`, `DistributedGradientTape`, `allreduce_gradients`, and the
Keras-style callbacks (`horovod/test/test_torch.py`, `test_keras.py` — SURVEY §4)."""
import numpy as np
import torch

from tests.dist_utils import run_distributed


def _worker(rank, world):
    import torch
    from parallax_b200 import callbacks as cbs
    from parallax_b200 import collectives as hvd
    hvd.init()
    out = {}
    assert hvd.mpi_threads_supported() and hvd.broadcast_object({"r": rank}, 1) == {"r": 1}
    assert not hvd.mpi_built() and not hvd.mpi_enabled() and hvd.gloo_built()
    assert isinstance(hvd.nccl_built(), bool) and hvd.cuda_built() == torch.cuda.is_available()

    # ---- DistributedOptimizer: gradient accumulation + fp16 compression ------------
    torch.manual_seed(0)
    model = torch.nn.Linear(4, 2)
    opt = hvd.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9),
                                   named_parameters=model.named_parameters(),
                                   compression=hvd.Compression.fp16, backward_passes_per_step=2)
    w0 = model.weight.detach().clone()
    xs = [torch.full((3, 4), float(rank + 1)), torch.full((3, 4), float(rank + 3))]
    # the Horovod idiom: N backward passes, then ONE step() that reduces and applies
    model(xs[0]).sum().backward()
    model(xs[1]).sum().backward()
    assert torch.equal(model.weight, w0)                               # nothing applied yet
    try:
        model(xs[1]).sum().backward()                                  # a third pass: error
        out["third_pass_raises"] = False
    except AssertionError as e:
        out["third_pass_raises"] = "more than backward_passes_per_step" in str(e)
        opt.zero_grad()                                                # redo the two passes
        opt._delay = {k: 2 for k in opt._delay}
        model(xs[0]).sum().backward()
        model(xs[1]).sum().backward()
    opt.step()
    # gradient of sum(Wx+b) wrt W = Σ_rows x; accumulated over 2 passes, averaged over ranks
    per_rank = [3.0 * ((r + 1) + (r + 3)) for r in range(world)]
    want = w0 - 0.1 * (sum(per_rank) / world)
    out["accum_ok"] = bool(torch.allclose(model.weight, want, atol=1e-2))
    opt.zero_grad()

    # ---- broadcast_optimizer_state: fresh optimizer on rank 1 gets rank 0's state ---
    torch.manual_seed(rank)
    m2 = torch.nn.Linear(3, 3)
    o2 = torch.optim.Adam(m2.parameters(), lr=0.01 * (rank + 1))
    if rank == 0:
        for _ in range(3):
            m2(torch.randn(2, 3)).sum().backward()
            o2.step()
            o2.zero_grad()
    hvd.broadcast_parameters(m2.state_dict(), 0)
    hvd.broadcast_optimizer_state(o2, 0)
    sd = o2.state_dict()
    out["opt_lr"] = sd["param_groups"][0]["lr"]
    out["opt_step"] = float(sd["state"][0]["step"])
    out["opt_exp_avg"] = sd["state"][0]["exp_avg"].clone()
    out["m2_w"] = m2.weight.detach().clone()

    # ---- DistributedGradientTape / allreduce_gradients ---------------------------------
    w = torch.ones(3, requires_grad=True)
    unused = torch.ones(2, requires_grad=True)
    loss = (w * float(rank + 1)).sum()
    g = hvd.DistributedGradientTape().gradient(loss, [w, unused])
    out["tape"] = g[0].clone()
    out["tape_unused_none"] = g[1] is None
    emb = torch.nn.Embedding(6, 2, sparse=True)
    emb(torch.tensor([rank, rank + 1])).sum().backward()
    lin = torch.nn.Linear(2, 1)
    lin(torch.full((1, 2), float(rank))).sum().backward()
    hvd.allreduce_gradients(list(emb.parameters()) + list(lin.parameters()), average=False,
                            sparse_as_dense=True)
    out["emb_grad_rows"] = emb.weight.grad.sum(1).clone()
    out["lin_grad"] = lin.weight.grad.clone()

    # ---- callbacks ----------------------------------------------------------------------
    torch.manual_seed(100 + rank)
    net = torch.nn.Linear(2, 2)
    sgd = torch.optim.SGD(net.parameters(), lr=0.4, momentum=0.9)
    cl = cbs.CallbackList([cbs.BroadcastGlobalVariablesCallback(0), cbs.MetricAverageCallback(),
                           cbs.LearningRateWarmupCallback(warmup_epochs=2, steps_per_epoch=4),
                           cbs.LearningRateScheduleCallback(0.1, start_epoch=3)],
                          model=net, optimizer=sgd)
    cl.on_train_begin()
    out["net_w"] = net.weight.detach().clone()
    lrs = []
    for epoch in range(5):
        cl.on_epoch_begin(epoch)
        for b in range(4):
            cl.on_batch_begin(b)
            lrs.append(sgd.param_groups[0]["lr"])
            net(torch.ones(1, 2)).sum().backward()
            sgd.step()
            sgd.zero_grad()
            cl.on_batch_end(b)
        logs = {"loss": float(rank), "acc": 10.0 * rank}
        cl.on_epoch_end(epoch, logs)
        out["logs%d" % epoch] = dict(logs)
    out["lrs"] = lrs
    hvd.shutdown()
    return out


def test_horovod_surface_two_ranks():
    r0, r1 = run_distributed(_worker, 2)
    assert r0["accum_ok"] and r1["accum_ok"]
    assert r0["third_pass_raises"] and r1["third_pass_raises"]
    # rank 1's fresh Adam now mirrors rank 0's: hyper-parameters, step, slots, weights
    assert r1["opt_lr"] == r0["opt_lr"] == 0.01 and r1["opt_step"] == r0["opt_step"] == 3.0
    assert torch.equal(r0["opt_exp_avg"], r1["opt_exp_avg"]) and torch.equal(r0["m2_w"], r1["m2_w"])
    for r in (r0, r1):
        assert torch.allclose(r["tape"], torch.full((3,), 1.5)) and r["tape_unused_none"]
        assert r["emb_grad_rows"].tolist() == [2.0, 4.0, 2.0, 0.0, 0.0, 0.0]   # rows 0,1 | 1,2
        assert torch.allclose(r["lin_grad"], torch.full((1, 2), 1.0))          # 0 + 1
    assert torch.equal(r0["net_w"], r1["net_w"])
    # metrics averaged over the two workers
    assert r0["logs0"]["loss"] == r1["logs0"]["loss"] == 0.5 and r0["logs4"]["acc"] == 5.0
    lrs = np.array(r0["lrs"]).reshape(5, 4)
    # warm-up: 1/size of the target at the start, linear per batch, exact target after it
    assert abs(lrs[0, 0] - 0.2) < 1e-9 and np.all(np.diff(lrs[:2].reshape(-1)) > 0)
    assert abs(lrs[1, 3] - 0.4 * 0.5 * (1.75 * 0.5 + 1)) < 1e-9
    assert np.allclose(lrs[2], 0.4) and np.allclose(lrs[3:], 0.04)
    assert abs(r0["logs3"]["lr"] - 0.04) < 1e-12


def test_schedule_callback_momentum_correction_single_process():
    from parallax_b200 import callbacks as cbs
    from parallax_b200 import collectives as hvd
    hvd.init()
    try:
        net = torch.nn.Linear(2, 1)
        sgd = torch.optim.SGD(net.parameters(), lr=1.0, momentum=0.5)
        net(torch.ones(1, 2)).sum().backward()
        sgd.step()
        buf = sgd.state[net.weight]["momentum_buffer"].clone()
        cb = cbs.LearningRateScheduleCallback(lambda e: 0.5 ** e, momentum_correction=True)
        cb.set_context(net, sgd)
        cb.on_train_begin()
        cb.on_epoch_begin(2)
        assert sgd.param_groups[0]["lr"] == 0.25
        assert torch.allclose(sgd.state[net.weight]["momentum_buffer"], buf * 0.25)
        import pytest
        with pytest.raises(ValueError):
            cbs.LearningRateScheduleCallback(1.0, staircase=False)
        with pytest.raises(AttributeError):
            cbs.CallbackList([]).no_such_hook
    finally:
        hvd.shutdown()


def _sweep_worker(rank, world):
    import itertools
    import torch
    from parallax_b200 import collectives as hvd
    hvd.init()
    out = {"ok": True, "msgs": []}

    def check(cond, msg):
        if not cond:
            out["ok"] = False
            out["msgs"].append(msg)
    # allreduce / allgather / broadcast over dtypes × ranks-of-tensor (test_torch.py sweeps)
    dtypes = [torch.int32, torch.int64, torch.float16, torch.float32, torch.float64]
    for dt, dim in itertools.product(dtypes, (1, 2, 3)):
        torch.manual_seed(1234)
        shape = (5,) * dim
        base = (torch.rand(shape) * 20 - 10).to(dt)
        x = base * (rank + 1) if dt.is_floating_point else base * (rank + 1)
        tag = "%s.%d" % (str(dt).split(".")[1], dim)
        s = hvd.allreduce(x, average=False, name="ar." + tag)
        want = base * sum(r + 1 for r in range(world))
        tol = 1e-2 if dt == torch.float16 else 1e-5
        check(s.dtype == dt and torch.allclose(s.double(), want.double(), atol=tol * 40),
              "allreduce sum " + tag)
        y = x.clone()
        hvd.allreduce_(y, average=False, name="ari." + tag)
        check(torch.equal(y, s), "allreduce_ in place " + tag)
        if dt.is_floating_point:
            a = hvd.allreduce(x, average=True, name="av." + tag)
            check(torch.allclose(a.double(), want.double() / world, atol=tol * 40), "avg " + tag)
        h = hvd.allreduce_async(x, average=False, name="as." + tag)
        while not hvd.poll(h):
            pass
        check(torch.equal(hvd.synchronize(h), s), "async " + tag)
        g = hvd.allgather(torch.full((rank + 1,) + shape[1:], rank, dtype=dt), name="ag." + tag)
        check(g.shape[0] == sum(r + 1 for r in range(world)) and g.dtype == dt and
              float(g[0].double().max()) == 0 and float(g[-1].double().min()) == world - 1,
              "allgather variable " + tag)
        b = hvd.broadcast(torch.full(shape, rank, dtype=dt), root_rank=world - 1, name="bc." + tag)
        check(b.dtype == dt and float(b.double().min()) == world - 1, "broadcast " + tag)
        z = torch.full(shape, rank, dtype=dt)
        hvd.broadcast_(z, root_rank=0, name="bci." + tag)
        check(float(z.double().max()) == 0, "broadcast_ " + tag)
    # dtype mismatch across ranks is an error on every rank
    try:
        hvd.allreduce(torch.zeros(3, dtype=torch.float32 if rank == 0 else torch.float64),
                      name="bad.dtype")
        check(False, "dtype mismatch not detected")
    except hvd.HorovodInternalError:
        pass
    # gradients of the collectives (horovod registers them for its ops)
    w = torch.arange(4, dtype=torch.float64, requires_grad=True)
    y = hvd.allreduce(w * (rank + 1), average=False, name="grad.ar")
    (y * torch.tensor([1.0, 2.0, 3.0, 4.0], dtype=torch.float64) * (rank + 2)).sum().backward()
    # dL_r/dy = (r+2)·c ; dy/dw_r = (r+1) ⇒ grad = (r+1)·Σ_r'(r'+2)·c
    c = torch.tensor([1.0, 2.0, 3.0, 4.0], dtype=torch.float64)
    check(torch.allclose(w.grad, (rank + 1) * sum(r + 2 for r in range(world)) * c), "grad ar")
    v = torch.ones(rank + 1, 2, dtype=torch.float64, requires_grad=True)
    g = hvd.allgather(v, name="grad.ag")
    (g * torch.arange(g.numel(), dtype=torch.float64).view_as(g)).sum().backward()
    off = sum(r + 1 for r in range(rank)) * 2
    want = world * torch.arange(off, off + v.numel(), dtype=torch.float64).view_as(v)
    check(torch.allclose(v.grad, want), "grad allgather")
    u = torch.full((3,), float(rank + 1), dtype=torch.float64, requires_grad=True)
    b = hvd.broadcast(u, root_rank=1, name="grad.bc")
    (b * (rank + 1)).sum().backward()
    want = torch.full((3,), float(sum(r + 1 for r in range(world)))) if rank == 1 \
        else torch.zeros(3)
    check(torch.allclose(u.grad, want.double()), "grad broadcast")
    hvd.shutdown()
    return out


def test_dtype_dim_sweep_and_gradients_two_ranks():
    for r in run_distributed(_sweep_worker, 2, timeout=400):
        assert r["ok"], r["msgs"]
