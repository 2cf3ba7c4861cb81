"""A few-second correctness probe of the live fabric: train BASELINE config 1 (2-layer
MLP + tiny embedding table) for a few steps through `parallel_run` on the current
world and compare every variable with a single-device oracle that sees the
concatenated batch.  `bench.py` runs it before timing anything (a fast fabric that
computes the wrong thing is not a result); `tests/test_multigpu.py` runs the full
matrix of run options through the same function.

Reference semantics being checked: synchronous data parallelism = one update with
the mean dense gradient and the (averaged or summed) sparse gradient of all workers
(`graph_transform_lib.py:330-582,1558-1946`, `horovod/tensorflow/__init__.py:62-82`).
"""
import torch

B, T, VOCAB = 8, 3, 64


def make_batch(step, world, rank=None):
    g = torch.Generator().manual_seed(100 + step)
    ids = torch.randint(0, VOCAB, (B * world, T), generator=g)
    ids[:, 0] = ids[0, 0]            # duplicates inside and across workers
    labels = torch.randint(0, 4, (B * world,), generator=g)
    if rank is None:
        return ids, labels
    return ids[rank * B:(rank + 1) * B], labels[rank * B:(rank + 1) * B]


def make_opt(name):
    from .. import optim
    return {"sgd": optim.GradientDescent(0.5),
            "adagrad": optim.Adagrad(0.2, initial_accumulator_value=1.0),
            "adam": optim.Adam(0.01),
            "momentum": optim.Momentum(0.1, 0.9),
            "rmsprop": optim.RMSProp(0.01, momentum=0.5),
            "ftrl": optim.Ftrl(0.2, l1_regularization_strength=0.001),
            "centered_rmsprop": optim.CenteredRMSProp(0.01, momentum=0.5, epsilon=1e-3)}[name]


def oracle(world, steps, opt, sparse_scale=1.0):
    """Single-device training on the concatenated batch; sparse grads scaled by
    `sparse_scale` (world for sum semantics, 1 for average)."""
    from .. import optim
    from ..models.simple import MLPWithEmbedding
    model = MLPWithEmbedding(VOCAB)
    model.emb.sparse = False
    params = dict(model.named_parameters())
    slots = {n: tuple(torch.full_like(p, v) for v in opt.slot_init())
             for n, p in params.items()}
    losses = []
    for s in range(steps):
        ids, labels = make_batch(s, world)
        out = model(ids, labels)
        model.zero_grad()
        out["loss"].backward()
        losses.append(out["loss"].item())
        hp = opt.hyper(s + 1)
        with torch.no_grad():
            for n, p in params.items():
                g = p.grad
                if n == "emb.weight":
                    rows = torch.unique(ids.reshape(-1))
                    optim.apply_sparse_rows_(opt.kind, p.data, rows, (g * sparse_scale)[rows],
                                             slots[n], hp)
                else:
                    optim.apply_dense_(opt.kind, p.data, g, slots[n], hp)
    return losses, {n: p.detach().clone() for n, p in params.items()}


def train(world, rank, run_option="HYBRID", opt_name="adagrad", steps=4, average=True,
          sync=True, sess_config=None, ps=None, nparts=5, ckpt_dir=None, save=False,
          resource="localhost"):
    """Run the probe model through the public API; returns (losses, {name: weight},
    backend).  With `ckpt_dir` an existing checkpoint is restored first (training resumes
    at its global step) and `save=True` writes one after the last step."""
    import parallax_b200 as parallax
    from ..models.simple import MLPWithEmbedding
    model = MLPWithEmbedding(VOCAB, partitioner=parallax.get_partitioner(nparts))
    g = parallax.Graph(model, optimizer=make_opt(opt_name))
    cfg = parallax.Config(run_option=run_option, average_sparse=average,
                          sess_config=dict(sess_config or {}), search_partitions=False)
    if ps is not None:
        cfg.communication_config = parallax.CommunicationConfig(ps)
    if ckpt_dir:
        cfg.ckpt_config = parallax.CheckPointConfig(ckpt_dir=ckpt_dir)
    # under torchrun the process is a worker whatever `resource` says; a stand-alone process
    # must name ONE GPU ("localhost:0"), or the launcher would start a worker per visible GPU
    sess, nw, wid, _ = parallax.parallel_run(g, resource, sync=sync, parallax_config=cfg)
    losses = []
    for s in range(sess.engine.global_step, steps):
        ids, labels = make_batch(s, world, rank)
        loss, _ = sess.run(["loss", "train_op"], {"ids": [ids], "labels": [labels]})
        losses.append(loss[0])
    if save:
        sess.save_checkpoint()
    sd = sess.engine.state_dict()
    backend = sess.engine.backend
    sess.close()
    w = dict(sd["dense"]["master"])
    w["emb.weight"] = sd["sparse"]["emb.weight"]["weight"]
    return losses, w, backend


def check(world, rank, run_option="HYBRID", opt_name="adagrad", steps=4, average=True,
          sess_config=None, ps=None, rtol=2e-4, atol=2e-5, resource="localhost"):
    """-> {"ok", "max_abs_err", "backend"} for this rank."""
    _, w, backend = train(world, rank, run_option, opt_name, steps, average,
                          sess_config=sess_config, ps=ps, resource=resource)
    _, ref = oracle(world, steps, make_opt(opt_name), 1.0 if average else float(world))
    err, ok = 0.0, True
    for n, r in ref.items():
        got = w[n].float().cpu()
        err = max(err, float((got - r).abs().max()))
        ok = ok and bool(torch.allclose(got, r, rtol=rtol, atol=atol))
    return {"ok": ok, "max_abs_err": err, "backend": backend}
