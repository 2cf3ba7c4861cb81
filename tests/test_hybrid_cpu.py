"""BASELINE config 1: 2-layer MLP + tiny embedding, CPU/gloo world_size=2 —
plumbing, dense/sparse routing and aggregation semantics."""

import numpy as np
import pytest
import torch

import parallax_b200 as parallax
from parallax_b200.models.simple import MLPWithEmbedding
from parallax_b200 import optim
from tests.dist_utils import run_distributed

B, T, VOCAB = 8, 3, 64


def make_batch(step, world, rank=None):
    g = torch.Generator().manual_seed(100 + step)
    ids = torch.randint(0, VOCAB, (B * world, T), generator=g)
    # force duplicates inside and across workers
    ids[:, 0] = ids[0, 0]
    labels = torch.randint(0, 4, (B * world,), generator=g)
    if rank is None:
        return ids, labels
    return ids[rank * B:(rank + 1) * B], labels[rank * B:(rank + 1) * B]


def oracle(world, steps, opt, sparse_scale):
    """Single-device training on the concatenated batch; sparse grads scaled by
    `sparse_scale` (world for sum semantics, 1 for average)."""
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
                    g = g * sparse_scale
                    rows = torch.nonzero(g.abs().sum(1) > 0).squeeze(1)
                    # rows touched: those looked up (even if grad is 0)
                    rows = torch.unique(ids.reshape(-1))
                    optim.apply_sparse_rows_(opt.kind, p.data, rows, g[rows],
                                             slots[n], hp)
                else:
                    optim.apply_dense_(opt.kind, p.data, g, slots[n], hp)
    return losses, {n: p.detach().clone() for n, p in params.items()}


def worker(rank, world, run_option, average_sparse, opt_name, steps, sync=True,
           local_agg=True, nparts=None, fabric=None):
    opt = make_opt(opt_name)
    part = parallax.get_partitioner(nparts) if nparts else None
    model = MLPWithEmbedding(VOCAB, partitioner=part)
    graph = parallax.Graph(model, optimizer=opt)
    cfg = parallax.Config()
    cfg.run_option = run_option
    cfg.average_sparse = average_sparse
    if fabric:
        cfg.sess_config = {"fabric": fabric}
    cfg.communication_config = parallax.CommunicationConfig(
        parallax.PSConfig(local_aggregation=local_agg))
    sess, nw, wid, nrep = parallax.parallel_run(graph, "localhost", sync=sync,
                                                parallax_config=cfg)
    assert (nw, wid, nrep) == (world, rank, 1)
    losses = []
    for s in range(steps):
        ids, labels = make_batch(s, world, rank)
        loss, gs, _ = sess.run(["loss", "global_step", "train_op"],
                               feed_dict={"ids": [ids], "labels": [labels]})
        assert gs == [s + 1]
        losses.append(loss[0])
    sd = sess.engine.state_dict()
    sess.close()
    weights = dict(sd["dense"]["master"])
    weights["emb.weight"] = sd["sparse"]["emb.weight"]["weight"]
    return losses, weights, sess.engine.run_option


def make_opt(name):
    return {"sgd": optim.GradientDescent(0.5),
            "adagrad": optim.Adagrad(0.2, initial_accumulator_value=1.0),
            "adam": optim.Adam(0.01),
            "momentum": optim.Momentum(0.1, 0.9),
            "rmsprop": optim.RMSProp(0.01, momentum=0.5)}[name]


@pytest.mark.parametrize("run_option", ["HYBRID", "MPI", "PS"])
@pytest.mark.parametrize("opt_name", ["sgd", "adagrad"])
def test_two_workers_match_single_device_average(run_option, opt_name):
    steps = 4
    res = run_distributed(worker, 2, run_option, True, opt_name, steps)
    ref_losses, ref_w = oracle(2, steps, make_opt(opt_name), sparse_scale=1.0)
    # per-worker losses average to the full-batch loss
    mean_losses = np.mean([r[0] for r in res], axis=0)
    np.testing.assert_allclose(mean_losses, ref_losses, rtol=1e-5, atol=1e-6)
    for losses, weights, eff in res:
        assert eff == run_option
        for n, w in ref_w.items():
            torch.testing.assert_close(weights[n], w, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("opt_name", ["sgd", "adagrad", "adam"])
def test_sparse_sum_semantics_default(opt_name):
    """Default average_sparse=False: dense = mean, sparse = SUM over workers."""
    steps = 3
    res = run_distributed(worker, 2, "HYBRID", False, opt_name, steps)
    _, ref_w = oracle(2, steps, make_opt(opt_name), sparse_scale=2.0)
    for _, weights, _ in res:
        for n, w in ref_w.items():
            torch.testing.assert_close(weights[n], w, rtol=1e-5, atol=1e-6)


def test_local_aggregation_off_and_partitions():
    steps = 3
    res = run_distributed(worker, 2, "HYBRID", True, "adagrad", steps, True,
                          False, 5)
    _, ref_w = oracle(2, steps, make_opt("adagrad"), sparse_scale=1.0)
    for _, weights, _ in res:
        for n, w in ref_w.items():
            torch.testing.assert_close(weights[n], w, rtol=1e-5, atol=1e-6)


def test_library_fabric_same_results():
    """`fabric="library"`: the device-agnostic library path (NCCL on GPUs, gloo here)."""
    steps = 3
    res = run_distributed(worker, 2, "HYBRID", True, "adagrad", steps, True, True, 4,
                          "library")
    _, ref_w = oracle(2, steps, make_opt("adagrad"), sparse_scale=1.0)
    for _, weights, _ in res:
        for n, w in ref_w.items():
            torch.testing.assert_close(weights[n], w, rtol=1e-5, atol=1e-6)


def test_single_worker_inprocess():
    losses, weights, eff = worker(0, 1, "HYBRID", False, "momentum", 3)
    ref_l, ref_w = oracle(1, 3, make_opt("momentum"), 1.0)
    np.testing.assert_allclose(losses, ref_l, rtol=1e-5)
    for n, w in ref_w.items():
        torch.testing.assert_close(weights[n], w, rtol=1e-5, atol=1e-6)


def test_async_ps_runs_and_learns():
    res = run_distributed(worker, 2, "PS", False, "sgd", 6, False)
    for losses, _, _ in res:
        assert losses[-1] < losses[0]
    # both workers see identical variables after each (serialised) async step
    for n in res[0][1]:
        torch.testing.assert_close(res[0][1][n], res[1][1][n])


def test_mpi_and_hybrid_require_sync():
    model = MLPWithEmbedding(VOCAB)
    graph = parallax.Graph(model, optimizer=optim.GradientDescent(0.1))
    for opt in ("MPI", "HYBRID"):
        cfg = parallax.Config(run_option=opt)
        with pytest.raises(ValueError):
            parallax.parallel_run(graph, "localhost", sync=False,
                                  parallax_config=cfg)


def _ckpt_worker(rank, world, ckpt_dir, run_option, nparts, steps, start_step):
    """train `steps` steps from whatever checkpoint `ckpt_dir` holds; save at the end"""
    part = parallax.get_partitioner(nparts)
    model = MLPWithEmbedding(VOCAB, partitioner=part)
    graph = parallax.Graph(model, optimizer=make_opt("adam"),
                           ema=parallax.ExponentialMovingAverage(0.9, ["fc2.*"]))
    # average_sparse: with the default SUM semantics the sparse update depends on the number
    # of workers (each contributes the gradient of its own mean loss)
    cfg = parallax.Config(run_option=run_option, search_partitions=False, average_sparse=True,
                          ckpt_config=parallax.CheckPointConfig(ckpt_dir=ckpt_dir,
                                                                save_ckpt_steps=10 ** 6))
    sess, nw, wid, _ = parallax.parallel_run(graph, "localhost", parallax_config=cfg)
    assert sess.engine.global_step == start_step            # restore-on-start
    for s in range(start_step, start_step + steps):
        # every world size consumes the same global batch, split evenly
        ids, labels = make_batch(s, 2)
        n = ids.shape[0] // world
        sess.run(["loss", "train_op"], {"ids": [ids[rank * n:(rank + 1) * n]],
                                        "labels": [labels[rank * n:(rank + 1) * n]]})
    sess.save_checkpoint()
    sd = sess.engine.state_dict()
    sess.close()
    return sd


def test_checkpoint_resumes_under_other_world_size_mode_and_partitioning(tmp_path):
    """2 workers HYBRID P=3 for 3 steps → save → ONE worker PS P=5 for 3 more steps equals an
    uninterrupted 2-worker run: checkpoints hold logical tensors, not a layout."""
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    run_distributed(_ckpt_worker, 2, a, "HYBRID", 3, 3, 0)
    resumed = run_distributed(_ckpt_worker, 1, a, "PS", 5, 3, 3)[0]
    straight = run_distributed(_ckpt_worker, 2, b, "HYBRID", 3, 6, 0)[0]
    assert resumed["global_step"] == straight["global_step"] == 6
    for n, w in straight["dense"]["master"].items():
        # the mean over 2 half-batches equals the full-batch mean, for dense and (with
        # average_sparse) sparse variables alike
        torch.testing.assert_close(resumed["dense"]["master"][n], w, rtol=2e-4, atol=2e-5)
        for s_r, s_s in zip(resumed["dense"]["slots"][n], straight["dense"]["slots"][n]):
            torch.testing.assert_close(s_r, s_s, rtol=2e-4, atol=2e-5)
    torch.testing.assert_close(resumed["dense"]["ema"]["fc2.weight"],
                               straight["dense"]["ema"]["fc2.weight"], rtol=2e-4, atol=2e-5)
    torch.testing.assert_close(resumed["sparse"]["emb.weight"]["weight"],
                               straight["sparse"]["emb.weight"]["weight"], rtol=2e-4, atol=2e-5)
    for s_r, s_s in zip(resumed["sparse"]["emb.weight"]["slots"],
                        straight["sparse"]["emb.weight"]["slots"]):
        torch.testing.assert_close(s_r, s_s, rtol=2e-4, atol=2e-5)
