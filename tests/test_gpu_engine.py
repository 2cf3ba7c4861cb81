"""End-to-end engine on the NVLink fabric (1 GPU) vs the host-fabric oracle."""
import numpy as np
import pytest
import torch

import parallax_b200 as parallax
from parallax_b200 import optim
from parallax_b200.models.simple import MLPWithEmbedding

pytestmark = pytest.mark.gpu


def _run(fabric, run_option, opt, steps, compute_dtype=None, sync=True,
         dense_update="sharded", clip=None, same_batch=False, graph=False):
    model = MLPWithEmbedding(64, partitioner=parallax.get_partitioner(3))
    rules = [parallax.ClipByGlobalNorm(clip, params=["fc1.*", "fc2.*"])] if clip else []
    graph = parallax.Graph(model, optimizer=opt, grad_rules=rules,
                           ema=parallax.ExponentialMovingAverage(0.9, ["fc2.*"]))
    sc = {"fabric": fabric, "dense_update": dense_update, "cuda_graph": graph}
    if compute_dtype:
        sc["compute_dtype"] = compute_dtype
    cfg = parallax.Config(run_option=run_option, sess_config=sc)
    sess, *_ = parallax.parallel_run(graph, "localhost:0", sync=sync,
                                     parallax_config=cfg)
    g = torch.Generator().manual_seed(0)
    losses = []
    for s in range(steps):
        if same_batch:
            g = torch.Generator().manual_seed(0)
        ids = torch.randint(0, 64, (8, 3), generator=g)
        ids[:, 0] = 5
        labels = torch.randint(0, 4, (8,), generator=g)
        loss, _ = sess.run(["loss", "train_op"], {"ids": [ids], "labels": [labels]})
        losses.append(loss[0])
    sd = sess.engine.state_dict()
    sess.close()
    return losses, sd


@pytest.mark.parametrize("run_option", ["HYBRID", "MPI", "PS"])
@pytest.mark.parametrize("opt_name", ["sgd", "adagrad", "adam"])
def test_engine_matches_host_oracle(run_option, opt_name):
    mk = lambda: {"sgd": optim.GradientDescent(0.3), "adagrad": optim.Adagrad(0.2, 1.0),
                  "adam": optim.Adam(0.01)}[opt_name]
    l_ref, sd_ref = _run("host", run_option, mk(), 5, clip=0.5)
    l_nv, sd_nv = _run("nvlink", run_option, mk(), 5, clip=0.5)
    np.testing.assert_allclose(l_nv, l_ref, rtol=1e-4, atol=1e-5)
    for n, w in sd_ref["dense"]["master"].items():
        torch.testing.assert_close(sd_nv["dense"]["master"][n], w, rtol=1e-4, atol=1e-5)
    for n, w in sd_ref["dense"]["ema"].items():
        torch.testing.assert_close(sd_nv["dense"]["ema"][n], w, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(sd_nv["sparse"]["emb.weight"]["weight"],
                               sd_ref["sparse"]["emb.weight"]["weight"],
                               rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("opt_name", ["ftrl", "centered_rmsprop", "adagrad_da"])
def test_engine_extended_optimizers_match_host_oracle(opt_name):
    """The long tail of the reference's recognised update ops runs on the fused kernels
    (rule family 1), dense and sparse."""
    mk = lambda: {"ftrl": optim.Ftrl(0.2, l1_regularization_strength=0.001),
                  "centered_rmsprop": optim.CenteredRMSProp(0.01, momentum=0.5, epsilon=1e-3),
                  "adagrad_da": optim.AdagradDA(0.2, l1_regularization_strength=0.001)}[opt_name]
    l_ref, sd_ref = _run("host", "HYBRID", mk(), 5)
    l_nv, sd_nv = _run("nvlink", "HYBRID", mk(), 5)
    np.testing.assert_allclose(l_nv, l_ref, rtol=2e-4, atol=2e-5)
    for n, w in sd_ref["dense"]["master"].items():
        torch.testing.assert_close(sd_nv["dense"]["master"][n], w, rtol=2e-4, atol=2e-5)
    torch.testing.assert_close(sd_nv["sparse"]["emb.weight"]["weight"],
                               sd_ref["sparse"]["emb.weight"]["weight"], rtol=2e-4, atol=2e-5)


def test_engine_replicated_update_and_async():
    l_ref, sd_ref = _run("host", "MPI", optim.Momentum(0.1, 0.9), 4)
    l_nv, sd_nv = _run("nvlink", "MPI", optim.Momentum(0.1, 0.9), 4,
                       dense_update="replicated")
    np.testing.assert_allclose(l_nv, l_ref, rtol=1e-4, atol=1e-5)
    l_ref, sd_ref = _run("host", "PS", optim.Adagrad(0.2, 1.0), 4, sync=False)
    l_nv, sd_nv = _run("nvlink", "PS", optim.Adagrad(0.2, 1.0), 4, sync=False)
    np.testing.assert_allclose(l_nv, l_ref, rtol=1e-4, atol=1e-5)


def test_engine_bf16_trains():
    losses, _ = _run("nvlink", "HYBRID", optim.Adagrad(0.2, 1.0), 12,
                     compute_dtype="bf16", same_batch=True)
    assert losses[-1] < losses[0]


def test_engine_cuda_graph_matches_eager():
    """The captured+replayed step must produce the same trajectory as eager."""
    l_e, sd_e = _run("nvlink", "HYBRID", optim.Adagrad(0.2, 1.0), 8, clip=0.5)
    l_g, sd_g = _run("nvlink", "HYBRID", optim.Adagrad(0.2, 1.0), 8, clip=0.5,
                     graph=True)
    np.testing.assert_allclose(l_g, l_e, rtol=1e-5, atol=1e-6)
    for n, w in sd_e["dense"]["master"].items():
        torch.testing.assert_close(sd_g["dense"]["master"][n], w, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(sd_g["sparse"]["emb.weight"]["weight"],
                               sd_e["sparse"]["emb.weight"]["weight"],
                               rtol=1e-5, atol=1e-6)


def test_smoke_entry():
    import __graft_entry__ as ge
    ge.smoke()


def test_engine_repartition_in_place_nvlink():
    """Online partition search support: tables are re-sharded in place on the
    fabric, keeping weights and optimizer slots; training continues."""
    model = MLPWithEmbedding(64, partitioner=parallax.get_partitioner(3))
    graph = parallax.Graph(model, optimizer=optim.Adagrad(0.2, 1.0))
    cfg = parallax.Config(sess_config={"fabric": "nvlink", "cuda_graph": True})
    sess, *_ = parallax.parallel_run(graph, "localhost:0", parallax_config=cfg)
    g = torch.Generator().manual_seed(0)

    def feed():
        return {"ids": [torch.randint(0, 64, (8, 3), generator=g)],
                "labels": [torch.randint(0, 4, (8,), generator=g)]}
    for _ in range(5):
        sess.run(["loss", "train_op"], feed())
    eng = sess.engine
    before = eng.state_dict()["sparse"]["emb.weight"]
    eng.repartition(7)
    assert eng.tables["emb.weight"].layout.P == 7
    after = eng.state_dict()["sparse"]["emb.weight"]
    torch.testing.assert_close(after["weight"], before["weight"])
    torch.testing.assert_close(after["slots"][0], before["slots"][0])
    losses = [sess.run(["loss", "train_op"], feed())[0][0] for _ in range(5)]
    assert np.isfinite(losses).all()
    sess.close()


def test_autotuner_scores_candidates_under_graph_replay(monkeypatch, tmp_path):
    """PARALLAX_AUTOTUNE=1: every candidate setting is applied, the step re-captured and
    scored by device-timed graph replays; the job settles on the best setting and keeps
    training correctly."""
    log = tmp_path / "autotune.csv"
    monkeypatch.setenv("PARALLAX_AUTOTUNE", "1")
    monkeypatch.setenv("PARALLAX_AUTOTUNE_LOG", str(log))
    model = MLPWithEmbedding(64, partitioner=parallax.get_partitioner(3))
    graph = parallax.Graph(model, optimizer=optim.Adagrad(0.2, 1.0))
    cfg = parallax.Config(run_option="HYBRID",
                          sess_config={"fabric": "nvlink", "cuda_graph": True})
    sess, *_ = parallax.parallel_run(graph, "localhost:0", sync=True, parallax_config=cfg)
    tuner = sess.engine.autotuner
    assert tuner is not None
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, 64, (8, 3), generator=g)
    labels = torch.randint(0, 4, (8,), generator=g)
    losses = []
    for s in range(400):
        loss, _ = sess.run(["loss", "train_op"], {"ids": [ids], "labels": [labels]})
        losses.append(loss[0])
        if tuner.done and s > 30:
            break
    assert tuner.done and tuner.best is not None and tuner.best[0] > 0
    assert sess.engine.fabric.max_blocks == tuner.best[1]["comm_blocks"]
    for _ in range(8):
        loss, _ = sess.run(["loss", "train_op"], {"ids": [ids], "labels": [labels]})
    assert getattr(sess.engine, "graph_captured", False)
    assert loss[0] < losses[0]
    lines = log.read_text().strip().splitlines()
    assert len(lines) >= 5 and "comm_blocks=" in lines[0] and "early_push=" in lines[0]
    sess.close()
