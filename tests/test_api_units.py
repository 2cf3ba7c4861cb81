"""Unit tests of the API surface: config validation, resource (de)serialisation,
shard arithmetic, partition search maths, mode degeneration, checkpoint
save/restore-on-start/resume-under-a-different-layout, profile steps."""
import json
import os

import numpy as np
import pytest
import torch

import parallax_b200 as parallax
from parallax_b200 import optim, shard
from parallax_b200.analyzer import analyze, greedy_load_balance
from parallax_b200.models.simple import MLPWithEmbedding
from parallax_b200.partitions import SearchState, find_optimal_p, fit_cost_model
from parallax_b200.resource import (parse_resource_info, serialize_resource_info,
                                    deserialize_resource_info, worker_layout,
                                    get_cluster_str_for_hosts)


def test_public_api_names():
    for n in ["parallel_run", "Config", "PSConfig", "MPIConfig", "CommunicationConfig",
              "CheckPointConfig", "ProfileConfig", "get_partitioner", "shard", "log"]:
        assert hasattr(parallax, n)
    import parallax as alias
    assert alias.parallel_run is parallax.parallel_run
    assert alias.shard.shard is parallax.shard.shard


def test_config_defaults_and_validation():
    c = parallax.Config()
    assert c.run_option == "HYBRID" and c.average_sparse is False
    assert c.search_partitions is True
    ps = c.communication_config.ps_config
    assert (ps.protocol, ps.replicate_variables, ps.local_aggregation) == ("grpc", True, True)
    assert ps.boundary_among_servers and ps.boundary_between_workers_and_servers
    assert parallax.Config(run_option="AR").normalized_run_option() == "MPI"
    cfg = parallax.Config()
    cfg.run_option = "PS",                       # the reference docs' stray tuple
    assert cfg.normalized_run_option() == "PS"
    with pytest.raises(ValueError):
        parallax.Config(run_option="TP").normalized_run_option()
    m = parallax.MPIConfig(mpirun_options=["-x", "NCCL_DEBUG=INFO", "-mca", "btl", "^openib"])
    assert m.mpirun_options == "-x NCCL_DEBUG=INFO -mca btl ^openib"
    assert m.exported_env() == {"NCCL_DEBUG": "INFO"}
    with pytest.raises(AssertionError):
        parallax.ProfileConfig(profile_dir="/tmp/x", profile_steps=[1], profile_range=(1, 2))
    with pytest.raises(AssertionError):
        parallax.CommunicationConfig(ps_config="nope")


def test_resource_info_roundtrip(tmp_path):
    f = tmp_path / "resource_info"
    f.write_text("hostA:0,1,2\nhostB:4,5\n\n")
    info = parse_resource_info(str(f), "HYBRID")
    assert info["master"][0]["hostname"] == "hostA"
    assert [w["gpus"] for w in info["worker"]] == [[0, 1, 2], [4, 5]]
    assert len(info["ps"]) == 2                                   # one PS entry per host
    assert len(info["worker"][0]["port"]) == 3                    # HYBRID: port per GPU
    assert len(parse_resource_info(str(f), "PS")["worker"][0]["port"]) == 1
    back = deserialize_resource_info(serialize_resource_info(info))
    assert back == info
    lay = worker_layout(info)
    assert [(h, m, l, g) for h, m, l, g in lay] == [
        ("hostA", 0, 0, 0), ("hostA", 0, 1, 1), ("hostA", 0, 2, 2),
        ("hostB", 1, 0, 4), ("hostB", 1, 1, 5)]
    assert get_cluster_str_for_hosts(info["worker"], True) == "hostA:3,hostB:2"


def test_shard_arithmetic():
    shard.reset()
    ns, sid = shard.create_num_shards_and_shard_id()
    with pytest.raises(ValueError):
        shard.create_num_shards_and_shard_id()
    ds = shard.shard(list(range(20)))
    assert list(ds) == list(range(20))                 # (1, 0) before parallel_run
    assert shard.update_shard_values_for_worker(4, 3, 1) == (4, 3)
    assert list(ds) == [3, 7, 11, 15, 19] and len(ds) == 5 and ds[1] == 7
    assert shard.update_shard_values_for_worker(2, 1, 3) == (6, 3)   # in-graph replicas
    samp = shard.DistributedShardSampler(list(range(12)))
    assert list(samp) == [3, 9]


def test_partition_search_math():
    f = lambda p: 0.01 * (p - 1) + 4.0 / p + 1.0
    a, b, c = fit_cost_model([2, 4, 8, 16, 32], [f(p) for p in [2, 4, 8, 16, 32]])
    assert abs(a - 0.01) < 1e-6 and abs(b - 4.0) < 1e-6 and abs(c - 1.0) < 1e-6
    assert find_optimal_p([8, 16, 32], [f(8), f(16), f(32)]) == 20
    s = SearchState(8)
    keep = True
    while keep:
        keep, p = s.report(f(s.p_to_test))
    assert s.p_list == [8, 16, 32] and p == 20
    # a candidate that crashes (OOM) doubles the minimum
    s = SearchState(4)
    keep, p = s.report(0.0, alive=False)
    assert keep and p == 8 and s.min_partitions == 8
    part = parallax.get_partitioner(16)
    assert part.num_partitions == 16 and os.environ["PARALLAX_MIN_PARTITIONS"] == "16"
    os.environ["PARALLAX_PARTITIONS"] = "64"
    try:
        assert parallax.get_partitioner(16).num_partitions == 64
    finally:
        del os.environ["PARALLAX_PARTITIONS"]


def test_analyzer_and_mode_degeneration():
    a = analyze(MLPWithEmbedding(64), world=4)
    assert [v.name for v in a.sparse] == ["emb.weight"] and len(a.dense) == 4
    assert a.variables["emb.weight"].partitions == 4
    assert a.effective_run_option("HYBRID") == "HYBRID"
    dense_only = analyze(torch.nn.Linear(3, 3))
    assert dense_only.effective_run_option("HYBRID") == "MPI"
    sparse_only = analyze(parallax.nn.Embedding(10, 4))
    assert sparse_only.effective_run_option("HYBRID") == "PS"
    assert greedy_load_balance([10, 1, 1, 1, 8], 2) == [0, 1, 1, 1, 1]


def _train(cfg, steps, seed=0, model_kw=None):
    torch.manual_seed(seed)
    model = MLPWithEmbedding(64, partitioner=parallax.get_partitioner(3), **(model_kw or {}))
    g = parallax.Graph(model, optimizer=optim.Adagrad(0.2, 1.0),
                       ema=parallax.ExponentialMovingAverage(0.9, ["fc2.*"]))
    sess, *_ = parallax.parallel_run(g, "localhost", parallax_config=cfg)
    gen = torch.Generator().manual_seed(1 + sess.engine.global_step)
    out = []
    for _ in range(steps):
        ids = torch.randint(0, 64, (4, 3), generator=gen)
        labels = torch.randint(0, 4, (4,), generator=gen)
        out.append(sess.run(["loss", "global_step", "train_op"],
                            {"ids": [ids], "labels": [labels]}))
    return sess, out


def test_checkpoint_save_and_restore_on_start(tmp_path):
    ck = parallax.CheckPointConfig(ckpt_dir=str(tmp_path / "ckpt"), save_ckpt_steps=2)
    cfg = parallax.Config(ckpt_config=ck, sess_config={"fabric": "host"})
    sess, out = _train(cfg, 5)
    files = sorted(os.listdir(tmp_path / "ckpt"))
    assert "checkpoint" in files and "model.ckpt-2.pt" in files and "model.ckpt-4.pt" in files
    assert open(tmp_path / "ckpt" / "checkpoint").read() == "model.ckpt-4.pt"
    ref = sess.engine.state_dict()
    sess.close()
    # a new job with the same ckpt_dir resumes from step 4 — under a different run option
    cfg2 = parallax.Config(run_option="PS", ckpt_config=ck, sess_config={"fabric": "host"})
    sess2, out2 = _train(cfg2, 1)
    assert out2[0][1] == [5]
    sd = torch.load(tmp_path / "ckpt" / "model.ckpt-4.pt", weights_only=False)
    assert sd["global_step"] == 4
    assert set(sd["sparse"]["emb.weight"]) == {"weight", "slots"}
    assert sd["sparse"]["emb.weight"]["weight"].shape == (64, 8)
    assert "fc2.weight" in sd["dense"]["ema"]
    sess2.close()


def test_profile_steps_dump(tmp_path):
    pc = parallax.ProfileConfig(profile_dir=str(tmp_path / "prof"), profile_steps=[1, 3])
    cfg = parallax.Config(profile_config=pc, sess_config={"fabric": "host"})
    sess, _ = _train(cfg, 4)
    sess.close()
    import socket
    d = tmp_path / "prof" / socket.gethostname() / "worker:0" / "run_meta"
    got = sorted(os.listdir(d))
    assert "run_meta_1.json" in got and "run_meta_3.json" in got and "run_meta_2.json" not in got
    assert "worker:0" in open(tmp_path / "prof" / socket.gethostname() / "task_info").read()


def test_session_feed_fetch_contract():
    cfg = parallax.Config(sess_config={"fabric": "host"})
    sess, out = _train(cfg, 1)
    loss, gs, train = out[0]
    assert isinstance(loss, list) and len(loss) == 1 and gs == [1] and train == [None]
    ids = torch.randint(0, 64, (4, 3))
    labels = torch.zeros(4, dtype=torch.long)
    r = sess.run({"l": "loss", "both": ["logits", "global_step"]},
                 {"ids": [ids], "labels": [labels]})
    assert set(r) == {"l", "both"} and r["both"][0][0].shape == (4, 4) and r["both"][1] == [1]
    with pytest.raises(ValueError):
        sess.run("loss", {"ids": [ids, ids], "labels": [labels]})
    with pytest.raises(KeyError):
        sess.run("loss", {"nope": [ids]})
    with pytest.raises(KeyError):
        sess.run("not_a_tensor", {"ids": [ids], "labels": [labels]})
    sess.close()


def test_inprocess_repartition_preserves_state_and_search_runs():
    from parallax_b200.partitions import search_inprocess
    cfg = parallax.Config(sess_config={"fabric": "host"})
    sess, _ = _train(cfg, 3)
    eng = sess.engine
    before = eng.state_dict()["sparse"]["emb.weight"]
    eng.repartition(7)
    assert eng.tables["emb.weight"].layout.P == 7
    after = eng.state_dict()["sparse"]["emb.weight"]
    torch.testing.assert_close(after["weight"], before["weight"])
    torch.testing.assert_close(after["slots"][0], before["slots"][0])
    gen = torch.Generator().manual_seed(3)

    def feed():
        return {"ids": [torch.randint(0, 64, (4, 3), generator=gen)],
                "labels": [torch.randint(0, 4, (4,), generator=gen)]}
    l0 = sess.run(["loss", "train_op"], feed())[0][0]
    p = search_inprocess(sess, feed, min_partitions=2, warmup=1, test=2)
    assert p >= 2 and eng.tables["emb.weight"].layout.P == p
    assert np.isfinite(sess.run(["loss", "train_op"], feed())[0][0]) and np.isfinite(l0)
    sess.close()


def test_profile_range_and_ckpt_secs(tmp_path):
    import socket
    import time
    pc = parallax.ProfileConfig(profile_dir=str(tmp_path / "prof"), profile_range=(1, 3))
    ck = parallax.CheckPointConfig(ckpt_dir=str(tmp_path / "ck"), save_ckpt_secs=0.0001)
    cfg = parallax.Config(profile_config=pc, ckpt_config=ck, sess_config={"fabric": "host"})
    sess, _ = _train(cfg, 4)
    sess.close()
    d = tmp_path / "prof" / socket.gethostname() / "worker:0" / "run_meta"
    got = sorted(f for f in os.listdir(d) if f.endswith(".json"))
    assert got == ["run_meta_1.json", "run_meta_2.json"]          # [start, end)
    saved = [f for f in os.listdir(tmp_path / "ck") if f.startswith("model.ckpt-")]
    assert len(saved) >= 2                                        # time-based saving fired


def test_div_partition_strategy_matches_mod_results():
    """The two TF partition strategies only differ in row placement."""
    outs = []
    for strat in ("mod", "div"):
        torch.manual_seed(0)
        model = MLPWithEmbedding(61, partitioner=parallax.get_partitioner(4, strategy=strat))
        g = parallax.Graph(model, optimizer=optim.Adagrad(0.2, 1.0))
        cfg = parallax.Config(sess_config={"fabric": "host"})
        sess, *_ = parallax.parallel_run(g, "localhost", parallax_config=cfg)
        gen = torch.Generator().manual_seed(5)
        for _ in range(3):
            ids = torch.randint(0, 61, (4, 3), generator=gen)
            sess.run(["loss", "train_op"], {"ids": [ids],
                                            "labels": [torch.randint(0, 4, (4,), generator=gen)]})
        outs.append(sess.engine.state_dict()["sparse"]["emb.weight"]["weight"])
        assert sess.engine.tables["emb.weight"].layout.strategy == strat
        sess.close()
    torch.testing.assert_close(outs[0], outs[1])


def test_inspect_checkpoint_tool(tmp_path, capsys):
    import io
    from parallax_b200.models.lm1b import LM1B, lm1b_graph
    from parallax_b200.tools import inspect_checkpoint as ic
    torch.manual_seed(0)
    m = LM1B(vocab_size=50, emb_size=8, state_size=16, projected_size=8, num_sampled=8,
             num_steps=3, num_shards=2, keep_prob=1.0)
    cfg = parallax.Config(sess_config={"fabric": "host"},
                          ckpt_config=parallax.CheckPointConfig(ckpt_dir=str(tmp_path),
                                                                save_ckpt_steps=2))
    sess, *_ = parallax.parallel_run(lm1b_graph(m, batch_size=4), "localhost", parallax_config=cfg)
    x = torch.randint(0, 50, (4, 3))
    for _ in range(2):
        sess.run(["loss", "train_op"], {"x": [x], "y": [x]})
    sess.close()
    path, sd = ic.load(str(tmp_path))
    assert path.endswith("model.ckpt-2.pt")
    kinds = {(k, n) for k, n, _ in ic.entries(sd)}
    assert ("dense", "W") in kinds and ("dense-ema", "W_P") in kinds
    assert ("sparse", "emb.weight") in kinds and ("sparse-slot0", "softmax_w.weight") in kinds
    buf = io.StringIO()
    total = ic.summarize(sd, buf)
    assert "global_step: 2" in buf.getvalue() and total > 0
    plain, ema = ic.to_state_dict(sd), ic.to_state_dict(sd, use_ema=True)
    assert plain["emb.weight"].shape == (50, 8) and not torch.equal(plain["W"], ema["W"])
    fresh = LM1B(vocab_size=50, emb_size=8, state_size=16, projected_size=8, num_sampled=8,
                 num_steps=3, num_shards=2, keep_prob=1.0)
    missing, unexpected = fresh.load_state_dict(plain, strict=False)
    assert not missing and not unexpected
    assert ic.main([str(tmp_path), "--tensor", "B"]) == 0 and "dense B" in capsys.readouterr().out
    out = str(tmp_path / "plain.pt")
    assert ic.main([str(tmp_path), "--to_state_dict", out, "--ema"]) == 0
    assert torch.equal(torch.load(out)["W_P"], ema["W_P"])
    with pytest.raises(FileNotFoundError):
        ic.load(str(tmp_path / "empty_dir_that_does_not_exist"))


@pytest.mark.parametrize("mode", ["sum", "mean", "max"])
def test_embedding_bag_matches_torch_and_trains_sharded(mode):
    torch.manual_seed(0)
    bag = parallax.nn.EmbeddingBag(30, 6, mode=mode, partitioner=parallax.get_partitioner(3))
    ref = torch.nn.EmbeddingBag(30, 6, mode=mode)
    with torch.no_grad():
        ref.weight.copy_(bag.weight)
    flat = torch.tensor([1, 2, 4, 5, 4, 3, 2, 9, 7])
    offsets = torch.tensor([0, 4, 4, 7])                       # the second bag is empty
    torch.testing.assert_close(bag(flat, offsets), ref(flat, offsets))
    two_d = torch.tensor([[1, 2, 3], [4, 4, 9]])
    torch.testing.assert_close(bag(two_d), ref(two_d))
    if mode == "sum":
        w = torch.rand(9)
        torch.testing.assert_close(bag(flat, offsets, per_sample_weights=w),
                                   ref(flat, offsets, per_sample_weights=w))
    else:
        with pytest.raises(NotImplementedError):
            bag(flat, offsets, per_sample_weights=torch.rand(9))
    with pytest.raises(ValueError):
        bag(flat)
    # under the engine the bag's table is a partitioned sparse variable
    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.bag = parallax.nn.EmbeddingBag(30, 6, mode=mode,
                                                partitioner=parallax.get_partitioner(3))
            self.fc = torch.nn.Linear(6, 2)

        def forward(self, ids, labels):
            return {"loss": torch.nn.functional.cross_entropy(self.fc(self.bag(ids)), labels)}
    net = Net()
    sess, *_ = parallax.parallel_run(
        parallax.Graph(net, optimizer=parallax.optim.Adagrad(0.5, 0.1)), "localhost",
        parallax_config=parallax.Config(sess_config={"fabric": "host"}, search_partitions=False))
    try:
        assert list(sess.engine.tables) == ["bag.table.weight"]
        assert sess.engine.tables["bag.table.weight"].layout.P == 3
        ids = torch.randint(0, 30, (16, 4))
        labels = (ids.sum(1) % 2).long()
        losses = [sess.run(["loss", "train_op"], {"ids": [ids], "labels": [labels]})[0][0]
                  for _ in range(40)]
        assert losses[-1] < 0.7 * losses[0]
    finally:
        sess.close()
    with pytest.raises(NotImplementedError, match="parallax.nn.EmbeddingBag"):
        from parallax_b200.analyzer import analyze
        analyze(torch.nn.Sequential(torch.nn.EmbeddingBag(5, 2, sparse=True)))


def test_remaining_tf_optimizers_math():
    """Adadelta / FTRL / proximal SGD+Adagrad / AdagradDA / centered RMSProp follow TF's
    Apply* kernels (`tensorflow/core/kernels/training_ops.cc`)."""
    from parallax_b200 import optim
    g = torch.tensor([0.5, -2.0, 0.01])
    # Adadelta = torch.optim.Adadelta (same recurrences)
    w = torch.tensor([1.0, -1.0, 0.3])
    ref = w.clone().requires_grad_(True)
    topt = torch.optim.Adadelta([ref], lr=0.7, rho=0.9, eps=1e-6)
    spec = optim.Adadelta(0.7, rho=0.9, epsilon=1e-6)
    slots = tuple(torch.full_like(w, v) for v in spec.slot_init())
    for step in (1, 2, 3):
        ref.grad = g.clone()
        topt.step()
        optim.apply_dense_("adadelta", w, g, slots, spec.hyper(step))
    torch.testing.assert_close(w, ref.detach())
    # proximal SGD: soft threshold then L2 shrink
    w = torch.tensor([1.0, -1.0, 0.001])
    spec = optim.ProximalGradientDescent(0.1, l1_regularization_strength=0.5,
                                         l2_regularization_strength=2.0)
    optim.apply_dense_("proximal_sgd", w, g, (), spec.hyper(1))
    prox = torch.tensor([1.0 - 0.05, -1.0 + 0.2, 0.001 - 0.001])
    want = torch.sign(prox) * (prox.abs() - 0.05).clamp(min=0) / 1.2
    torch.testing.assert_close(w, want)
    assert w[2] == 0.0                                   # L1 produces exact zeros
    # proximal Adagrad with l1 = l2 = 0 is Adagrad
    w1, w2 = torch.tensor([1.0, -1.0, 0.3]), torch.tensor([1.0, -1.0, 0.3])
    a1, a2 = torch.full_like(w1, 0.1), torch.full_like(w2, 0.1)
    optim.apply_dense_("proximal_adagrad", w1, g, (a1,), optim.ProximalAdagrad(0.2).hyper(1))
    optim.apply_dense_("adagrad", w2, g, (a2,), optim.Adagrad(0.2, 0.1).hyper(1))
    torch.testing.assert_close(w1, w2)
    # FTRL, first step from w = 0 with lr_power -0.5: closed form
    spec = optim.Ftrl(0.5, initial_accumulator_value=0.1, l1_regularization_strength=0.1,
                      l2_regularization_strength=0.01)
    w = torch.zeros(3)
    slots = tuple(torch.full_like(w, v) for v in spec.slot_init())
    optim.apply_dense_("ftrl", w, g, slots, spec.hyper(1))
    n = 0.1 + g * g
    want = torch.where(g.abs() > 0.1, (torch.sign(g) * 0.1 - g) / (n.sqrt() / 0.5 + 0.02),
                       torch.zeros(3))
    torch.testing.assert_close(w, want)
    assert w[2] == 0.0 and torch.equal(slots[0], n) and torch.equal(slots[1], g)
    with pytest.raises(ValueError):
        optim.Ftrl(0.1, learning_rate_power=0.5)
    # AdagradDA: w = -lr·shrink(Σg) / (l2·t·lr + sqrt(Σg²))
    spec = optim.AdagradDA(0.3, initial_gradient_squared_accumulator_value=0.1,
                           l1_regularization_strength=0.2, l2_regularization_strength=0.5)
    w = torch.tensor([9.0, 9.0, 9.0])
    slots = tuple(torch.full_like(w, v) for v in spec.slot_init())
    for step in (1, 2):
        optim.apply_dense_("adagrad_da", w, g, slots, spec.hyper(step))
    gs, gg = 2 * g, 0.1 + 2 * g * g
    want = -0.3 * torch.sign(gs) * (gs.abs() - 0.2 * 2).clamp(min=0) / (0.5 * 2 * 0.3 + gg.sqrt())
    torch.testing.assert_close(w, want)
    # centered RMSProp subtracts the squared running mean
    spec = optim.CenteredRMSProp(0.01, decay=0.9, momentum=0.5, epsilon=1e-3)
    w = torch.tensor([1.0, -1.0, 0.3])
    slots = tuple(torch.zeros(3) for _ in spec.slot_init())
    optim.apply_dense_("centered_rmsprop", w, g, slots, spec.hyper(1))
    ms, mg = 0.1 * g * g, 0.1 * g
    torch.testing.assert_close(w, torch.tensor([1.0, -1.0, 0.3]) -
                               0.01 * g / (ms - mg * mg + 1e-3).sqrt())
    assert len(spec.slot_init()) == 3
    with pytest.raises(NotImplementedError, match="no fused kernel"):
        optim.require_fused("lion", "NVLink fabric")
    for kind in optim.KINDS + optim.EXT_KINDS:       # every recognised update op is fused
        optim.require_fused(kind, "NVLink fabric")


@pytest.mark.parametrize("name", ["adadelta", "ftrl", "proximal_sgd", "proximal_adagrad",
                                  "adagrad_da", "centered_rmsprop"])
def test_remaining_tf_optimizers_train_dense_and_sparse(name):
    from parallax_b200 import optim
    from parallax_b200.models.simple import MLPWithEmbedding
    mk = {"adadelta": lambda: optim.Adadelta(1.0, epsilon=1e-2), "ftrl": lambda: optim.Ftrl(0.5),
          "proximal_sgd": lambda: optim.ProximalGradientDescent(0.5, 1e-4, 1e-4),
          "proximal_adagrad": lambda: optim.ProximalAdagrad(0.5, 0.1,
                                                            l1_regularization_strength=1e-4),
          "adagrad_da": lambda: optim.AdagradDA(0.5),
          "centered_rmsprop": lambda: optim.CenteredRMSProp(0.01, momentum=0.5, epsilon=1e-3)}[name]
    torch.manual_seed(0)
    model = MLPWithEmbedding(32, partitioner=parallax.get_partitioner(2))
    sess, *_ = parallax.parallel_run(
        parallax.Graph(model, optimizer=mk()), "localhost",
        parallax_config=parallax.Config(sess_config={"fabric": "host"}, search_partitions=False))
    try:
        g = torch.Generator().manual_seed(0)
        ids = torch.randint(0, 32, (16, 3), generator=g)
        labels = torch.randint(0, 4, (16,), generator=g)
        losses = [sess.run(["loss", "train_op"], {"ids": [ids], "labels": [labels]})[0][0]
                  for _ in range(60)]
        # (FTRL and dual averaging rebuild the weights from accumulated gradients, i.e. they
        # forget a non-zero initialisation at their first step — TF's kernels do too)
        bound = 0.95 if name in ("ftrl", "adagrad_da") else 0.9
        assert np.isfinite(losses).all() and losses[-1] < bound * losses[0], (name, losses[::10])
        sd = sess.engine.state_dict()
        nslots = optim.NUM_SLOTS[name]
        assert len(sd["dense"]["slots"]["fc1.weight"]) == nslots
        assert len(sd["sparse"]["emb.weight"]["slots"]) == nslots
    finally:
        sess.close()


def test_learning_rate_schedules_follow_tf_formulas():
    S = parallax.optim.schedules
    f = S.exponential_decay(0.1, 100, 0.5, staircase=True)
    assert [f(1), f(100), f(101), f(201)] == [0.1, 0.1, 0.05, 0.025]
    assert abs(S.exponential_decay(0.1, 100, 0.5)(51) - 0.1 * 0.5 ** 0.5) < 1e-12
    assert abs(S.natural_exp_decay(1.0, 10, 0.5)(11) - np.exp(-0.5)) < 1e-12
    assert abs(S.inverse_time_decay(1.0, 10, 0.5, staircase=True)(26) - 0.5) < 1e-12
    p = S.piecewise_constant([10, 20], [1.0, 0.5, 0.1])
    assert [p(1), p(11), p(12), p(21), p(22), p(999)] == [1.0, 1.0, 0.5, 0.5, 0.1, 0.1]
    with pytest.raises(ValueError):
        S.piecewise_constant([10], [1.0])
    q = S.polynomial_decay(1.0, 100, end_learning_rate=0.1, power=2.0)
    assert abs(q(51) - (0.9 * 0.25 + 0.1)) < 1e-12 and q(1000) == 0.1
    c = S.polynomial_decay(1.0, 100, end_learning_rate=0.0, cycle=True)
    assert abs(c(151) - (1 - 150 / 200.0)) < 1e-12
    cos = S.cosine_decay(2.0, 100, alpha=0.1)
    assert abs(cos(1) - 2.0) < 1e-12 and abs(cos(51) - 2.0 * (0.9 * 0.5 + 0.1)) < 1e-12
    assert abs(cos(500) - 0.2) < 1e-12
    w = S.warmup(S.piecewise_constant([100], [1.0, 0.1]), 10, start_factor=0.1)
    assert abs(w(1) - 0.1) < 1e-12 and abs(w(6) - 0.55) < 1e-12 and w(11) == 1.0 and w(200) == 0.1
    # a schedule is accepted wherever a learning rate is
    opt = parallax.optim.Momentum(S.exponential_decay(0.1, 10, 0.5, staircase=True), 0.9)
    assert opt.hyper(1)[0] == 0.1 and opt.hyper(11)[0] == 0.05


def test_clip_by_value_rule():
    from parallax_b200.models.simple import LinearRegression
    torch.manual_seed(0)
    m = LinearRegression(2)
    with torch.no_grad():
        m.linear.weight.fill_(0.0)
        m.linear.bias.fill_(0.0)
    g = parallax.Graph(m, optimizer=parallax.optim.GradientDescent(1.0),
                       grad_rules=[parallax.ClipByValue(0.25, params=["linear.weight"])])
    sess, *_ = parallax.parallel_run(g, "localhost",
                                     parallax_config=parallax.Config(sess_config={"fabric": "host"}))
    try:
        x = torch.tensor([[10.0, 0.01], [10.0, 0.01]])
        y = torch.tensor([100.0, 100.0])
        sess.run(["loss", "train_op"], {"x": [x], "y": [y]})
        sd = sess.engine.state_dict()["dense"]["master"]
        # d loss/d w = 2·mean((0-100)·x) = [-2000, -2]: both clamp to -0.25 ⇒ w = +0.25
        torch.testing.assert_close(sd["linear.weight"].view(-1), torch.tensor([0.25, 0.25]))
        assert abs(float(sd["linear.bias"]) - 200.0) < 1e-3        # bias is not clipped
    finally:
        sess.close()


def test_horovod_environment_names_are_honoured():
    from parallax_b200 import consts
    env = {"HOROVOD_TIMELINE": "/tmp/t.json", "HOROVOD_STALL_CHECK_TIME_SECONDS": "5",
           "PARALLAX_STALL_CHECK_TIME_SECONDS": "9", "HOROVOD_CACHE_CAPACITY": "64"}
    adopted = consts.adopt_horovod_env(env)
    assert env["PARALLAX_TIMELINE"] == "/tmp/t.json" and env["PARALLAX_CACHE_CAPACITY"] == "64"
    assert env["PARALLAX_STALL_CHECK_TIME_SECONDS"] == "9"        # an explicit setting wins
    assert sorted(adopted) == ["PARALLAX_CACHE_CAPACITY", "PARALLAX_TIMELINE"]
    # knobs of Horovod's negotiation loop are accepted and reported as having no effect
    inert = consts.inert_horovod_env({"HOROVOD_CYCLE_TIME": "3.5", "HOROVOD_TIMELINE": "x",
                                      "HOROVOD_HIERARCHICAL_ALLREDUCE": "1"})
    assert sorted(inert) == ["HOROVOD_CYCLE_TIME", "HOROVOD_HIERARCHICAL_ALLREDUCE"]
    assert "no negotiation cycle" in inert["HOROVOD_CYCLE_TIME"]


def test_byte_greedy_placement_and_owner_chunks():
    """`layout.assign_owners` = the reference's GreedyLoadBalancingStrategy over ALL sparse
    variables; `TableLayout` with an explicit owner map enumerates exactly the same rows chunk
    by chunk (`owner_chunks`) as all at once, for both partition strategies."""
    from parallax_b200.parallel.layout import TableLayout, assign_owners
    placed = assign_owners([("big", 5, 100), ("small", 5, 10), ("tiny", 3, 1)], 2)
    load = [0, 0]
    for key, nbytes in (("big", 100), ("small", 10), ("tiny", 1)):
        for o in placed[key]:
            load[o] += nbytes
    assert placed["big"] == [0, 1, 0, 1, 0]            # equal sizes: round-robin
    assert abs(load[0] - load[1]) <= 100               # the small tables fill the lighter owner
    assert placed["small"].count(1) > placed["small"].count(0)
    for strategy in ("mod", "div"):
        for V, P, W in ((1003, 8, 3), (17, 32, 8), (64, 5, 2)):
            owners = assign_owners([("a", P, 7), ("b", P, 3)], W)["b"]
            L = TableLayout(V, P, W, strategy, owners=owners)
            seen = 0
            for o in range(W):
                g, l = L.global_ids_of_owner(o)
                gs, ls = [], []
                for g2, l2 in L.owner_chunks(o, chunk=13):
                    gs.append(g2)
                    ls.append(l2)
                g3 = torch.cat(gs) if gs else torch.zeros(0, dtype=torch.int64)
                l3 = torch.cat(ls) if ls else torch.zeros(0, dtype=torch.int64)
                assert sorted(zip(g.tolist(), l.tolist())) == sorted(zip(g3.tolist(), l3.tolist()))
                assert bool((L.owner_of(g3) == o).all()) and (l3.numel() == 0 or
                                                              int(l3.max()) < L.rows_local)
                seen += g3.numel()
            assert seen == V
    with pytest.raises(ValueError):
        TableLayout(10, 4, 2, owners=[0, 1, 2, 0])      # owner out of range


def test_placement_planner_cli(capsys):
    from parallax_b200.tools import launch_ps
    rc = launch_ps.main(["--owners", "3", "a:1000:64:5:1", "w:793470:512:32:1+b:793470:1:32:1"])
    out = capsys.readouterr().out
    assert rc == 0 and "byte-greedy" in out and "owner 2:" in out and "w+b" in out
    assert launch_ps.main(["--owners", "2", "a:100:8:4+b:100:8:5"]) == 2      # group mismatch


def test_lookup_many_defer_on_plain_modules():
    """Without an engine (plain torch modules / host fabric) `lookup_many` is the sequence of
    lookups; `defer=True` returns a handle with the same rows and gradients still flow."""
    import parallax_b200 as parallax
    a = parallax.nn.Embedding(10, 4)
    b = parallax.nn.Embedding(10, 1)
    ids = torch.tensor([1, 3, 3])
    ra, rb = parallax.nn.lookup_many([a, b], ids)
    h = parallax.nn.lookup_many([a, b], ids, defer=True)
    ra2, rb2 = h.rows()
    assert torch.equal(ra, ra2) and torch.equal(rb, rb2)
    (ra2.sum() + rb2.sum()).backward()
    assert a.weight.grad is not None and b.weight.grad is not None


def test_sharded_checkpoint_resharding_logic(tmp_path):
    """`checkpoint.save_sharded` / `load_sharded` with stand-in tables: one shard file per owner,
    manifest with the placement, own-file fast path when the placement is unchanged and a full
    re-scatter into a different partitioning otherwise."""
    from parallax_b200 import checkpoint as ckpt
    from parallax_b200.parallel.layout import TableLayout

    class _Comm(object):
        distributed, is_cuda, device = False, False, torch.device("cpu")

        def __init__(self, rank, world):
            self.rank, self.world = rank, world

        def barrier(self):
            pass

    class _Table(object):
        def __init__(self, V, D, P, W, rank, full=None, slot=None):
            self.V, self.D, self.nslots, self.replicated, self.rank = V, D, 1, False, rank
            self.layout = TableLayout(V, P, W, "mod")
            self.w = torch.zeros(self.layout.rows_local, D)
            self.s = torch.zeros(self.layout.rows_local, D)
            if full is not None:
                for g, l in self.layout.owner_chunks(rank):
                    self.w[l] = full[g]
                    self.s[l] = slot[g]

        def local_rows(self, what="weight"):
            src = self.w if what == "weight" else self.s
            gs, rows = [], []
            for g, l in self.layout.owner_chunks(self.rank):
                gs.append(g)
                rows.append(src[l])
            return torch.cat(gs), torch.cat(rows)

        def load_rows(self, ids, rows, what="weight"):
            own = self.layout.owner_of(ids) == self.rank
            dst = self.w if what == "weight" else self.s
            dst[self.layout.local_row_of(ids[own])] = rows[own]

    class _Engine(object):
        dense, global_step, run_option = None, 7, "HYBRID"

        def __init__(self, comm, table):
            self.comm, self.tables = comm, {"emb.weight": table}
            self.model = torch.nn.Linear(1, 1)

    V, D = 101, 3
    full, slot = torch.randn(V, D), torch.rand(V, D)
    d = str(tmp_path / "model.ckpt-7")
    os.makedirs(d)
    for r in range(2):                                   # two ranks save, one after the other
        ckpt.save_sharded(_Engine(_Comm(r, 2), _Table(V, D, 4, 2, r, full, slot)), d, r == 0)
    files = sorted(os.listdir(d))
    assert files == ["dense.pt", "manifest.json", "sparse-emb.weight-rank0.pt",
                     "sparse-emb.weight-rank1.pt"]
    assert ckpt.latest_checkpoint(str(tmp_path)) == d
    # same placement: every rank needs only its own shard
    t0 = _Table(V, D, 4, 2, 1)
    os.rename(os.path.join(d, files[2]), os.path.join(d, files[2] + ".away"))
    ckpt.load_sharded(_Engine(_Comm(1, 2), t0), d)
    g, l = t0.layout.global_ids_of_owner(1)
    torch.testing.assert_close(t0.w[l], full[g])
    torch.testing.assert_close(t0.s[l], slot[g])
    os.rename(os.path.join(d, files[2] + ".away"), os.path.join(d, files[2]))
    # different world size and partition count: rows are re-scattered to their new owners
    got_w, got_s = torch.zeros(V, D), torch.zeros(V, D)
    for r in range(3):
        t = _Table(V, D, 5, 3, r)
        eng = _Engine(_Comm(r, 3), t)
        ckpt.load_sharded(eng, d)
        assert eng.global_step == 7
        g, l = t.layout.global_ids_of_owner(r)
        got_w[g], got_s[g] = t.w[l], t.s[l]
    torch.testing.assert_close(got_w, full)
    torch.testing.assert_close(got_s, slot)
    # offline: the inspection tool assembles the shards without an engine
    import io
    from parallax_b200.tools import inspect_checkpoint as ic
    for target in (d, str(tmp_path)):                    # the checkpoint itself / its directory
        path, sd = ic.load(target)
        assert path == d and sd["global_step"] == 7 and sd["skipped"] == []
        torch.testing.assert_close(sd["sparse"]["emb.weight"]["weight"], full)
        torch.testing.assert_close(sd["sparse"]["emb.weight"]["slots"][0], slot)
    buf = io.StringIO()
    ic.summarize(sd, buf)
    assert "sparse-slot0" in buf.getvalue() and "101x3" in buf.getvalue()
    torch.testing.assert_close(ic.to_state_dict(sd)["emb.weight"], full)
    _, small = ic.load(d, max_table_bytes=100)           # too large to assemble: only listed
    assert small["sparse"] == {} and small["skipped"] == ["emb.weight"]
    out = str(tmp_path / "plain.pt")
    assert ic.main([d, "--to_state_dict", out]) == 0
    torch.testing.assert_close(torch.load(out)["emb.weight"], full)
    # offline consumers read either format: a second (sharded) checkpoint, then the average
    full2 = full + 2.0
    d2 = str(tmp_path / "model.ckpt-9")
    os.makedirs(d2)
    for r in range(2):
        e = _Engine(_Comm(r, 2), _Table(V, D, 4, 2, r, full2, slot))
        e.global_step = 9
        ckpt.save_sharded(e, d2, r == 0)
    torch.save({"global_step": 3, "dense": None, "buffers": {},
                "sparse": {"emb.weight": {"weight": full - 2.0, "slots": [slot]}}},
               str(tmp_path / "model.ckpt-3.pt"))           # and an older single-file one
    assert [st for st, _ in ckpt.list_checkpoints(str(tmp_path))] == [3, 7, 9]
    assert ckpt.load_logical(str(tmp_path / "model.ckpt-3.pt"))["global_step"] == 3
    from parallax_b200.models.nmt.train import avg_checkpoints
    avg_path = avg_checkpoints(str(tmp_path), 3)
    avg = torch.load(avg_path, weights_only=False)
    assert avg["global_step"] == 9
    torch.testing.assert_close(avg["sparse"]["emb.weight"]["weight"], full)     # mean of -2, 0, +2
    os.remove(os.path.join(d, files[3]))                 # a missing shard is an error, not zeros
    with pytest.raises((RuntimeError, FileNotFoundError)):
        ckpt.assemble_table(d, "emb.weight")


def test_sharded_checkpoint_restores_on_the_host_fabric(tmp_path):
    """A checkpoint directory as GPU owners write it (manifest + one shard file per owner) restores
    into a host-fabric engine with another partition count — train on the NVLink fabric, evaluate
    anywhere — and `CheckpointSaver.restore_if_present` picks it up from the checkpoint directory."""
    import json
    from parallax_b200 import checkpoint as ckpt
    from parallax_b200.models.simple import MLPWithEmbedding
    from parallax_b200.utils import selfcheck as sc

    def make(nparts, ckpt_dir=None):
        torch.manual_seed(0)
        model = MLPWithEmbedding(sc.VOCAB, partitioner=parallax.get_partitioner(nparts))
        g = parallax.Graph(model, optimizer=optim.Adagrad(0.2, initial_accumulator_value=1.0))
        cfg = parallax.Config(run_option="HYBRID", search_partitions=False,
                              sess_config={"fabric": "host"})
        if ckpt_dir:
            cfg.ckpt_config = parallax.CheckPointConfig(ckpt_dir=ckpt_dir)
        sess, *_ = parallax.parallel_run(g, "localhost", sync=True, parallax_config=cfg)
        return sess

    s1 = make(5)
    for s in range(3):
        ids, labels = sc.make_batch(s, 1, 0)
        s1.run(["loss", "train_op"], {"ids": [ids], "labels": [labels]})
    sd = s1.engine.state_dict()
    s1.close()
    d = str(tmp_path / "model.ckpt-3")
    os.makedirs(d)
    man = {"format": 2, "global_step": 3, "world": 2, "run_option": "HYBRID", "sparse": {}}
    for name, ent in sd["sparse"].items():
        V, D = ent["weight"].shape
        files = []
        for r in range(2):                                   # two owners, rows dealt round-robin
            rows = torch.arange(r, V, 2)
            fn = "sparse-%s-rank%d.pt" % (name, r)
            torch.save({"ids": rows, "weight": ent["weight"][rows],
                        "slots": [s_[rows] for s_ in ent["slots"]]}, os.path.join(d, fn))
            files.append(fn)
        man["sparse"][name] = {"V": V, "D": D, "nslots": len(ent["slots"]), "files": files,
                               "placement": [2, "mod", 2, [0, 1], False]}
    torch.save({"global_step": 3, "dense": sd["dense"], "buffers": sd["buffers"]},
               os.path.join(d, "dense.pt"))
    with open(os.path.join(d, "manifest.json"), "w") as f:
        json.dump(man, f)
    s2 = make(3, ckpt_dir=str(tmp_path))                     # restore-on-start finds the directory
    sd2 = s2.engine.state_dict()
    s2.close()
    assert sd2["global_step"] == 3
    for name in sd["sparse"]:
        assert torch.equal(sd2["sparse"][name]["weight"], sd["sparse"][name]["weight"])
        for a, b in zip(sd2["sparse"][name]["slots"], sd["sparse"][name]["slots"]):
            assert torch.equal(a, b)
    for n, v in sd["dense"]["master"].items():
        assert torch.equal(sd2["dense"]["master"][n], v)
