"""Property-based tests (hypothesis) of the pure logic: table layout is a
bijection onto owner shards, the partition search always terminates with a
tested candidate range, resource info survives serialisation, shard handles
partition any dataset."""
import torch
from hypothesis import given, settings, strategies as st

from parallax_b200 import shard
from parallax_b200.parallel.layout import TableLayout
from parallax_b200.partitions import SearchState
from parallax_b200.resource import serialize_resource_info, deserialize_resource_info


@settings(max_examples=60, deadline=None)
@given(V=st.integers(1, 400), P=st.integers(1, 40), W=st.integers(1, 8),
       strat=st.sampled_from(["mod", "div"]))
def test_layout_is_a_bijection(V, P, W, strat):
    L = TableLayout(V, P, W, strat)
    ids = torch.arange(V)
    owner, local = L.owner_of(ids), L.local_row_of(ids)
    assert int(owner.min()) >= 0 and int(owner.max()) < W
    assert int(local.min()) >= 0 and int(local.max()) < L.rows_local
    keys = set(zip(owner.tolist(), local.tolist()))
    assert len(keys) == V                                # no two ids share a slot
    part = L.partition_of(ids)
    assert int(part.max()) < P and torch.equal(owner, part % W)
    if strat == "div":                                   # contiguous ranges, sizes differ by <= 1
        sizes = torch.bincount(part, minlength=P)
        assert int(sizes.max()) - int(sizes.min()) <= 1
        assert torch.equal(part, torch.sort(part).values)


@settings(max_examples=60, deadline=None)
@given(p0=st.integers(1, 64), a=st.floats(1e-4, 0.1), b=st.floats(0.1, 50.0),
       c=st.floats(0.0, 5.0))
def test_partition_search_terminates(p0, a, b, c):
    f = lambda p: a * (p - 1) + b / p + c
    s = SearchState(p0)
    keep, n = True, 0
    while keep and n < 64:
        keep, p = s.report(f(s.p_to_test))
        n += 1
    assert not keep and n < 64
    assert min(s.p_list) <= p <= max(s.p_list) and p >= 1
    # the choice is never worse than the worst tested candidate
    assert f(p) <= max(f(q) for q in s.p_list) + 1e-9


hosts = st.lists(st.tuples(st.from_regex(r"[a-z][a-z0-9]{0,8}", fullmatch=True),
                           st.lists(st.integers(0, 15), max_size=8, unique=True)),
                 min_size=1, max_size=4)


@settings(max_examples=40, deadline=None)
@given(hs=hosts)
def test_resource_info_serialisation_roundtrip(hs):
    info = {"master": [{"hostname": hs[0][0], "port": [1234], "gpus": []}],
            "ps": [{"hostname": h, "port": [2000 + i], "gpus": []} for i, (h, _) in enumerate(hs)],
            "worker": [{"hostname": h, "port": [3000 + i], "gpus": sorted(g)}
                       for i, (h, g) in enumerate(hs)]}
    assert deserialize_resource_info(serialize_resource_info(info)) == info


@settings(max_examples=40, deadline=None)
@given(n=st.integers(0, 200), workers=st.integers(1, 9))
def test_shards_partition_the_dataset(n, workers):
    data = list(range(n))
    seen = []
    for wid in range(workers):
        shard.reset()
        ds = shard.shard(data)
        shard.update_shard_values_for_worker(workers, wid, 1)
        part = list(ds)
        assert len(part) == len(ds)
        seen += part
    assert sorted(seen) == data
    shard.reset()


# ---------------------------------------------------------------------------
# engine semantics: random optimizer / gradient-rule / partition settings on the
# host fabric must reproduce a plain single-device torch oracle
# ---------------------------------------------------------------------------
def _oracle_step(model, opt, slots, ema, ids, labels, step, loss_scale, emb_scale, dense_scale,
                 clip, ema_decay):
    import torch
    from parallax_b200 import optim
    out = model(ids, labels)
    model.zero_grad()
    (out["loss"] * loss_scale).backward()
    params = dict(model.named_parameters())
    grads = {n: p.grad.clone() for n, p in params.items()}
    grads["emb.weight"] = grads["emb.weight"] * emb_scale
    dense = [n for n in params if n != "emb.weight"]
    for n in dense:
        grads[n] = grads[n] * dense_scale
    if clip:
        norm = torch.sqrt(sum((grads[n].double() ** 2).sum() for n in dense)).float()
        factor = min(1.0, clip / max(float(norm), 1e-30))
        for n in dense:
            grads[n] = grads[n] * factor
    hp = opt.hyper(step)
    with torch.no_grad():
        for n, p in params.items():
            if n == "emb.weight":
                rows = torch.unique(ids.reshape(-1))
                optim.apply_sparse_rows_(opt.kind, p.data, rows, grads[n][rows], slots[n], hp)
            else:
                optim.apply_dense_(opt.kind, p.data, grads[n], slots[n], hp)
                if ema_decay is not None:
                    ema[n].sub_((1.0 - ema_decay) * (ema[n] - p.data))
    return out["loss"].item()


@settings(max_examples=20, deadline=None)
@given(kind=st.sampled_from(["sgd", "momentum", "adagrad", "adam", "rmsprop"]),
       run_option=st.sampled_from(["HYBRID", "PS", "MPI"]),
       nparts=st.sampled_from([None, 1, 3, 7]), strategy=st.sampled_from(["mod", "div"]),
       clip=st.sampled_from([None, 0.05, 10.0]), loss_scale=st.sampled_from([1.0, 4.0]),
       emb_scale=st.sampled_from([1.0, 8.0]), wd=st.sampled_from([0.0, 0.01]),
       ema_decay=st.sampled_from([None, 0.9]), local_agg=st.booleans(), seed=st.integers(0, 50))
def test_host_engine_matches_plain_torch(kind, run_option, nparts, strategy, clip, loss_scale,
                                         emb_scale, wd, ema_decay, local_agg, seed):
    import torch
    import parallax_b200 as parallax
    from parallax_b200 import optim
    from parallax_b200 import shard as _shard
    from parallax_b200.models.simple import MLPWithEmbedding
    _shard.reset()
    mk = {"sgd": lambda: optim.GradientDescent(0.3, weight_decay=wd),
          "momentum": lambda: optim.Momentum(0.1, 0.9, use_nesterov=bool(seed % 2),
                                             weight_decay=wd),
          "adagrad": lambda: optim.Adagrad(0.2, 0.5, weight_decay=wd),
          "adam": lambda: optim.Adam(0.01, weight_decay=wd),
          "rmsprop": lambda: optim.RMSProp(0.01, 0.9, 0.5, 1e-3, weight_decay=wd)}[kind]
    part = parallax.get_partitioner(nparts, strategy) if nparts else None
    model = MLPWithEmbedding(40, partitioner=part, seed=seed)
    ref = MLPWithEmbedding(40, seed=seed)
    ref.emb.sparse = False
    dense_names = ["fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"]
    rules = [parallax.ScaleGradients(emb_scale, params=["emb.weight"])]
    if clip:
        rules.append(parallax.ClipByGlobalNorm(clip, params=["fc1.*", "fc2.*"]))
    graph = parallax.Graph(model, optimizer=mk(), grad_rules=rules, loss_scale=loss_scale,
                           ema=parallax.ExponentialMovingAverage(ema_decay, ["fc*"])
                           if ema_decay else None)
    cfg = parallax.Config(run_option=run_option, search_partitions=False,
                          sess_config={"fabric": "host"})
    cfg.communication_config = parallax.CommunicationConfig(
        parallax.PSConfig(local_aggregation=local_agg))
    sess, *_ = parallax.parallel_run(graph, "localhost", parallax_config=cfg)
    opt = mk()
    slots = {n: tuple(torch.full_like(p, v) for v in opt.slot_init())
             for n, p in ref.named_parameters()}
    ema = {n: p.detach().clone() for n, p in ref.named_parameters() if n in dense_names}
    g = torch.Generator().manual_seed(seed)
    try:
        for step in range(1, 4):
            ids = torch.randint(0, 40, (6, 3), generator=g)
            ids[:, 0] = ids[0, 0]                       # duplicates inside the batch
            labels = torch.randint(0, 4, (6,), generator=g)
            want = _oracle_step(ref, opt, slots, ema, ids, labels, step, loss_scale, emb_scale,
                                1.0, clip, ema_decay)
            got = sess.run(["loss", "train_op"], {"ids": [ids], "labels": [labels]})[0][0]
            assert abs(got - want) < 1e-4 * max(1.0, abs(want))
        sd = sess.engine.state_dict()
        for n in dense_names:
            torch.testing.assert_close(sd["dense"]["master"][n].view_as(ref.state_dict()[n]),
                                       ref.state_dict()[n], rtol=2e-4, atol=2e-5)
            if ema_decay:
                torch.testing.assert_close(sd["dense"]["ema"][n].view_as(ema[n]), ema[n],
                                           rtol=2e-4, atol=2e-5)
        torch.testing.assert_close(sd["sparse"]["emb.weight"]["weight"], ref.emb.weight.detach(),
                                   rtol=2e-4, atol=2e-5)
    finally:
        sess.close()
