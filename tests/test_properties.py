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
