"""Model-level checks on the host fabric: the LM1B training graph (clip, scale,
EMA, sampled softmax, partitioned tables) learns; every example model builds and
takes a step; the unique log-uniform sampler has the right law."""

import numpy as np
import pytest
import torch

import parallax_b200 as parallax
from parallax_b200.models.lm1b import (LM1B, lm1b_graph, log_uniform_sample_unique,
                                       log_uniform_logq_unique)


def test_lm1b_tiny_learns_and_tracks_ema():
    torch.manual_seed(0)
    m = LM1B(vocab_size=300, emb_size=16, state_size=32, projected_size=16, num_sampled=32,
             num_steps=4, num_shards=3, keep_prob=1.0)
    g = lm1b_graph(m, batch_size=8, learning_rate=0.2)
    sess, *_ = parallax.parallel_run(g, "localhost",
                                     parallax_config=parallax.Config(sess_config={"fabric": "host"}))
    gen = torch.Generator().manual_seed(1)
    x = torch.randint(0, 300, (8, 4), generator=gen)
    y = torch.roll(x, -1, dims=1)                       # a learnable mapping
    losses = [sess.run(["loss", "train_op"], {"x": [x], "y": [y]})[0][0] for _ in range(60)]
    assert np.mean(losses[-5:]) < 0.5 * np.mean(losses[:5])
    sd = sess.engine.state_dict()
    assert set(sd["dense"]["ema"]) == {"W", "B", "W_P"}
    # EMA lags the weights but moved away from the initial values
    assert not torch.allclose(sd["dense"]["ema"]["W"], sd["dense"]["master"]["W"])
    assert set(sd["sparse"]) == {"emb.weight", "softmax_w.weight", "softmax_b.weight"}
    # recurrent state round-trips through feeds like the reference driver does
    out = sess.run({"c": "final_state_c", "h": "final_state_h"}, {"x": [x], "y": [y]})
    out2 = sess.run("loss", {"x": [x], "y": [y], "initial_state_c": out["c"],
                             "initial_state_h": out["h"]})
    assert np.isfinite(out2[0])
    sess.close()


def test_unique_log_uniform_sampler_law():
    torch.manual_seed(0)
    V, S = 5000, 256
    ids, tries = log_uniform_sample_unique(S, V, "cpu")
    assert ids.numel() == S and ids.unique().numel() == S
    assert S <= float(tries) <= 3 * S
    # small ids are (almost) always drawn, large ones rarely: log-uniform head
    hits = torch.zeros(V)
    for _ in range(30):
        i, _ = log_uniform_sample_unique(S, V, "cpu")
        hits[i] += 1
    assert hits[:5].min() >= 28 and hits[-1000:].mean() < 3
    q = torch.exp(log_uniform_logq_unique(torch.tensor([0, 10, 4000]), tries, V))
    assert q[0] > q[1] > q[2] and q[0] <= 1.0 + 1e-6


@pytest.mark.parametrize("name,hw", [("lenet", 28), ("resnet20", 32), ("densenet40_k12", 32),
                                     ("trivial", 224)])
def test_cnn_zoo_takes_a_step(name, hw):
    from parallax_b200.models import cnn
    torch.manual_seed(0)
    m = cnn.get_model(name, 10)
    g = cnn.cnn_graph(m, "momentum", 0.01)
    sess, *_ = parallax.parallel_run(g, "localhost",
                                     parallax_config=parallax.Config(sess_config={"fabric": "host"}))
    x, y = torch.randn(2, 3, hw, hw), torch.randint(0, 10, (2,))
    l = [sess.run(["loss", "train_op"], {"images": [x], "labels": [y]})[0][0] for _ in range(3)]
    assert np.isfinite(l).all() and sess.engine.run_option == "MPI"     # dense-only ⇒ AR
    sess.close()


def test_ncf_and_bert_shapes_run():
    from parallax_b200.models.ncf import NeuMF, ncf_graph
    from parallax_b200.models.bert import Bert, bert_graph
    torch.manual_seed(0)
    cfgs = lambda ro: parallax.Config(run_option=ro, sess_config={"fabric": "host"})
    m = NeuMF(500, 200, 8, (32, 16, 8), 4, lazy=False)
    sess, *_ = parallax.parallel_run(ncf_graph(m, 0.01), "localhost", parallax_config=cfgs("PS"))
    f = {"users": [torch.randint(0, 500, (32,))], "items": [torch.randint(0, 200, (32,))],
         "labels": [torch.randint(0, 2, (32,))]}
    l = [sess.run(["loss", "train_op"], f)[0][0] for _ in range(20)]
    assert l[-1] < l[0] and sess.engine.run_option == "PS"
    sess.close()
    m = Bert(100, 32, 2, 4, 64, 16, 2)
    sess, *_ = parallax.parallel_run(bert_graph(m, 1e-3), "localhost",
                                     parallax_config=cfgs("HYBRID"))
    f = {"input_ids": [torch.randint(0, 100, (2, 16))],
         "mlm_positions": [torch.randint(0, 16, (2, 3))],
         "mlm_labels": [torch.randint(0, 100, (2, 3))]}
    l = [sess.run(["loss", "train_op"], f)[0][0] for _ in range(10)]
    assert l[-1] < l[0] and sess.engine.run_option == "HYBRID"
    sess.close()


def test_lm1b_vocabulary_and_dataset(tmp_path):
    from parallax_b200.models.lm1b_data import Dataset, Vocabulary
    (tmp_path / "vocab.txt").write_text("<S> 100\n<UNK> 50\nthe 40\ncat 30\nsat 20\nmat 10\n")
    v = Vocabulary.from_file(str(tmp_path / "vocab.txt"))
    assert (v.num_tokens, v.s_id, v.unk_id) == (6, 0, 1) and v.get_count("cat") == 30
    assert v.get_id("dog") == 1 and v.get_token(3) == "cat"
    assert Vocabulary.from_file(str(tmp_path / "vocab.txt"), num_tokens_limit=4).num_tokens == 4
    # native and python tokenisation agree
    line = "the  cat\tsat on the mat"
    assert v.encode_line(line) == v.encode_line(line, native=False) == [0, 2, 3, 4, 1, 2, 5, 0]
    for i, text in enumerate(["the cat sat\nthe mat\n", "cat sat\nmat the cat sat the\n"]):
        (tmp_path / ("news-%05d" % i)).write_text(text)
    ds = Dataset(v, str(tmp_path / "news-*"), deterministic=True)
    batches = list(ds.iterate_once(2, 4))
    x, y, w = batches[0]
    # row 0 streams sentence 1 then continues with the next unread sentence
    assert x[0].tolist() == [0, 2, 3, 4] and y[0].tolist() == [2, 3, 4, 0] and w[0].tolist() == [1] * 4
    assert x[1].tolist() == [0, 2, 5, 0] and y[1].tolist() == [2, 5, 0, 3]
    total = sum(int(b[2].sum()) for b in batches)
    assert total == sum(len(s.split()) + 1 for s in
                        ["the cat sat", "the mat", "cat sat", "mat the cat sat the"])
    # files are split over workers; an empty share is an error, not a silent hang
    assert [len(ds.files(2, k)) for k in (0, 1)] == [1, 1]
    one = list(Dataset(v, str(tmp_path / "news-*"), deterministic=True).iterate_once(2, 4, 2, 1))
    assert sum(int(b[2].sum()) for b in one) == 3 + 6
    with pytest.raises(ValueError):
        next(Dataset(v, str(tmp_path / "nothing-*")).iterate_forever(2, 4))
    it = Dataset(v, str(tmp_path / "news-*"), seed=1).iterate_forever(2, 4)
    assert all(next(it)[0].shape == (2, 4) for _ in range(12))      # wraps around epochs
