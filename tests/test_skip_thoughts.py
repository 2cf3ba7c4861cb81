"""skip-thoughts example: LN-GRU cell, input shards, preprocessing, model, sentence
encoder / manager, vocabulary expansion, perplexity tracking, training through
`parallel_run` on the host fabric."""
import collections
import math
import os

import numpy as np
import pytest
import torch

import parallax_b200 as parallax
from parallax_b200.models import skip_thoughts as st
from parallax_b200.models.skip_thoughts.gru_cell import LayerNormGRU
from parallax_b200.models.skip_thoughts.input_ops import (files_for_shard, read_shard, write_shard,
                                                          parse_example_batch)


def _corpus(d, books=4, sentences=120, vocab=60, seed=0):
    """every sentence continues a counting pattern: neighbours are predictable"""
    import random
    rng = random.Random(seed)
    os.makedirs(d, exist_ok=True)
    files = []
    for b in range(books):
        fn = os.path.join(d, "book%d.txt" % b)
        files.append(fn)
        with open(fn, "w") as f:
            start = rng.randrange(vocab)
            for s in range(sentences):
                n = rng.randint(3, 7)
                f.write(" ".join("w%d" % ((start + 2 * s + k) % vocab) for k in range(n)) + "\n")
            f.write("\n")                                   # an empty line breaks triples
    return files


# ---------------------------------------------------------------------- cell
def test_layer_norm_gru_matches_cell_equations():
    torch.manual_seed(0)
    g = LayerNormGRU(5, 7)
    # recurrent blocks are orthonormal, input weights inside the uniform range
    for k in range(3):
        blk = g.w_hu[:, 7 * k:7 * (k + 1)]
        assert torch.allclose(blk.t() @ blk, torch.eye(7), atol=1e-5)
    assert g.w_x.abs().max() <= 0.1
    x = torch.randn(3, 6, 5)
    lengths = torch.tensor([6, 4, 1])
    out, h = g(x, lengths)

    def ln(m, v):
        return torch.nn.functional.layer_norm(v, (v.shape[-1],), m.weight, m.bias)
    for b in range(3):
        hb = torch.zeros(7)
        for t in range(int(lengths[b])):
            xt = x[b, t]
            zr = torch.sigmoid(ln(g.ln_wh, hb @ g.w_hu[:, :14]) + ln(g.ln_wx, xt @ g.w_x))
            z, r = zr[:7], zr[7:]
            cand = torch.tanh(r * ln(g.ln_u, hb @ g.w_hu[:, 14:]) + ln(g.ln_w, xt @ g.w))
            hb = (1 - z) * hb + z * cand
            assert torch.allclose(out[b, t], hb, atol=1e-5)
        assert torch.allclose(h[b], hb, atol=1e-5) and (out[b, int(lengths[b]):] == 0).all()
    # reverse = forward over the length-reversed sequence
    xr = x.clone()
    for b in range(3):
        n = int(lengths[b])
        xr[b, :n] = x[b, :n].flip(0)
    _, h_rev = g(x, lengths, reverse=True)
    _, h_fwd = g(xr, lengths)
    assert torch.allclose(h_rev, h_fwd, atol=1e-5)


# --------------------------------------------------------------------- input
def test_shard_files_roundtrip_and_file_split(tmp_path):
    exs = [([1, 2, 0], [3, 0], [4, 5, 6, 0]), ([7, 0], [8, 9, 0], [1, 0])]
    fn = write_shard(str(tmp_path / "train-00000-of-00001"), exs)
    back = read_shard(fn)
    assert [[a.tolist() for a in ex] for ex in back] == [[list(a) for a in ex] for ex in exs]
    enc, pre, post = parse_example_batch(back)
    assert enc.ids.tolist() == [[1, 2, 0], [7, 0, 0]] and enc.mask.tolist() == [[1, 1, 1], [1, 1, 0]]
    assert post.ids.shape == (2, 4) and int(post.mask.sum()) == 6
    # contiguous slices; the first `remainder` shards get one more file
    files = ["f%02d" % i for i in range(10)]
    parts = [files_for_shard(files, 4, k) for k in range(4)]
    assert [len(p) for p in parts] == [3, 3, 2, 2] and sum(parts, []) == files
    assert files_for_shard(files[:3], 4, 3) == []


def test_preprocess_builds_vocab_and_triples(tmp_path):
    files = _corpus(str(tmp_path / "books"))
    out = st.preprocess_dataset.preprocess(files, str(tmp_path / "data"), num_words=40,
                                           train_output_shards=4, num_validation_sentences=30,
                                           max_sentence_length=6)
    vocab = out["vocab"]
    assert list(vocab)[:2] == ["<eos>", "<unk>"] and len(vocab) == 40
    assert open(tmp_path / "data" / "vocab.txt").read().split("\n") == list(vocab)
    counts = [int(l.split()[1]) for l in open(tmp_path / "data" / "word_counts.txt")]
    assert counts == sorted(counts, reverse=True)
    assert len(out["train"]) == 4 and len(out["validation"]) == 1
    stats = out["stats"]
    assert stats["sentences_too_long"] > 0 and stats["sentences_skipped_empty"] > 0
    exs = sum((read_shard(f) for f in out["train"] + out["validation"]), [])
    assert len(exs) == stats["sentences_output"]
    for enc, pre, post in exs[:50]:
        assert enc[-1] == pre[-1] == post[-1] == 0 and max(len(enc), len(pre), len(post)) <= 7
    # a loaded vocabulary is used as is; duplicates are rejected
    again = st.preprocess_dataset.build_vocabulary([], vocab_file=str(tmp_path / "data" / "vocab.txt"))
    assert again == vocab
    dup = tmp_path / "dup.txt"
    dup.write_text("a\nb\na\n")
    with pytest.raises(AssertionError):
        st.preprocess_dataset.build_vocabulary([], vocab_file=str(dup))
    assert st.preprocess_dataset.sentence_to_ids("w1 zzz", {"w1": 5}) == [5, 1, 0]


def test_input_queue_shards_by_file(tmp_path):
    for i in range(5):
        write_shard(str(tmp_path / ("train-%05d-of-00005" % i)), [([i + 1, 0], [9, 0], [9, 0])] * 4)
    pat = str(tmp_path / "train-*.npz")
    q = st.input_ops.prefetch_input_data(pat, 2, shuffle=False, epochs=1)     # planted handles …
    parallax.shard.update_shard_values_for_worker(2, 1, 1)                    # … resolved later
    assert [os.path.basename(f)[6:11] for f in q.my_files()] == ["00003", "00004"]
    seen = [int(enc.ids[0, 0]) for enc, _, _ in q]
    assert seen == [4, 4, 5, 5]
    with pytest.raises(ValueError):
        st.input_ops.prefetch_input_data(str(tmp_path / "nope-*"), 2)
    q3 = st.input_ops.prefetch_input_data(pat, 3, shuffle=True, capacity=8, seed=1, num_shards=1,
                                          shard_id=0, epochs=2)
    firsts = [int(v) for enc, _, _ in q3 for v in enc.ids[:, 0]]
    assert len(firsts) == 39 and collections.Counter(firsts).most_common(1)[0][1] <= 8


# --------------------------------------------------------------------- model
def _cfg(**kw):
    base = dict(vocab_size=40, word_embedding_dim=12, encoder_dim=16, batch_size=16)
    base.update(kw)
    return st.model_config(**base)


@pytest.mark.parametrize("bidirectional", [False, True])
def test_model_loss_and_padding_invariance(bidirectional):
    torch.manual_seed(0)
    m = st.SkipThoughtsModel(_cfg(bidirectional_encoder=bidirectional))
    batch = parse_example_batch([([3, 4, 5, 0], [6, 7, 0], [8, 0]), ([9, 0], [3, 0], [4, 5, 6, 0])])
    feed = {k: v[0] for k, v in st.feed_from_batch(batch).items()}
    out = m(**feed)
    assert out["thought_vectors"].shape == (2, 16) and float(out["sum_weights"]) == 11
    assert abs(out["loss"].item() - (out["loss_pre"] + out["loss_post"]).item()) < 1e-4
    out["loss"].backward()
    assert all(p.grad is not None for p in m.parameters())
    # ids under the mask's zeros do not matter
    feed2 = dict(feed)
    feed2["encode_ids"] = feed["encode_ids"].clone()
    feed2["encode_ids"][1, 2:] = 7
    feed2["decode_pre_ids"] = feed["decode_pre_ids"].clone()
    feed2["decode_pre_ids"][1, 2:] = 9
    out2 = m(**feed2)
    assert torch.allclose(out2["loss"], out["loss"], atol=1e-4)
    assert torch.allclose(out2["thought_vectors"], out["thought_vectors"], atol=1e-6)
    # encode mode on embeddings = encode on ids
    emb = m.word_embedding(feed["encode_ids"])
    assert torch.allclose(m.encode_embeddings(emb, feed["encode_mask"]),
                          out["thought_vectors"], atol=1e-6)
    with pytest.raises(ValueError):
        st.SkipThoughtsModel(_cfg(bidirectional_encoder=True, encoder_dim=15))


def test_training_config_and_schedule():
    with pytest.raises(ValueError):
        st.training_config(learning_rate_decay_factor=0.5, learning_rate_decay_steps=0)
    tc = st.training_config(learning_rate=0.01, learning_rate_decay_factor=0.5,
                            learning_rate_decay_steps=100)
    from parallax_b200.models.skip_thoughts.model import learning_rate_fn
    lr = learning_rate_fn(tc)
    assert lr(1) == lr(100) == 0.01 and lr(101) == 0.005 and lr(201) == 0.0025
    assert learning_rate_fn(st.training_config(learning_rate_decay_factor=0))(999) == 0.0008


def test_skip_thoughts_trains_encodes_and_tracks_perplexity(tmp_path):
    files = _corpus(str(tmp_path / "books"), books=6, sentences=150)
    data = str(tmp_path / "data")
    out = st.preprocess_dataset.preprocess(files, data, num_words=62, train_output_shards=3,
                                           num_validation_sentences=60)
    mc = _cfg(vocab_size=62, word_embedding_dim=16, encoder_dim=32, num_embedding_partitions=2,
              input_file_pattern=os.path.join(data, "train-?????-of-00003.npz"))
    tc = st.training_config(learning_rate=0.01, learning_rate_decay_steps=1000)
    torch.manual_seed(0)
    model = st.SkipThoughtsModel(mc)
    ck = str(tmp_path / "ckpt")
    cfg = parallax.Config(run_option="HYBRID", search_partitions=False,
                          sess_config={"fabric": "host"},
                          ckpt_config=parallax.CheckPointConfig(ckpt_dir=ck, save_ckpt_steps=60))
    queue = st.input_ops.prefetch_input_data(mc.input_file_pattern, mc.batch_size, seed=3,
                                             capacity=256)
    sess, *_ = parallax.parallel_run(st.skip_thoughts_graph(model, tc), "localhost",
                                     parallax_config=cfg)
    try:
        assert sess.engine.run_option == "HYBRID" and list(sess.engine.tables) == [
            "word_embedding.weight"]
        ppl = []
        for batch in queue:
            loss, w, gs, _ = sess.run(["loss", "sum_weights", "global_step", "train_op"],
                                      st.feed_from_batch(batch))
            ppl.append(math.exp(loss[0] / w[0]))
            if gs[0] >= 120:
                break
        assert np.mean(ppl[:5]) > 30 and np.mean(ppl[-5:]) < 8, (ppl[:5], ppl[-5:])
    finally:
        sess.close()
    # validation perplexity from the saved checkpoint
    vc = _cfg(vocab_size=62, word_embedding_dim=16, encoder_dim=32,
              input_file_pattern=os.path.join(data, "validation-?????-of-00001.npz"))
    step, vppl = st.track_perplexity.run_once(vc, ck, num_eval_examples=48, min_global_step=10)
    assert step == 120 and vppl < 12
    assert st.track_perplexity.run_once(vc, ck, last_step=120) is None           # nothing new
    assert st.track_perplexity.run_once(vc, str(tmp_path / "empty")) is None
    assert st.track_perplexity.run(vc, ck, eval_dir=str(tmp_path / "eval"), max_evals=1,
                                   num_eval_examples=16, min_global_step=10) == 120
    assert open(tmp_path / "eval" / "perplexity.tsv").read().startswith("120\t")
    # sentence encoding with the trained model, restored from the Parallax checkpoint
    state = torch.load(os.path.join(ck, "model.ckpt-120.pt"), weights_only=False)
    vocab = list(out["vocab"])
    emb = st.encoder.embeddings_from_model(state, vocab)
    enc = st.SkipThoughtsEncoder(emb).build_from_config(vc, state)
    sents = ["w1 w2 w3 w4", "w1 w2 w3 w4", "w30 w31", "never seen words !"]
    v = enc.encode(sents, use_norm=True, batch_size=3)
    assert v.shape == (4, 32) and np.allclose(np.linalg.norm(v, axis=1), 1.0, atol=1e-5)
    assert np.allclose(v[0], v[1], atol=1e-6) and not np.allclose(v[0], v[2], atol=1e-3)
    raw = enc.encode(sents[:1], use_norm=False, use_eos=True)
    assert raw.shape == (1, 32) and not np.allclose(np.linalg.norm(raw), 1.0)
    # vocabulary expansion + manager with two models
    rng = np.random.RandomState(0)
    true_map = rng.randn(5, 16).astype(np.float32)
    w2v = collections.OrderedDict((w, np.linalg.lstsq(true_map.T, emb[w], rcond=None)[0])
                                  for w in vocab[2:30])
    w2v["brandnew"] = rng.randn(5).astype(np.float32)
    w2v["a_phrase"] = rng.randn(5).astype(np.float32)
    mat = np.stack([emb[w] for w in vocab])
    combined = st.vocabulary_expansion.expand_vocabulary(
        mat, collections.OrderedDict((w, i) for i, w in enumerate(vocab)), w2v)
    assert "brandnew" in combined and "a_phrase" not in combined
    assert np.allclose(combined["w5"], emb["w5"]) and len(combined) == len(vocab) + 1
    vf, ef = st.vocabulary_expansion.save_expanded(combined, str(tmp_path / "exp"))
    mgr = st.EncoderManager()
    with pytest.raises(ValueError):
        mgr.encode(["x"])
    mgr.load_model(vc, vf, ef, state)
    mgr.load_model(vc, vocab, mat, state)
    both = mgr.encode(["w1 w2 brandnew"])
    assert both.shape == (1, 64) and not np.allclose(both[0, :32], both[0, 32:])
    mgr.close()


def test_fit_linear_map_recovers_affine_map():
    rng = np.random.RandomState(1)
    x = rng.randn(50, 4).astype(np.float32)
    w, b = rng.randn(4, 3).astype(np.float32), rng.randn(3).astype(np.float32)
    w2, b2 = st.vocabulary_expansion.fit_linear_map(x, x @ w + b)
    assert np.allclose(w2, w, atol=1e-4) and np.allclose(b2, b, atol=1e-4)
    with pytest.raises(ValueError):
        st.vocabulary_expansion.expand_vocabulary(np.zeros((3, 2)), {"a": 0}, {"zzz": [1.0]})


class _BagEncoder(object):
    """deterministic stand-in for a trained encoder: hashed bag of words"""

    def encode(self, data, use_norm=True, verbose=False, batch_size=128, use_eos=False):
        out = np.zeros((len(data), 64), np.float32)
        for i, s in enumerate(data):
            for w in s.lower().split():
                out[i, sum(map(ord, w)) % 64] += 1.0
        return out / np.maximum(np.linalg.norm(out, axis=1, keepdims=True), 1e-9)


def test_downstream_evaluation_protocols(tmp_path):
    from parallax_b200.models.skip_thoughts import evaluate as ev
    rng = np.random.RandomState(0)
    good, bad = ["great", "fine", "lovely", "superb"], ["awful", "poor", "boring", "bad"]
    filler = ["movie", "plot", "actor", "scene", "the", "a"]

    def sent(words):
        return " ".join(rng.choice(filler, 3).tolist() + rng.choice(words, 2).tolist())
    (tmp_path / "custrev.pos").write_text("\n".join(sent(good) for _ in range(60)))
    (tmp_path / "custrev.neg").write_text("\n".join(sent(bad) for _ in range(60)))
    enc = _BagEncoder()
    acc = ev.eval_nested_kfold(enc, "CR", str(tmp_path), k=3, scan=[1, 16])
    assert acc > 0.9 and ev.evaluate(enc, "CR", str(tmp_path))["accuracy"] > 0.9
    # TREC: coarse label = text before ':'
    kinds = {"NUM": ["how", "many", "count"], "LOC": ["where", "city", "country"],
             "HUM": ["who", "person", "name"]}
    def trec(n):
        return "\n".join("%s:x %s" % (k, " ".join(rng.choice(w, 3).tolist()))
                         for _ in range(n) for k, w in kinds.items())
    (tmp_path / "train_5500.label").write_text(trec(30))
    (tmp_path / "TREC_10.label").write_text(trec(8))
    assert ev.eval_trec(enc, str(tmp_path), k=3, scan=[1, 16]) > 0.9
    # MSRP: paraphrase = same words reordered
    def pair(same):
        a = rng.choice(filler + good + bad, 5).tolist()
        b = list(rng.permutation(a)) if same else rng.choice(filler + good + bad, 5).tolist()
        return "%d\t1\t2\t%s\t%s" % (int(same), " ".join(a), " ".join(b))
    for fn, n in (("msr_paraphrase_train.txt", 80), ("msr_paraphrase_test.txt", 30)):
        (tmp_path / fn).write_text("Quality\tid1\tid2\ts1\ts2\n" +
                                   "\n".join(pair(i % 2 == 0) for i in range(n)))
    acc, f1 = ev.eval_msrp(enc, str(tmp_path), k=3, scan=[1, 16])
    assert acc > 0.8 and f1 > 0.8
    # SICK: relatedness grows with word overlap
    def sick(i):
        a = rng.choice(filler + good + bad, 6, replace=False).tolist()
        keep = i % 6
        b = a[:keep] + rng.choice(["zz1", "zz2", "zz3", "zz4", "zz5", "zz6"], 6 - keep,
                                  replace=False).tolist()
        return "%d\t%s\t%s\t%.1f\tNEUTRAL" % (i, " ".join(a), " ".join(b), 1 + 4 * keep / 5.0)
    for fn, n in (("SICK_train.txt", 120), ("SICK_test_annotated.txt", 40)):
        (tmp_path / fn).write_text("pair_ID\tA\tB\tscore\tjudgment\n" +
                                   "\n".join(sick(i) for i in range(n)))
    p, sp, mse = ev.eval_sick(enc, str(tmp_path))
    assert p > 0.8 and sp > 0.8 and mse < 1.0
    y = ev.encode_score_labels([1.0, 3.6, 5.0])
    assert y[0].tolist() == [1, 0, 0, 0, 0] and np.allclose(y[1], [0, 0, 0.4, 0.6, 0])
    assert y[2].tolist() == [0, 0, 0, 0, 1]
    with pytest.raises(ValueError):
        ev.evaluate(enc, "SST", str(tmp_path))
