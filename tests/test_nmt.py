"""NMT example: hyper-parameters, vocabulary, iterator, model variants, decoding,
metrics and end-to-end training on the host fabric.

Mirrors the reference's vendored NMT test-suite:
`nmt_test.py` (short training run + inference), `model_test.py` (every
encoder/attention/architecture combination builds and steps),
`inference_test.py` (greedy/beam, multi-worker inference),
`utils/iterator_utils_test.py` (bucketing, sharding, skip_count),
`utils/vocab_utils_test.py`, `utils/evaluation_utils_test.py`,
`utils/misc_utils_test.py`."""
import math
import os

import pytest
import torch

import parallax_b200 as parallax
import parallax_b200.models.nmt as nmt
from parallax_b200.models.nmt import (evaluation_utils, inference, iterator_utils,
                                      misc_utils, vocab_utils)
from parallax_b200.models.nmt.model import reverse_by_length
from tests.nmt_fixture import make_fixture


def _hp(**kw):
    kw.setdefault("num_units", 16)
    kw.setdefault("dropout", 0.0)
    hp = nmt.create_hparams(**kw)
    return nmt.extend_hparams(hp, 30, 30)


# ------------------------------------------------------------------ hparams
def test_standard_hparams_and_overrides():
    hp = nmt.create_hparams("wmt16_gnmt_8_layer", overrides="num_units=64,dropout=0.1")
    assert (hp.num_layers, hp.encoder_type, hp.attention_architecture) == (8, "gnmt", "gnmt_v2")
    assert hp.num_units == 64 and hp.dropout == 0.1 and hp.beam_width == 10
    nmt.extend_hparams(hp, 100, 120)
    assert hp.num_encoder_layers == hp.num_decoder_layers == 8
    # GNMT: bidirectional bottom layer and the layer above it are not residual
    assert hp.num_encoder_residual_layers == hp.num_decoder_residual_layers == 6
    assert (hp.src_vocab_size, hp.tgt_vocab_size) == (100, 120)
    assert "best_bleu" in hp
    i15 = nmt.extend_hparams(nmt.create_hparams("iwslt15"))
    assert (i15.encoder_type, i15.attention, i15.decay_scheme) == ("bi", "scaled_luong", "luong234")
    with pytest.raises(ValueError):
        nmt.create_hparams("no_such_file")


def test_extend_hparams_validation(tmp_path):
    with pytest.raises(ValueError, match="should be even"):
        nmt.extend_hparams(nmt.create_hparams(encoder_type="bi", num_layers=3))
    with pytest.raises(ValueError, match=">= 2"):
        nmt.extend_hparams(nmt.create_hparams(encoder_type="gnmt", num_layers=1,
                                              attention_architecture="gnmt"))
    with pytest.raises(ValueError, match="subword"):
        nmt.extend_hparams(nmt.create_hparams(subword_option="wordpiece"))
    hp = nmt.extend_hparams(nmt.create_hparams(num_encoder_layers=4, num_decoder_layers=2,
                                               residual=True))
    assert not hp.pass_hidden_state            # different depths ⇒ no state passing
    assert (hp.num_encoder_residual_layers, hp.num_decoder_residual_layers) == (3, 1)
    nmt.save_hparams(str(tmp_path), hp)
    back = nmt.load_hparams(str(tmp_path))
    assert back.values() == hp.values()
    assert nmt.load_hparams(str(tmp_path / "missing")) is None


def test_learning_rate_schedule():
    hp = nmt.create_hparams(learning_rate=1.0, num_train_steps=1200, decay_scheme="luong234",
                            warmup_steps=0)
    lr = nmt.learning_rate_fn(hp)
    # luong234: decay starts at 2/3 of training; staircase halving every
    # (remaining steps / 4) completed steps after that
    assert lr(1) == lr(801) == lr(900) == 1.0
    assert lr(901) == 0.5 and lr(1001) == 0.25 and lr(1200) == 0.125
    hp10 = nmt.create_hparams(learning_rate=1.0, num_train_steps=1000, decay_scheme="luong10")
    assert nmt.learning_rate_fn(hp10)(551) == 0.5 and nmt.learning_rate_fn(hp10)(601) == 0.25
    warm = nmt.learning_rate_fn(nmt.create_hparams(learning_rate=1.0, warmup_steps=100,
                                                   decay_scheme="", num_train_steps=1000))
    assert abs(warm(1) - 0.01) < 1e-9 and warm(50) < warm(100) < warm(101) == 1.0
    with pytest.raises(ValueError):
        nmt.learning_rate_fn(nmt.create_hparams(decay_scheme="cosine"))


# --------------------------------------------------------------- vocabulary
def test_check_vocab_moves_special_tokens_first(tmp_path):
    d = make_fixture(str(tmp_path / "fx"), n_train=10, special_first=False)
    out = str(tmp_path / "out")
    size, path = vocab_utils.check_vocab(os.path.join(d, "vocab.src"), out)
    assert size == 19 and path.startswith(out)
    vocab, _ = vocab_utils.load_vocab(path)
    assert vocab[:3] == ["<unk>", "<s>", "</s>"] and sorted(vocab[3:]) == sorted(
        "w%d" % i for i in range(16))
    # an already well-formed file is used as is
    size2, path2 = vocab_utils.check_vocab(path, str(tmp_path / "out2"))
    assert (size2, path2) == (19, path)
    with pytest.raises(ValueError):
        vocab_utils.check_vocab(str(tmp_path / "nope"), out)
    table = vocab_utils.VocabTable.from_file(path)
    assert table.encode(["<s>", "w3", "zzz"]) == [1, table.lookup("w3"), 0]
    assert table.decode([2, 999]) == ["</s>", "<unk>"]


def test_pretrained_embeddings(tmp_path):
    d = make_fixture(str(tmp_path / "fx"), n_train=10)
    emb = tmp_path / "emb.txt"
    emb.write_text("2 4\nw1 1 2 3 4\nw5 5 6 7 8\n")
    mat, ntrain = vocab_utils.pretrained_embedding_matrix(os.path.join(d, "vocab.src"), str(emb))
    table = vocab_utils.VocabTable.from_file(os.path.join(d, "vocab.src"))
    assert mat.shape == (19, 4) and ntrain == 3
    assert mat[table.lookup("w5")].tolist() == [5, 6, 7, 8] and mat[table.lookup("w2")].sum() == 0
    hp = _hp(num_units=4, attention="")
    hp.src_vocab_size = hp.tgt_vocab_size = 19
    m = nmt.create_model(hp)
    m.load_pretrained_embeddings("encoder", mat)
    assert m.embedding_encoder.weight[table.lookup("w1")].tolist() == [1, 2, 3, 4]


# ----------------------------------------------------------------- iterator
def _tables(d):
    return vocab_utils.create_vocab_tables(os.path.join(d, "vocab.src"),
                                           os.path.join(d, "vocab.tgt"), False)


def test_iterator_batches_and_targets(tmp_path):
    d = make_fixture(str(tmp_path), n_train=50)
    sv, tv = _tables(d)
    it = iterator_utils.get_iterator(d + "/train.src", d + "/train.tgt", sv, tv, 8, "<s>", "</s>",
                                     random_seed=0, num_buckets=1, src_max_len=4, tgt_max_len=3,
                                     num_shards=1, shard_index=0, shuffle=False)
    batches = list(it)
    assert [b.batch_size for b in batches] == [8] * 6 + [2]
    b = batches[0]
    assert int(b.source_sequence_length.max()) <= 4 and b.target_input.shape[1] <= 4
    with open(d + "/train.src") as f:
        first = f.readline().split()[:4]
    assert b.source[0, :len(first)].tolist() == sv.encode(first)
    for i in range(b.batch_size):
        n = int(b.target_sequence_length[i])
        assert b.target_input[i, 0] == tv.lookup("<s>") and b.target_output[i, n - 1] == tv.lookup("</s>")
        assert b.target_input[i, 1:n].tolist() == b.target_output[i, :n - 1].tolist()
        # padding uses the </s> id
        assert (b.target_output[i, n:] == tv.lookup("</s>")).all()
    # a second epoch yields the same number of examples; skip_count drops a prefix
    assert sum(x.batch_size for x in it.initialize()) == 50
    assert sum(x.batch_size for x in it.initialize(skip_count=20)) == 30


def test_iterator_buckets_group_similar_lengths(tmp_path):
    d = make_fixture(str(tmp_path), n_train=400, min_len=1, max_len=12)
    sv, tv = _tables(d)
    it = iterator_utils.get_iterator(d + "/train.src", d + "/train.tgt", sv, tv, 16, "<s>", "</s>",
                                     random_seed=3, num_buckets=4, src_max_len=12, tgt_max_len=12,
                                     num_shards=1, shard_index=0)
    spans, total = [], 0
    for b in it:
        total += b.batch_size
        if b.batch_size == 16:
            spans.append(int(b.source_sequence_length.max() - b.source_sequence_length.min()))
    assert total == 400
    assert max(spans) <= 3          # bucket_width = 3 ⇒ lengths inside a bucket differ by < 3
    static = iterator_utils.get_iterator(d + "/train.src", d + "/train.tgt", sv, tv, 16, "<s>",
                                         "</s>", random_seed=3, num_buckets=4, src_max_len=12,
                                         tgt_max_len=12, num_shards=1, shard_index=0,
                                         static_shapes=True)
    shapes = {(tuple(b.source.shape), tuple(b.target_input.shape)) for b in static}
    assert len(shapes) <= 5         # at most one feed signature per bucket


def test_iterator_shards_follow_parallax_assignment(tmp_path):
    d = make_fixture(str(tmp_path), n_train=40)
    sv, tv = _tables(d)
    mk = lambda: iterator_utils.get_iterator(d + "/train.src", d + "/train.tgt", sv, tv, 40,
                                             "<s>", "</s>", shuffle=False)
    it = mk()                                  # handles planted before parallel_run …
    parallax.shard.update_shard_values_for_worker(4, 1, 1)       # … resolved by it
    (b,) = list(it)
    assert b.batch_size == 10
    with open(d + "/train.src") as f:
        lines = f.read().splitlines()
    assert b.source[0, :int(b.source_sequence_length[0])].tolist() == sv.encode(lines[1].split())
    assert b.source[1, :int(b.source_sequence_length[1])].tolist() == sv.encode(lines[5].split())


def test_infer_iterator_keeps_file_order(tmp_path):
    d = make_fixture(str(tmp_path), n_train=5)
    sv, _ = _tables(d)
    data = inference.load_data(d + "/dev.src")
    got = list(iterator_utils.get_infer_iterator(data[:7], sv, 3, "</s>", src_max_len=4))
    assert [b.batch_size for b in got] == [3, 3, 1]
    assert got[0].target_input is None and int(got[0].source_sequence_length.max()) <= 4
    hp = nmt.create_hparams(inference_indices=[2, 0])
    assert inference.load_data(d + "/dev.src", hp) == [data[2], data[0]]


# -------------------------------------------------------------------- model
VARIANTS = [
    dict(encoder_type="uni", attention="", num_layers=2),
    dict(encoder_type="bi", attention="", num_layers=2),
    dict(encoder_type="bi", attention="scaled_luong", num_layers=2),
    dict(encoder_type="bi", attention="luong", num_layers=4, residual=True),
    dict(encoder_type="uni", attention="bahdanau", num_layers=2, unit_type="gru"),
    dict(encoder_type="uni", attention="normed_bahdanau", num_layers=3, residual=True,
         output_attention=False),
    dict(encoder_type="gnmt", attention="normed_bahdanau", attention_architecture="gnmt",
         num_layers=3, residual=True),
    dict(encoder_type="gnmt", attention="normed_bahdanau", attention_architecture="gnmt_v2",
         num_layers=4, residual=True),
    dict(encoder_type="gnmt", attention="scaled_luong", attention_architecture="gnmt_v2",
         num_layers=2, unit_type="gru"),
    dict(encoder_type="uni", attention="luong", num_layers=2, unit_type="layer_norm_lstm"),
    dict(encoder_type="bi", attention="luong", num_layers=2, pass_hidden_state=False),
    dict(encoder_type="uni", attention="scaled_luong", num_layers=2, share_vocab=True,
         init_op="glorot_uniform"),
]


def _batch(B=4, S=7, T=6, V=30, seed=0):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(3, V, (B, S), generator=g)
    tin = torch.randint(3, V, (B, T), generator=g)
    tout = torch.randint(3, V, (B, T), generator=g)
    return src, tin, tout, torch.tensor([7, 5, 3, 6][:B]), torch.tensor([6, 4, 6, 2][:B])


@pytest.mark.parametrize("cfg", VARIANTS, ids=lambda c: "-".join(str(v) for v in c.values()))
def test_model_variants_train_and_decode_consistently(cfg):
    """every variant: finite loss, gradients for every parameter, padding has no
    influence, and step-by-step decoding reproduces the teacher-forced logits
    (the GNMT decoder batches its upper layers over time in training)."""
    torch.manual_seed(0)
    hp = _hp(init_weight=0.5, **cfg)
    m = nmt.create_model(hp)
    src, tin, tout, sl, tl = _batch()
    out = m(src, tin, tout, sl, tl)
    assert math.isfinite(out["loss"].item()) and int(out["predict_count"]) == int(tl.sum())
    out["loss"].backward()
    missing = [n for n, p in m.named_parameters() if p.grad is None]
    assert not missing, missing
    m.eval()
    with torch.no_grad():
        full = m.logits(src, tin, sl)
        # tokens past source_sequence_length are invisible
        src2 = src.clone()
        src2[1, 5:] = 9
        src2[2, 3:] = 11
        assert torch.allclose(m.logits(src2, tin, sl), full, atol=1e-6)
        memory, state = m.encode(src, sl)
        steps = []
        for t in range(tin.shape[1]):
            lg, state = m.decode_step(tin[:, t], state, memory)
            steps.append(lg)
        assert torch.allclose(torch.stack(steps, 1), full, atol=1e-5)
        greedy, glen = inference.greedy_decode(m, src, sl, 1, 2, 9)
        beam, scores, blen = inference.beam_search_decode(m, src, sl, 1, 2, 9, 1)
        assert greedy.shape[1] == beam.shape[2] and (greedy == beam[:, 0]).all()
        assert (glen == blen[:, 0]).all()


def test_unknown_options_raise():
    for bad in (dict(unit_type="nas"), dict(encoder_type="tri"), dict(attention="dot"),
                dict(attention_architecture="deep"), dict(init_op="zeros")):
        with pytest.raises(ValueError):
            nmt.create_model(_hp(**bad))
    with pytest.raises(ValueError):
        nmt.nmt_graph(nmt.create_model(_hp()), _hp(optimizer="lamb"))
    with pytest.raises(ValueError):
        hp = _hp(share_vocab=True)
        hp.tgt_vocab_size = 31
        nmt.create_model(hp)


def test_reverse_by_length_and_bidirectional_state():
    x = torch.arange(12.0).view(2, 6, 1)
    r = reverse_by_length(x, torch.tensor([4, 6]))
    assert r[0, :, 0].tolist() == [3, 2, 1, 0, 4, 5] and r[1, :, 0].tolist() == [11, 10, 9, 8, 7, 6]
    torch.manual_seed(0)
    m = nmt.create_model(_hp(encoder_type="bi", num_layers=4, attention="luong"))
    src, _, _, sl, _ = _batch()
    emb = m.embedding_encoder(src)
    out, states = m.encoder(emb, sl)
    assert out.shape == (4, 7, 32) and len(states) == 4         # fw0, bw0, fw1, bw1
    # forward half of the top output at a sequence's last step = final fw state
    for b in range(4):
        assert torch.allclose(out[b, sl[b] - 1, :16], states[2][0][b], atol=1e-6)
        assert torch.allclose(out[b, 0, 16:], states[3][0][b], atol=1e-6)
        assert (out[b, sl[b]:] == 0).all()


# ----------------------------------------------------------------- decoding
def test_beam_search_properties():
    torch.manual_seed(1)
    m = nmt.create_model(_hp(init_weight=1.0, attention="scaled_luong", encoder_type="uni"))
    m.eval()
    src, _, _, sl, _ = _batch()
    ids, scores, lens = inference.beam_search_decode(m, src, sl, 1, 2, 8, 4, 0.0)
    assert ids.shape[:2] == (4, 4) and (scores[:, :-1] >= scores[:, 1:]).all()
    # the reported score of every hypothesis is its model log-probability
    with torch.no_grad():
        for b in range(2):
            for k in range(4):
                hyp = ids[b, k]
                n = int(lens[b, k])
                memory, state = m.encode(src[b:b + 1], sl[b:b + 1])
                tok, lp = torch.tensor([1]), 0.0
                for t in range(n):
                    lg, state = m.decode_step(tok, state, memory)
                    lp += float(torch.log_softmax(lg, -1)[0, hyp[t]])
                    tok = hyp[t:t + 1]
                assert abs(lp - float(scores[b, k])) < 1e-3
                assert (hyp[n:] == 2).all()
    # beams of one sentence are distinct hypotheses
    assert len({tuple(ids[0, k].tolist()) for k in range(4)}) == 4
    # the width-4 search never does worse than greedy
    g, gl = inference.greedy_decode(m, src, sl, 1, 2, 8)
    best1, s1, _ = inference.beam_search_decode(m, src, sl, 1, 2, 8, 1, 0.0)
    assert (scores[:, 0] >= s1[:, 0] - 1e-5).all()
    # length penalty only re-ranks: scores become log_prob / ((5+len)/6)^alpha
    _, s_lp, l_lp = inference.beam_search_decode(m, src, sl, 1, 2, 8, 4, 1.0)
    assert torch.isfinite(s_lp).all()


def test_infer_batch_modes_and_translation_text():
    torch.manual_seed(0)
    m = nmt.create_model(_hp(attention="luong", encoder_type="uni", init_weight=1.0))
    src, _, _, sl, _ = _batch()
    hp = m.hp
    hp.tgt_max_len_infer = 5
    greedy = inference.infer_batch(m, hp, src, sl, 1, 2)
    assert greedy.shape[:2] == (1, 4) and greedy.shape[2] <= 5 and m.training
    hp.beam_width, hp.num_translations_per_input = 3, 2
    assert inference.infer_batch(m, hp, src, sl, 1, 2).shape[:2] == (2, 4)
    hp.beam_width, hp.sampling_temperature, hp.num_translations_per_input = 0, 1.0, 3
    g = torch.Generator().manual_seed(0)
    samp = inference.infer_batch(m, hp, src, sl, 1, 2, generator=g)
    assert samp.shape[:2] == (3, 4) and not (samp[0] == samp[1]).all()
    vocab = vocab_utils.VocabTable(["<unk>", "<s>", "</s>", "new@@", "er", "▁a", "b", "▁c"])
    assert inference.get_translation([3, 4, 2, 4], vocab, "</s>", "bpe") == "newer"
    assert inference.get_translation(torch.tensor([5, 6, 7, 2]), vocab, "</s>", "spm") == "ab c"
    assert inference.get_translation([4, 6], vocab, "</s>", "") == "er b"


def test_worker_slices_cover_input():
    for n, w in ((10, 3), (7, 7), (5, 8), (100, 4)):
        spans = [inference.worker_slice(n, w, j) for j in range(w)]
        got = [i for s, e in spans for i in range(s, max(e, s))]
        assert got == list(range(n))


# ------------------------------------------------------------------ metrics
def test_bleu_rouge_accuracy(tmp_path):
    ref, hyp = tmp_path / "ref", tmp_path / "hyp"
    ref.write_text("the cat sat on the mat\nhello world again my friend\n")
    hyp.write_text("the cat sat on the mat\nhello world again my friend\n")
    for metric in ("bleu", "rouge", "accuracy", "word_accuracy"):
        assert abs(evaluation_utils.evaluate(str(ref), str(hyp), metric) - 100.0) < 1e-6
    hyp.write_text("the cat sat on a mat\nhello world my friend\n")
    bleu = evaluation_utils.evaluate(str(ref), str(hyp), "bleu")
    # hand-computed: clipped n-gram matches 9/10, 5/8, 2/6, 1/4; BP = exp(1-11/10)
    p = [9 / 10, 5 / 8, 2 / 6, 1 / 4]
    want = 100 * math.exp(sum(math.log(x) for x in p) / 4) * math.exp(1 - 11 / 10)
    assert abs(bleu - want) < 1e-6
    assert 0 < evaluation_utils.evaluate(str(ref), str(hyp), "rouge") < 100
    assert evaluation_utils.evaluate(str(ref), str(hyp), "accuracy") == 0.0
    wa = evaluation_utils.evaluate(str(ref), str(hyp), "word_accuracy")
    assert abs(wa - 100 * (5 / 6 + 2 / 5) / 2) < 1e-6
    with pytest.raises(ValueError):
        evaluation_utils.evaluate(str(ref), str(hyp), "meteor")
    # sub-word references are merged before scoring
    ref.write_text("new@@ er hou@@ se\n")
    hyp.write_text("newer house\n")
    assert evaluation_utils.evaluate(str(ref), str(hyp), "accuracy") == 0.0
    assert abs(evaluation_utils.evaluate(str(ref), str(hyp), "rouge", "bpe") - 100.0) < 1e-6
    b, prec, bp, ratio, hl, rl = evaluation_utils.compute_bleu(
        [[["a", "b", "c", "d"]]], [["a", "b", "x", "d"]], smooth=True)
    assert 0 < b < 1 and bp == 1.0 and (hl, rl) == (4, 4)
    r = evaluation_utils.rouge(["a b c d"], ["a b x d"])
    assert abs(r["rouge_1/f_score"] - 0.75) < 1e-9 and abs(r["rouge_2/r_score"] - 1 / 3) < 1e-9


def test_misc_utils():
    assert misc_utils.safe_exp(1e6) == float("inf") and misc_utils.safe_exp(0.0) == 1.0
    assert misc_utils.format_bpe_text(["Th@@", "is", "is", "te@@", "s@@", "t"]) == "This is test"
    assert misc_utils.format_spm_text("▁This ▁is ▁a ▁ te st".split()) == "This is a test"
    assert misc_utils.format_text([b"a", "b"]) == "a b"
    st = misc_utils.Stats()
    st.update(0.5, loss=2.0, predict_count=40, word_count=100, batch_size=10)
    st.update(0.5, loss=4.0, predict_count=20, word_count=100, batch_size=10)
    info = st.process()
    assert abs(info["train_ppl"] - math.exp(60.0 / 60.0)) < 1e-9
    assert abs(info["speed"] - 0.2) < 1e-9 and not info["overflow"]


# --------------------------------------------------------- end-to-end train
def _train_hp(d, out, **kw):
    base = dict(src="src", tgt="tgt", train_prefix=d + "/train", dev_prefix=d + "/dev",
                test_prefix=d + "/test", vocab_prefix=d + "/vocab", out_dir=out, num_units=64,
                num_layers=1, encoder_type="uni", attention="scaled_luong", dropout=0.0,
                batch_size=32, num_buckets=2, src_max_len=10, tgt_max_len=10, learning_rate=0.5,
                num_train_steps=700, decay_scheme="", steps_per_stats=100, random_seed=1,
                metrics=["bleu", "accuracy"], infer_batch_size=20, num_embeddings_partitions=2)
    base.update(kw)
    return nmt.create_hparams(**base)


def test_nmt_learns_reversal_through_parallax(tmp_path):
    """`nmt_test.py` analogue: train through `parallel_run` (HYBRID: dense
    variables + two partitioned sparse embeddings), perplexity → 1, then decode
    the dev set with greedy and beam search and score it."""
    d = make_fixture(str(tmp_path / "data"))
    hp = _train_hp(d, str(tmp_path / "out"))
    cfg = parallax.Config(run_option="HYBRID", search_partitions=False,
                          sess_config={"fabric": "host"})
    tr = nmt.train.train(hp, "localhost", cfg)
    try:
        assert tr.sess.engine.run_option == "HYBRID"
        assert sorted(tr.sess.engine.tables) == ["embedding_decoder.weight",
                                                 "embedding_encoder.weight"]
        ppl = [h["train_ppl"] for h in tr.history]
        assert ppl[0] > 8 and ppl[-1] < 1.5, ppl
        assert tr.final_ppl["dev"] < 1.5 and tr.final_ppl["test"] < 1.5
        assert tr.final_scores["dev"]["bleu"] > 80 and tr.final_scores["test"]["accuracy"] > 70
        assert hp.best_bleu == tr.final_scores["dev"]["bleu"]
        assert os.path.exists(os.path.join(hp.best_bleu_dir, "best_step"))
        saved = nmt.load_hparams(hp.out_dir)
        assert saved.best_bleu == hp.best_bleu and saved.src_vocab_size == 19
        # beam search over the trained model agrees with the references too
        hp.beam_width, hp.length_penalty_weight = 4, 1.0
        scores = inference.decode_and_evaluate(
            "dev-beam", tr.model, hp, inference.load_data(d + "/dev.src"), tr.src_vocab,
            tr.tgt_vocab, str(tmp_path / "out" / "beam"), ref_file=d + "/dev.tgt")
        assert scores["bleu"] > 80
        assert isinstance(tr.sample_decode(), str)
        # multi-worker inference: three workers translate slices, worker 0 merges
        hp.beam_width = 0
        outf = str(tmp_path / "out" / "multi")
        for job in (2, 1, 0):
            res = inference.multi_worker_inference(tr.model, hp, d + "/test.src", outf, tr.src_vocab,
                                                   tr.tgt_vocab, num_workers=3, jobid=job)
        assert res == outf
        single = str(tmp_path / "out" / "single")
        inference.single_worker_inference(tr.model, hp, d + "/test.src", single, tr.src_vocab,
                                          tr.tgt_vocab)
        assert open(outf).read() == open(single).read()
        assert evaluation_utils.evaluate(d + "/test.tgt", outf, "bleu") > 80
    finally:
        tr.sess.close()


def test_gnmt_trains_through_parallax_and_averages_checkpoints(tmp_path):
    d = make_fixture(str(tmp_path / "data"), n_train=300)
    hp = _train_hp(d, str(tmp_path / "out"), encoder_type="gnmt", num_layers=2,
                   attention="normed_bahdanau", attention_architecture="gnmt_v2", residual=True,
                   num_units=32, num_train_steps=120, steps_per_stats=40, optimizer="adam",
                   learning_rate=0.001, dev_prefix="", test_prefix="", share_vocab=False)
    ck = str(tmp_path / "ckpt")
    cfg = parallax.Config(run_option="HYBRID", search_partitions=False,
                          sess_config={"fabric": "host"},
                          ckpt_config=parallax.CheckPointConfig(ckpt_dir=ck, save_ckpt_steps=40))
    tr = nmt.train.train(hp, "localhost", cfg, final_eval=False)
    try:
        ppl = [h["train_ppl"] for h in tr.history]
        assert len(ppl) == 3 and ppl[-1] < ppl[0]
        path = nmt.train.avg_checkpoints(ck, 3)
        avg = torch.load(path, weights_only=False)
        parts = [torch.load(os.path.join(ck, "model.ckpt-%d.pt" % s), weights_only=False)
                 for s in (40, 80, 120)]
        k = "output_layer.weight"
        want = sum(p["dense"]["master"][k].double() for p in parts) / 3
        assert torch.allclose(avg["dense"]["master"][k].double(), want, atol=1e-6)
        e = "embedding_encoder.weight"
        want_e = sum(p["sparse"][e]["weight"].double() for p in parts) / 3
        assert torch.allclose(avg["sparse"][e]["weight"].double(), want_e, atol=1e-6)
        assert avg["global_step"] == 120
        assert nmt.train.avg_checkpoints(ck, 5) is None
    finally:
        tr.sess.close()


def test_shared_vocabulary_is_one_table_in_the_engine():
    """`share_vocab`: encoder and decoder use ONE embedding module; the engine must
    replace every alias by the same sharded table and train it from both uses."""
    torch.manual_seed(0)
    hp = _hp(share_vocab=True, attention="luong", encoder_type="uni", num_layers=1,
             num_embeddings_partitions=2, learning_rate=0.5)
    m = nmt.create_model(hp)
    sess, *_ = parallax.parallel_run(nmt.nmt_graph(m, hp), "localhost",
                                     parallax_config=parallax.Config(
                                         search_partitions=False, sess_config={"fabric": "host"}))
    try:
        assert sorted(sess.engine.tables) == ["embedding_encoder.weight"]
        assert m.embedding_encoder is m.embedding_decoder
        assert type(m.embedding_encoder).__name__ == "ShardedEmbedding"
        src, tin, tout, sl, tl = _batch()
        feed = {"source": [src], "target_input": [tin], "target_output": [tout],
                "source_sequence_length": [sl], "target_sequence_length": [tl]}
        w0 = sess.engine.tables["embedding_encoder.weight"].full_weight().clone()
        losses = [sess.run(["loss", "train_op"], feed)[0][0] for _ in range(25)]
        assert losses[-1] < losses[0] - 1.0
        w1 = sess.engine.tables["embedding_encoder.weight"].full_weight()
        moved = (w1 - w0).abs().sum(1) > 0
        # rows used only by the source side and rows used only by the target side both moved
        only_src = set(src.reshape(-1).tolist()) - set(tin.reshape(-1).tolist())
        only_tgt = set(tin[:, :2].reshape(-1).tolist()) - set(src.reshape(-1).tolist())
        assert all(moved[i] for i in only_tgt) and any(moved[i] for i in only_src)
    finally:
        sess.close()


@pytest.mark.parametrize("unit_type", ["lstm", "gru", "layer_norm_lstm"])
def test_rnn_layer_step_equals_sequence_call(unit_type):
    """the explicit-GEMM step used by the decoders = the (cuDNN / native) sequence kernel"""
    from parallax_b200.models.nmt.model import RNNLayer
    torch.manual_seed(0)
    layer = RNNLayer(unit_type, 12, 10, forget_bias=1.0, dropout=0.0, residual=True,
                     init_weight=0.5).eval()
    x = torch.randn(3, 5, 12)
    seq, st_seq = layer(x)
    st, outs = layer.zero_state(3, x.device, x.dtype), []
    for t in range(5):
        o, st = layer.step(x[:, t], st)
        outs.append(o)
    assert torch.allclose(torch.stack(outs, 1), seq, atol=1e-5)
    for a, b in zip(st if isinstance(st, tuple) else (st,),
                    st_seq if isinstance(st_seq, tuple) else (st_seq,)):
        assert torch.allclose(a, b, atol=1e-5)
    # gradients agree too
    layer.train()
    x1 = x.clone().requires_grad_(True)
    layer(x1)[0].sum().backward()
    g_seq = [p.grad.clone() for p in layer.parameters()]
    layer.zero_grad()
    x2 = x.clone().requires_grad_(True)
    st, tot = layer.zero_state(3, x.device, x.dtype), 0.0
    for t in range(5):
        o, st = layer.step(x2[:, t], st)
        tot = tot + o.sum()
    tot.backward()
    for a, p in zip(g_seq, layer.parameters()):
        assert torch.allclose(a, p.grad, atol=1e-4)
    assert torch.allclose(x1.grad, x2.grad, atol=1e-5)
