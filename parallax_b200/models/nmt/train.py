"""NMT training / evaluation loop on a Parallax session.

Parity: `examples/nmt/train.py:45-584` (statistics per `steps_per_stats`
window — step time, perplexity, words/s, gradient norm —, internal evaluation =
dev/test perplexity, external evaluation = decode dev/test and score
BLEU/ROUGE/accuracy, best-metric checkpoints, sample decodes, epoch handling
with `skip_count`), `examples/nmt/nmt_distributed_driver.py:76-189` (only
worker 0 logs statistics), `examples/nmt/model_helper.py:473-607`
(`avg_checkpoints`, `compute_perplexity`).

Every worker runs the same evaluation passes in lock-step — on the host fabric
a sharded-embedding lookup is a collective — and only worker 0 prints and
writes files.
"""
import os
import random
import time

import torch

from ... import checkpoint as _ckpt
from ...log import parallax_log as log
from . import inference, misc_utils, vocab_utils
from .hparams import extend_hparams, save_hparams
from .iterator_utils import get_iterator
from .model import create_model, nmt_graph


def prepare_vocab(hp):
    """check both vocabulary files, fill in sizes (`nmt.py:421-458`)"""
    src_file = hp.vocab_prefix + "." + hp.src
    tgt_file = hp.vocab_prefix + "." + hp.tgt
    out = hp.out_dir or "."
    src_size, src_file = vocab_utils.check_vocab(src_file, out, hp.check_special_token,
                                                 hp.sos, hp.eos)
    if hp.share_vocab:
        tgt_file, tgt_size = src_file, src_size
    else:
        tgt_size, tgt_file = vocab_utils.check_vocab(tgt_file, out, hp.check_special_token,
                                                     hp.sos, hp.eos)
    hp.src_vocab_file, hp.tgt_vocab_file = src_file, tgt_file
    return extend_hparams(hp, src_size, tgt_size)


def compute_perplexity(run_eval, batches):
    """exp(Σ loss·batch / Σ predict_count) over `batches` (`model_helper.py:585-607`)"""
    total_loss = total_count = 0.0
    for b in batches:
        loss, count = run_eval(b)
        total_loss += float(loss) * b.batch_size
        total_count += float(count)
    return misc_utils.safe_exp(total_loss / max(total_count, 1.0))


def avg_checkpoints(ckpt_dir, num_last_checkpoints, out_dir=None):
    """Average the dense/sparse weights of the last N checkpoints
    (`model_helper.py:473-536`); optimizer slots and the step are taken from the
    newest.  Returns the path written (``<ckpt_dir>/avg_checkpoints`` by default)."""
    found = _ckpt.list_checkpoints(ckpt_dir)[-int(num_last_checkpoints):]   # either format
    if len(found) < num_last_checkpoints:
        log.info("skipping averaging: only %d checkpoints", len(found))
        return None
    states = [_ckpt.load_logical(path) for _, path in found]
    avg = states[-1]
    avg.pop("skipped", None)

    def mean_into(dst, srcs):
        for k, v in dst.items():
            if torch.is_tensor(v) and v.is_floating_point():
                dst[k] = sum(s[k].to(torch.float64) for s in srcs).div_(len(srcs)).to(v.dtype)
    if avg.get("dense") and "master" in avg["dense"]:
        mean_into(avg["dense"]["master"], [s["dense"]["master"] for s in states])
    for name in avg.get("sparse", {}):
        w = sum(s["sparse"][name]["weight"].to(torch.float64) for s in states) / len(states)
        avg["sparse"][name]["weight"] = w.to(avg["sparse"][name]["weight"].dtype)
    out_dir = out_dir or os.path.join(ckpt_dir, "avg_checkpoints")
    os.makedirs(out_dir, exist_ok=True)
    name = "%s%d.pt" % (_ckpt.PREFIX, int(avg["global_step"]))
    torch.save(avg, os.path.join(out_dir, name))
    with open(os.path.join(out_dir, _ckpt.INDEX), "w") as f:
        f.write(name)
    return os.path.join(out_dir, name)


def sparse_capacity(hp):
    """Upper bound of gradient rows per step for each embedding table — on the
    NVLink fabric the receive rings of a sparse table are sized once, at the first
    step, and NMT batches vary in length (`sess_config["sparse_capacity"]`)."""
    src = hp.batch_size * int(hp.src_max_len or 100)
    tgt = hp.batch_size * (int(hp.tgt_max_len or 100) + 1)
    if hp.share_vocab:
        return {"embedding_encoder.weight": src + tgt}
    return {"embedding_encoder.weight": src, "embedding_decoder.weight": tgt}


def with_sparse_capacity(config, hp):
    """copy of `config` whose sess_config carries `sparse_capacity(hp)` unless the
    user set one"""
    import copy
    cfg = copy.copy(config)
    sc = dict(cfg.sess_config) if isinstance(cfg.sess_config, dict) else {}
    sc.setdefault("sparse_capacity", sparse_capacity(hp))
    if sc.get("cuda_graph"):
        # the encoders pack by length (a host read of the lengths) and batch shapes vary
        log.warning("cuda_graph is not supported for the NMT models; running eagerly")
        sc["cuda_graph"] = False
    cfg.sess_config = sc
    return cfg


class Trainer(object):
    """Drives one worker's session; see `train()`."""

    def __init__(self, hp, sess, num_workers, worker_id, model, src_vocab, tgt_vocab):
        self.hp, self.sess, self.model = hp, sess, model
        self.num_workers, self.worker_id = num_workers, worker_id
        self.src_vocab, self.tgt_vocab = src_vocab, tgt_vocab
        self.chief = worker_id == 0
        self.stats = misc_utils.Stats()
        self.history = []

    # -- data -----------------------------------------------------------------
    def _file(self, prefix, lang):
        return "%s.%s" % (prefix, lang)

    def train_iterator(self, skip_count=0):
        hp = self.hp
        return get_iterator(
            self._file(hp.train_prefix, hp.src), self._file(hp.train_prefix, hp.tgt),
            self.src_vocab, self.tgt_vocab, hp.batch_size, hp.sos, hp.eos,
            random_seed=hp.random_seed, num_buckets=hp.num_buckets,
            src_max_len=hp.src_max_len, tgt_max_len=hp.tgt_max_len, skip_count=skip_count,
            pin_memory=True)

    def eval_iterator(self, prefix):
        hp = self.hp
        return get_iterator(
            self._file(prefix, hp.src), self._file(prefix, hp.tgt), self.src_vocab,
            self.tgt_vocab, hp.batch_size, hp.sos, hp.eos, random_seed=hp.random_seed,
            num_buckets=hp.num_buckets,
            src_max_len=hp.get("src_max_len_infer") or None,
            tgt_max_len=hp.get("tgt_max_len_infer") or None,
            num_shards=1, shard_index=0, shuffle=False)

    # -- evaluation -------------------------------------------------------------
    def _run_eval(self, batch):
        loss, count = self.sess.run(["loss", "predict_count"], batch.as_feed())
        return loss[0], count[0]

    def internal_eval(self):
        """dev / test perplexity (`train.py:59-95`)"""
        out = {}
        for name, prefix in (("dev", self.hp.dev_prefix), ("test", self.hp.test_prefix)):
            if prefix:
                was = self.model.training
                self.model.eval()
                try:
                    out[name] = compute_perplexity(self._run_eval, self.eval_iterator(prefix))
                finally:
                    self.model.train(was)
                if self.chief:
                    log.info("  eval %s: perplexity %.2f", name, out[name])
        return out

    def external_eval(self, save_on_best=True):
        """decode dev / test, score every metric, remember the best
        (`train.py:97-160,514-584`)"""
        hp, out = self.hp, {}
        gs = self.sess.engine.global_step
        for name, prefix in (("dev", hp.dev_prefix), ("test", hp.test_prefix)):
            if not prefix:
                continue
            data = inference.load_data(self._file(prefix, hp.src))
            trans = os.path.join(hp.out_dir or ".", "output_%s" % name)
            if not self.chief:
                trans = os.devnull
            scores = inference.decode_and_evaluate(
                name, self.model, hp, data, self.src_vocab, self.tgt_vocab, trans,
                ref_file=self._file(prefix, hp.tgt) if self.chief else None)
            out[name] = scores
            if name == "dev" and save_on_best:
                for metric, v in scores.items():
                    if v > hp.get("best_" + metric, 0.0):
                        setattr(hp, "best_" + metric, v)
                        self._save_best(metric, gs)
        if self.chief and hp.out_dir:
            save_hparams(hp.out_dir, hp)
        return out

    def _save_best(self, metric, step):
        d = self.hp.get("best_%s_dir" % metric)
        if not (self.chief and d):
            return
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "best_step"), "w") as f:
            f.write("%d\n" % step)

    def sample_decode(self, prefix=None):
        """translate one random dev sentence (`train.py:443-468`); the pick is a
        function of the global step so every worker decodes the same one"""
        hp = self.hp
        prefix = prefix or hp.dev_prefix
        if not prefix:
            return None
        src = inference.load_data(self._file(prefix, hp.src))
        tgt = inference.load_data(self._file(prefix, hp.tgt))
        i = random.Random(self.sess.engine.global_step).randint(0, len(src) - 1)
        from .iterator_utils import get_infer_iterator
        dev = next(self.model.parameters()).device
        batch = next(iter(get_infer_iterator([src[i]], self.src_vocab, 1, hp.eos)))
        ids = inference.infer_batch(self.model, hp, batch.source.to(dev),
                                    batch.source_sequence_length.to(dev),
                                    self.tgt_vocab.lookup(hp.sos), self.tgt_vocab.lookup(hp.eos))
        text = inference.get_translation(ids[0, 0].cpu(), self.tgt_vocab, hp.eos,
                                         hp.subword_option)
        if self.chief:
            log.info("  # %d\n    src: %s\n    ref: %s\n    nmt: %s", i, src[i], tgt[i], text)
        return text

    # -- training ---------------------------------------------------------------
    def train(self, num_train_steps=None, steps_per_eval=None):
        hp = self.hp
        total = int(hp.num_train_steps if num_train_steps is None else num_train_steps)
        steps_per_stats = int(hp.steps_per_stats)
        steps_per_eval = int(steps_per_eval or hp.get("steps_per_eval") or 10 * steps_per_stats)
        steps_per_external = int(hp.get("steps_per_external_eval") or 5 * steps_per_eval)
        eng = self.sess.engine
        skip = hp.batch_size * int(hp.get("epoch_step", 0) or 0)
        it = self.train_iterator(skip).initialize()
        last_stats = last_eval = last_ext = eng.global_step
        while eng.global_step < total:
            try:
                batch = next(it)
            except StopIteration:       # finished an epoch
                hp.epoch_step = 0
                if self.chief:
                    log.info("# Finished an epoch, step %d.", eng.global_step)
                it.initialize(0)
                continue
            t0 = time.time()
            loss, pc, wc, gs, _ = self.sess.run(
                ["loss", "predict_count", "word_count", "global_step", "train_op"],
                batch.as_feed())
            hp.epoch_step = int(hp.get("epoch_step", 0) or 0) + 1
            self.stats.update(time.time() - t0, loss[0], pc[0], wc[0], batch.batch_size)
            gs = gs[0]
            if gs - last_stats >= steps_per_stats:
                last_stats = gs
                info = self.stats.process()
                info["global_step"] = gs
                info["learning_rate"] = self.sess.graph.optimizer.lr_at(gs)
                self.history.append(info)
                if self.chief:
                    log.info("  step %d lr %g step-time %.2fs wps %.2fK ppl %.2f",
                             gs, info["learning_rate"], info["avg_step_time"],
                             info["speed"] * self.num_workers, info["train_ppl"])
                self.stats.reset()
                if info["overflow"]:
                    break
            if gs - last_eval >= steps_per_eval:
                last_eval = gs
                self.sample_decode()
                self.internal_eval()
            if gs - last_ext >= steps_per_external:
                last_ext = gs
                self.external_eval()
        return self.history


def train(hp, resource_info="localhost", parallax_config=None, sync=True,
          num_train_steps=None, final_eval=True):
    """Build model + graph, start the Parallax session, train.  Returns the
    `Trainer` (with `.history`, and the session still open)."""
    import parallax_b200 as parallax
    if "src_vocab_size" not in hp:
        prepare_vocab(hp)
    elif "num_encoder_residual_layers" not in hp:
        extend_hparams(hp)
    if hp.random_seed is not None:
        torch.manual_seed(int(hp.random_seed))
    model = create_model(hp)
    if hp.get("embed_prefix"):
        for which, lang, vf in (("encoder", hp.src, hp.src_vocab_file),
                                ("decoder", hp.tgt, hp.tgt_vocab_file)):
            fn = "%s.%s" % (hp.embed_prefix, lang)
            if os.path.exists(fn) and not (which == "decoder" and hp.share_vocab):
                mat, n = vocab_utils.pretrained_embedding_matrix(vf, fn)
                model.load_pretrained_embeddings(which, mat, n)
                log.info("embedding_%s initialised from %s", which, fn)
    graph = nmt_graph(model, hp)
    src_vocab, tgt_vocab = vocab_utils.create_vocab_tables(
        hp.src_vocab_file, hp.tgt_vocab_file, hp.share_vocab)
    parallax_config = with_sparse_capacity(parallax_config or parallax.Config(), hp)
    sess, num_workers, worker_id, _ = parallax.parallel_run(
        graph, resource_info, sync=sync, parallax_config=parallax_config)
    tr = Trainer(hp, sess, num_workers, worker_id, model, src_vocab, tgt_vocab)
    if worker_id == 0 and hp.out_dir:
        save_hparams(hp.out_dir, hp)
    tr.train(num_train_steps)
    if final_eval:
        tr.final_ppl = tr.internal_eval()
        tr.final_scores = tr.external_eval()
    return tr
