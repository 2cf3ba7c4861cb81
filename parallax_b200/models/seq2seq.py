"""Sequence models of the reference's `nmt` and `skip_thoughts` examples.

* `NMT` — GNMT-style encoder/decoder LSTM stack with Luong attention
  (`examples/nmt/model.py`, `gnmt_model.py`, `standard_hparams/*.json`); source
  and target embeddings are *partitioned sparse variables* created under
  `parallax.get_partitioner` (`examples/nmt/model_helper.py:309-311`).
* `SkipThoughts` — GRU sentence encoder with two GRU decoders (previous / next
  sentence), Adam (`examples/skip_thoughts/skip_thoughts_model.py`).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import nn as pnn
from .. import optim
from ..graph import Graph, ClipByGlobalNorm
from ..partitions import get_partitioner


class NMT(nn.Module):
    def __init__(self, src_vocab=32000, tgt_vocab=32000, units=512, layers=2,
                 num_partitions=4, dropout=0.2):
        super().__init__()
        part = get_partitioner(num_partitions)
        self.src_emb = pnn.Embedding(src_vocab, units, partitioner=part)
        self.tgt_emb = pnn.Embedding(tgt_vocab, units, partitioner=part, seed=7)
        self.encoder = nn.LSTM(units, units, layers, batch_first=True, dropout=dropout)
        self.decoder = nn.LSTM(units, units, layers, batch_first=True, dropout=dropout)
        self.attn_w = nn.Linear(units, units, bias=False)
        self.attn_out = nn.Linear(2 * units, units, bias=False)
        self.proj = nn.Linear(units, tgt_vocab, bias=False)

    def forward(self, src, tgt_in, tgt_out, tgt_weight=None):
        dt = self.proj.weight.dtype
        enc, state = self.encoder(self.src_emb(src).to(dt))
        dec, _ = self.decoder(self.tgt_emb(tgt_in).to(dt), state)
        score = torch.bmm(self.attn_w(dec), enc.transpose(1, 2))            # Luong "general"
        ctx = torch.bmm(F.softmax(score.float(), -1).to(dt), enc)
        att = torch.tanh(self.attn_out(torch.cat([ctx, dec], -1)))
        logits = self.proj(att).float()
        loss = F.cross_entropy(logits.view(-1, logits.shape[-1]), tgt_out.reshape(-1),
                               reduction="none")
        if tgt_weight is not None:
            loss = loss * tgt_weight.reshape(-1).float()
            return {"loss": loss.sum() / src.shape[0], "logits": logits}
        return {"loss": loss.mean() * tgt_out.shape[1], "logits": logits}


def nmt_graph(model, learning_rate=1.0, max_gradient_norm=5.0):
    """SGD + global-norm clipping of the dense variables (`nmt/model.py:160-190`)."""
    dense = lambda n: not n.endswith("_emb.weight")
    return Graph(model, optimizer=optim.GradientDescent(learning_rate),
                 grad_rules=[ClipByGlobalNorm(max_gradient_norm, params=dense)], name="nmt")


class SkipThoughts(nn.Module):
    def __init__(self, vocab=20000, word_dim=620, units=2400, num_partitions=2):
        super().__init__()
        part = get_partitioner(num_partitions)
        self.word_emb = pnn.Embedding(vocab, word_dim, partitioner=part, init_scale=0.1)
        self.encoder = nn.GRU(word_dim, units, batch_first=True)
        self.dec_pre = nn.GRU(word_dim + units, units, batch_first=True)
        self.dec_post = nn.GRU(word_dim + units, units, batch_first=True)
        self.logits = nn.Linear(units, vocab)

    def _decode(self, gru, thought, ids_in, ids_out, mask):
        dt = self.logits.weight.dtype
        e = self.word_emb(ids_in).to(dt)
        inp = torch.cat([e, thought.unsqueeze(1).expand(-1, e.shape[1], -1)], -1)
        out, _ = gru(inp)
        lg = self.logits(out).float()
        l = F.cross_entropy(lg.view(-1, lg.shape[-1]), ids_out.reshape(-1), reduction="none")
        m = mask.reshape(-1).float()
        return (l * m).sum() / m.sum().clamp(min=1.0)

    def forward(self, encode_ids, pre_in, pre_out, pre_mask, post_in, post_out, post_mask):
        dt = self.logits.weight.dtype
        _, h = self.encoder(self.word_emb(encode_ids).to(dt))
        thought = h[-1]
        loss = self._decode(self.dec_pre, thought, pre_in, pre_out, pre_mask) + \
            self._decode(self.dec_post, thought, post_in, post_out, post_mask)
        return {"loss": loss, "thought_vectors": thought}


def skip_thoughts_graph(model, learning_rate=0.0008, clip_gradient_norm=5.0):
    dense = lambda n: n != "word_emb.weight"
    return Graph(model, optimizer=optim.Adam(learning_rate),
                 grad_rules=[ClipByGlobalNorm(clip_gradient_norm, params=dense)],
                 name="skip_thoughts")
