"""Sequence models of the reference's `nmt` and `skip_thoughts` examples.

* NMT lives in its own package, `parallax_b200.models.nmt`.
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
