"""Tiny models for plumbing tests and the quick-start example.

* `LinearRegression` — the reference's `examples/simple/simple_driver.py:93-136`
  (2→1 linear regression trained with SGD).
* `MLPWithEmbedding` — BASELINE.json config 1: a 2-layer MLP on top of a tiny
  embedding table (one sparse + several dense variables), the smallest model
  that exercises both routes.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import nn as pnn


class LinearRegression(nn.Module):
    def __init__(self, in_features=2):
        super().__init__()
        self.linear = nn.Linear(in_features, 1)

    def forward(self, x, y):
        pred = self.linear(x).squeeze(-1)
        return {"loss": F.mse_loss(pred, y), "pred": pred}


class MLPWithEmbedding(nn.Module):
    def __init__(self, vocab=64, emb=8, hidden=16, classes=4, partitioner=None,
                 seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.emb = pnn.Embedding(vocab, emb, partitioner=partitioner)
        with torch.no_grad():
            self.emb.weight.copy_(torch.randn(vocab, emb, generator=g) * 0.1)
        self.fc1 = nn.Linear(emb, hidden)
        self.fc2 = nn.Linear(hidden, classes)
        with torch.no_grad():
            for p in list(self.fc1.parameters()) + list(self.fc2.parameters()):
                p.copy_(torch.randn(p.shape, generator=g) * 0.2)

    def forward(self, ids, labels):
        # ids: [B, T] -> mean-pooled embedding
        h = self.emb(ids).mean(dim=1)
        h = torch.tanh(self.fc1(h))
        logits = self.fc2(h)
        return {"loss": F.cross_entropy(logits.float(), labels), "logits": logits}
