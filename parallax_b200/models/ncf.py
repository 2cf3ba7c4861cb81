"""NCF / NeuMF recommender with very large embedding tables (BASELINE.json
config 4: "NCF / 100M-row embedding table, PS-mode with 8-way sparse-variable
partitioning + local aggregation").  Not part of the reference's examples; it
exercises the same PS-partitioned sparse path at a scale where the table does
not fit one replica comfortably and lookups are dominated by remote rows.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import nn as pnn
from .. import optim
from ..graph import Graph
from ..partitions import get_partitioner


class NeuMF(nn.Module):
    def __init__(self, num_users=100_000_000, num_items=1_000_000, mf_dim=32,
                 mlp_layers=(64, 32, 16), num_partitions=8, lazy=True):
        super().__init__()
        part = get_partitioner(num_partitions)
        d_mlp = mlp_layers[0] // 2
        # one fused row per entity: [GMF part | MLP part] -> a single lookup per table
        self.user_emb = pnn.Embedding(num_users, mf_dim + d_mlp, partitioner=part, lazy=lazy,
                                      init_scale=0.01)
        self.item_emb = pnn.Embedding(num_items, mf_dim + d_mlp, partitioner=part, lazy=lazy,
                                      init_scale=0.01, seed=77)
        self.mf_dim = mf_dim
        layers, d = [], mlp_layers[0]
        for h in mlp_layers[1:]:
            layers += [nn.Linear(d, h), nn.ReLU()]
            d = h
        self.mlp = nn.Sequential(*layers)
        self.out = nn.Linear(mf_dim + d, 1)

    def forward(self, users, items, labels):
        dt = self.out.weight.dtype
        u = self.user_emb(users).to(dt)
        i = self.item_emb(items).to(dt)
        gmf = u[:, :self.mf_dim] * i[:, :self.mf_dim]
        mlp = self.mlp(torch.cat([u[:, self.mf_dim:], i[:, self.mf_dim:]], 1))
        logit = self.out(torch.cat([gmf, mlp], 1)).squeeze(-1).float()
        return {"loss": F.binary_cross_entropy_with_logits(logit, labels.float()),
                "logits": logit}


def ncf_graph(model, learning_rate=0.001):
    return Graph(model, optimizer=optim.Adam(learning_rate), name="ncf")
