"""Attention mechanisms of the NMT example.

Parity: `examples/nmt/attention_model.py:150-183` `create_attention_mechanism`
— ``luong`` / ``scaled_luong`` (`tf.contrib.seq2seq.LuongAttention`) and
``bahdanau`` / ``normed_bahdanau`` (`BahdanauAttention(normalize=True)`).  All
four project the memory (encoder outputs) once per batch with a bias-free
``memory_layer``; positions past `source_sequence_length` score −inf.

* luong           : score = q · k                  (× g when scaled, g₀ = 1)
* bahdanau        : score = v · tanh(k + W_q q)
* normed_bahdanau : score = g·v/‖v‖ · tanh(k + W_q q + b),  g₀ = √(1/units)
"""
import math

import torch
import torch.nn as nn

ATTENTION_OPTIONS = ("luong", "scaled_luong", "bahdanau", "normed_bahdanau")


class AttentionMechanism(nn.Module):
    def __init__(self, option, num_units, memory_size, query_size=None, init_weight=0.1):
        super().__init__()
        if option not in ATTENTION_OPTIONS:
            raise ValueError("Unknown attention option %s" % option)
        self.option, self.num_units = option, num_units
        self.memory_layer = nn.Linear(memory_size, num_units, bias=False)
        if option in ("bahdanau", "normed_bahdanau"):
            self.query_layer = nn.Linear(query_size or num_units, num_units, bias=False)
            self.v = nn.Parameter(torch.empty(num_units).uniform_(-init_weight, init_weight))
            if option == "normed_bahdanau":
                self.g = nn.Parameter(torch.tensor(math.sqrt(1.0 / num_units)))
                self.b = nn.Parameter(torch.zeros(num_units))
        elif option == "scaled_luong":
            self.g = nn.Parameter(torch.tensor(1.0))

    def prepare(self, memory, lengths):
        """memory [B,S,M], lengths [B] → (keys [B,S,U], values [B,S,M], pad mask [B,S])"""
        S = memory.shape[1]
        pad = torch.arange(S, device=memory.device)[None, :] >= lengths.to(memory.device)[:, None]
        # TF zeroes the memory past each sequence's end before projecting it
        values = memory.masked_fill(pad[..., None], 0.0)
        return self.memory_layer(values), values, pad

    def score(self, query, keys):
        """query [B,U] (or [B,T,U]) , keys [B,S,U] → scores [B,S] (or [B,T,S])"""
        single = query.dim() == 2
        q = query[:, None, :] if single else query
        if self.option in ("luong", "scaled_luong"):
            s = torch.bmm(q, keys.transpose(1, 2))
            if self.option == "scaled_luong":
                s = s * self.g.to(s.dtype)
        else:
            pq = self.query_layer(q)                                   # [B,T,U]
            v = self.v
            h = keys[:, None, :, :] + pq[:, :, None, :]
            if self.option == "normed_bahdanau":
                v = self.g * self.v / self.v.norm()
                h = h + self.b
            s = (torch.tanh(h) * v.to(h.dtype)).sum(-1)                # [B,T,S]
        return s[:, 0] if single else s

    def forward(self, query, keys, values, pad):
        """→ (context [B,M], alignments [B,S]) for a single decoding step"""
        s = self.score(query, keys).float().masked_fill(pad, float("-inf"))
        align = torch.softmax(s, -1)
        ctx = torch.bmm(align[:, None, :].to(values.dtype), values)[:, 0]
        return ctx, align
