"""User-facing modules for declaring sparse (row-indexed) variables.

In the reference a variable is *sparse* when its gradient is an
`IndexedSlices` (produced by `tf.gather`/`embedding_lookup`), recorded through
the `GRADIENTS_INFO` collection (`tensorflow/python/ops/gradients_impl.py:943-947`,
`common/runner.py:40-60`), and it is *partitioned* when created under a
`parallax.get_partitioner(...)` variable scope (`examples/lm1b/language_model.py:34-45`).
The torch analogue: an `nn.Embedding` with ``sparse=True`` produces row-sparse
gradients; `parallax.nn.Embedding` is that module plus an optional
partitioner and an optional *lazy* initialiser so a 100M-row table is never
materialised on the host.
"""
import math

import torch
import torch.nn as tnn


class Embedding(tnn.Embedding):
    """``nn.Embedding(sparse=True)`` + partitioner + lazy init.

    Args:
      partitioner: result of `parallax.get_partitioner(min_p)` (optional).
      lazy: if True the weight lives on the ``meta`` device; each owner
        initialises only its shard (uniform in ``[-init_scale, init_scale]``).
      init_scale: default ``sqrt(3/embedding_dim)`` — TF's
        `uniform_unit_scaling_initializer` used by LM1B.
    """

    def __init__(self, num_embeddings, embedding_dim, partitioner=None,
                 lazy=False, init_scale=None, seed=1234, **kw):
        kw["sparse"] = True
        if lazy:
            kw["device"] = "meta"
        super().__init__(num_embeddings, embedding_dim, **kw)
        self.partitioner = partitioner
        self.lazy = bool(lazy)
        self.init_scale = float(init_scale) if init_scale is not None \
            else math.sqrt(3.0 / embedding_dim)
        self.init_seed = int(seed)
        if not lazy:
            with torch.no_grad():
                self.weight.uniform_(-self.init_scale, self.init_scale)


class EmbeddingBag(tnn.Module):
    """``nn.EmbeddingBag`` semantics (sum / mean / max over bags of ids, optional
    per-sample weights) on top of a sparse `Embedding`, so the table is an ordinary
    sparse variable: partitionable, looked up over the fabric, updated by its row
    owners.  Accepts 2-D input (fixed-size bags) or 1-D input + `offsets`."""

    def __init__(self, num_embeddings, embedding_dim, mode="sum", partitioner=None,
                 lazy=False, init_scale=None, seed=1234, include_last_offset=False,
                 padding_idx=None):
        super().__init__()
        if mode not in ("sum", "mean", "max"):
            raise ValueError("mode must be sum, mean or max")
        self.mode, self.include_last_offset, self.padding_idx = mode, include_last_offset, \
            padding_idx
        self.table = Embedding(num_embeddings, embedding_dim, partitioner=partitioner, lazy=lazy,
                               init_scale=init_scale, seed=seed)

    @property
    def weight(self):
        return self.table.weight

    def forward(self, input, offsets=None, per_sample_weights=None):
        if per_sample_weights is not None and self.mode != "sum":
            raise NotImplementedError("per_sample_weights needs mode='sum'")
        if input.dim() == 2:
            if offsets is not None:
                raise ValueError("offsets must be None for 2-D input")
            B, L = input.shape
            flat = input.reshape(-1)
            bag = torch.arange(B, device=input.device).repeat_interleave(L)
            psw = per_sample_weights.reshape(-1) if per_sample_weights is not None else None
        else:
            if offsets is None:
                raise ValueError("offsets are required for 1-D input")
            flat, psw = input, per_sample_weights
            off = offsets.to(input.device)
            B = off.numel() - (1 if self.include_last_offset else 0)
            ends = off[1:] if self.include_last_offset else \
                torch.cat([off[1:], off.new_tensor([flat.numel()])])
            lens = ends - off[:B]
            bag = torch.arange(B, device=input.device).repeat_interleave(lens)
        rows = self.table(flat)                               # [N, D] — the sparse lookup
        keep = None
        if self.padding_idx is not None:
            keep = (flat != self.padding_idx)
            rows = rows * keep[:, None].to(rows.dtype)
        if psw is not None:
            rows = rows * psw[:, None].to(rows.dtype)
        bag = bag.to(rows.device)
        D = rows.shape[1]
        if self.mode == "max":
            out = torch.full((B, D), float("-inf"), dtype=rows.dtype, device=rows.device)
            out = out.scatter_reduce(0, bag[:, None].expand(-1, D), rows, "amax",
                                     include_self=True)
            return torch.where(torch.isinf(out), torch.zeros_like(out), out)   # empty bags → 0
        out = torch.zeros(B, D, dtype=rows.dtype, device=rows.device).index_add_(0, bag, rows)
        if self.mode == "mean":
            ones = torch.ones(flat.numel(), device=rows.device, dtype=rows.dtype) if keep is None \
                else keep.to(rows.dtype)
            cnt = torch.zeros(B, dtype=rows.dtype, device=rows.device).index_add_(0, bag, ones)
            out = out / cnt.clamp(min=1.0)[:, None]
        return out


def partition(module, partitioner):
    """Attach a partitioner to an existing ``nn.Embedding(sparse=True)``."""
    assert isinstance(module, tnn.Embedding), "only nn.Embedding is partitionable"
    module.sparse = True
    module.partitioner = partitioner
    return module


def lookup_many(modules, ids, defer=False):
    """Rows of several embedding modules for the SAME ids (e.g. a softmax weight table
    and its bias table).  Declare the modules as a group on the model
    (``co_lookup_groups = [("softmax_w", "softmax_b")]``) and the NVLink fabric serves
    them with one lookup kernel, one push kernel and one owner kernel per step.
    ``defer=True`` returns a handle whose ``.rows()`` yields the tensors: issue the lookup
    early (e.g. on a side stream) and call ``.rows()`` where the rows are consumed."""
    from .parallel.engine import lookup_many as _lm
    return _lm(list(modules), ids, defer=defer)
