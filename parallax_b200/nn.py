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


def partition(module, partitioner):
    """Attach a partitioner to an existing ``nn.Embedding(sparse=True)``."""
    assert isinstance(module, tnn.Embedding), "only nn.Embedding is partitionable"
    module.sparse = True
    module.partitioner = partitioner
    return module
