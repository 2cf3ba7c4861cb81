"""Dense / sparse classification of a single-device model.

Parity: reference `common/runner.py:40-60` (`_get_grads`: a variable is
*sparse* iff its gradient in the `GRADIENTS_INFO` collection is an
`IndexedSlices`) and the mode degeneration rules of `:93-119` (HYBRID with no
sparse variable runs as MPI, with no dense variable as PS).

Torch mapping: parameters of ``nn.Embedding(sparse=True)`` produce row-sparse
gradients — those are the sparse variables; every other trainable parameter
is dense.  `probe=True` additionally runs one backward on example feeds and
reports any parameter whose ``.grad`` comes back as a ``torch.sparse`` tensor
from somewhere the static walk missed (e.g. ``F.embedding(..., sparse=True)``
on a bare Parameter), which the engine cannot shard and therefore rejects
loudly.
"""
from collections import OrderedDict

import torch
import torch.nn as tnn


class VariableInfo(object):
    __slots__ = ("name", "shape", "numel", "dtype", "sparse", "module_path",
                 "partitions", "owner", "bucket", "requires_grad")

    def __init__(self, name, p, sparse, module_path=None, partitions=None):
        self.name = name
        self.shape = tuple(p.shape)
        self.numel = int(p.numel())
        self.dtype = p.dtype
        self.sparse = sparse
        self.module_path = module_path
        self.partitions = partitions
        self.owner = None
        self.bucket = None
        self.requires_grad = bool(p.requires_grad)

    def nbytes(self):
        return self.numel * torch.empty((), dtype=self.dtype).element_size()

    def as_dict(self):
        return {k: (str(getattr(self, k)) if k == "dtype" else getattr(self, k))
                for k in self.__slots__}


class Analysis(object):
    def __init__(self, variables, sparse_modules):
        self.variables = variables            # OrderedDict name -> VariableInfo
        self.sparse_modules = sparse_modules  # OrderedDict module path -> module

    @property
    def dense(self):
        return [v for v in self.variables.values()
                if not v.sparse and v.requires_grad]

    @property
    def sparse(self):
        return [v for v in self.variables.values() if v.sparse]

    def effective_run_option(self, run_option):
        """Reference `common/runner.py:93-119`."""
        if run_option == "HYBRID":
            if not self.sparse:
                return "MPI"
            if not self.dense:
                return "PS"
        return run_option

    def report(self):
        return {"variables": [v.as_dict() for v in self.variables.values()],
                "num_dense": len(self.dense), "num_sparse": len(self.sparse),
                "dense_bytes": sum(v.nbytes() for v in self.dense),
                "sparse_bytes": sum(v.nbytes() for v in self.sparse)}


def analyze(model, world=1):
    """Static walk: tag every parameter dense or sparse."""
    sparse_param_ids = {}
    sparse_modules = OrderedDict()
    for path, m in model.named_modules():
        if isinstance(m, tnn.Embedding) and m.sparse:
            sparse_param_ids[id(m.weight)] = path
            sparse_modules[path] = m
        elif isinstance(m, tnn.EmbeddingBag) and m.sparse:
            raise NotImplementedError(
                "nn.EmbeddingBag(sparse=True) at %r: use parallax.nn.EmbeddingBag (same "
                "semantics on a shardable table)" % path)
    variables = OrderedDict()
    for name, p in model.named_parameters():
        if id(p) in sparse_param_ids:
            path = sparse_param_ids[id(p)]
            part = getattr(sparse_modules[path], "partitioner", None)
            nparts = part.num_partitions if part is not None else max(world, 1)
            variables[name] = VariableInfo(name, p, True, path, nparts)
        else:
            variables[name] = VariableInfo(name, p, False)
    return Analysis(variables, sparse_modules)


def probe_sparse_grads(model, loss_fn):
    """Run one backward and return names of parameters whose gradient is a
    torch sparse tensor (dynamic counterpart of GRADIENTS_INFO)."""
    model.zero_grad(set_to_none=True)
    loss = loss_fn()
    loss.backward()
    out = [n for n, p in model.named_parameters()
           if p.grad is not None and p.grad.is_sparse]
    model.zero_grad(set_to_none=True)
    return out


def greedy_load_balance(sizes, num_owners):
    """Assign items to owners minimising the running byte load — the
    reference's `GreedyLoadBalancingStrategy` with `byte_size_load_fn`
    (`ps/between_graph_parallel.py:49-70`).  Returns a list of owner ids,
    processed in the given order."""
    loads = [0] * num_owners
    owners = []
    for s in sizes:
        o = min(range(num_owners), key=lambda i: (loads[i], i))
        owners.append(o)
        loads[o] += int(s)
    return owners
