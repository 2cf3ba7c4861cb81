"""Row → (partition, owner, local row) arithmetic for partitioned tables.

Parity: TF `embedding_lookup` partition strategies used by the reference
(`tensorflow/python/ops/embedding_ops.py:151-167`): "mod" ``p = id % P,
new_id = id // P``; "div" contiguous ranges where the first ``V % P``
partitions hold one extra row.  Partitions are placed on owners round-robin
(``owner = p % world``) which is what byte-greedy balancing yields for the
equal-sized partitions of one table (`ps/between_graph_parallel.py:49-70`).

A rank stores its partitions back to back: partition p is the
``p // world``-th partition on its owner and starts at local row
``(p // world) * rows_per_part`` (``rows_per_part = ceil(V / P)``).  The same
arithmetic is implemented on the device in `ops/csrc/kernels/sparse_common.cuh`.
"""
import torch


class TableLayout(object):
    def __init__(self, num_rows, num_partitions, world, strategy="mod",
                 replicated=False):
        self.V = int(num_rows)
        self.world = int(world)
        self.replicated = bool(replicated)
        if replicated:
            num_partitions, strategy = 1, "mod"
        self.P = int(num_partitions)
        self.strategy = strategy
        self.rows_per_part = (self.V + self.P - 1) // self.P
        # number of partitions each owner holds (max over owners)
        eff_world = 1 if replicated else self.world
        self.parts_per_owner = (self.P + eff_world - 1) // eff_world
        self.rows_local = self.parts_per_owner * self.rows_per_part
        if strategy == "div":
            self._extras = self.V % self.P
            self._base = self.V // self.P

    # -- tensor arithmetic (int64 tensors) --------------------------------------
    def partition_of(self, ids):
        if self.strategy == "mod":
            return ids % self.P
        # div: first `extras` partitions hold base+1 rows
        thr = self._extras * (self._base + 1)
        return torch.where(ids < thr, ids // (self._base + 1),
                           (ids - self._extras) // max(self._base, 1))

    def index_in_partition(self, ids):
        if self.strategy == "mod":
            return ids // self.P
        p = self.partition_of(ids)
        start = torch.where(p < self._extras, p * (self._base + 1),
                            p * self._base + self._extras)
        return ids - start

    def owner_of(self, ids):
        if self.replicated:
            return torch.zeros_like(ids)
        return self.partition_of(ids) % self.world

    def local_row_of(self, ids):
        if self.replicated:
            return ids
        p = self.partition_of(ids)
        return (p // self.world) * self.rows_per_part + \
            self.index_in_partition(ids)

    def global_ids_of_owner(self, owner):
        """(global ids, local rows) of every real row stored on `owner`."""
        ids = torch.arange(self.V, dtype=torch.int64)
        if self.replicated:
            return ids, ids
        mask = self.owner_of(ids) == owner
        g = ids[mask]
        return g, self.local_row_of(g)
