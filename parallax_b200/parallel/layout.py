"""Row → (partition, owner, local row) arithmetic for partitioned tables.

Parity: TF `embedding_lookup` partition strategies used by the reference
(`tensorflow/python/ops/embedding_ops.py:151-167`): "mod" ``p = id % P,
new_id = id // P``; "div" contiguous ranges where the first ``V % P``
partitions hold one extra row.  Partitions are placed on owners by an explicit
map ``owners[p]`` — the engine fills it with the byte-greedy placement over ALL
sparse variables' partitions (`ps/between_graph_parallel.py:49-70`,
`assign_owners` below); without a map the placement is round-robin
(``owner = p % world``), which is what the greedy rule yields for one table.

A rank stores its partitions back to back: partition p is the ``slot[p]``-th
partition on its owner (in increasing p) and starts at local row
``slot[p] * rows_per_part`` (``rows_per_part = ceil(V / P)``).  The same
arithmetic runs on the device (`geom_map` in `ops/csrc/kernels/sparse.cu`,
which reads the two small maps from device memory).
"""
import torch


class TableLayout(object):
    def __init__(self, num_rows, num_partitions, world, strategy="mod",
                 replicated=False, owners=None):
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
        if owners is None or replicated:
            owners = [p % eff_world for p in range(self.P)]
        owners = [int(o) for o in owners]
        if len(owners) != self.P or any(o < 0 or o >= eff_world for o in owners):
            raise ValueError("owners must map each of the %d partitions to a rank in "
                             "[0, %d)" % (self.P, eff_world))
        self.owners = owners
        seen = [0] * eff_world
        self.slots = []
        for o in owners:
            self.slots.append(seen[o])
            seen[o] += 1
        self._owners_t = torch.tensor(owners, dtype=torch.int64)
        self._slots_t = torch.tensor(self.slots, dtype=torch.int64)
        self.parts_per_owner = max(seen)
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
        return self._owners_t.to(ids.device)[self.partition_of(ids)]

    def local_row_of(self, ids):
        if self.replicated:
            return ids
        p = self.partition_of(ids)
        return self._slots_t.to(ids.device)[p] * self.rows_per_part + \
            self.index_in_partition(ids)

    def same_placement(self, other):
        return (self.V, self.P, self.world, self.strategy, self.replicated, self.owners) == \
            (other.V, other.P, other.world, other.strategy, other.replicated, other.owners)

    def partition_rows(self, p):
        """Number of real rows in partition p."""
        if self.strategy == "mod":
            return (self.V - p + self.P - 1) // self.P if p < self.V else 0
        return self._base + (1 if p < self._extras else 0)

    def owner_chunks(self, owner, chunk=1 << 20):
        """Yield (global ids, local rows) of the real rows stored on `owner`, one
        partition (or a piece of one) at a time — O(chunk) host memory, so a 100 M-row
        table is never enumerated at once."""
        for p in range(self.P):
            if self.owners[p] != owner:
                continue
            rows = self.partition_rows(p)
            base_local = (0 if self.replicated else self.slots[p]) * self.rows_per_part
            for s in range(0, rows, chunk):
                idx = torch.arange(s, min(s + chunk, rows), dtype=torch.int64)
                if self.strategy == "mod":
                    g = idx * self.P + p
                else:
                    start = p * (self._base + 1) if p < self._extras else \
                        p * self._base + self._extras
                    g = idx + start
                yield g, (g if self.replicated else idx + base_local)

    def global_ids_of_owner(self, owner):
        """(global ids, local rows) of every real row stored on `owner`."""
        ids = torch.arange(self.V, dtype=torch.int64)
        if self.replicated:
            return ids, ids
        mask = self.owner_of(ids) == owner
        g = ids[mask]
        return g, self.local_row_of(g)


def assign_owners(items, world):
    """Byte-greedy placement of partitions on owners across ALL sparse variables
    (the reference's `GreedyLoadBalancingStrategy` + `byte_size_load_fn`,
    `ps/between_graph_parallel.py:49-70`): `items` is a list of
    ``(key, num_partitions, bytes_per_partition)`` in a deterministic order; every
    partition goes to the currently least-loaded owner (ties → lowest rank).
    Returns ``{key: [owner of partition 0, 1, …]}``.  Tables that are looked up
    together (a co-indexed group) must be passed as ONE item whose bytes are the
    sum over the members, so they end up with the same map."""
    from ..analyzer import greedy_load_balance
    sizes, index = [], []
    for key, nparts, nbytes in items:
        for p in range(int(nparts)):
            sizes.append(int(nbytes))
            index.append((key, p))
    placed = greedy_load_balance(sizes, max(int(world), 1))
    out = {}
    for (key, p), o in zip(index, placed):
        out.setdefault(key, []).append(o)
    return out
