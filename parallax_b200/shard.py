"""Data sharding helpers.

Parity: reference `common/shard.py:26-87` (`create_num_shards_and_shard_id`,
`shard`) and the transform that later overwrites the planted constants,
`graph_transform_lib.py:707-773` (`update_shard_values_for_worker`:
``num_shards *= num_workers; shard_id += num_replicas_per_worker*worker_id``).

In the reference the two values are int64 graph constants that the graph
transform patches.  Here they are tiny mutable handles (`ShardValue`) that
`parallel_run` updates in place on each worker before the user's `run`
function iterates its data, so a dataset built *before* `parallel_run`
(the reference contract) still shards correctly.
"""
import itertools

NUM_SHARDS = "num_shards"
SHARD_ID = "shard_id"
SHARD_FILTER_PRED = "shard_filter_predicate"


class ShardValue(object):
    """An int-like late-bound value."""
    __slots__ = ("value", "name")

    def __init__(self, value, name):
        self.value = int(value)
        self.name = name

    def __int__(self):
        return self.value

    __index__ = __int__

    def __repr__(self):
        return "ShardValue(%s=%d)" % (self.name, self.value)

    def __eq__(self, other):
        return int(self) == int(other)

    def __hash__(self):
        return hash((self.name, id(self)))


class _Registry(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.num_shards = None
        self.shard_id = None
        self.filters = []
        self.assignment = None      # (num_workers, worker_id, replicas) once known


_registry = _Registry()


def reset():
    """Forget planted shard values (new "graph")."""
    _registry.reset()


def create_num_shards_and_shard_id():
    """Create and return the (num_shards, shard_id) handles, initialised to
    (1, 0).  Raises ValueError if they already exist
    (reference `common/shard.py:38-45`)."""
    if _registry.num_shards is not None:
        raise ValueError('"num_shards" already exists.')
    if _registry.shard_id is not None:
        raise ValueError('"shard_id" already exists.')
    _registry.num_shards = ShardValue(1, NUM_SHARDS)
    _registry.shard_id = ShardValue(0, SHARD_ID)
    if _registry.assignment is not None:     # created after parallel_run
        update_shard_values_for_worker(*_registry.assignment)
    return _registry.num_shards, _registry.shard_id


def _get_or_create_num_shards_and_shard_id():
    if _registry.num_shards is None:
        create_num_shards_and_shard_id()
    return _registry.num_shards, _registry.shard_id


class ShardedDataset(object):
    """`ds.shard(num_shards, shard_id)` with late-bound values: element `i`
    is kept iff ``i % num_shards == shard_id``.  Works on any iterable; if the
    source supports ``__len__``/``__getitem__`` so does the result."""

    def __init__(self, ds, num_shards, shard_id):
        self._ds = ds
        self._num_shards = num_shards
        self._shard_id = shard_id

    def __iter__(self):
        n, k = int(self._num_shards), int(self._shard_id)
        return itertools.islice(iter(self._ds), k, None, n)

    def __len__(self):
        n, k = int(self._num_shards), int(self._shard_id)
        total = len(self._ds)
        return max(0, (total - k + n - 1) // n)

    def __getitem__(self, i):
        n, k = int(self._num_shards), int(self._shard_id)
        if i < 0 or i >= len(self):
            raise IndexError(i)
        return self._ds[k + i * n]


def shard(ds):
    """Same effect as ``ds.shard(num_shards, index)`` with the values Parallax
    assigns to this worker (reference `common/shard.py:69-87`)."""
    num_shards, shard_id = _get_or_create_num_shards_and_shard_id()
    out = ShardedDataset(ds, num_shards, shard_id)
    _registry.filters.append(out)
    return out


def update_shard_values_for_worker(num_workers, worker_id,
                                   num_replicas_per_worker=1):
    """``num_shards *= num_workers`` and
    ``shard_id += num_replicas_per_worker * worker_id``
    (reference `graph_transform_lib.py:707-722`).  Idempotent per process:
    values are recomputed from the planted base (1, 0)."""
    _registry.assignment = (num_workers, worker_id, num_replicas_per_worker)
    if _registry.num_shards is None:
        return None
    _registry.num_shards.value = 1 * num_workers * num_replicas_per_worker
    _registry.shard_id.value = 0 + num_replicas_per_worker * worker_id
    return int(_registry.num_shards), int(_registry.shard_id)


class DistributedShardSampler(object):
    """torch ``Sampler`` built on the same handles (for ``DataLoader``)."""

    def __init__(self, data_source):
        self._n = len(data_source)
        self._num_shards, self._shard_id = \
            _get_or_create_num_shards_and_shard_id()

    def __iter__(self):
        return iter(range(int(self._shard_id), self._n, int(self._num_shards)))

    def __len__(self):
        n, k = int(self._num_shards), int(self._shard_id)
        return max(0, (self._n - k + n - 1) // n)
