"""Input pipeline of skip-thoughts: sharded record files → padded batches.

Parity: `examples/skip_thoughts/ops/input_ops.py:35-131` — `SentenceBatch(ids,
mask)`, `parse_example_batch` (three variable-length id lists per example →
dense padded ids + 0/1 masks) and `prefetch_input_data` (glob the file pattern,
sort, give each worker a contiguous slice of the files via
`parallax.shard.create_num_shards_and_shard_id` — the first `remainder` shards
get one more file —, shuffle through a bounded buffer).

Record files: the reference stores `tf.Example` protos in TFRecord shards; here a
shard is one ``.npz`` holding, for each of the three sentences, a flat int32 id
array plus offsets — loadable with `mmap_mode` and sliceable without parsing.
"""
import collections
import glob
import random

import numpy as np
import torch

from ... import shard as _shard
from ...log import parallax_log

SentenceBatch = collections.namedtuple("SentenceBatch", ("ids", "mask"))
FIELDS = ("encode", "decode_pre", "decode_post")


def write_shard(filename, examples):
    """examples: iterable of (encode_ids, decode_pre_ids, decode_post_ids)"""
    cols = {f: [] for f in FIELDS}
    for ex in examples:
        for f, ids in zip(FIELDS, ex):
            cols[f].append(np.asarray(ids, dtype=np.int32))
    out = {}
    for f in FIELDS:
        lens = np.array([len(a) for a in cols[f]], dtype=np.int64)
        out[f + "_offsets"] = np.concatenate([[0], np.cumsum(lens)])
        out[f + "_ids"] = np.concatenate(cols[f]) if cols[f] else np.zeros(0, np.int32)
    if not filename.endswith(".npz"):
        filename += ".npz"
    np.savez(filename, **out)
    return filename


def read_shard(filename):
    """→ list of (encode, decode_pre, decode_post) int32 arrays"""
    with np.load(filename) as z:
        cols = []
        for f in FIELDS:
            ids, off = z[f + "_ids"], z[f + "_offsets"]
            cols.append([ids[off[i]:off[i + 1]] for i in range(len(off) - 1)])
    return list(zip(*cols))


def _to_batch(seqs, pin=False):
    n, width = len(seqs), max(max((len(s) for s in seqs), default=0), 1)
    ids = torch.zeros(n, width, dtype=torch.int64)
    mask = torch.zeros(n, width, dtype=torch.int8)
    for i, s in enumerate(seqs):
        if len(s):
            ids[i, :len(s)] = torch.from_numpy(np.asarray(s, dtype=np.int64))
            mask[i, :len(s)] = 1
    if pin and torch.cuda.is_available():
        ids, mask = ids.pin_memory(), mask.pin_memory()
    return SentenceBatch(ids=ids, mask=mask)


def parse_example_batch(examples, pin=False):
    """list of id triples → (encode, decode_pre, decode_post) `SentenceBatch`es"""
    return tuple(_to_batch([ex[k] for ex in examples], pin) for k in range(3))


def files_for_shard(data_files, num_shards, shard_id):
    """contiguous slice of the sorted file list for one shard
    (`input_ops.py:91-103`)."""
    files = sorted(data_files)
    n, k, sid = len(files), int(num_shards), int(shard_id)
    size, rem = n // k, n % k
    begin = (size + 1) * sid if sid < rem + 1 else size * sid + rem
    count = size + 1 if sid < rem else size
    return files[begin:begin + count]


class InputQueue(object):
    """`prefetch_input_data` + `dequeue_many(batch_size)`: an endless stream of
    example batches read from this worker's files."""

    def __init__(self, file_pattern, batch_size, shuffle=True, capacity=640000, seed=None,
                 num_shards=None, shard_id=None, epochs=None, pin_memory=False):
        files = []
        for pattern in file_pattern.split(","):
            files.extend(glob.glob(pattern))
        if not files:
            raise ValueError("Found no input files matching %s" % file_pattern)
        parallax_log.info("Prefetching values from %d files matching %s", len(files),
                          file_pattern)
        self.files = sorted(files)
        if num_shards is None:
            num_shards, shard_id = _shard._get_or_create_num_shards_and_shard_id()
        self.num_shards, self.shard_id = num_shards, shard_id
        self.batch_size, self.shuffle, self.capacity = int(batch_size), shuffle, int(capacity)
        self.epochs, self.pin = epochs, pin_memory
        self.rng = random.Random(seed)

    def my_files(self):
        return files_for_shard(self.files, self.num_shards, self.shard_id)

    def _records(self):
        epoch = 0
        while self.epochs is None or epoch < self.epochs:
            files = list(self.my_files())
            if not files:
                raise ValueError("shard %d of %d has no input files (%d files in total)" %
                                 (int(self.shard_id), int(self.num_shards), len(self.files)))
            if self.shuffle:
                self.rng.shuffle(files)
            for fn in files:
                for ex in read_shard(fn):
                    yield ex
            epoch += 1

    def _shuffled(self):
        if not self.shuffle:
            yield from self._records()
            return
        buf, min_after = [], int(0.6 * self.capacity)
        for ex in self._records():
            buf.append(ex)
            if len(buf) > min_after:
                j = self.rng.randrange(len(buf))
                buf[j], buf[-1] = buf[-1], buf[j]
                yield buf.pop()
        self.rng.shuffle(buf)
        yield from buf

    def __iter__(self):
        batch = []
        for ex in self._shuffled():
            batch.append(ex)
            if len(batch) == self.batch_size:
                yield parse_example_batch(batch, self.pin)
                batch = []


def prefetch_input_data(file_pattern, batch_size, shuffle=True, capacity=640000, **kw):
    return InputQueue(file_pattern, batch_size, shuffle, capacity, **kw)
