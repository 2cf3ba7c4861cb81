"""Input pipeline of the NMT example: sharded, shuffled, length-bucketed batches.

Parity: `examples/nmt/utils/iterator_utils.py:31-204` — `get_iterator`
(zip source/target lines → `parallax.shard.shard` (`:103`) → skip → shuffle →
split → drop empty pairs → truncate to `src_max_len`/`tgt_max_len` → vocabulary
lookup → ``tgt_in = <s> + tgt``, ``tgt_out = tgt + </s>`` → lengths →
`group_by_window` bucketing by ``max(len_src, len_tgt) // bucket_width`` with
batches padded with the ``</s>`` ids) and `get_infer_iterator` (`:31-70`).

What is different on purpose: batches are batch-major int64 tensors built on
the host (optionally in pinned memory, ready for the engine's non-blocking
H2D copy), and `static_shapes=True` pads every batch of a bucket to the
bucket's upper bound and drops ragged final windows, so a CUDA-graph-captured
training step sees at most `num_buckets` distinct feed signatures instead of
one per batch.
"""
import collections
import random

import torch

from ... import shard as _shard


class BatchedInput(collections.namedtuple(
        "BatchedInput", ("source", "target_input", "target_output",
                         "source_sequence_length", "target_sequence_length"))):
    """One batch.  `as_feed()` gives the feed_dict of `Seq2Seq.forward`."""

    def as_feed(self):
        return {k: [v] for k, v in self._asdict().items() if v is not None}

    @property
    def batch_size(self):
        return int(self.source.shape[0])


def _pad(rows, width, pad_id, pin):
    out = torch.full((len(rows), width), pad_id, dtype=torch.int64)
    for i, r in enumerate(rows):
        if r:
            out[i, :len(r)] = torch.tensor(r, dtype=torch.int64)
    return out.pin_memory() if pin else out


def _lines(dataset):
    if isinstance(dataset, str):
        with open(dataset, encoding="utf-8") as f:
            for line in f:
                yield line.rstrip("\n")
    else:
        for line in dataset:
            yield line


class NMTIterator(object):
    """Re-initialisable iterator over training/eval batches."""

    def __init__(self, src_dataset, tgt_dataset, src_vocab_table, tgt_vocab_table,
                 batch_size, sos, eos, random_seed=None, num_buckets=1,
                 src_max_len=None, tgt_max_len=None, output_buffer_size=None,
                 skip_count=None, num_shards=None, shard_index=None,
                 reshuffle_each_iteration=True, shuffle=True, static_shapes=False,
                 pin_memory=False):
        self.src_dataset, self.tgt_dataset = src_dataset, tgt_dataset
        self.src_vocab, self.tgt_vocab = src_vocab_table, tgt_vocab_table
        self.batch_size = int(batch_size)
        self.src_eos_id = src_vocab_table.lookup(eos)
        self.tgt_sos_id = tgt_vocab_table.lookup(sos)
        self.tgt_eos_id = tgt_vocab_table.lookup(eos)
        self.num_buckets = max(int(num_buckets), 1)
        self.src_max_len, self.tgt_max_len = src_max_len, tgt_max_len
        self.output_buffer_size = output_buffer_size or self.batch_size * 1000
        self.skip_count = skip_count or 0
        # shards default to the worker's (num_shards, shard_id) pair planted by
        # `parallax.shard` — late-bound handles resolved by `parallel_run`
        if num_shards is None:
            num_shards, shard_index = _shard._get_or_create_num_shards_and_shard_id()
        self.num_shards, self.shard_index = num_shards, shard_index
        self.reshuffle = reshuffle_each_iteration
        self.shuffle = shuffle
        self.static_shapes = static_shapes
        self.pin = pin_memory and torch.cuda.is_available()
        self._seed = random_seed
        self._epoch = 0
        self._it = None
        if src_max_len:
            self.bucket_width = (src_max_len + self.num_buckets - 1) // self.num_buckets
        else:
            self.bucket_width = 10

    # ----------------------------------------------------------- pipeline
    def _shard_ids(self):
        return int(self.num_shards), int(self.shard_index or 0)

    def _examples(self, skip_count):
        ns, si = self._shard_ids()
        pairs = zip(_lines(self.src_dataset), _lines(self.tgt_dataset))
        n = 0
        for i, (s, t) in enumerate(pairs):
            if ns > 1 and i % ns != si:
                continue
            n += 1
            if n <= skip_count:
                continue
            yield s, t

    def _shuffled(self, it):
        """streaming shuffle with a bounded buffer (`tf.data.Dataset.shuffle`)"""
        if not self.shuffle:
            for x in it:
                yield x
            return
        seed = self._seed if self._seed is not None else random.randrange(1 << 30)
        if self.reshuffle:
            seed += self._epoch
        rng = random.Random(seed)
        buf = []
        for x in it:
            if len(buf) < self.output_buffer_size:
                buf.append(x)
                continue
            j = rng.randrange(len(buf))
            buf[j], x = x, buf[j]
            yield x
        rng.shuffle(buf)
        for x in buf:
            yield x

    def _encoded(self, it):
        for s, t in it:
            sw, tw = s.split(), t.split()
            if not sw or not tw:
                continue
            if self.src_max_len:
                sw = sw[:self.src_max_len]
            if self.tgt_max_len:
                tw = tw[:self.tgt_max_len]
            si, ti = self.src_vocab.encode(sw), self.tgt_vocab.encode(tw)
            yield si, [self.tgt_sos_id] + ti, ti + [self.tgt_eos_id]

    def _bucket_id(self, ex):
        if self.num_buckets <= 1:
            return 0
        b = max(len(ex[0]) // self.bucket_width, len(ex[1]) // self.bucket_width)
        return min(self.num_buckets, b)

    def _make_batch(self, exs, bucket):
        src_len = torch.tensor([len(e[0]) for e in exs], dtype=torch.int64)
        tgt_len = torch.tensor([len(e[1]) for e in exs], dtype=torch.int64)
        if self.static_shapes and self.num_buckets > 1:
            sw = tw = (bucket + 1) * self.bucket_width
            sw = max(sw, int(src_len.max()))
            tw = max(tw + 1, int(tgt_len.max()))
        elif self.static_shapes:
            sw = self.src_max_len or int(src_len.max())
            tw = (self.tgt_max_len + 1) if self.tgt_max_len else int(tgt_len.max())
        else:
            sw, tw = int(src_len.max()), int(tgt_len.max())
        return BatchedInput(
            source=_pad([e[0] for e in exs], sw, self.src_eos_id, self.pin),
            target_input=_pad([e[1] for e in exs], tw, self.tgt_eos_id, self.pin),
            target_output=_pad([e[2] for e in exs], tw, self.tgt_eos_id, self.pin),
            source_sequence_length=src_len, target_sequence_length=tgt_len)

    def _batches(self, skip_count):
        windows = collections.defaultdict(list)
        for ex in self._encoded(self._shuffled(self._examples(skip_count))):
            b = self._bucket_id(ex)
            w = windows[b]
            w.append(ex)
            if len(w) == self.batch_size:
                yield self._make_batch(w, b)
                windows[b] = []
        if not self.static_shapes:
            for b in sorted(windows):        # ragged final windows
                if windows[b]:
                    yield self._make_batch(windows[b], b)

    # ---------------------------------------------------------- iterator
    def initialize(self, skip_count=None):
        """(Re)start the epoch — the counterpart of running
        `iterator.initializer` with the `skip_count` placeholder."""
        sc = self.skip_count if skip_count is None else skip_count
        self._it = self._batches(sc)
        self._epoch += 1
        return self

    def __iter__(self):
        if self._it is None:
            self.initialize()
        return self

    def __next__(self):
        if self._it is None:
            self.initialize()
        try:
            return next(self._it)
        except StopIteration:
            self._it = None
            raise


def get_iterator(src_dataset, tgt_dataset, src_vocab_table, tgt_vocab_table, batch_size,
                 sos, eos, random_seed=None, num_buckets=1, src_max_len=None,
                 tgt_max_len=None, **kw):
    return NMTIterator(src_dataset, tgt_dataset, src_vocab_table, tgt_vocab_table,
                       batch_size, sos, eos, random_seed, num_buckets, src_max_len,
                       tgt_max_len, **kw)


class InferIterator(object):
    """Source-only batches in file order (`get_infer_iterator`, `:31-70`)."""

    def __init__(self, src_dataset, src_vocab_table, batch_size, eos, src_max_len=None):
        self.src_dataset, self.src_vocab = src_dataset, src_vocab_table
        self.batch_size, self.src_max_len = int(batch_size), src_max_len
        self.src_eos_id = src_vocab_table.lookup(eos)

    def __iter__(self):
        rows = []
        for line in _lines(self.src_dataset):
            w = line.split()
            if self.src_max_len:
                w = w[:self.src_max_len]
            rows.append(self.src_vocab.encode(w))
            if len(rows) == self.batch_size:
                yield self._batch(rows)
                rows = []
        if rows:
            yield self._batch(rows)

    def _batch(self, rows):
        lens = torch.tensor([len(r) for r in rows], dtype=torch.int64)
        return BatchedInput(source=_pad(rows, max(int(lens.max()), 1), self.src_eos_id, False),
                            target_input=None, target_output=None,
                            source_sequence_length=lens, target_sequence_length=None)


def get_infer_iterator(src_dataset, src_vocab_table, batch_size, eos, src_max_len=None):
    return InferIterator(src_dataset, src_vocab_table, batch_size, eos, src_max_len)
