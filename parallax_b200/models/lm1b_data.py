"""LM1B input pipeline: vocabulary + sharded sentence stream packed into
``[batch, num_steps]`` windows.

Parity: `examples/lm1b/data_utils.py:18-148` / `lm1b_input.py` — `Vocabulary`
(``word count`` per line, ``<S>`` / ``<UNK>``, optional `num_tokens_limit`),
`Dataset` (files of the pattern are split over the workers, sentences of a file
are shuffled unless `deterministic`, every sentence is ``<S> w… <S>``; each of
the `batch_size` rows is a stream of concatenated sentences cut into windows of
`num_steps`; ``y`` is ``x`` shifted by one, ``w`` masks padding;
`iterate_once` / `iterate_forever`).

Tokenisation and word → id lookup run in native code (`NativeVocab`) when the
library is built; the pure-python path is the fallback and the test oracle.
"""
import glob
import random

import numpy as np


class Vocabulary(object):
    def __init__(self):
        self._token_to_id, self._token_to_count, self._id_to_token = {}, {}, []
        self._s_id = self._unk_id = None
        self._native = None

    unk, s = "<UNK>", "<S>"

    @property
    def num_tokens(self):
        return len(self._id_to_token)

    @property
    def unk_id(self):
        return self._unk_id

    @property
    def s_id(self):
        return self._s_id

    def add(self, token, count=0):
        self._token_to_id[token] = len(self._id_to_token)
        self._token_to_count[token] = count
        self._id_to_token.append(token)

    def finalize(self):
        for special in (self.s, self.unk):
            if special not in self._token_to_id:
                self.add(special)
        self._s_id, self._unk_id = self._token_to_id[self.s], self._token_to_id[self.unk]
        self._native = None
        return self

    def get_id(self, token):
        return self._token_to_id.get(token, self._unk_id)

    def get_token(self, id_):
        return self._id_to_token[id_]

    def get_count(self, token):
        return self._token_to_count.get(token, 0)

    def encode_line(self, line, native=True):
        """``<S> w1 w2 … <S>`` as a list of ids"""
        if native:
            nv = self._native_vocab()
            if nv is not None:
                return [self._s_id] + nv.encode(line) + [self._s_id]
        return [self._s_id] + [self.get_id(w) for w in line.split()] + [self._s_id]

    def _native_vocab(self):
        if self._native is None:
            try:
                from ..utils.dataloader import NativeVocab
                self._native = NativeVocab(self._id_to_token, self._unk_id)
            except Exception:            # library not built: python lookup
                self._native = False
        return self._native or None

    @staticmethod
    def from_file(filename, num_tokens_limit=None):
        v = Vocabulary()
        with open(filename, encoding="utf-8") as f:
            for line in f:
                parts = line.split()
                if not parts:
                    continue
                v.add(parts[0], int(parts[1]) if len(parts) > 1 and parts[1].isdigit() else 0)
                if num_tokens_limit is not None and v.num_tokens == num_tokens_limit:
                    break
        return v.finalize()


class Dataset(object):
    def __init__(self, vocab, file_pattern, deterministic=False, seed=None, native=True):
        self._vocab, self._pattern, self._det = vocab, file_pattern, deterministic
        self._rng = random.Random(seed)
        self._native = native

    def files(self, num_workers=1, worker_id=0):
        names = self._pattern if isinstance(self._pattern, (list, tuple)) \
            else glob.glob(self._pattern)
        return sorted(names)[worker_id::num_workers]

    def _sentences(self, files):
        for fn in files:
            with open(fn, encoding="utf-8") as f:
                lines = [l.strip() for l in f]
            if not self._det:
                self._rng.shuffle(lines)
            for line in lines:
                yield self._vocab.encode_line(line, self._native)

    @staticmethod
    def _iterate(sentences, batch_size, num_steps):
        streams = [None] * batch_size
        x = np.zeros([batch_size, num_steps], np.int64)
        y = np.zeros([batch_size, num_steps], np.int64)
        w = np.zeros([batch_size, num_steps], np.float32)
        while True:
            x[:], y[:], w[:] = 0, 0, 0
            for i in range(batch_size):
                pos = 0
                while pos < num_steps:
                    if streams[i] is None or len(streams[i]) <= 1:
                        try:
                            streams[i] = next(sentences)
                        except StopIteration:
                            break
                    n = min(len(streams[i]) - 1, num_steps - pos)
                    x[i, pos:pos + n] = streams[i][:n]
                    y[i, pos:pos + n] = streams[i][1:n + 1]
                    w[i, pos:pos + n] = 1
                    streams[i] = streams[i][n:]
                    pos += n
            if not w.any():
                return
            yield x.copy(), y.copy(), w.copy()

    def iterate_once(self, batch_size, num_steps, num_workers=1, worker_id=0):
        files = self.files(num_workers, worker_id)
        if not self._det:
            self._rng.shuffle(files)
        for v in self._iterate(self._sentences(files), batch_size, num_steps):
            yield v

    def iterate_forever(self, batch_size, num_steps, num_workers=1, worker_id=0):
        while True:
            got = False
            for v in self.iterate_once(batch_size, num_steps, num_workers, worker_id):
                got = True
                yield v
            if not got:
                raise ValueError("no data for worker %d of %d in %r" %
                                 (worker_id, num_workers, self._pattern))
