"""LM1B input pipeline: vocabulary + sharded sentence stream packed into
[batch, num_steps] windows (the role of reference `examples/lm1b/data_utils.py:66-148`:
files are split over workers, sentences are concatenated with <S> boundaries
and cut into windows; `w` masks padding)."""
import glob
import random

import numpy as np


class Vocabulary(object):
    def __init__(self):
        self._token_to_id, self._id_to_token = {}, []
        self._s_id = self._unk_id = None

    def add(self, token):
        self._token_to_id[token] = len(self._id_to_token)
        self._id_to_token.append(token)

    @property
    def num_tokens(self):
        return len(self._id_to_token)

    @property
    def s_id(self):
        return self._s_id

    def get_id(self, token):
        return self._token_to_id.get(token, self._unk_id)

    def get_token(self, id_):
        return self._id_to_token[id_]

    @staticmethod
    def from_file(filename):
        v = Vocabulary()
        with open(filename, encoding="utf-8") as f:
            for line in f:
                tok = line.split()[0] if line.strip() else None
                if tok is not None:
                    v.add(tok)
        for special in ("<S>", "<UNK>"):
            if special not in v._token_to_id:
                v.add(special)
        v._s_id, v._unk_id = v._token_to_id["<S>"], v._token_to_id["<UNK>"]
        return v


class Dataset(object):
    def __init__(self, vocab, file_pattern, deterministic=False):
        self._vocab, self._pattern, self._det = vocab, file_pattern, deterministic

    def _sentences(self, files):
        for fn in files:
            with open(fn, encoding="utf-8") as f:
                lines = [l.strip() for l in f]
            if not self._det:
                random.shuffle(lines)
            for line in lines:
                ids = [self._vocab.get_id(w) for w in line.split()]
                yield [self._vocab.s_id] + ids + [self._vocab.s_id]

    def _iterate(self, sentences, batch_size, num_steps):
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
        files = sorted(glob.glob(self._pattern))[worker_id::num_workers]
        if not self._det:
            random.shuffle(files)
        for v in self._iterate(self._sentences(files), batch_size, num_steps):
            yield v

    def iterate_forever(self, batch_size, num_steps, num_workers=1, worker_id=0):
        while True:
            for v in self.iterate_once(batch_size, num_steps, num_workers, worker_id):
                yield v
