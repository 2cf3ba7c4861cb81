"""Vocabulary handling for the NMT example.

Parity: `examples/nmt/utils/vocab_utils.py:29-118` — `load_vocab`,
`check_vocab` (the first three entries must be ``<unk>``, ``<s>``, ``</s>``;
otherwise a corrected copy is written to `out_dir`), `create_vocab_tables`
(string→id lookup tables with ``<unk>`` = id 0 as default; one shared table
under `share_vocab`) and `load_embed_txt` (Glove / word2vec text format).
"""
import os

import torch

from .hparams import UNK, SOS, EOS

UNK_ID = 0


def load_vocab(vocab_file):
    with open(vocab_file, encoding="utf-8") as f:
        vocab = [line.rstrip("\n") for line in f]
    # a trailing empty line is not a token
    while vocab and vocab[-1] == "":
        vocab.pop()
    return vocab, len(vocab)


def check_vocab(vocab_file, out_dir, check_special_token=True, sos=None, eos=None, unk=None):
    """Returns (vocab_size, vocab_file) — `vocab_file` is the possibly rewritten
    file that starts with the three special tokens."""
    if not os.path.exists(vocab_file):
        raise ValueError("vocab_file '%s' does not exist." % vocab_file)
    vocab, size = load_vocab(vocab_file)
    if check_special_token:
        unk, sos, eos = unk or UNK, sos or SOS, eos or EOS
        assert size >= 3, "vocabulary needs at least the three special tokens"
        if vocab[:3] != [unk, sos, eos]:
            vocab = [unk, sos, eos] + [w for w in vocab if w not in (unk, sos, eos)]
            size = len(vocab)
            os.makedirs(out_dir, exist_ok=True)
            vocab_file = os.path.join(out_dir, os.path.basename(vocab_file))
            with open(vocab_file, "w", encoding="utf-8") as f:
                for w in vocab:
                    f.write(w + "\n")
    return size, vocab_file


class VocabTable(object):
    """token ↔ id with a default id for out-of-vocabulary tokens
    (`lookup_ops.index_table_from_file(default_value=UNK_ID)` and its reverse
    `index_to_string_table_from_file(default_value=UNK)`)."""

    def __init__(self, tokens, default_id=UNK_ID, default_token=UNK):
        self.tokens = list(tokens)
        self.index = {}
        for i, w in enumerate(self.tokens):
            self.index.setdefault(w, i)
        self.default_id, self.default_token = default_id, default_token

    @classmethod
    def from_file(cls, vocab_file, **kw):
        return cls(load_vocab(vocab_file)[0], **kw)

    def __len__(self):
        return len(self.tokens)

    def lookup(self, word):
        return self.index.get(word, self.default_id)

    def encode(self, words):
        get, d = self.index.get, self.default_id
        return [get(w, d) for w in words]

    def decode(self, ids):
        n, t, d = len(self.tokens), self.tokens, self.default_token
        return [t[i] if 0 <= i < n else d for i in (int(x) for x in ids)]


def create_vocab_tables(src_vocab_file, tgt_vocab_file, share_vocab):
    src = VocabTable.from_file(src_vocab_file)
    tgt = src if share_vocab else VocabTable.from_file(tgt_vocab_file)
    return src, tgt


def load_embed_txt(embed_file):
    """``word v1 v2 …`` per line (an optional ``count dim`` header line is
    skipped) → (dict word → list[float], dim)."""
    emb, dim = {}, None
    with open(embed_file, encoding="utf-8") as f:
        for i, line in enumerate(f):
            parts = line.rstrip().split(" ")
            if i == 0 and len(parts) == 2:
                continue
            vec = [float(x) for x in parts[1:]]
            if dim is None:
                dim = len(vec)
            assert len(vec) == dim, "All embedding size should be same."
            emb[parts[0]] = vec
    return emb, dim


def pretrained_embedding_matrix(vocab_file, embed_file, num_trainable_tokens=3):
    """Embedding matrix for `vocab_file` initialised from `embed_file`
    (`examples/nmt/model_helper.py:248-281`): tokens missing from the file get
    zero rows; the first `num_trainable_tokens` (special) rows are not taken from
    the file (`Seq2Seq.load_pretrained_embeddings`)."""
    vocab, _ = load_vocab(vocab_file)
    emb, dim = load_embed_txt(embed_file)
    mat = torch.zeros(len(vocab), dim)
    for i, w in enumerate(vocab):
        if w in emb:
            mat[i] = torch.tensor(emb[w])
    return mat, num_trainable_tokens
