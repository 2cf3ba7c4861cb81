"""Vocabulary expansion: map a large word2vec space into the model's embedding
space.

Parity: `examples/skip_thoughts/vocabulary_expansion.py:71-209`: take the words
the skip-thoughts vocabulary shares with a word2vec model, fit a linear
regression word2vec → skip-thoughts embedding on them, then embed *every*
word2vec word with it (words the model was trained on keep their trained
vectors).  The reference uses sklearn + gensim; here the regression is a
least-squares solve and the "word2vec model" is any dict word → vector.
"""
import collections
import os

import numpy as np


def fit_linear_map(x, y):
    """least-squares W, b with  y ≈ x·W + b"""
    xa = np.concatenate([x, np.ones((x.shape[0], 1), x.dtype)], 1)
    sol, *_ = np.linalg.lstsq(xa.astype(np.float64), y.astype(np.float64), rcond=None)
    return sol[:-1].astype(np.float32), sol[-1].astype(np.float32)


def expand_vocabulary(skip_thoughts_emb, skip_thoughts_vocab, word2vec):
    """skip_thoughts_emb [V,D]; skip_thoughts_vocab: OrderedDict word → id;
    word2vec: dict word → vector.  → OrderedDict word → embedding [D]."""
    shared = [w for w in word2vec if w in skip_thoughts_vocab]
    if len(shared) < 2:
        raise ValueError("need at least two shared words to fit the mapping")
    x = np.stack([np.asarray(word2vec[w], np.float32) for w in shared])
    y = np.stack([skip_thoughts_emb[skip_thoughts_vocab[w]] for w in shared])
    w_map, b_map = fit_linear_map(x, y)
    combined = collections.OrderedDict()
    for w, v in word2vec.items():
        if "_" in w:                      # word2vec phrases are skipped
            continue
        combined[w] = np.asarray(v, np.float32) @ w_map + b_map
    for w, i in skip_thoughts_vocab.items():   # trained words keep their vectors
        combined[w] = skip_thoughts_emb[i]
    return combined


def save_expanded(combined, output_dir):
    os.makedirs(output_dir, exist_ok=True)
    with open(os.path.join(output_dir, "vocab.txt"), "w", encoding="utf-8") as f:
        f.write("\n".join(combined.keys()))
    np.save(os.path.join(output_dir, "embeddings.npy"), np.stack(list(combined.values())))
    return os.path.join(output_dir, "vocab.txt"), os.path.join(output_dir, "embeddings.npy")


def load_word2vec_text(path, limit=None):
    """``word v1 v2 …`` per line (optional ``count dim`` header)"""
    out = collections.OrderedDict()
    with open(path, encoding="utf-8") as f:
        for i, line in enumerate(f):
            parts = line.rstrip().split(" ")
            if i == 0 and len(parts) == 2:
                continue
            out[parts[0]] = np.asarray([float(v) for v in parts[1:]], np.float32)
            if limit and len(out) >= limit:
                break
    return out
