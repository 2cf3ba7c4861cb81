"""Sentence → skip-thought vector, with one or several trained models.

Parity: `examples/skip_thoughts/skip_thoughts_encoder.py:51-271`
(`SkipThoughtsEncoder`: tokenise, map words to an (expanded) embedding matrix —
unknown words use ``<unk>`` —, sort by length, batch, pad, run the encoder in
"encode" mode feeding *embeddings* not ids, optional L2 normalisation) and
`examples/skip_thoughts/encoder_manager.py:44-139` (`EncoderManager`: load
several models, e.g. a uni- and a bi-directional one, and concatenate their
vectors).
"""
import re

import numpy as np
import torch

from . import special_words
from .model import SkipThoughtsModel

_TOKEN = re.compile(r"\w+|[^\w\s]", re.UNICODE)


def tokenize(sentence):
    """word / punctuation tokens (the reference uses NLTK's punkt + word tokenizer,
    which is not available offline)"""
    return _TOKEN.findall(sentence)


def _batch_and_pad(sequences):
    """list of [len_i, D] arrays → (padded [N, max_len, D], mask [N, max_len])"""
    n, width, dim = len(sequences), max(len(s) for s in sequences), sequences[0].shape[1]
    emb = np.zeros((n, width, dim), dtype=np.float32)
    mask = np.zeros((n, width), dtype=np.int8)
    for i, s in enumerate(sequences):
        emb[i, :len(s)] = s
        mask[i, :len(s)] = 1
    return emb, mask


class SkipThoughtsEncoder(object):
    """`embeddings`: dict word → vector (np.ndarray [D]); must contain ``<unk>``."""

    def __init__(self, embeddings, model=None):
        self._embeddings = embeddings
        self.model = model

    def build_from_config(self, model_config, state=None, device="cpu"):
        """instantiate the model in encode mode; `state` is either a plain
        ``state_dict`` of `SkipThoughtsModel` or a Parallax checkpoint dict"""
        self.model = SkipThoughtsModel(model_config).to(device).eval()
        if state is not None:
            load_weights(self.model, state)
        return self

    def _word_to_embedding(self, w):
        return self._embeddings.get(w, self._embeddings[special_words.UNK])

    def _preprocess(self, data, use_eos):
        out = []
        for item in data:
            toks = tokenize(item)
            if use_eos:
                toks.append(special_words.EOS)
            out.append(np.stack([self._word_to_embedding(w) for w in toks]) if toks else
                       self._word_to_embedding(special_words.EOS)[None])
        return out

    @torch.no_grad()
    def encode(self, data, use_norm=True, verbose=False, batch_size=128, use_eos=False):
        data = self._preprocess(data, use_eos)
        order = np.argsort([len(d) for d in data], kind="stable")
        dev = next(self.model.parameters()).device
        dt = self.model.compute_dtype
        thought = [None] * len(data)
        for start in range(0, len(data), batch_size):
            idx = order[start:start + batch_size]
            emb, mask = _batch_and_pad([data[i] for i in idx])
            vec = self.model.encode_embeddings(torch.from_numpy(emb).to(dev, dt),
                                               torch.from_numpy(mask).to(dev)).float().cpu().numpy()
            for i, v in zip(idx, vec):
                thought[i] = v
        thought = np.stack(thought)
        if use_norm:
            thought = thought / np.maximum(np.linalg.norm(thought, axis=1, keepdims=True), 1e-12)
        return thought


def load_weights(model, state):
    """plain ``state_dict`` or Parallax checkpoint (`engine.state_dict()` layout)"""
    if "dense" in state and "sparse" in state:
        with torch.no_grad():
            params = dict(model.named_parameters())
            for n, v in state["dense"]["master"].items():
                if n in params:
                    params[n].copy_(v.view_as(params[n]))
            for n, t in state["sparse"].items():
                if n in params:
                    params[n].copy_(t["weight"])
    else:
        model.load_state_dict(state)
    return model


def embeddings_from_model(model_or_state, vocab):
    """dict word → trained embedding row for a vocabulary list"""
    if isinstance(model_or_state, dict):
        st = model_or_state
        w = st["sparse"]["word_embedding.weight"]["weight"] if "sparse" in st \
            else st["word_embedding.weight"]
    else:
        w = model_or_state.word_embedding.weight
    w = w.detach().float().cpu().numpy()
    return {word: w[i] for i, word in enumerate(vocab)}


class EncoderManager(object):
    def __init__(self):
        self.encoders = []

    def load_model(self, model_config, vocabulary, embedding_matrix, state=None, device="cpu"):
        """`vocabulary`: list of words (or a file with one per line);
        `embedding_matrix`: np.ndarray [len(vocabulary), D] (or a ``.npy`` file) —
        typically the output of `vocabulary_expansion.expand_vocabulary`."""
        if isinstance(vocabulary, str):
            with open(vocabulary, encoding="utf-8") as f:
                vocabulary = [line.strip() for line in f]
        if isinstance(embedding_matrix, str):
            embedding_matrix = np.load(embedding_matrix)
        assert len(vocabulary) == embedding_matrix.shape[0]
        emb = dict(zip(vocabulary, embedding_matrix))
        enc = SkipThoughtsEncoder(emb).build_from_config(model_config, state, device)
        self.encoders.append(enc)
        return enc

    def encode(self, data, use_norm=True, verbose=False, batch_size=128, use_eos=False):
        if not self.encoders:
            raise ValueError("Must call load_model at least once before calling encode.")
        return np.concatenate([e.encode(data, use_norm, verbose, batch_size, use_eos)
                               for e in self.encoders], axis=1)

    def close(self):
        self.encoders = []
