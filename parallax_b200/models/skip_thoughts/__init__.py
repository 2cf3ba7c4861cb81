"""The skip-thoughts example as a library (reference
`parallax/parallax/examples/skip_thoughts/`): configuration, layer-normalised
GRU, sharded input pipeline, model, corpus preprocessing, sentence encoder /
encoder manager, vocabulary expansion, validation-perplexity tracking, downstream evaluation
(MR / CR / SUBJ / MPQA / TREC / MSRP / SICK).  The
driver is `examples/skip_thoughts/skip_distributed_driver.py`."""
from . import (configuration, encoder, evaluate, gru_cell, input_ops, preprocess_dataset,
               special_words, track_perplexity, vocabulary_expansion)
from .configuration import model_config, training_config
from .encoder import EncoderManager, SkipThoughtsEncoder
from .model import SkipThoughtsModel, feed_from_batch, skip_thoughts_graph

__all__ = ["configuration", "encoder", "evaluate", "gru_cell", "input_ops", "preprocess_dataset",
           "special_words", "track_perplexity", "vocabulary_expansion", "model_config",
           "training_config", "EncoderManager", "SkipThoughtsEncoder", "SkipThoughtsModel",
           "feed_from_batch", "skip_thoughts_graph"]
