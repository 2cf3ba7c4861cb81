"""The NMT example as a library (reference `parallax/parallax/examples/nmt/`):
hyper-parameters, vocabulary and iterator utilities, vanilla / attention /
GNMT sequence-to-sequence models, greedy-sampling-beam inference, BLEU / ROUGE /
accuracy scoring and the training loop.  The driver is
`examples/nmt/nmt_distributed_driver.py`."""
from .hparams import (HParams, create_hparams, create_standard_hparams, extend_hparams,
                      maybe_parse_standard_hparams, load_hparams, save_hparams)
from .model import Seq2Seq, create_model, nmt_graph, learning_rate_fn
from . import (attention, evaluation_utils, inference, iterator_utils, misc_utils, train,
               vocab_utils)

__all__ = ["HParams", "create_hparams", "create_standard_hparams", "extend_hparams",
           "maybe_parse_standard_hparams", "load_hparams", "save_hparams", "Seq2Seq",
           "create_model", "nmt_graph", "learning_rate_fn", "attention", "evaluation_utils",
           "inference", "iterator_utils", "misc_utils", "train", "vocab_utils"]
