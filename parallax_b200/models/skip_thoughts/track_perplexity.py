"""Validation perplexity of checkpoints.

Parity: `examples/skip_thoughts/track_perplexity.py:57-202`: for the newest
checkpoint (if it is new and past `min_global_step`) run `num_eval_examples /
batch_size` validation batches, perplexity = exp(Σ loss·weight / Σ weight);
`run` polls the checkpoint directory every `eval_interval_secs`.
"""
import math
import os
import time

import numpy as np
import torch

from ... import checkpoint as _ckpt
from ...log import parallax_log as log
from .encoder import load_weights
from .input_ops import prefetch_input_data
from .model import SkipThoughtsModel


@torch.no_grad()
def evaluate_model(model, batches, num_batches):
    """→ perplexity over `num_batches` batches"""
    model.eval()
    dev = next(model.parameters()).device
    sum_losses = sum_weights = 0.0
    for i, (enc, pre, post) in enumerate(batches):
        if i >= num_batches:
            break
        out = model(enc.ids.to(dev), enc.mask.to(dev), pre.ids.to(dev), pre.mask.to(dev),
                    post.ids.to(dev), post.mask.to(dev))
        sum_losses += float(out["loss"])
        sum_weights += float(out["sum_weights"])
    return math.exp(sum_losses / max(sum_weights, 1.0))


def run_once(model_config, checkpoint_dir, num_eval_examples=50000, min_global_step=100,
             last_step=None, device="cpu"):
    """→ (global_step, perplexity) or None when there is nothing new to evaluate"""
    path = _ckpt.latest_checkpoint(checkpoint_dir)
    if path is None:
        log.info("Skipping evaluation. No checkpoint found in: %s", checkpoint_dir)
        return None
    state = _ckpt.load_logical(path)          # single file or the NVLink fabric's sharded directory
    step = int(state["global_step"])
    if step < min_global_step or step == last_step:
        log.info("Skipping evaluation. Global step = %d (min %d, last %s)", step,
                 min_global_step, last_step)
        return None
    model = load_weights(SkipThoughtsModel(model_config).to(device), state)
    batches = prefetch_input_data(model_config.input_file_pattern, model_config.batch_size,
                                  shuffle=False, num_shards=1, shard_id=0, epochs=1)
    nb = int(np.ceil(num_eval_examples / float(model_config.batch_size)))
    ppl = evaluate_model(model, batches, nb)
    log.info("Perplexity = %f (global step %d)", ppl, step)
    return step, ppl


def run(model_config, checkpoint_dir, eval_dir=None, eval_interval_secs=600, max_evals=None,
        **kw):
    """poll forever (or `max_evals` times); appends ``step\\tppl`` to eval_dir/perplexity.tsv"""
    last, done = None, 0
    while max_evals is None or done < max_evals:
        start = time.time()
        res = run_once(model_config, checkpoint_dir, last_step=last, **kw)
        if res is not None:
            last = res[0]
            if eval_dir:
                os.makedirs(eval_dir, exist_ok=True)
                with open(os.path.join(eval_dir, "perplexity.tsv"), "a") as f:
                    f.write("%d\t%f\n" % res)
        done += 1
        wait = start + eval_interval_secs - time.time()
        if wait > 0 and (max_evals is None or done < max_evals):
            time.sleep(wait)
    return last
