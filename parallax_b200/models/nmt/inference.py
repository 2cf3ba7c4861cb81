"""Decoding (greedy / sampling / beam search) and file-level inference.

Parity: `examples/nmt/model.py:395-470` (decoder in INFER mode:
`GreedyEmbeddingHelper`, `SampleEmbeddingHelper(softmax_temperature)`,
`BeamSearchDecoder(beam_width, length_penalty_weight)`, maximum iterations =
`tgt_max_len_infer` or 2 × the longest source), `examples/nmt/inference.py:
31-237` (`load_data`, `single_worker_inference`, `multi_worker_inference`:
the input file is split evenly over the workers, every worker writes its
slice, worker 0 concatenates them) and `utils/nmt_utils.py:29-109`
(`decode_and_evaluate`, `get_translation`).

Beam search follows `tf.contrib.seq2seq.BeamSearchDecoder` step for step:
finished beams may only continue with ``</s>`` at zero cost; candidates are
ranked by ``log_prob / ((5+len)/6)^α`` where `len` does not count the closing
``</s>``; the surviving hypotheses are read back through the parent pointers
(`gather_tree`) and come out best first.
"""
import os
import time

import torch

from . import evaluation_utils, misc_utils


def _max_iterations(hp, source_sequence_length):
    if hp.get("tgt_max_len_infer"):
        return int(hp.tgt_max_len_infer)
    return int(round(float(source_sequence_length.max()) * 2.0))


@torch.no_grad()
def greedy_decode(model, source, source_sequence_length, sos_id, eos_id, max_iterations,
                  sampling_temperature=0.0, generator=None):
    """→ sample ids [B, T'] (``</s>``-filled after the end), lengths [B]"""
    memory, state = model.encode(source, source_sequence_length)
    B, dev = source.shape[0], source.device
    tok = torch.full((B,), sos_id, dtype=torch.int64, device=dev)
    finished = torch.zeros(B, dtype=torch.bool, device=dev)
    lengths = torch.zeros(B, dtype=torch.int64, device=dev)
    out = []
    for _ in range(max_iterations):
        logits, state = model.decode_step(tok, state, memory)
        if sampling_temperature > 0.0:
            probs = torch.softmax(logits / sampling_temperature, -1)
            tok = torch.multinomial(probs, 1, generator=generator)[:, 0]
        else:
            tok = logits.argmax(-1)
        tok = torch.where(finished, torch.full_like(tok, eos_id), tok)
        out.append(tok)
        lengths += (~finished).to(torch.int64)
        finished = finished | (tok == eos_id)
        if bool(finished.all()):
            break
    ids = torch.stack(out, 1) if out else torch.zeros(B, 0, dtype=torch.int64, device=dev)
    return ids, lengths


def _length_penalty(lengths, weight):
    if weight == 0.0:
        return torch.ones_like(lengths, dtype=torch.float32)
    return ((5.0 + lengths.to(torch.float32)) / 6.0) ** weight


@torch.no_grad()
def beam_search_decode(model, source, source_sequence_length, sos_id, eos_id, max_iterations,
                       beam_width, length_penalty_weight=0.0):
    """→ predicted ids [B, W, T'] best beam first, scores [B, W], lengths [B, W]"""
    B, W, dev = source.shape[0], int(beam_width), source.device
    memory, state = model.encode(source, source_sequence_length)
    tile = torch.arange(B, device=dev).repeat_interleave(W)
    state = model.decoder.reorder_state(state, tile)
    if memory is not None:
        memory = tuple(m.index_select(0, tile) for m in memory)
    NEG = float("-inf")
    log_probs = torch.full((B, W), NEG, device=dev)
    log_probs[:, 0] = 0.0
    finished = torch.zeros(B, W, dtype=torch.bool, device=dev)
    lengths = torch.zeros(B, W, dtype=torch.int64, device=dev)
    tok = torch.full((B * W,), sos_id, dtype=torch.int64, device=dev)
    step_ids, step_parents = [], []
    base = (torch.arange(B, device=dev) * W)[:, None]
    scores = log_probs.clone()
    for _ in range(max_iterations):
        logits, state = model.decode_step(tok, state, memory)
        V = logits.shape[-1]
        step_lp = torch.log_softmax(logits, -1).view(B, W, V)
        # finished beams: only </s>, at no cost
        fin_row = torch.full((V,), NEG, device=dev)
        fin_row[eos_id] = 0.0
        step_lp = torch.where(finished[..., None], fin_row, step_lp)
        total = log_probs[..., None] + step_lp                          # [B,W,V]
        add = torch.ones(V, dtype=torch.int64, device=dev)
        add[eos_id] = 0
        cand_len = lengths[..., None] + add * (~finished)[..., None].to(torch.int64)
        cand_scores = total / _length_penalty(cand_len, length_penalty_weight)
        scores, idx = cand_scores.view(B, W * V).topk(W, dim=-1)
        parent, word = idx // V, idx % V
        log_probs = total.view(B, W * V).gather(1, idx)
        prev_fin = finished.gather(1, parent)
        lengths = lengths.gather(1, parent) + (~prev_fin).to(torch.int64)
        finished = prev_fin | (word == eos_id)
        step_ids.append(word)
        step_parents.append(parent)
        state = model.decoder.reorder_state(state, (base + parent).view(-1))
        tok = word.view(-1)
        if bool(finished.all()):
            break
    # gather_tree: follow parent pointers from the last step backwards
    T = len(step_ids)
    out = torch.full((B, W, T), eos_id, dtype=torch.int64, device=dev)
    beam = torch.arange(W, device=dev)[None, :].expand(B, W)
    for t in range(T - 1, -1, -1):
        out[:, :, t] = step_ids[t].gather(1, beam)
        beam = step_parents[t].gather(1, beam)
    # everything after a hypothesis' first </s> is </s>
    after = (out == eos_id).to(torch.int64).cumsum(-1) > 0
    first = after & ~torch.cat([torch.zeros_like(after[..., :1]), after[..., :-1]], -1)
    out = torch.where(after & ~first, torch.full_like(out, eos_id), out)
    return out, scores, lengths


@torch.no_grad()
def infer_batch(model, hp, source, source_sequence_length, sos_id, eos_id, generator=None):
    """`model.decode`-equivalent: [N, B, T'] with N = number of hypotheses kept
    (`num_translations_per_input`; beams for beam search, independent samples
    otherwise)."""
    was_training = model.training
    model.eval()
    try:
        max_it = _max_iterations(hp, source_sequence_length)
        n = max(int(hp.get("num_translations_per_input", 1) or 1), 1)
        if hp.beam_width and hp.beam_width > 0:
            ids, _, _ = beam_search_decode(model, source, source_sequence_length, sos_id,
                                           eos_id, max_it, hp.beam_width,
                                           float(hp.length_penalty_weight or 0.0))
            return ids.transpose(0, 1)[:min(n, hp.beam_width)]
        outs = []
        for _ in range(n if hp.sampling_temperature > 0.0 else 1):
            ids, _ = greedy_decode(model, source, source_sequence_length, sos_id, eos_id,
                                   max_it, float(hp.sampling_temperature or 0.0), generator)
            outs.append(ids)
        T = max(o.shape[1] for o in outs)
        outs = [torch.cat([o, torch.full((o.shape[0], T - o.shape[1]), eos_id,
                                         dtype=o.dtype, device=o.device)], 1) for o in outs]
        return torch.stack(outs, 0)
    finally:
        model.train(was_training)


def get_translation(ids, tgt_vocab, tgt_eos, subword_option=""):
    """one hypothesis (1-D ids) → text: cut at the first ``</s>``, undo sub-words"""
    words = tgt_vocab.decode(ids.tolist() if torch.is_tensor(ids) else ids)
    if tgt_eos and tgt_eos in words:
        words = words[:words.index(tgt_eos)]
    if subword_option == "bpe":
        return misc_utils.format_bpe_text(words)
    if subword_option == "spm":
        return misc_utils.format_spm_text(words)
    return misc_utils.format_text(words)


def load_data(inference_input_file, hparams=None):
    """lines of the input file; `hparams.inference_indices` selects a subset"""
    with open(inference_input_file, encoding="utf-8") as f:
        data = f.read().splitlines()
    if hparams is not None and hparams.get("inference_indices"):
        data = [data[i] for i in hparams.inference_indices]
    return data


def decode_to_file(model, hp, infer_data, src_vocab, tgt_vocab, trans_file, device=None):
    """Translate `infer_data` (list of source lines) into `trans_file`; with
    `num_translations_per_input` > 1 the hypotheses of a sentence are written on
    consecutive lines."""
    from .iterator_utils import get_infer_iterator
    device = device or next(model.parameters()).device
    sos_id, eos_id = tgt_vocab.lookup(hp.sos), tgt_vocab.lookup(hp.eos)
    it = get_infer_iterator(infer_data, src_vocab, hp.infer_batch_size, hp.eos,
                            hp.get("src_max_len_infer") or None)
    start, n = time.time(), 0
    os.makedirs(os.path.dirname(os.path.abspath(trans_file)), exist_ok=True)
    with open(trans_file, "w", encoding="utf-8") as f:
        for batch in it:
            ids = infer_batch(model, hp, batch.source.to(device),
                              batch.source_sequence_length.to(device), sos_id, eos_id)
            ids = ids.cpu()
            for b in range(ids.shape[1]):
                for k in range(ids.shape[0]):
                    f.write(get_translation(ids[k, b], tgt_vocab, hp.eos,
                                            hp.subword_option) + "\n")
                n += 1
    return n, time.time() - start


def decode_and_evaluate(name, model, hp, infer_data, src_vocab, tgt_vocab, trans_file,
                        ref_file=None, metrics=None, device=None, decode=True):
    """`nmt_utils.decode_and_evaluate`: write translations, then score every
    metric against `ref_file` → dict metric → score."""
    if decode:
        n, secs = decode_to_file(model, hp, infer_data, src_vocab, tgt_vocab, trans_file, device)
        misc_utils.print_out("  done, num sentences %d, time %ds" % (n, secs))
    scores = {}
    if ref_file and os.path.exists(trans_file):
        for metric in (metrics if metrics is not None else hp.metrics):
            scores[metric] = evaluation_utils.evaluate(ref_file, trans_file, metric,
                                                       subword_option=hp.subword_option)
            misc_utils.print_out("  %s %s: %.1f" % (metric, name, scores[metric]))
    return scores


def single_worker_inference(model, hp, inference_input_file, inference_output_file,
                            src_vocab, tgt_vocab, device=None):
    data = load_data(inference_input_file, hp)
    return decode_to_file(model, hp, data, src_vocab, tgt_vocab, inference_output_file, device)


def worker_slice(num_lines, num_workers, jobid):
    """[start, end) of worker `jobid` (`inference.py:192-195`: equal slices of
    ⌊(n-1)/workers⌋+1 lines)."""
    per = (num_lines - 1) // num_workers + 1
    return jobid * per, min((jobid + 1) * per, num_lines)


def multi_worker_inference(model, hp, inference_input_file, inference_output_file,
                           src_vocab, tgt_vocab, num_workers, jobid, device=None,
                           wait_secs=0.2, timeout=3600.0):
    """Every worker translates its slice into ``<output>_<jobid>`` and marks it
    done; worker 0 then concatenates the parts in order into `<output>`."""
    assert num_workers > 1
    data = load_data(inference_input_file, hp)
    s, e = worker_slice(len(data), num_workers, jobid)
    part = "%s_%d" % (inference_output_file, jobid)
    decode_to_file(model, hp, data[s:e], src_vocab, tgt_vocab, part + ".tmp", device)
    os.replace(part + ".tmp", part)
    with open(part + "_done", "w") as f:
        f.write("%d" % (e - s))
    if jobid != 0:
        return None
    t0 = time.time()
    with open(inference_output_file, "w", encoding="utf-8") as out:
        for w in range(num_workers):
            pw = "%s_%d" % (inference_output_file, w)
            while not os.path.exists(pw + "_done"):
                if time.time() - t0 > timeout:
                    raise RuntimeError("worker %d did not finish its inference slice" % w)
                time.sleep(wait_secs)
            with open(pw, encoding="utf-8") as f:
                out.write(f.read())
            os.remove(pw)
            os.remove(pw + "_done")
    return inference_output_file
