"""Translation-quality metrics for the NMT example.

Parity: `examples/nmt/utils/evaluation_utils.py:32-184` (`evaluate` dispatching
on metric name; sub-word clean-up before scoring), `scripts/bleu.py`
(corpus BLEU-4 with brevity penalty, optional +1 smoothing) and
`scripts/rouge.py` (ROUGE-1/2 n-gram F-scores and ROUGE-L from the longest
common subsequence, averaged over sentences).  Scores are on a 0–100 scale
like the reference's.
"""
import collections
import math
import re


def _clean(sentence, subword_option):
    """strip, undo BPE (``@@ ``) or SentencePiece (``▁``) segmentation"""
    sentence = sentence.strip()
    if subword_option == "bpe":
        sentence = re.sub("@@ ", "", sentence)
        sentence = re.sub("@@$", "", sentence)
    elif subword_option == "spm":
        sentence = "".join(sentence.split()).replace("▁", " ").lstrip()
    return sentence


def _read(path, subword_option):
    with open(path, encoding="utf-8") as f:
        return [_clean(line, subword_option) for line in f]


# --------------------------------------------------------------------- BLEU
def _ngrams(tokens, max_order):
    c = collections.Counter()
    for n in range(1, max_order + 1):
        for i in range(len(tokens) - n + 1):
            c[tuple(tokens[i:i + n])] += 1
    return c


def compute_bleu(reference_corpus, translation_corpus, max_order=4, smooth=False):
    """Corpus BLEU.

    reference_corpus: list (per segment) of lists of reference token lists.
    translation_corpus: list of token lists.
    Returns (bleu, precisions, brevity_penalty, length_ratio,
             translation_length, reference_length)."""
    match = [0] * max_order
    possible = [0] * max_order
    ref_len = hyp_len = 0
    for refs, hyp in zip(reference_corpus, translation_corpus):
        ref_len += min(len(r) for r in refs)
        hyp_len += len(hyp)
        merged = collections.Counter()
        for r in refs:
            merged |= _ngrams(r, max_order)
        hyp_ng = _ngrams(hyp, max_order)
        for ng, cnt in (hyp_ng & merged).items():
            match[len(ng) - 1] += cnt
        for n in range(1, max_order + 1):
            possible[n - 1] += max(len(hyp) - n + 1, 0)
    prec = [0.0] * max_order
    for n in range(max_order):
        if smooth:
            prec[n] = (match[n] + 1.0) / (possible[n] + 1.0)
        elif possible[n] > 0:
            prec[n] = match[n] / possible[n]
    geo = math.exp(sum(math.log(p) for p in prec) / max_order) if min(prec) > 0 else 0.0
    ratio = hyp_len / ref_len if ref_len else 0.0
    bp = 1.0 if ratio > 1.0 else (math.exp(1.0 - 1.0 / ratio) if ratio > 0 else 0.0)
    return geo * bp, prec, bp, ratio, hyp_len, ref_len


def _bleu(ref_file, trans_file, subword_option=None, max_order=4, smooth=False):
    refs = [[r.split()] for r in _read(ref_file, subword_option)]
    hyps = [h.split() for h in _read(trans_file, None)]
    return 100.0 * compute_bleu(refs, hyps, max_order, smooth)[0]


# -------------------------------------------------------------------- ROUGE
def _lcs_len(a, b):
    if not a or not b:
        return 0
    prev = [0] * (len(b) + 1)
    for x in a:
        cur = [0]
        for j, y in enumerate(b):
            cur.append(prev[j] + 1 if x == y else max(prev[j + 1], cur[j]))
        prev = cur
    return prev[-1]


def _f(p, r, beta=1.0):
    if p + r == 0:
        return 0.0
    return (1 + beta ** 2) * p * r / (r + beta ** 2 * p)


def rouge_n(hyp_tokens, ref_tokens, n):
    h = collections.Counter(tuple(hyp_tokens[i:i + n]) for i in range(len(hyp_tokens) - n + 1))
    r = collections.Counter(tuple(ref_tokens[i:i + n]) for i in range(len(ref_tokens) - n + 1))
    overlap = sum((h & r).values())
    p = overlap / max(sum(h.values()), 1)
    rc = overlap / max(sum(r.values()), 1)
    return _f(p, rc), p, rc


def rouge_l(hyp_tokens, ref_tokens):
    """sentence-level ROUGE-L with β = P/R weighting like the reference script
    (`scripts/rouge.py:181-205`: β = P/(R+ε), F = (1+β²)RP/(R+β²P))."""
    lcs = _lcs_len(hyp_tokens, ref_tokens)
    p = lcs / max(len(hyp_tokens), 1)
    r = lcs / max(len(ref_tokens), 1)
    beta = p / (r + 1e-12)
    num, den = (1 + beta ** 2) * r * p, r + beta ** 2 * p
    return num / (den + 1e-12), p, r


def rouge(hypotheses, references):
    """dict with the mean over sentence pairs of rouge_{1,2,l}/{f,p,r}_score."""
    acc = collections.defaultdict(float)
    pairs = [(h.split(), r.split()) for h, r in zip(hypotheses, references)]
    for h, r in pairs:
        for key, val in (("rouge_1", rouge_n(h, r, 1)), ("rouge_2", rouge_n(h, r, 2)),
                         ("rouge_l", rouge_l(h, r))):
            for suffix, v in zip(("f_score", "p_score", "r_score"), val):
                acc["%s/%s" % (key, suffix)] += v
    n = max(len(pairs), 1)
    return {k: v / n for k, v in acc.items()}


def _rouge(ref_file, summarization_file, subword_option=None):
    refs = _read(ref_file, subword_option)
    hyps = _read(summarization_file, None)
    return 100.0 * rouge(hyps, refs)["rouge_l/f_score"]


# ----------------------------------------------------------------- accuracy
def _accuracy(label_file, pred_file):
    """sentence-level exact match (%)"""
    with open(label_file, encoding="utf-8") as fl, open(pred_file, encoding="utf-8") as fp:
        count = match = 0.0
        for label in fl:
            pred = fp.readline()
            count += 1
            match += label.strip() == pred.strip()
    return 100.0 * match / max(count, 1.0)


def _word_accuracy(label_file, pred_file):
    """mean over sentences of position-wise word matches / longer length (%)"""
    with open(label_file, encoding="utf-8") as fl, open(pred_file, encoding="utf-8") as fp:
        total_acc = total = 0.0
        for sentence in fl:
            labels = sentence.strip().split(" ")
            preds = fp.readline().strip().split(" ")
            match = sum(1.0 for a, b in zip(labels, preds) if a == b)
            total_acc += 100.0 * match / max(len(labels), len(preds))
            total += 1
    return total_acc / max(total, 1.0)


def evaluate(ref_file, trans_file, metric, subword_option=None):
    """Score `trans_file` against `ref_file` — `metric` ∈ bleu | rouge | accuracy |
    word_accuracy (case-insensitive)."""
    m = metric.lower()
    if m == "bleu":
        return _bleu(ref_file, trans_file, subword_option)
    if m == "rouge":
        return _rouge(ref_file, trans_file, subword_option)
    if m == "accuracy":
        return _accuracy(ref_file, trans_file)
    if m == "word_accuracy":
        return _word_accuracy(ref_file, trans_file)
    raise ValueError("Unknown metric %s" % metric)
