"""Small helpers of the NMT example (`examples/nmt/utils/misc_utils.py:33-183`):
overflow-safe exp, timing/log helpers, text post-processing for BPE and
SentencePiece sub-word output."""
import math
import sys
import time

SPM_SPACE = "▁"


def safe_exp(value):
    """exp that saturates to inf instead of raising OverflowError."""
    try:
        return math.exp(value)
    except OverflowError:
        return float("inf")


def print_time(s, start_time, out=None):
    msg = "%s, time %ds, %s." % (s, time.time() - start_time, time.ctime())
    print_out(msg, out)
    return time.time()


def print_out(s, f=None, new_line=True):
    if isinstance(s, bytes):
        s = s.decode("utf-8")
    if f is not None:
        f.write(s + ("\n" if new_line else ""))
    sys.stdout.write(s + ("\n" if new_line else ""))
    sys.stdout.flush()


def format_text(words):
    """list of tokens (str or bytes) → sentence."""
    if not hasattr(words, "__len__") or isinstance(words, (str, bytes)):
        words = [words]
    return " ".join(w.decode("utf-8") if isinstance(w, bytes) else str(w) for w in words)


def format_bpe_text(symbols, delimiter="@@"):
    """Merge BPE pieces: a piece ending in `delimiter` continues the word."""
    words, word = [], ""
    if isinstance(symbols, str):
        symbols = symbols.split()
    dl = len(delimiter)
    for s in symbols:
        if isinstance(s, bytes):
            s = s.decode("utf-8")
        if len(s) >= dl and s[-dl:] == delimiter:
            word += s[:-dl]
        else:
            words.append(word + s)
            word = ""
    return " ".join(words)


def format_spm_text(symbols):
    """Merge SentencePiece pieces: ``▁`` marks a word start."""
    if isinstance(symbols, str):
        symbols = symbols.split()
    joined = "".join(s.decode("utf-8") if isinstance(s, bytes) else s for s in symbols)
    return joined.replace(SPM_SPACE, " ").strip()


class Stats(object):
    """Running training statistics of one logging window
    (`examples/nmt/train.py:208-263` `init_stats`/`update_stats`/`process_stats`)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.step_time = self.loss = self.predict_count = 0.0
        self.total_count = self.grad_norm = 0.0
        self.steps = 0

    def update(self, step_time, loss, predict_count, word_count, batch_size, grad_norm=0.0):
        self.step_time += step_time
        self.loss += loss * batch_size
        self.predict_count += predict_count
        self.total_count += word_count
        self.grad_norm += grad_norm
        self.steps += 1

    def process(self):
        """→ dict(avg_step_time, avg_grad_norm, train_ppl, speed[k words/s], overflow)"""
        n = max(self.steps, 1)
        ppl = safe_exp(self.loss / max(self.predict_count, 1.0))
        return {"avg_step_time": self.step_time / n, "avg_grad_norm": self.grad_norm / n,
                "train_ppl": ppl,
                "speed": self.total_count / (1000.0 * max(self.step_time, 1e-9)),
                "overflow": (not math.isfinite(ppl)) or ppl > 1e20}
