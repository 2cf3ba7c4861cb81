"""Synthetic parallel corpus for the NMT tests: the "translation" of a sentence
is its reversal with every word ``w<i>`` renamed ``t<i>`` — learnable by a tiny
attention model in seconds (the reference tests train on a 100-sentence
IWSLT15 excerpt, `examples/nmt/nmt_test.py:48-100`; no corpus is shipped here)."""
import os
import random


def make_fixture(d, n_train=2000, n_dev=40, vocab=16, min_len=3, max_len=6, seed=0,
                 special_first=True):
    os.makedirs(d, exist_ok=True)
    rng = random.Random(seed)

    def sent():
        return [rng.randrange(vocab) for _ in range(rng.randint(min_len, max_len))]
    for name, n in (("train", n_train), ("dev", n_dev), ("test", n_dev)):
        with open(os.path.join(d, name + ".src"), "w") as fs, \
                open(os.path.join(d, name + ".tgt"), "w") as ft:
            for _ in range(n):
                s = sent()
                fs.write(" ".join("w%d" % i for i in s) + "\n")
                ft.write(" ".join("t%d" % i for i in reversed(s)) + "\n")
    special = ["<unk>", "<s>", "</s>"]
    for lang, pre in (("src", "w"), ("tgt", "t")):
        words = [pre + str(i) for i in range(vocab)]
        toks = special + words if special_first else words + special
        with open(os.path.join(d, "vocab." + lang), "w") as f:
            f.write("\n".join(toks) + "\n")
    return d
