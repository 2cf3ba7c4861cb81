"""Downstream evaluation of skip-thought vectors.

Parity: `examples/skip_thoughts/evaluate.py:30-122`, which feeds an
`EncoderManager` to the evaluation suite of the original skip-thoughts release
(`skipthoughts.eval_classification / eval_trec / eval_msrp / eval_sick`).  Those
modules are not vendored by the reference; the protocols are reproduced here on
top of scikit-learn:

* MR / CR / SUBJ / MPQA — binary sentence classification, logistic regression with
  the regularisation strength chosen by inner k-fold CV inside an outer 10-fold CV
  (`eval_nested_kfold`);
* TREC — 6-way question classification, C chosen by k-fold CV on the training set,
  accuracy on the test set;
* MSRP — paraphrase detection on pair features ``[|u − v|, u ⊙ v]``: accuracy + F1;
* SICK — semantic relatedness: the score distribution over 1…5 is regressed from the
  pair features, Pearson / Spearman / MSE of the expected score.

`encoder` is anything with ``encode(list_of_sentences, …) → np.ndarray [N, D]``.
"""
import argparse
import os

import numpy as np

TASK_FILES = {
    "MR": ("rt-polarity.pos", "rt-polarity.neg"),
    "CR": ("custrev.pos", "custrev.neg"),
    "SUBJ": ("plot.tok.gt9.5000", "quote.tok.gt9.5000"),
    "MPQA": ("mpqa.pos", "mpqa.neg"),
}


def _lr(C, seed=1234):
    from sklearn.linear_model import LogisticRegression
    return LogisticRegression(C=C, max_iter=1000, random_state=seed)


def _read_lines(path):
    with open(path, encoding="latin-1") as f:
        return [line.strip() for line in f if line.strip()]


# ------------------------------------------------------------ classification
def load_binary_task(name, loc):
    pos_f, neg_f = TASK_FILES[name]
    pos, neg = _read_lines(os.path.join(loc, pos_f)), _read_lines(os.path.join(loc, neg_f))
    return pos + neg, np.array([1] * len(pos) + [0] * len(neg))


def _choose_c(x, y, k, scan, seed):
    from sklearn.model_selection import StratifiedKFold
    scores = []
    for C in scan:
        accs = []
        for tr, te in StratifiedKFold(k, shuffle=True, random_state=seed).split(x, y):
            accs.append(_lr(C, seed).fit(x[tr], y[tr]).score(x[te], y[te]))
        scores.append(np.mean(accs))
    return scan[int(np.argmax(scores))]


def eval_nested_kfold(encoder, name, loc, k=10, inner_k=None, scan=None, seed=1234,
                      features=None):
    """→ mean outer-fold accuracy.  `features` skips the encoder (tests)."""
    from sklearn.model_selection import StratifiedKFold
    text, labels = load_binary_task(name, loc)
    x = np.asarray(features if features is not None else encoder.encode(text, verbose=False))
    scan = scan or [2 ** t for t in range(0, 9)]
    accs = []
    for tr, te in StratifiedKFold(k, shuffle=True, random_state=seed).split(x, labels):
        C = _choose_c(x[tr], labels[tr], inner_k or k, scan, seed)
        accs.append(_lr(C, seed).fit(x[tr], labels[tr]).score(x[te], labels[te]))
    return float(np.mean(accs))


def load_trec(loc):
    """``LABEL:fine w1 w2 …`` per line → (sentences, coarse labels) for train and test"""
    def rd(fn):
        text, labels = [], []
        for line in _read_lines(os.path.join(loc, fn)):
            head, _, rest = line.partition(" ")
            labels.append(head.split(":")[0])
            text.append(rest)
        return text, labels
    return rd("train_5500.label"), rd("TREC_10.label")


def eval_trec(encoder, loc, k=10, scan=None, seed=1234):
    (tr_x, tr_y), (te_x, te_y) = load_trec(loc)
    classes = sorted(set(tr_y))
    ytr = np.array([classes.index(c) for c in tr_y])
    yte = np.array([classes.index(c) for c in te_y])
    xtr, xte = encoder.encode(tr_x, verbose=False), encoder.encode(te_x, verbose=False)
    C = _choose_c(xtr, ytr, k, scan or [2 ** t for t in range(0, 9)], seed)
    return float(_lr(C, seed).fit(xtr, ytr).score(xte, yte))


# -------------------------------------------------------------------- pairs
def pair_features(u, v):
    return np.concatenate([np.abs(u - v), u * v], axis=1)


def load_msrp(loc):
    def rd(fn):
        a, b, y = [], [], []
        for i, line in enumerate(_read_lines(os.path.join(loc, fn))):
            if i == 0:
                continue                                   # header
            cols = line.split("\t")
            y.append(int(cols[0]))
            a.append(cols[3])
            b.append(cols[4])
        return a, b, np.array(y)
    return rd("msr_paraphrase_train.txt"), rd("msr_paraphrase_test.txt")


def eval_msrp(encoder, loc, k=10, scan=None, seed=1234):
    """→ (accuracy, F1) on the test split"""
    from sklearn.metrics import f1_score
    (a, b, y), (ta, tb, ty) = load_msrp(loc)
    enc = lambda s: encoder.encode(s, verbose=False)
    xtr, xte = pair_features(enc(a), enc(b)), pair_features(enc(ta), enc(tb))
    C = _choose_c(xtr, y, k, scan or [2 ** t for t in range(0, 9)], seed)
    clf = _lr(C, seed).fit(xtr, y)
    pred = clf.predict(xte)
    return float((pred == ty).mean()), float(f1_score(ty, pred))


def load_sick(loc):
    def rd(fn):
        a, b, s = [], [], []
        for i, line in enumerate(_read_lines(os.path.join(loc, fn))):
            if i == 0:
                continue
            cols = line.split("\t")
            a.append(cols[1])
            b.append(cols[2])
            s.append(float(cols[3]))
        return a, b, np.array(s)
    return rd("SICK_train.txt"), rd("SICK_test_annotated.txt")


def encode_score_labels(scores, nclass=5):
    """relatedness s ∈ [1, 5] → sparse target distribution over the integer scores
    (mass split between ⌊s⌋ and ⌊s⌋+1, as in Tai et al. / the original eval_sick)"""
    y = np.zeros((len(scores), nclass), np.float32)
    for i, s in enumerate(scores):
        lo = int(np.floor(s))
        if lo >= nclass:
            y[i, nclass - 1] = 1.0
        else:
            y[i, lo - 1] = lo + 1 - s
            y[i, lo] = s - lo
    return y


def eval_sick(encoder, loc, seed=1234, epochs=300):
    """→ (pearson, spearman, mse) of the expected relatedness on the test split"""
    import torch
    from scipy.stats import pearsonr, spearmanr
    (a, b, s), (ta, tb, ts) = load_sick(loc)
    enc = lambda x: encoder.encode(x, verbose=False)
    xtr = torch.tensor(pair_features(enc(a), enc(b)), dtype=torch.float32)
    xte = torch.tensor(pair_features(enc(ta), enc(tb)), dtype=torch.float32)
    ytr = torch.tensor(encode_score_labels(s))
    torch.manual_seed(seed)
    lin = torch.nn.Linear(xtr.shape[1], 5)
    opt = torch.optim.Adam(lin.parameters(), lr=0.01)
    for _ in range(epochs):                                # KL(target ‖ softmax) regression
        opt.zero_grad()
        loss = -(ytr * torch.log_softmax(lin(xtr), -1)).sum(1).mean()
        loss.backward()
        opt.step()
    with torch.no_grad():
        pred = (torch.softmax(lin(xte), -1) * torch.arange(1.0, 6.0)).sum(1).numpy()
    return float(pearsonr(pred, ts)[0]), float(spearmanr(pred, ts)[0]), \
        float(np.mean((pred - ts) ** 2))


def evaluate(encoder, eval_task, data_dir):
    if eval_task in TASK_FILES:
        return {"accuracy": eval_nested_kfold(encoder, eval_task, data_dir)}
    if eval_task == "TREC":
        return {"accuracy": eval_trec(encoder, data_dir)}
    if eval_task == "MSRP":
        acc, f1 = eval_msrp(encoder, data_dir)
        return {"accuracy": acc, "f1": f1}
    if eval_task == "SICK":
        p, sp, mse = eval_sick(encoder, data_dir)
        return {"pearson": p, "spearman": sp, "mse": mse}
    raise ValueError("Unrecognized eval_task: %s" % eval_task)


def main(argv=None):          # pragma: no cover - thin CLI
    import torch
    from . import configuration
    from .encoder import EncoderManager
    ap = argparse.ArgumentParser()
    ap.add_argument("--eval_task", default="CR", help="MR, CR, SUBJ, MPQA, SICK, MSRP, TREC")
    ap.add_argument("--data_dir", required=True)
    for kind in ("uni", "bi"):
        ap.add_argument("--%s_vocab_file" % kind)
        ap.add_argument("--%s_embeddings_file" % kind)
        ap.add_argument("--%s_checkpoint_path" % kind)
    a = ap.parse_args(argv)
    mgr = EncoderManager()
    for kind, bidi in (("uni", False), ("bi", True)):
        ck = getattr(a, kind + "_checkpoint_path")
        if ck:
            from ... import checkpoint as _ckpt
            path = _ckpt.latest_checkpoint(ck) if os.path.isdir(ck) else ck
            state = _ckpt.load_logical(path)
            mgr.load_model(configuration.model_config(bidirectional_encoder=bidi),
                           getattr(a, kind + "_vocab_file"),
                           getattr(a, kind + "_embeddings_file"), state)
    print(evaluate(mgr, a.eval_task, a.data_dir))
    mgr.close()


if __name__ == "__main__":    # pragma: no cover
    main()
