"""Corpus → vocabulary + sharded training records.

Parity: `examples/skip_thoughts/data/preprocess_dataset.py:89-301`: build (or
load) the vocabulary — ``<eos>`` = 0, ``<unk>`` = 1, then words by decreasing
frequency up to `num_words` — write ``vocab.txt`` / ``word_counts.txt``; turn every
run of three consecutive sentences of a pre-tokenised ``.txt`` file into one
(predecessor, current, successor) example (ids end with ``<eos>``; triples with
an empty or over-long sentence are skipped); shuffle; split off a validation
set; write ``train-?????-of-?????`` / ``validation-…`` shards.

    python -m parallax_b200.models.skip_thoughts.preprocess_dataset \
        --input_files "books/*.txt" --output_dir /data/skip --num_words 20000
"""
import argparse
import collections
import glob
import os
import random

from . import special_words
from .input_ops import write_shard


def build_vocabulary(input_files, num_words=20000, vocab_file=None, output_dir=None):
    """→ OrderedDict word → id"""
    if vocab_file:
        vocab = collections.OrderedDict()
        with open(vocab_file, encoding="utf-8") as f:
            for i, line in enumerate(f):
                word = line.strip()
                assert word not in vocab, "Attempting to add word twice: %s" % word
                vocab[word] = i
        return vocab
    counts = collections.Counter()
    for fn in input_files:
        with open(fn, encoding="utf-8") as f:
            for sentence in f:
                counts.update(sentence.split())
    ordered = sorted(counts.items(), key=lambda kv: (-kv[1], kv[0]))
    vocab = collections.OrderedDict()
    vocab[special_words.EOS] = special_words.EOS_ID
    vocab[special_words.UNK] = special_words.UNK_ID
    for w, _ in ordered[:max(num_words - 2, 0)]:
        if w not in vocab:
            vocab[w] = len(vocab)
    if output_dir:
        os.makedirs(output_dir, exist_ok=True)
        with open(os.path.join(output_dir, "vocab.txt"), "w", encoding="utf-8") as f:
            f.write("\n".join(vocab.keys()))
        with open(os.path.join(output_dir, "word_counts.txt"), "w", encoding="utf-8") as f:
            for w, c in ordered:
                f.write("%s %d\n" % (w, c))
    return vocab


def sentence_to_ids(sentence, vocab):
    ids = [vocab.get(w, special_words.UNK_ID) for w in sentence.split()]
    ids.append(special_words.EOS_ID)
    return ids


def process_input_file(filename, vocab, max_sentence_length=30, stats=None):
    """every run of three consecutive sentences → one example"""
    stats = stats if stats is not None else collections.Counter()
    examples = []
    pre = cur = None
    with open(filename, encoding="utf-8") as f:
        for line in f:
            nxt = line.strip()
            stats["sentences_seen"] += 1
            if pre is not None and cur is not None:
                if pre and cur and nxt:
                    ids = [sentence_to_ids(s, vocab) for s in (pre, cur, nxt)]
                    if max(len(i) for i in ids) - 1 <= max_sentence_length:
                        examples.append((ids[1], ids[0], ids[2]))
                        stats["sentences_output"] += 1
                    else:
                        stats["sentences_too_long"] += 1
                else:
                    stats["sentences_skipped_empty"] += 1
            pre, cur = cur, nxt
    return examples


def write_dataset(name, dataset, indices, num_shards, output_dir):
    """round-robin-free contiguous split of `indices` into `num_shards` files"""
    files = []
    borders = [int(round(i * len(indices) / float(num_shards))) for i in range(num_shards + 1)]
    for i in range(num_shards):
        fn = os.path.join(output_dir, "%s-%.5d-of-%.5d" % (name, i, num_shards))
        files.append(write_shard(fn, [dataset[j] for j in indices[borders[i]:borders[i + 1]]]))
    return files


def preprocess(input_files, output_dir, num_words=20000, vocab_file=None,
               max_sentence_length=30, num_validation_sentences=50000, train_output_shards=100,
               validation_output_shards=1, seed=123):
    os.makedirs(output_dir, exist_ok=True)
    vocab = build_vocabulary(input_files, num_words, vocab_file, output_dir)
    stats = collections.Counter()
    dataset = []
    for fn in input_files:
        dataset.extend(process_input_file(fn, vocab, max_sentence_length, stats))
    indices = list(range(len(dataset)))
    random.Random(seed).shuffle(indices)
    nval = min(num_validation_sentences, len(indices) // 10)
    val, train = indices[:nval], indices[nval:]
    out = {"vocab": vocab, "stats": stats,
           "train": write_dataset("train", dataset, train, train_output_shards, output_dir),
           "validation": write_dataset("validation", dataset, val, validation_output_shards,
                                       output_dir) if nval else []}
    return out


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--input_files", required=True, help="comma separated glob patterns")
    ap.add_argument("--vocab_file", default="")
    ap.add_argument("--output_dir", required=True)
    ap.add_argument("--train_output_shards", type=int, default=100)
    ap.add_argument("--validation_output_shards", type=int, default=1)
    ap.add_argument("--num_validation_sentences", type=int, default=50000)
    ap.add_argument("--num_words", type=int, default=20000)
    ap.add_argument("--max_sentences", type=int, default=0)
    ap.add_argument("--max_sentence_length", type=int, default=30)
    a = ap.parse_args(argv)
    files = []
    for p in a.input_files.split(","):
        files.extend(sorted(glob.glob(p)))
    if not files:
        raise ValueError("Found no files matching %s" % a.input_files)
    out = preprocess(files, a.output_dir, a.num_words, a.vocab_file or None,
                     a.max_sentence_length, a.num_validation_sentences,
                     a.train_output_shards, a.validation_output_shards)
    print("wrote %d train / %d validation shards; %s" %
          (len(out["train"]), len(out["validation"]), dict(out["stats"])))


if __name__ == "__main__":
    main()
