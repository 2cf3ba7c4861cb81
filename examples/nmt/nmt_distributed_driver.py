"""NMT driver (reference `examples/nmt/nmt_distributed_driver.py:76-189`):
GNMT-style seq2seq with partitioned embeddings; the parallel corpus is sharded
across workers with `parallax.shard.shard` (`utils/iterator_utils.py:103`);
only worker 0 logs statistics (`:147-163`).

    python examples/nmt/nmt_distributed_driver.py --synthetic --resource_info_file localhost
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

import parallax_b200 as parallax
from parallax_b200.models.seq2seq import NMT, nmt_graph
import parallax_config

ap = parallax_config.add_flags(argparse.ArgumentParser())
ap.add_argument("--src_file", default=None)
ap.add_argument("--tgt_file", default=None)
ap.add_argument("--synthetic", action="store_true")
ap.add_argument("--src_vocab_size", type=int, default=32000)
ap.add_argument("--tgt_vocab_size", type=int, default=32000)
ap.add_argument("--num_units", type=int, default=512)
ap.add_argument("--num_layers", type=int, default=4)
ap.add_argument("--num_embeddings_partitions", type=int, default=4)
ap.add_argument("--batch_size", type=int, default=128)
ap.add_argument("--max_len", type=int, default=50)
ap.add_argument("--learning_rate", type=float, default=1.0)
ap.add_argument("--max_gradient_norm", type=float, default=5.0)
ap.add_argument("--max_steps", type=int, default=200)
ap.add_argument("--log_frequency", type=int, default=20)
FLAGS = ap.parse_args()


def corpus():
    """(src_ids, tgt_ids) pairs — hashed-token ids when reading real files."""
    if FLAGS.synthetic or not FLAGS.src_file:
        g = torch.Generator().manual_seed(0)
        for _ in range(100000):
            n = int(torch.randint(5, FLAGS.max_len, (1,), generator=g))
            yield (torch.randint(3, FLAGS.src_vocab_size, (n,), generator=g),
                   torch.randint(3, FLAGS.tgt_vocab_size, (n,), generator=g))
    else:
        with open(FLAGS.src_file) as fs, open(FLAGS.tgt_file) as ft:
            for s, t in zip(fs, ft):
                si = [3 + hash(w) % (FLAGS.src_vocab_size - 3) for w in s.split()][:FLAGS.max_len]
                ti = [3 + hash(w) % (FLAGS.tgt_vocab_size - 3) for w in t.split()][:FLAGS.max_len]
                if si and ti:
                    yield torch.tensor(si), torch.tensor(ti)


def batches(ds):
    buf = []
    for pair in ds:
        buf.append(pair)
        if len(buf) == FLAGS.batch_size:
            L = FLAGS.max_len
            src = torch.zeros(len(buf), L, dtype=torch.long)
            tin = torch.zeros(len(buf), L + 1, dtype=torch.long)
            tout = torch.zeros(len(buf), L + 1, dtype=torch.long)
            w = torch.zeros(len(buf), L + 1)
            for i, (s, t) in enumerate(buf):
                src[i, :len(s)] = s
                tin[i, 0] = 1
                tin[i, 1:len(t) + 1] = t
                tout[i, :len(t)] = t
                tout[i, len(t)] = 2
                w[i, :len(t) + 1] = 1
            yield src, tin, tout, w
            buf = []


def main():
    model = NMT(FLAGS.src_vocab_size, FLAGS.tgt_vocab_size, FLAGS.num_units, FLAGS.num_layers,
                FLAGS.num_embeddings_partitions)
    graph = nmt_graph(model, FLAGS.learning_rate, FLAGS.max_gradient_norm)
    ds = parallax.shard.shard(corpus())        # sharded after parallel_run assigns ids

    def run(sess, num_workers, worker_id, num_replicas_per_worker):
        t0, words = time.time(), 0
        for step, (src, tin, tout, w) in enumerate(batches(ds)):
            if step >= FLAGS.max_steps:
                break
            loss, gs, _ = sess.run(["loss", "global_step", "train_op"],
                                   {"src": [src], "tgt_in": [tin], "tgt_out": [tout],
                                    "tgt_weight": [w]})
            words += int(w.sum())
            if worker_id == 0 and (step + 1) % FLAGS.log_frequency == 0:
                dt = time.time() - t0
                parallax.log.info("global step %d  loss %.3f  wps %.0f", gs[0], loss[0],
                                  words * num_workers / dt)
                t0, words = time.time(), 0
        sess.close()

    sess, nw, wid, nrep = parallax.parallel_run(
        graph, FLAGS.resource_info_file, sync=FLAGS.sync,
        parallax_config=parallax_config.build_config(FLAGS))
    run(sess, nw, wid, nrep)


if __name__ == "__main__":
    main()
