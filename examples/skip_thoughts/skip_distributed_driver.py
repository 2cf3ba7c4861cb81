"""Skip-thoughts driver (reference `examples/skip_thoughts/skip_distributed_driver.py:48-104`):
GRU encoder + previous/next-sentence decoders, Adam; the input is sharded with
`parallax.shard.create_num_shards_and_shard_id` (`ops/input_ops.py:92`).

    python examples/skip_thoughts/skip_distributed_driver.py --resource_info_file localhost
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

import parallax_b200 as parallax
from parallax_b200.models.seq2seq import SkipThoughts, skip_thoughts_graph
import parallax_config

ap = parallax_config.add_flags(argparse.ArgumentParser())
ap.add_argument("--vocab_size", type=int, default=20000)
ap.add_argument("--word_dim", type=int, default=620)
ap.add_argument("--encoder_dim", type=int, default=2400)
ap.add_argument("--batch_size", type=int, default=128)
ap.add_argument("--seq_len", type=int, default=30)
ap.add_argument("--learning_rate", type=float, default=0.0008)
ap.add_argument("--max_steps", type=int, default=200)
ap.add_argument("--log_frequency", type=int, default=20)
FLAGS = ap.parse_args()


def main():
    model = SkipThoughts(FLAGS.vocab_size, FLAGS.word_dim, FLAGS.encoder_dim)
    graph = skip_thoughts_graph(model, FLAGS.learning_rate)
    num_shards, shard_id = parallax.shard.create_num_shards_and_shard_id()

    def triples():
        """sentence triples (prev, cur, next); example i belongs to shard i % num_shards"""
        g = torch.Generator().manual_seed(0)
        i = 0
        while True:
            t = torch.randint(1, FLAGS.vocab_size, (3, FLAGS.seq_len + 1), generator=g)
            if i % int(num_shards) == int(shard_id):
                yield t
            i += 1

    def run(sess, num_workers, worker_id, num_replicas_per_worker):
        it = triples()
        mask = torch.ones(FLAGS.batch_size, FLAGS.seq_len)
        t0 = time.time()
        for step in range(FLAGS.max_steps):
            b = torch.stack([next(it) for _ in range(FLAGS.batch_size)])     # [B,3,L+1]
            feeds = {"encode_ids": [b[:, 1, :-1]],
                     "pre_in": [b[:, 0, :-1]], "pre_out": [b[:, 0, 1:]], "pre_mask": [mask],
                     "post_in": [b[:, 2, :-1]], "post_out": [b[:, 2, 1:]], "post_mask": [mask]}
            loss, gs, _ = sess.run(["loss", "global_step", "train_op"], feeds)
            if worker_id == 0 and (step + 1) % FLAGS.log_frequency == 0:
                parallax.log.info("global step %d: loss = %.4f (%.2f sec/step)", gs[0], loss[0],
                                  (time.time() - t0) / FLAGS.log_frequency)
                t0 = time.time()
        sess.close()

    sess, nw, wid, nrep = parallax.parallel_run(
        graph, FLAGS.resource_info_file, sync=FLAGS.sync,
        parallax_config=parallax_config.build_config(FLAGS))
    run(sess, nw, wid, nrep)


if __name__ == "__main__":
    main()
