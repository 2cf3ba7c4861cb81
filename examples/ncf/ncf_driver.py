"""NCF / NeuMF driver (BASELINE.json config 4: very large user table, PS mode,
8-way sparse-variable partitioning, local aggregation).  Not a reference example —
the same `parallel_run` contract on a recommender shape.

    python examples/ncf/ncf_driver.py --num_users 100000000 --run_option PS \
        --resource_info_file localhost:0,1,2,3,4,5,6,7 --compute_dtype bf16 --cuda_graph
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

import parallax_b200 as parallax
from parallax_b200.models.ncf import NeuMF, ncf_graph
import parallax_config

ap = parallax_config.add_flags(argparse.ArgumentParser())
ap.add_argument("--num_users", type=int, default=1000000)
ap.add_argument("--num_items", type=int, default=100000)
ap.add_argument("--mf_dim", type=int, default=32)
ap.add_argument("--num_partitions", type=int, default=8)
ap.add_argument("--batch_size", type=int, default=65536)
ap.add_argument("--learning_rate", type=float, default=0.001)
ap.add_argument("--zipf", type=float, default=1.05, help="skew of the synthetic id stream")
ap.add_argument("--max_steps", type=int, default=200)
ap.add_argument("--log_frequency", type=int, default=50)
ap.set_defaults(run_option="PS")
FLAGS = ap.parse_args()


def zipf_ids(n, size, alpha, gen):
    """heavy-tailed ids in [0, n): a few hot users/items, like real interaction logs"""
    u = torch.rand(size, generator=gen).clamp_(min=1e-9)
    ids = (n ** u - 1.0) if alpha <= 1.0 else ((1.0 - u) ** (-1.0 / (alpha - 1.0 + 1e-6)) - 1.0)
    return ids.to(torch.int64).remainder_(n)


def main():
    model = NeuMF(FLAGS.num_users, FLAGS.num_items, FLAGS.mf_dim,
                  num_partitions=FLAGS.num_partitions, lazy=FLAGS.num_users > 5_000_000)
    graph = ncf_graph(model, FLAGS.learning_rate)
    sess, num_workers, worker_id, _ = parallax.parallel_run(
        graph, FLAGS.resource_info_file, sync=FLAGS.sync,
        parallax_config=parallax_config.build_config(FLAGS))
    gen = torch.Generator().manual_seed(17 + worker_id)
    B, t0 = FLAGS.batch_size, time.time()
    for step in range(1, FLAGS.max_steps + 1):
        users = zipf_ids(FLAGS.num_users, B, FLAGS.zipf, gen)
        items = zipf_ids(FLAGS.num_items, B, FLAGS.zipf, gen)
        labels = ((users + items) % 2).to(torch.int64)          # a learnable parity rule
        loss, _ = sess.run(["loss", "train_op"],
                           {"users": [users], "items": [items], "labels": [labels]})
        if worker_id == 0 and step % FLAGS.log_frequency == 0:
            dt = time.time() - t0
            t0 = time.time()
            parallax.log.info("step %d  loss %.4f  %.2fM samples/sec (total)", step, loss[0],
                              FLAGS.log_frequency * B * num_workers / dt / 1e6)
    sess.close()


if __name__ == "__main__":
    main()
