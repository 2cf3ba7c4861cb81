"""Quick-start example: 2→1 linear regression trained with SGD
(reference `examples/simple/simple_driver.py:93-136`, `doc/quick_start.md`).

    python examples/simple/simple_driver.py --resource_info_file localhost:0,1
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import parallax_b200 as parallax
from parallax_b200.models.simple import LinearRegression
import parallax_config

ap = parallax_config.add_flags(argparse.ArgumentParser())
ap.add_argument("--max_steps", type=int, default=200)
ap.add_argument("--learning_rate", type=float, default=0.05)
FLAGS = ap.parse_args()


def main():
    torch.manual_seed(0)
    # ---- the single-device program -----------------------------------------------
    single_gpu_graph = parallax.Graph(
        LinearRegression(2), optimizer=parallax.optim.GradientDescent(FLAGS.learning_rate))

    def run(sess, num_workers, worker_id, num_replicas_per_worker):
        rng = np.random.RandomState(worker_id)
        for i in range(FLAGS.max_steps):
            xs, ys = [], []
            for _ in range(num_replicas_per_worker):
                x = rng.rand(16, 2).astype(np.float32)
                xs.append(x)
                ys.append((x @ np.array([10.0, -3.0], np.float32) + 2.0).astype(np.float32))
            loss, step, _ = sess.run(["loss", "global_step", "train_op"],
                                     feed_dict={"x": xs, "y": ys})
            if i % 50 == 0:
                parallax.log.info("worker %d step %d loss %.5f", worker_id, step[0], loss[0])
        w = sess.engine.state_dict()["dense"]["master"]
        if worker_id == 0:
            parallax.log.info("learned: w=%s b=%s", w["linear.weight"].tolist(),
                              w["linear.bias"].tolist())
        sess.close()

    sess, num_workers, worker_id, num_replicas_per_worker = parallax.parallel_run(
        single_gpu_graph, FLAGS.resource_info_file, sync=FLAGS.sync,
        parallax_config=parallax_config.build_config(FLAGS))
    run(sess, num_workers, worker_id, num_replicas_per_worker)


if __name__ == "__main__":
    main()
