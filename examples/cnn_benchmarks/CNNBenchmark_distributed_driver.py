"""CNN benchmark driver (reference
`examples/tf_cnn_benchmarks/CNNBenchmark_distributed_driver.py:50-91`,
`benchmark_cnn.py:487-1014`): synthetic images, momentum/sgd/rmsprop, reports
images/sec and steps/sec.

    python examples/cnn_benchmarks/CNNBenchmark_distributed_driver.py --model resnet50 \
        --run_option MPI --compute_dtype bf16 --resource_info_file localhost:0,1,2,3
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

import parallax_b200 as parallax
from parallax_b200.models import cnn
import parallax_config

ap = parallax_config.add_flags(argparse.ArgumentParser())
ap.add_argument("--model", default="resnet50", choices=sorted(cnn.MODELS))
ap.add_argument("--batch_size", type=int, default=64)
ap.add_argument("--optimizer", default="momentum", choices=["momentum", "sgd", "rmsprop"])
ap.add_argument("--learning_rate", type=float, default=0.01)
ap.add_argument("--num_classes", type=int, default=1000)
ap.add_argument("--max_steps", type=int, default=500)
ap.add_argument("--log_frequency", type=int, default=50)
ap.add_argument("--params_stat", action="store_true",
                help="print total parameter/gradient element counts")
FLAGS = ap.parse_args()


def main():
    model = cnn.get_model(FLAGS.model, FLAGS.num_classes)
    graph = cnn.cnn_graph(model, FLAGS.optimizer, FLAGS.learning_rate)
    hw = cnn.image_size(model)

    def run(sess, num_workers, worker_id, num_replicas_per_worker):
        if FLAGS.params_stat and worker_id == 0:
            n = sum(v.numel for v in sess.engine.analysis.variables.values())
            parallax.log.info("total parameters / gradient elements: %d", n)
        images = torch.randn(FLAGS.batch_size, 3, hw, hw)
        labels = torch.randint(0, FLAGS.num_classes, (FLAGS.batch_size,))
        if torch.cuda.is_available():
            images, labels = images.pin_memory(), labels.pin_memory()
        start = time.time()
        for step in range(FLAGS.max_steps):
            loss, _ = sess.run(["loss", "train_op"],
                               feed_dict={"images": [images], "labels": [labels]})
            if (step + 1) % FLAGS.log_frequency == 0 and worker_id == 0:
                dt = time.time() - start
                start = time.time()
                sps = FLAGS.log_frequency / dt
                parallax.log.info("step %d  loss %.3f  %.2f steps/sec  %.1f images/sec (total)",
                                  step + 1, loss[0], sps,
                                  sps * FLAGS.batch_size * num_workers)
        sess.close()

    sess, num_workers, worker_id, num_replicas_per_worker = parallax.parallel_run(
        graph, FLAGS.resource_info_file, sync=FLAGS.sync,
        parallax_config=parallax_config.build_config(FLAGS))
    run(sess, num_workers, worker_id, num_replicas_per_worker)


if __name__ == "__main__":
    main()
