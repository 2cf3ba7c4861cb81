"""CNN benchmark driver (reference
`examples/tf_cnn_benchmarks/CNNBenchmark_distributed_driver.py:50-91`,
`benchmark_cnn.py:487-1014`): every `tf_cnn_benchmarks` flag of the harness
(`--model`, `--batch_size`, `--optimizer`, `--use_fp16`, `--data_dir`, `--eval`,
…) plus the Parallax flags; reports images/sec and steps/sec.

    # synthetic ImageNet-shaped data, ResNet-50, AR mode, bf16, 8 GPUs
    python examples/cnn_benchmarks/CNNBenchmark_distributed_driver.py --model resnet50 \
        --optimizer momentum --use_fp16 --run_option MPI \
        --resource_info_file localhost:0,1,2,3,4,5,6,7
    # CIFAR-10 from disk, evaluation of the checkpoint in --ckpt_dir
    python examples/cnn_benchmarks/CNNBenchmark_distributed_driver.py --model resnet56 \
        --data_dir /data/cifar-10-batches-py --data_name cifar10 --eval --ckpt_dir /tmp/ck
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import parallax_b200 as parallax
from parallax_b200.models.cnn_benchmarks import benchmark_cnn
import parallax_config

ap = parallax_config.add_flags(argparse.ArgumentParser(conflict_handler="resolve"))
benchmark_cnn.add_arguments(ap)          # its --cuda_graph (default on) replaces the shared flag
ap.add_argument("--max_steps", type=int, default=None, help="alias of --num_batches")
ap.add_argument("--log_frequency", type=int, default=None, help="alias of --display_every")
ap.set_defaults(run_option="MPI", model="resnet50")
FLAGS = ap.parse_args()


def main():
    if FLAGS.max_steps:
        FLAGS.num_batches = FLAGS.max_steps
    if FLAGS.log_frequency:
        FLAGS.display_every = FLAGS.log_frequency
    if FLAGS.compute_dtype in ("bf16", "bfloat16"):
        FLAGS.use_fp16 = True
    bench = benchmark_cnn.BenchmarkCNN(benchmark_cnn.make_params_from_flags(FLAGS))
    cfg = parallax_config.build_config(FLAGS)
    sc = dict(cfg.sess_config or {})
    sc.update(bench.sess_config())
    cfg.sess_config = sc
    if FLAGS.checkpoint_dir and not cfg.ckpt_config.ckpt_dir:
        cfg.ckpt_config = parallax.CheckPointConfig(ckpt_dir=FLAGS.checkpoint_dir)
    # the LR schedule of the reference depends on the GLOBAL batch; the resource spec
    # gives the worker count before parallel_run re-executes this script per worker
    from parallax_b200.resource import parse_resource_info, worker_layout
    try:
        nw = len(worker_layout(parse_resource_info(FLAGS.resource_info_file,
                                                   cfg.normalized_run_option())))
    except Exception:
        nw = int(os.environ.get("WORLD_SIZE", "1"))
    graph = bench.build_graph(num_workers=nw)
    sess, num_workers, worker_id, _ = parallax.parallel_run(
        graph, FLAGS.resource_info_file, sync=FLAGS.sync, parallax_config=cfg)
    res = bench.run(sess, num_workers, worker_id)
    if worker_id == 0:
        parallax.log.info("result: %s", res)
    sess.close()


if __name__ == "__main__":
    main()
