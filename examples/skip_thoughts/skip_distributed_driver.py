"""Skip-thoughts driver (reference
`examples/skip_thoughts/skip_distributed_driver.py:48-104`, `train.py:44-99`):
layer-normalised GRU encoder + previous/next-sentence decoders, Adam with
staircase decay; the input shards are split across workers with
`parallax.shard.create_num_shards_and_shard_id` (`ops/input_ops.py:92`).

    # synthetic corpus
    python examples/skip_thoughts/skip_distributed_driver.py --synthetic --max_steps 100
    # preprocessed shards (python -m parallax_b200.models.skip_thoughts.preprocess_dataset …)
    python examples/skip_thoughts/skip_distributed_driver.py \
        --input_file_pattern "/data/skip/train-?????-of-00100.npz" \
        --resource_info_file localhost:0,1,2,3
"""
import argparse
import math
import os
import random
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import parallax_b200 as parallax
from parallax_b200.models import skip_thoughts as st
import parallax_config

ap = parallax_config.add_flags(argparse.ArgumentParser())
ap.add_argument("--input_file_pattern", default=None)
ap.add_argument("--train_dir", default="/tmp/skip_thoughts")
ap.add_argument("--synthetic", action="store_true",
                help="write a toy corpus + shards under train_dir/synthetic")
ap.add_argument("--vocab_size", type=int, default=20000)
ap.add_argument("--word_embedding_dim", type=int, default=620)
ap.add_argument("--encoder_dim", type=int, default=2400)
ap.add_argument("--bidirectional_encoder", action="store_true")
ap.add_argument("--num_embedding_partitions", type=int, default=0)
ap.add_argument("--batch_size", type=int, default=128)
ap.add_argument("--learning_rate", type=float, default=0.0008)
ap.add_argument("--learning_rate_decay_factor", type=float, default=0.5)
ap.add_argument("--learning_rate_decay_steps", type=int, default=400000)
ap.add_argument("--clip_gradient_norm", type=float, default=5.0)
ap.add_argument("--max_len", type=int, default=31,
                help="longest sentence incl. <eos> (preprocess_dataset --max_sentence_length+1)")
ap.add_argument("--max_steps", type=int, default=500000)
ap.add_argument("--log_frequency", type=int, default=100)
FLAGS = ap.parse_args()


def synthetic_shards(d, vocab_size, n_books=8, sentences=400):
    """toy 'books': every sentence continues a counting pattern, so the next and the
    previous sentence are predictable from the current one"""
    if not os.path.isdir(d):
        os.makedirs(d)
        rng = random.Random(0)
        for b in range(n_books):
            with open(os.path.join(d, "book%d.txt" % b), "w") as f:
                start = rng.randrange(vocab_size)
                for s in range(sentences):
                    n = rng.randint(4, 12)
                    f.write(" ".join("w%d" % ((start + s * 3 + k) % (vocab_size - 2))
                                     for k in range(n)) + "\n")
        files = sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".txt"))
        st.preprocess_dataset.preprocess(files, d, num_words=vocab_size, train_output_shards=8,
                                         num_validation_sentences=200)
    return os.path.join(d, "train-?????-of-00008.npz")


def main():
    pattern = FLAGS.input_file_pattern
    if FLAGS.synthetic or not pattern:
        pattern = synthetic_shards(os.path.join(FLAGS.train_dir, "synthetic"), FLAGS.vocab_size)
    mc = st.model_config(input_file_pattern=pattern, vocab_size=FLAGS.vocab_size,
                         batch_size=FLAGS.batch_size, word_embedding_dim=FLAGS.word_embedding_dim,
                         bidirectional_encoder=FLAGS.bidirectional_encoder,
                         encoder_dim=FLAGS.encoder_dim,
                         num_embedding_partitions=FLAGS.num_embedding_partitions)
    tc = st.training_config(FLAGS.learning_rate, FLAGS.learning_rate_decay_factor,
                            FLAGS.learning_rate_decay_steps, FLAGS.max_steps,
                            FLAGS.clip_gradient_norm)
    model = st.SkipThoughtsModel(mc)
    graph = st.skip_thoughts_graph(model, tc)
    # handles planted before parallel_run, resolved per worker by it
    queue = st.input_ops.prefetch_input_data(pattern, mc.batch_size, mc.shuffle_input_data,
                                             mc.input_queue_capacity, pin_memory=True)
    cfg = parallax_config.build_config(FLAGS)
    if cfg.ckpt_config.ckpt_dir is None:
        cfg.ckpt_config = parallax.CheckPointConfig(ckpt_dir=FLAGS.train_dir,
                                                    save_ckpt_secs=tc.save_model_secs)
    # three lookups per step (encode, previous, next) of at most max_len ids each; the
    # NVLink fabric sizes a table's receive rings once, at the first step
    sc = dict(cfg.sess_config or {})
    sc.setdefault("sparse_capacity", {"word_embedding.weight": 3 * mc.batch_size * FLAGS.max_len})
    cfg.sess_config = sc
    sess, num_workers, worker_id, _ = parallax.parallel_run(
        graph, FLAGS.resource_info_file, sync=FLAGS.sync, parallax_config=cfg)
    t0 = time.time()
    for batch in queue:
        loss, w, gs, _ = sess.run(["loss", "sum_weights", "global_step", "train_op"],
                                  st.feed_from_batch(batch))
        if worker_id == 0 and gs[0] % FLAGS.log_frequency == 0:
            parallax.log.info("global step %d: loss = %.4f  ppl = %.2f (%.3f sec/step)", gs[0],
                              loss[0], math.exp(loss[0] / max(w[0], 1.0)),
                              (time.time() - t0) / FLAGS.log_frequency)
            t0 = time.time()
        if gs[0] >= FLAGS.max_steps:
            break
    sess.close()


if __name__ == "__main__":
    main()
