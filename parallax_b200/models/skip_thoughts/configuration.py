"""Default configuration of the skip-thoughts model and training.

Parity: `examples/skip_thoughts/configuration.py:28-114` — `model_config`
(vocab 20 000, word dim 620, encoder dim 2 400, batch 128, uniform init scale
0.1, optional bidirectional encoder) and `training_config` (Adam-style lr
0.0008 halved every 400 000 steps, 500 000 steps, gradient clip 5)."""


class _HParams(object):
    """attribute bag"""

    def __repr__(self):
        return "_HParams(%s)" % ", ".join("%s=%r" % kv for kv in sorted(vars(self).items()))


def model_config(input_file_pattern=None, input_queue_capacity=640000,
                 num_input_reader_threads=1, shuffle_input_data=True,
                 uniform_init_scale=0.1, vocab_size=20000, batch_size=128,
                 word_embedding_dim=620, bidirectional_encoder=False, encoder_dim=2400,
                 num_embedding_partitions=0):
    """`num_embedding_partitions` > 1 creates the word-embedding table under
    `parallax.get_partitioner` (a partitioned sparse variable)."""
    c = _HParams()
    c.input_file_pattern = input_file_pattern
    c.input_queue_capacity = input_queue_capacity
    c.num_input_reader_threads = num_input_reader_threads
    c.shuffle_input_data = shuffle_input_data
    c.uniform_init_scale = uniform_init_scale
    c.vocab_size = vocab_size
    c.batch_size = batch_size
    c.word_embedding_dim = word_embedding_dim
    c.bidirectional_encoder = bidirectional_encoder
    c.encoder_dim = encoder_dim
    c.num_embedding_partitions = num_embedding_partitions
    return c


def training_config(learning_rate=0.0008, learning_rate_decay_factor=0.5,
                    learning_rate_decay_steps=400000, number_of_steps=500000,
                    clip_gradient_norm=5.0, save_model_secs=600, save_summaries_secs=600):
    if learning_rate_decay_factor and not learning_rate_decay_steps:
        raise ValueError("learning_rate_decay_factor requires learning_rate_decay_steps.")
    c = _HParams()
    c.learning_rate = learning_rate
    c.learning_rate_decay_factor = learning_rate_decay_factor
    c.learning_rate_decay_steps = learning_rate_decay_steps
    c.number_of_steps = number_of_steps
    c.clip_gradient_norm = clip_gradient_norm
    c.save_model_secs = save_model_secs
    c.save_summaries_secs = save_summaries_secs
    return c
