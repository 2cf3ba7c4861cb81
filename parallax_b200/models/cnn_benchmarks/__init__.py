"""The CNN benchmark example as a library (reference
`parallax/parallax/examples/tf_cnn_benchmarks/`): parameter table, datasets
(ImageNet TFRecords, CIFAR-10, synthetic), image preprocessing, per-model
defaults and the `BenchmarkCNN` harness.  Networks live in `models/cnn.py` and
`models/resnet.py`; the driver is
`examples/cnn_benchmarks/CNNBenchmark_distributed_driver.py`."""
from . import benchmark_cnn, datasets, model_config, preprocessing
from .benchmark_cnn import BenchmarkCNN, Params, get_learning_rate, make_params

__all__ = ["benchmark_cnn", "datasets", "model_config", "preprocessing", "BenchmarkCNN",
           "Params", "get_learning_rate", "make_params"]
