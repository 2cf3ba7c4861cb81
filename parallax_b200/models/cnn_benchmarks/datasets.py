"""Benchmark datasets (`examples/tf_cnn_benchmarks/datasets.py:34-177`):
`ImagenetData` (TFRecord shards ``train-?????-of-01024`` / ``validation-…``, 1 000
classes, 1 281 167 / 50 000 examples), `Cifar10Data` (python pickle batches,
50 000 / 10 000) and the synthetic stand-ins used when no `data_dir` is given
(`Dataset.use_synthetic_gpu_images`)."""
import glob
import os
import pickle

import numpy as np


class Dataset(object):
    def __init__(self, name, height=None, width=None, depth=None, data_dir=None,
                 queue_runner_required=False, num_classes=1000):
        self.name, self.height, self.width = name, height, width
        self.depth = depth or 3
        self.data_dir = data_dir
        self._queue_runner_required = queue_runner_required
        self._num_classes = num_classes

    def tf_record_pattern(self, subset):
        return os.path.join(self.data_dir, "%s-*-of-*" % subset)

    @property
    def num_classes(self):
        return self._num_classes

    @num_classes.setter
    def num_classes(self, val):
        self._num_classes = val

    def num_examples_per_epoch(self, subset="train"):
        raise NotImplementedError

    def use_synthetic_gpu_images(self):
        return not self.data_dir

    def queue_runner_required(self):
        return self._queue_runner_required

    def get_image_preprocessor(self):
        raise NotImplementedError

    def __str__(self):
        return self.name


class ImagenetData(Dataset):
    def __init__(self, data_dir=None):
        super().__init__("imagenet", 300, 300, data_dir=data_dir)

    def num_examples_per_epoch(self, subset="train"):
        if subset == "train":
            return 1281167
        if subset == "validation":
            return 50000
        raise ValueError('Invalid data subset "%s"' % subset)

    def files(self, subset):
        files = sorted(glob.glob(self.tf_record_pattern(subset)))
        if not files:
            raise ValueError("no %s records under %s" % (subset, self.data_dir))
        return files

    def get_image_preprocessor(self):
        from . import preprocessing
        if self.use_synthetic_gpu_images():
            return preprocessing.SyntheticImagePreprocessor
        return preprocessing.RecordInputImagePreprocessor


class Cifar10Data(Dataset):
    def __init__(self, data_dir=None):
        super().__init__("cifar10", 32, 32, data_dir=data_dir, queue_runner_required=True,
                         num_classes=10)

    def read_data_files(self, subset="train"):
        """→ (images uint8 [N,3,32,32], labels int64 [N]) from the python-version batches"""
        assert self.data_dir, "Cannot call `read_data_files` when using synthetic data"
        if subset == "train":
            names = [os.path.join(self.data_dir, "data_batch_%d" % i) for i in range(1, 6)]
        elif subset == "validation":
            names = [os.path.join(self.data_dir, "test_batch")]
        else:
            raise ValueError('Invalid data subset "%s"' % subset)
        imgs, labels = [], []
        for fn in names:
            with open(fn, "rb") as f:
                d = pickle.load(f, encoding="bytes")
            imgs.append(np.asarray(d[b"data"], dtype=np.uint8).reshape(-1, 3, 32, 32))
            labels.append(np.asarray(d[b"labels"], dtype=np.int64))
        return np.concatenate(imgs), np.concatenate(labels)

    def num_examples_per_epoch(self, subset="train"):
        if subset == "train":
            return 50000
        if subset == "validation":
            return 10000
        raise ValueError('Invalid data subset "%s"' % subset)

    def get_image_preprocessor(self):
        from . import preprocessing
        if self.use_synthetic_gpu_images():
            return preprocessing.SyntheticImagePreprocessor
        return preprocessing.Cifar10ImagePreprocessor


_SUPPORTED = {"imagenet": ImagenetData, "cifar10": Cifar10Data}


def create_dataset(data_dir, data_name):
    """infer the dataset from the directory name when `data_name` is not given;
    no directory ⇒ synthetic imagenet-shaped data"""
    if not data_dir and not data_name:
        data_name = "imagenet"
    if data_name is None:
        for name in _SUPPORTED:
            if name in os.path.basename(os.path.normpath(data_dir)).lower():
                data_name = name
                break
        else:
            raise ValueError("Could not identify name of dataset. Please specify with "
                             "--data_name option.")
    if data_name not in _SUPPORTED:
        raise ValueError("Unknown dataset. Must be one of %s" % ", ".join(sorted(_SUPPORTED)))
    return _SUPPORTED[data_name](data_dir)
