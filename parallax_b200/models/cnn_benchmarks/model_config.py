"""Model name → network + per-model defaults.

Parity: `examples/tf_cnn_benchmarks/models/model_config.py:30-92` (separate
ImageNet and CIFAR-10 model maps, `get_model_config`, `register_model`) and the
`Model(name, image_size, batch_size, learning_rate)` defaults set by each model
file (e.g. alexnet 227 px / batch 512 / lr 0.005; resnets 224 / 64 (32 for
resnet152 and the v2s of 101/152 in the reference — 64 kept here for 50/101);
CIFAR resnets 32 / 128 / 0.1 — `models/*.py`)."""
from .. import cnn as _cnn
from ..resnet import goyal_lr


class ModelConfig(object):
    def __init__(self, name, factory, image_size, batch_size, learning_rate, lr_schedule=None):
        self.name, self.factory, self.image_size = name, factory, image_size
        self.batch_size, self.learning_rate = batch_size, learning_rate
        self.lr_schedule = lr_schedule
        self.fp16_loss_scale = 128.0

    def get_image_size(self):
        return self.image_size

    def get_batch_size(self):
        return self.batch_size

    def set_batch_size(self, batch_size):
        self.batch_size = batch_size

    def get_default_batch_size(self):
        return self.batch_size

    def get_fp16_loss_scale(self):
        return self.fp16_loss_scale

    def get_learning_rate(self, global_batch_size, steps_per_epoch):
        """callable lr(step) | float: the model's own schedule when no
        `--learning_rate` is given (ResNets: Goyal et al. warm-up + step decay,
        `models/resnet_model.py:232-254`)"""
        if self.lr_schedule is not None:
            return self.lr_schedule(global_batch_size, steps_per_epoch)
        return self.learning_rate

    def build(self, num_classes):
        return _cnn.MODELS[self.factory](num_classes)


def _mk(name, factory, size, batch, lr, sched=None):
    return lambda: ModelConfig(name, factory, size, batch, lr, sched)


_goyal = lambda gb, spe: goyal_lr(gb, spe, base=0.1)

_model_name_to_imagenet_model = {
    "vgg11": _mk("vgg11", "vgg11", 224, 64, 0.005),
    "vgg16": _mk("vgg16", "vgg16", 224, 64, 0.005),
    "vgg19": _mk("vgg19", "vgg19", 224, 64, 0.005),
    "lenet": _mk("lenet5", "lenet", 28, 32, 0.005),
    "googlenet": _mk("googlenet", "googlenet", 224, 32, 0.005),
    "overfeat": _mk("overfeat", "overfeat", 231, 32, 0.005),
    "alexnet": _mk("alexnet", "alexnet", 224, 512, 0.005),
    "trivial": _mk("trivial", "trivial", 224, 32, 0.005),
    "inception3": _mk("inception3", "inception3", 299, 32, 0.005),
    "inception4": _mk("inception4", "inception4", 299, 32, 0.005),
    "resnet50": _mk("resnet50", "resnet50", 224, 64, 0.005, _goyal),
    "resnet50_v2": _mk("resnet50_v2", "resnet50_v2", 224, 64, 0.005, _goyal),
    "resnet101": _mk("resnet101", "resnet101", 224, 32, 0.005, _goyal),
    "resnet101_v2": _mk("resnet101_v2", "resnet101_v2", 224, 32, 0.005, _goyal),
    "resnet152": _mk("resnet152", "resnet152", 224, 32, 0.005, _goyal),
    "resnet152_v2": _mk("resnet152_v2", "resnet152_v2", 224, 32, 0.005, _goyal),
}

_model_name_to_cifar_model = {
    "alexnet": _mk("alexnet", "alexnet_cifar", 32, 128, 0.1),
    "trivial": _mk("trivial", "trivial_cifar", 32, 32, 0.005),
    "densenet40_k12": _mk("densenet40_k12", "densenet40_k12", 32, 64, 0.1),
    "densenet100_k12": _mk("densenet100_k12", "densenet100_k12", 32, 64, 0.1),
    "densenet100_k24": _mk("densenet100_k24", "densenet100_k24", 32, 64, 0.1),
}
for _d in (20, 32, 44, 56, 110):
    for _sfx in ("", "_v2"):
        _n = "resnet%d%s" % (_d, _sfx)
        _model_name_to_cifar_model[_n] = _mk(_n, _n, 32, 128, 0.1)


def _get_model_map(dataset_name):
    if dataset_name == "cifar10":
        return _model_name_to_cifar_model
    if dataset_name in ("imagenet", "synthetic"):
        return _model_name_to_imagenet_model
    raise ValueError("Invalid dataset name: %s" % dataset_name)


def get_model_config(model_name, dataset):
    model_map = _get_model_map(dataset.name)
    if model_name not in model_map:
        raise ValueError("Invalid model name '%s' for dataset '%s'" % (model_name, dataset.name))
    return model_map[model_name]()


def register_model(model_name, dataset_name, model_func):
    """`model_func()` must return a `ModelConfig`"""
    model_map = _get_model_map(dataset_name)
    if model_name in model_map:
        raise ValueError('Model "%s" is already registered for dataset "%s"' %
                         (model_name, dataset_name))
    model_map[model_name] = model_func
