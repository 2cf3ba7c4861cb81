"""CNN model zoo of the reference's `tf_cnn_benchmarks` example
(`examples/tf_cnn_benchmarks/models/model_config.py:30-64`): trivial, lenet,
alexnet, overfeat, vgg11/16/19, googlenet, resnet50/101/152 (resnet in
`models/resnet.py`).  Architectures follow the reference's builders
(`models/{alexnet,vgg,lenet,googlenet,overfeat,trivial}_model.py`); the math
runs on cuDNN/cuBLAS through PyTorch — the product under test is the dense
gradient path."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import optim
from ..graph import Graph
from .resnet import resnet50, resnet101, resnet152


class _Classifier(nn.Module):
    image_size = 224

    def forward(self, images, labels):
        x = images.to(next(self.parameters()).dtype)
        logits = self.net(x)
        return {"loss": F.cross_entropy(logits.float(), labels), "logits": logits}


class Trivial(_Classifier):
    def __init__(self, num_classes=1000):
        super().__init__()
        self.net = nn.Sequential(nn.Flatten(), nn.Linear(3 * 224 * 224, 1), nn.ReLU(),
                                 nn.Linear(1, 4096), nn.ReLU(), nn.Linear(4096, num_classes))


class LeNet(_Classifier):
    image_size = 28

    def __init__(self, num_classes=1000):
        super().__init__()
        self.net = nn.Sequential(
            nn.Conv2d(3, 32, 5, padding=2), nn.ReLU(), nn.MaxPool2d(2),
            nn.Conv2d(32, 64, 5, padding=2), nn.ReLU(), nn.MaxPool2d(2), nn.Flatten(),
            nn.Linear(64 * 7 * 7, 512), nn.ReLU(), nn.Linear(512, num_classes))


class AlexNet(_Classifier):
    def __init__(self, num_classes=1000):
        super().__init__()
        self.net = nn.Sequential(
            nn.Conv2d(3, 64, 11, 4, 2), nn.ReLU(), nn.MaxPool2d(3, 2),
            nn.Conv2d(64, 192, 5, padding=2), nn.ReLU(), nn.MaxPool2d(3, 2),
            nn.Conv2d(192, 384, 3, padding=1), nn.ReLU(),
            nn.Conv2d(384, 384, 3, padding=1), nn.ReLU(),
            nn.Conv2d(384, 256, 3, padding=1), nn.ReLU(), nn.MaxPool2d(3, 2), nn.Flatten(),
            nn.Linear(256 * 6 * 6, 4096), nn.ReLU(), nn.Dropout(0.5),
            nn.Linear(4096, 4096), nn.ReLU(), nn.Dropout(0.5), nn.Linear(4096, num_classes))


class Overfeat(_Classifier):
    image_size = 231

    def __init__(self, num_classes=1000):
        super().__init__()
        self.net = nn.Sequential(
            nn.Conv2d(3, 96, 11, 4), nn.ReLU(), nn.MaxPool2d(2),
            nn.Conv2d(96, 256, 5), nn.ReLU(), nn.MaxPool2d(2),
            nn.Conv2d(256, 512, 3, padding=1), nn.ReLU(),
            nn.Conv2d(512, 1024, 3, padding=1), nn.ReLU(),
            nn.Conv2d(1024, 1024, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2), nn.Flatten(),
            nn.Linear(1024 * 6 * 6, 3072), nn.ReLU(), nn.Linear(3072, 4096), nn.ReLU(),
            nn.Linear(4096, num_classes))


def _vgg_features(cfg):
    layers, c = [], 3
    for v in cfg:
        if v == "M":
            layers.append(nn.MaxPool2d(2))
        else:
            layers += [nn.Conv2d(c, v, 3, padding=1), nn.ReLU()]
            c = v
    return layers


_VGG = {11: [64, "M", 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"],
        16: [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M",
             512, 512, 512, "M"],
        19: [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M",
             512, 512, 512, 512, "M"]}


class VGG(_Classifier):
    def __init__(self, depth=16, num_classes=1000):
        super().__init__()
        self.net = nn.Sequential(*_vgg_features(_VGG[depth]), nn.Flatten(),
                                 nn.Linear(512 * 7 * 7, 4096), nn.ReLU(), nn.Dropout(0.5),
                                 nn.Linear(4096, 4096), nn.ReLU(), nn.Dropout(0.5),
                                 nn.Linear(4096, num_classes))


class _Inception(nn.Module):
    def __init__(self, cin, c1, c3r, c3, c5r, c5, cp):
        super().__init__()
        self.b1 = nn.Sequential(nn.Conv2d(cin, c1, 1), nn.ReLU())
        self.b2 = nn.Sequential(nn.Conv2d(cin, c3r, 1), nn.ReLU(),
                                nn.Conv2d(c3r, c3, 3, padding=1), nn.ReLU())
        self.b3 = nn.Sequential(nn.Conv2d(cin, c5r, 1), nn.ReLU(),
                                nn.Conv2d(c5r, c5, 5, padding=2), nn.ReLU())
        self.b4 = nn.Sequential(nn.MaxPool2d(3, 1, 1), nn.Conv2d(cin, cp, 1), nn.ReLU())

    def forward(self, x):
        return torch.cat([self.b1(x), self.b2(x), self.b3(x), self.b4(x)], 1)


class GoogLeNet(_Classifier):
    def __init__(self, num_classes=1000):
        super().__init__()
        self.net = nn.Sequential(
            nn.Conv2d(3, 64, 7, 2, 3), nn.ReLU(), nn.MaxPool2d(3, 2, 1),
            nn.Conv2d(64, 64, 1), nn.ReLU(), nn.Conv2d(64, 192, 3, padding=1), nn.ReLU(),
            nn.MaxPool2d(3, 2, 1),
            _Inception(192, 64, 96, 128, 16, 32, 32), _Inception(256, 128, 128, 192, 32, 96, 64),
            nn.MaxPool2d(3, 2, 1),
            _Inception(480, 192, 96, 208, 16, 48, 64), _Inception(512, 160, 112, 224, 24, 64, 64),
            _Inception(512, 128, 128, 256, 24, 64, 64), _Inception(512, 112, 144, 288, 32, 64, 64),
            _Inception(528, 256, 160, 320, 32, 128, 128), nn.MaxPool2d(3, 2, 1),
            _Inception(832, 256, 160, 320, 32, 128, 128),
            _Inception(832, 384, 192, 384, 48, 128, 128),
            nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(1024, num_classes))


MODELS = {
    "trivial": Trivial, "lenet": LeNet, "alexnet": AlexNet, "overfeat": Overfeat,
    "vgg11": lambda n=1000: VGG(11, n), "vgg16": lambda n=1000: VGG(16, n),
    "vgg19": lambda n=1000: VGG(19, n), "googlenet": GoogLeNet,
    "resnet50": resnet50, "resnet101": resnet101, "resnet152": resnet152,
}


def get_model(name, num_classes=1000):
    if name not in MODELS:
        raise ValueError("unknown model %r (have %s)" % (name, sorted(MODELS)))
    return MODELS[name](num_classes)


def image_size(model):
    return getattr(model, "image_size", 224)


def cnn_graph(model, optimizer="momentum", learning_rate=0.01, momentum=0.9,
              weight_decay=4e-5):
    """Optimizer choice of `benchmark_cnn.py:805-818` (momentum | sgd | rmsprop)."""
    opt = {"momentum": lambda: optim.Momentum(learning_rate, momentum, weight_decay=weight_decay),
           "sgd": lambda: optim.GradientDescent(learning_rate, weight_decay=weight_decay),
           "rmsprop": lambda: optim.RMSProp(learning_rate, 0.9, momentum, 1.0,
                                            weight_decay=weight_decay)}[optimizer]()
    return Graph(model, optimizer=opt, loss="loss", name="cnn")
