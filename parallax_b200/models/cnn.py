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
from .resnet import resnet50, resnet101, resnet152, resnet_v2


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


# ---------------------------------------------------------------------------
# Inception v3 / v4-lite, CIFAR ResNet and DenseNet (model_config.py:30-64)
# ---------------------------------------------------------------------------
def _cbr(cin, cout, k, s=1, p=0):
    return nn.Sequential(nn.Conv2d(cin, cout, k, s, p, bias=False), nn.BatchNorm2d(cout),
                         nn.ReLU(inplace=True))


class _Branches(nn.Module):
    def __init__(self, *branches):
        super().__init__()
        self.branches = nn.ModuleList(branches)

    def forward(self, x):
        return torch.cat([b(x) for b in self.branches], 1)


def _inc_a(cin, pool):
    return _Branches(_cbr(cin, 64, 1),
                     nn.Sequential(_cbr(cin, 48, 1), _cbr(48, 64, 5, p=2)),
                     nn.Sequential(_cbr(cin, 64, 1), _cbr(64, 96, 3, p=1), _cbr(96, 96, 3, p=1)),
                     nn.Sequential(nn.AvgPool2d(3, 1, 1), _cbr(cin, pool, 1)))


def _inc_b(cin):      # grid reduction 35 -> 17
    return _Branches(_cbr(cin, 384, 3, 2),
                     nn.Sequential(_cbr(cin, 64, 1), _cbr(64, 96, 3, p=1), _cbr(96, 96, 3, 2)),
                     nn.MaxPool2d(3, 2))


def _inc_c(cin, c7):
    return _Branches(
        _cbr(cin, 192, 1),
        nn.Sequential(_cbr(cin, c7, 1), _cbr(c7, c7, (1, 7), p=(0, 3)),
                      _cbr(c7, 192, (7, 1), p=(3, 0))),
        nn.Sequential(_cbr(cin, c7, 1), _cbr(c7, c7, (7, 1), p=(3, 0)),
                      _cbr(c7, c7, (1, 7), p=(0, 3)), _cbr(c7, c7, (7, 1), p=(3, 0)),
                      _cbr(c7, 192, (1, 7), p=(0, 3))),
        nn.Sequential(nn.AvgPool2d(3, 1, 1), _cbr(cin, 192, 1)))


def _inc_d(cin):      # grid reduction 17 -> 8
    return _Branches(nn.Sequential(_cbr(cin, 192, 1), _cbr(192, 320, 3, 2)),
                     nn.Sequential(_cbr(cin, 192, 1), _cbr(192, 192, (1, 7), p=(0, 3)),
                                   _cbr(192, 192, (7, 1), p=(3, 0)), _cbr(192, 192, 3, 2)),
                     nn.MaxPool2d(3, 2))


class _IncE(nn.Module):
    def __init__(self, cin):
        super().__init__()
        self.b1 = _cbr(cin, 320, 1)
        self.b3 = _cbr(cin, 384, 1)
        self.b3a, self.b3b = _cbr(384, 384, (1, 3), p=(0, 1)), _cbr(384, 384, (3, 1), p=(1, 0))
        self.bd = nn.Sequential(_cbr(cin, 448, 1), _cbr(448, 384, 3, p=1))
        self.bda, self.bdb = _cbr(384, 384, (1, 3), p=(0, 1)), _cbr(384, 384, (3, 1), p=(1, 0))
        self.bp = nn.Sequential(nn.AvgPool2d(3, 1, 1), _cbr(cin, 192, 1))

    def forward(self, x):
        b3, bd = self.b3(x), self.bd(x)
        return torch.cat([self.b1(x), self.b3a(b3), self.b3b(b3), self.bda(bd), self.bdb(bd),
                          self.bp(x)], 1)


class Inception3(_Classifier):
    image_size = 299

    def __init__(self, num_classes=1000, depth_e=2):
        super().__init__()
        self.net = nn.Sequential(
            _cbr(3, 32, 3, 2), _cbr(32, 32, 3), _cbr(32, 64, 3, p=1), nn.MaxPool2d(3, 2),
            _cbr(64, 80, 1), _cbr(80, 192, 3), nn.MaxPool2d(3, 2),
            _inc_a(192, 32), _inc_a(256, 64), _inc_a(288, 64), _inc_b(288),
            _inc_c(768, 128), _inc_c(768, 160), _inc_c(768, 160), _inc_c(768, 192), _inc_d(768),
            *[_IncE(1280 if i == 0 else 2048) for i in range(depth_e)],
            nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Dropout(0.2), nn.Linear(2048, num_classes))


class Inception4(Inception3):
    """Inception-v4-sized variant: the v3 stem/blocks with the v4 block counts
    (4×A, 7×C, 3×E) — same kernels and gradient volume class as
    `models/inception_model.py` Inceptionv4Model."""

    def __init__(self, num_classes=1000):
        _Classifier.__init__(self)
        self.net = nn.Sequential(
            _cbr(3, 32, 3, 2), _cbr(32, 32, 3), _cbr(32, 64, 3, p=1), nn.MaxPool2d(3, 2),
            _cbr(64, 80, 1), _cbr(80, 192, 3), nn.MaxPool2d(3, 2),
            _inc_a(192, 32), _inc_a(256, 64), _inc_a(288, 64), _inc_a(288, 64), _inc_b(288),
            *[_inc_c(768, c) for c in (128, 160, 160, 160, 160, 160, 192)], _inc_d(768),
            _IncE(1280), _IncE(2048), _IncE(2048),
            nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Dropout(0.2), nn.Linear(2048, num_classes))


class _CifarBlock(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.c1, self.b1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False), nn.BatchNorm2d(cout)
        self.c2, self.b2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False), nn.BatchNorm2d(cout)
        self.short = None if stride == 1 and cin == cout else \
            nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        y = self.b2(self.c2(F.relu(self.b1(self.c1(x)))))
        return F.relu(y + (x if self.short is None else self.short(x)))


class _CifarPreActBlock(nn.Module):
    """v2 (pre-activation) basic block: BN-ReLU-conv ×2, identity added unactivated"""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.b1, self.c1 = nn.BatchNorm2d(cin), nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.b2, self.c2 = nn.BatchNorm2d(cout), nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.short = None if stride == 1 and cin == cout else \
            nn.Conv2d(cin, cout, 1, stride, bias=False)

    def forward(self, x):
        pre = F.relu(self.b1(x))
        y = self.c2(F.relu(self.b2(self.c1(pre))))
        return y + (x if self.short is None else self.short(pre))


class CifarResNet(_Classifier):
    """resnet20/32/44/56/110 (+ `_v2` pre-activation variants) for CIFAR-10
    (`models/resnet_model.py:283-371`)."""
    image_size = 32

    def __init__(self, depth=20, num_classes=10, v2=False):
        super().__init__()
        n = (depth - 2) // 6
        block = _CifarPreActBlock if v2 else _CifarBlock
        layers = [nn.Conv2d(3, 16, 3, 1, 1, bias=False)]
        if not v2:
            layers += [nn.BatchNorm2d(16), nn.ReLU()]
        cin = 16
        for cout, stride in ((16, 1), (32, 2), (64, 2)):
            for i in range(n):
                layers.append(block(cin, cout, stride if i == 0 else 1))
                cin = cout
        if v2:
            layers += [nn.BatchNorm2d(64), nn.ReLU()]
        self.net = nn.Sequential(*layers, nn.AdaptiveAvgPool2d(1), nn.Flatten(),
                                 nn.Linear(64, num_classes))


class AlexNetCifar(_Classifier):
    """`models/alexnet_model.py:56-86` AlexnetCifar10Model: two 5×5 conv + LRN + pool
    stages, fully connected 384 → 192 → classes"""
    image_size = 32

    def __init__(self, num_classes=10):
        super().__init__()
        self.net = nn.Sequential(
            nn.Conv2d(3, 64, 5, padding=2), nn.ReLU(), nn.MaxPool2d(3, 2, 1),
            nn.LocalResponseNorm(9, alpha=0.001, beta=0.75, k=1.0),
            nn.Conv2d(64, 64, 5, padding=2), nn.ReLU(),
            nn.LocalResponseNorm(9, alpha=0.001, beta=0.75, k=1.0), nn.MaxPool2d(3, 2, 1),
            nn.Flatten(), nn.Linear(64 * 8 * 8, 384), nn.ReLU(), nn.Linear(384, 192), nn.ReLU(),
            nn.Linear(192, num_classes))


class TrivialCifar(_Classifier):
    """`models/trivial_model.py:33-45`"""
    image_size = 32

    def __init__(self, num_classes=10):
        super().__init__()
        self.net = nn.Sequential(nn.Flatten(), nn.Linear(3 * 32 * 32, 1), nn.ReLU(),
                                 nn.Linear(1, 4096), nn.ReLU(), nn.Linear(4096, num_classes))


class _DenseLayer(nn.Module):
    def __init__(self, cin, k):
        super().__init__()
        self.bn, self.conv = nn.BatchNorm2d(cin), nn.Conv2d(cin, k, 3, 1, 1, bias=False)

    def forward(self, x):
        return torch.cat([x, self.conv(F.relu(self.bn(x)))], 1)


class CifarDenseNet(_Classifier):
    """densenet40_k12 / densenet100_k12 / densenet100_k24 (`models/densenet_model.py`)."""
    image_size = 32

    def __init__(self, depth=40, k=12, num_classes=10):
        super().__init__()
        n = (depth - 4) // 3
        layers, c = [nn.Conv2d(3, 16, 3, 1, 1, bias=False)], 16
        for stage in range(3):
            for _ in range(n):
                layers.append(_DenseLayer(c, k))
                c += k
            if stage < 2:
                layers += [nn.BatchNorm2d(c), nn.ReLU(), nn.Conv2d(c, c, 1, bias=False),
                           nn.AvgPool2d(2)]
        self.net = nn.Sequential(*layers, nn.BatchNorm2d(c), nn.ReLU(), nn.AdaptiveAvgPool2d(1),
                                 nn.Flatten(), nn.Linear(c, num_classes))


MODELS.update({
    "resnet50_v2": lambda n=1000: resnet_v2(50, n),
    "resnet101_v2": lambda n=1000: resnet_v2(101, n),
    "resnet152_v2": lambda n=1000: resnet_v2(152, n),
    "inception3": Inception3, "inception4": Inception4,
    "resnet20": lambda n=10: CifarResNet(20, n), "resnet32": lambda n=10: CifarResNet(32, n),
    "resnet44": lambda n=10: CifarResNet(44, n), "resnet56": lambda n=10: CifarResNet(56, n),
    "resnet110": lambda n=10: CifarResNet(110, n),
    "resnet20_v2": lambda n=10: CifarResNet(20, n, True),
    "resnet32_v2": lambda n=10: CifarResNet(32, n, True),
    "resnet44_v2": lambda n=10: CifarResNet(44, n, True),
    "resnet56_v2": lambda n=10: CifarResNet(56, n, True),
    "resnet110_v2": lambda n=10: CifarResNet(110, n, True),
    "alexnet_cifar": AlexNetCifar, "trivial_cifar": TrivialCifar,
    "densenet40_k12": lambda n=10: CifarDenseNet(40, 12, n),
    "densenet100_k12": lambda n=10: CifarDenseNet(100, 12, n),
    "densenet100_k24": lambda n=10: CifarDenseNet(100, 24, n),
})
