"""ResNet v1 (50/101/152) — the reference's dense headline workload.

Parity: `parallax/parallax/examples/tf_cnn_benchmarks/models/resnet_model.py:192-258`
(ResNet-50 v1, bottleneck counts (3,4,6,3), batch 64/GPU, momentum SGD,
Goyal et al. LR schedule `:232-254`) driven by
`CNNBenchmark_distributed_driver.py:50-91` with synthetic images
(`benchmark_cnn.py:836-860`).  Model math stays on cuDNN/cuBLAS via PyTorch
(channels_last, bf16); the product here is the dense aggregation path.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import optim
from ..graph import Graph


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = downsample

    def forward(self, x):
        idt = x
        out = F.relu(self.bn1(self.conv1(x)), inplace=True)
        out = F.relu(self.bn2(self.conv2(out)), inplace=True)
        out = self.bn3(self.conv3(out))
        if self.downsample is not None:
            idt = self.downsample(x)
        return F.relu(out + idt, inplace=True)


class PreActBottleneck(nn.Module):
    """ResNet v2 (pre-activation) bottleneck (`resnet_model.py` bottleneck_block_v2)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.bn1 = nn.BatchNorm2d(inplanes)
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.downsample = downsample

    def forward(self, x):
        pre = F.relu(self.bn1(x))
        idt = x if self.downsample is None else self.downsample(pre)
        out = self.conv1(pre)
        out = self.conv2(F.relu(self.bn2(out)))
        out = self.conv3(F.relu(self.bn3(out)))
        return out + idt


class ResNet(nn.Module):
    def __init__(self, layers, num_classes=1000, v2=False):
        super().__init__()
        self.v2 = v2
        self.block = PreActBottleneck if v2 else Bottleneck
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.layer1 = self._make(64, layers[0], 1)
        self.layer2 = self._make(128, layers[1], 2)
        self.layer3 = self._make(256, layers[2], 2)
        self.layer4 = self._make(512, layers[3], 2)
        self.fc = nn.Linear(2048, num_classes)
        self.final_bn = nn.BatchNorm2d(2048) if v2 else None
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        for m in self.modules():
            if isinstance(m, Bottleneck):
                nn.init.zeros_(m.bn3.weight)

    def _make(self, planes, blocks, stride):
        down = None
        if stride != 1 or self.inplanes != planes * 4:
            conv = nn.Conv2d(self.inplanes, planes * 4, 1, stride, bias=False)
            down = conv if self.v2 else nn.Sequential(conv, nn.BatchNorm2d(planes * 4))
        layers = [self.block(self.inplanes, planes, stride, down)]
        self.inplanes = planes * 4
        layers += [self.block(self.inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, images, labels):
        x = images.to(self.conv1.weight.dtype).contiguous(memory_format=torch.channels_last)
        x = F.relu(self.bn1(self.conv1(x)), inplace=True)
        x = F.max_pool2d(x, 3, 2, 1)
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        if self.final_bn is not None:
            x = F.relu(self.final_bn(x))
        x = torch.flatten(F.adaptive_avg_pool2d(x, 1), 1)
        logits = self.fc(x)
        return {"loss": F.cross_entropy(logits.float(), labels), "logits": logits}


def resnet50(num_classes=1000):
    return ResNet((3, 4, 6, 3), num_classes).to(memory_format=torch.channels_last)


def resnet101(num_classes=1000):
    return ResNet((3, 4, 23, 3), num_classes).to(memory_format=torch.channels_last)


def resnet152(num_classes=1000):
    return ResNet((3, 8, 36, 3), num_classes).to(memory_format=torch.channels_last)


def resnet_v2(depth=50, num_classes=1000):
    cfg = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}[depth]
    return ResNet(cfg, num_classes, v2=True).to(memory_format=torch.channels_last)


def goyal_lr(batch_size_global, steps_per_epoch, base=0.1):
    """Linear warm-up for 5 epochs to base·(batch/256), ÷10 at epochs 30/60/80
    (`resnet_model.py:232-254`)."""
    peak = base * batch_size_global / 256.0

    def lr(step):
        e = step / float(max(steps_per_epoch, 1))
        if e < 5:
            return peak * (0.1 + 0.9 * e / 5.0) if batch_size_global > 256 else peak
        return peak * (1.0 if e < 30 else 0.1 if e < 60 else 0.01 if e < 80 else 0.001)
    return lr


def resnet_graph(model, learning_rate=0.1, momentum=0.9, weight_decay=1e-4):
    return Graph(model, optimizer=optim.Momentum(learning_rate, momentum,
                                                 weight_decay=weight_decay),
                 loss="loss", name="resnet")
