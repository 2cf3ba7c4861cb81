"""MNIST-style training on the Horovod-style API with the Keras-like callbacks (the role of
Horovod's `examples/pytorch_mnist.py` / `keras_mnist_advanced.py`): data sharded by rank,
parameters and optimizer state broadcast from rank 0, learning-rate warm-up over the first epochs,
metrics averaged over the workers at the end of each epoch.  Uses torchvision's MNIST when
`--data-dir` has it, otherwise a synthetic digit-like dataset (no network in the sandbox).

    python -m parallax_b200.run -np 2 examples/horovod/pytorch_mnist.py --epochs 3
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import torch.nn as nn
import torch.nn.functional as F

from parallax_b200 import callbacks as cbs
from parallax_b200 import collectives as hvd

ap = argparse.ArgumentParser()
ap.add_argument("--batch-size", type=int, default=64)
ap.add_argument("--epochs", type=int, default=3)
ap.add_argument("--lr", type=float, default=0.01)
ap.add_argument("--momentum", type=float, default=0.5)
ap.add_argument("--warmup-epochs", type=int, default=1)
ap.add_argument("--data-dir", default=None)
ap.add_argument("--num-synthetic", type=int, default=2048)
ap.add_argument("--no-cuda", action="store_true")
args = ap.parse_args()


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1, self.conv2 = nn.Conv2d(1, 10, 5), nn.Conv2d(10, 20, 5)
        self.fc1, self.fc2 = nn.Linear(320, 50), nn.Linear(50, 10)

    def forward(self, x):
        x = F.relu(F.max_pool2d(self.conv1(x), 2))
        x = F.relu(F.max_pool2d(self.conv2(x), 2))
        return self.fc2(F.relu(self.fc1(x.flatten(1))))


def dataset():
    if args.data_dir:
        from torchvision import datasets, transforms
        ds = datasets.MNIST(args.data_dir, train=True, download=False,
                            transform=transforms.ToTensor())
        x = torch.stack([ds[i][0] for i in range(len(ds))])
        y = torch.tensor([ds[i][1] for i in range(len(ds))])
        return x, y
    g = torch.Generator().manual_seed(0)            # class k = a bright 8×8 patch at position k
    y = torch.randint(0, 10, (args.num_synthetic,), generator=g)
    x = 0.1 * torch.rand(args.num_synthetic, 1, 28, 28, generator=g)
    for i, k in enumerate(y.tolist()):
        r, c = divmod(k, 4)
        x[i, 0, 2 + 8 * r:10 + 8 * r, 2 + 6 * c:10 + 6 * c] += 0.9
    return x, y


def main():
    hvd.init()
    cuda = not args.no_cuda and torch.cuda.is_available()
    device = torch.device("cuda", hvd.local_rank()) if cuda else torch.device("cpu")
    torch.manual_seed(42 + hvd.rank())              # different init per rank: broadcast fixes it
    model = Net().to(device)
    x, y = dataset()
    x, y = x[hvd.rank()::hvd.size()].to(device), y[hvd.rank()::hvd.size()].to(device)
    steps = len(x) // args.batch_size
    opt = hvd.DistributedOptimizer(
        torch.optim.SGD(model.parameters(), lr=args.lr * hvd.size(), momentum=args.momentum),
        named_parameters=model.named_parameters())
    cl = cbs.CallbackList([cbs.BroadcastGlobalVariablesCallback(0), cbs.MetricAverageCallback(),
                           cbs.LearningRateWarmupCallback(args.warmup_epochs,
                                                          steps_per_epoch=steps, verbose=1)],
                          model=model, optimizer=opt)
    cl.on_train_begin()
    for epoch in range(args.epochs):
        cl.on_epoch_begin(epoch)
        perm = torch.randperm(len(x), device=device)
        tot_loss = tot_acc = 0.0
        for b in range(steps):
            cl.on_batch_begin(b)
            idx = perm[b * args.batch_size:(b + 1) * args.batch_size]
            opt.zero_grad()
            out = model(x[idx])
            loss = F.cross_entropy(out, y[idx])
            loss.backward()
            opt.step()
            tot_loss += loss.item()
            tot_acc += (out.argmax(1) == y[idx]).float().mean().item()
            cl.on_batch_end(b)
        logs = {"loss": tot_loss / steps, "accuracy": tot_acc / steps}
        cl.on_epoch_end(epoch, logs)
        if hvd.rank() == 0:
            print("Epoch %d: loss %.4f  accuracy %.3f  lr %.4f" %
                  (epoch + 1, logs["loss"], logs["accuracy"], logs["lr"]), flush=True)
    hvd.shutdown()


if __name__ == "__main__":
    main()
