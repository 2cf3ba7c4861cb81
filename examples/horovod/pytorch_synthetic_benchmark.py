"""Synthetic data-parallel benchmark on the Horovod-style API (the role of Horovod's
`examples/pytorch_synthetic_benchmark.py`): a torchvision-shaped CNN, a plain
``torch.optim`` optimizer wrapped in `hvd.DistributedOptimizer`, images/sec per worker and in
total.  Gradients are reduced by the fabric kernels (fused per dtype), not by NCCL.

    python -m parallax_b200.run -np 8 examples/horovod/pytorch_synthetic_benchmark.py \
        --model resnet50 --batch-size 64 --fp16-allreduce
"""
import argparse
import os
import sys
import timeit

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
import torch
import torch.nn.functional as F

from parallax_b200 import collectives as hvd
from parallax_b200.models import cnn

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="resnet50", choices=sorted(cnn.MODELS))
ap.add_argument("--batch-size", type=int, default=32)
ap.add_argument("--num-classes", type=int, default=1000)
ap.add_argument("--fp16-allreduce", action="store_true", help="compress gradients to fp16")
ap.add_argument("--num-warmup-batches", type=int, default=10)
ap.add_argument("--num-batches-per-iter", type=int, default=10)
ap.add_argument("--num-iters", type=int, default=10)
ap.add_argument("--no-cuda", action="store_true")
args = ap.parse_args()


def log(s):
    if hvd.rank() == 0:
        print(s, flush=True)


def main():
    hvd.init()
    cuda = not args.no_cuda and torch.cuda.is_available()
    device = torch.device("cuda", hvd.local_rank()) if cuda else torch.device("cpu")
    torch.manual_seed(0)
    model = cnn.get_model(args.model, args.num_classes).to(device)
    hw = cnn.image_size(model)
    opt = torch.optim.SGD(model.parameters(), lr=0.01 * hvd.size(), momentum=0.9)
    opt = hvd.DistributedOptimizer(
        opt, named_parameters=model.named_parameters(),
        compression=hvd.Compression.fp16 if args.fp16_allreduce else hvd.Compression.none)
    hvd.broadcast_parameters(model.state_dict(), root_rank=0)
    hvd.broadcast_optimizer_state(opt, root_rank=0)
    data = torch.randn(args.batch_size, 3, hw, hw, device=device)
    target = torch.randint(0, args.num_classes, (args.batch_size,), device=device)

    def step():
        opt.zero_grad()
        loss = F.cross_entropy(model(data, target)["logits"].float(), target)
        loss.backward()
        opt.step()
        if cuda:
            torch.cuda.synchronize()

    log("Model: %s  Batch size: %d  Number of %ss: %d" %
        (args.model, args.batch_size, "GPU" if cuda else "CPU", hvd.size()))
    timeit.timeit(step, number=args.num_warmup_batches)
    rates = []
    for i in range(args.num_iters):
        t = timeit.timeit(step, number=args.num_batches_per_iter)
        rates.append(args.batch_size * args.num_batches_per_iter / t)
        log("Iter #%d: %.1f img/sec per worker" % (i, rates[-1]))
    mean, conf = np.mean(rates), 1.96 * np.std(rates)
    log("Img/sec per worker: %.1f +-%.1f" % (mean, conf))
    log("Total img/sec on %d worker(s): %.1f +-%.1f" % (hvd.size(), hvd.size() * mean,
                                                        hvd.size() * conf))
    hvd.shutdown()


if __name__ == "__main__":
    main()
