"""Skip-gram word2vec with negative sampling on the Horovod-style API (the role of Horovod's
`examples/tensorflow_word2vec.py`): the embedding tables are `nn.Embedding(sparse=True)`, so each
worker's gradient is a row-sparse tensor and `DistributedOptimizer` exchanges it as an all-gather of
(indices, values) — Horovod's IndexedSlices path (`horovod/tensorflow/__init__.py:62-82`) — or,
with `--sparse-as-dense`, as a dense all-reduce.  The corpus is synthetic (no network in the
sandbox): a Markov chain over the vocabulary in which word w is followed by one of a few fixed
"neighbours", so the learned vectors of neighbours end up close.

    python -m parallax_b200.run -np 2 examples/horovod/pytorch_word2vec.py --steps 300
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import torch.nn as nn
import torch.nn.functional as F

from parallax_b200 import collectives as hvd

ap = argparse.ArgumentParser()
ap.add_argument("--vocab", type=int, default=512)
ap.add_argument("--dim", type=int, default=32)
ap.add_argument("--batch-size", type=int, default=256)
ap.add_argument("--negatives", type=int, default=8)
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--lr", type=float, default=0.2)
ap.add_argument("--corpus-len", type=int, default=40000)
ap.add_argument("--sparse-as-dense", action="store_true")
ap.add_argument("--no-cuda", action="store_true")
args = ap.parse_args()


def corpus(V, n, seed=0):
    """word w is followed by (w*7 + k) % V for a random k in 0..3"""
    g = torch.Generator().manual_seed(seed)
    w = torch.empty(n, dtype=torch.int64)
    w[0] = 0
    k = torch.randint(0, 4, (n,), generator=g)
    restart = torch.rand(n, generator=g) < 0.05
    jump = torch.randint(0, V, (n,), generator=g)
    for i in range(1, n):
        w[i] = jump[i] if restart[i] else (w[i - 1] * 7 + k[i]) % V
    return w


class SkipGram(nn.Module):
    def __init__(self, V, D):
        super().__init__()
        self.inp = nn.Embedding(V, D, sparse=True)
        self.out = nn.Embedding(V, D, sparse=True)
        nn.init.uniform_(self.inp.weight, -0.5, 0.5)
        nn.init.zeros_(self.out.weight)

    def forward(self, center, context, negatives):
        v = self.inp(center)                                   # [B, D]
        pos = (v * self.out(context)).sum(-1)                  # [B]
        neg = torch.bmm(self.out(negatives), v.unsqueeze(2)).squeeze(2)   # [B, K]
        # per-example loss, summed: a row that occurs once in the batch moves by lr x its own
        # gradient, as in word2vec's per-pair SGD (a batch mean would shrink it by 1/B)
        return -(F.logsigmoid(pos).sum() + F.logsigmoid(-neg).sum())


def main():
    hvd.init()
    cuda = torch.cuda.is_available() and not args.no_cuda
    dev = torch.device("cuda", hvd.local_rank()) if cuda else torch.device("cpu")
    torch.manual_seed(1234)
    model = SkipGram(args.vocab, args.dim).to(dev)
    opt = torch.optim.SGD(model.parameters(), lr=args.lr)
    opt = hvd.DistributedOptimizer(opt, named_parameters=model.named_parameters(),
                                   sparse_as_dense=args.sparse_as_dense)
    hvd.broadcast_parameters(model.state_dict(), root_rank=0)
    words = corpus(args.vocab, args.corpus_len)
    # every worker reads its own shard of (center, next word) pairs
    pairs = torch.stack([words[:-1], words[1:]], 1)[hvd.rank()::hvd.size()]
    g = torch.Generator().manual_seed(100 + hvd.rank())
    first = last = None
    for step in range(args.steps):
        idx = torch.randint(0, pairs.shape[0], (args.batch_size,), generator=g)
        center, context = pairs[idx, 0].to(dev), pairs[idx, 1].to(dev)
        negatives = torch.randint(0, args.vocab, (args.batch_size, args.negatives),
                                  generator=g).to(dev)
        opt.zero_grad()
        loss = model(center, context, negatives)
        loss.backward()
        assert model.inp.weight.grad.is_sparse
        opt.step()
        avg = float(hvd.allreduce(loss.detach() / args.batch_size, average=True, name="loss"))
        first = avg if first is None else first
        last = avg
        if hvd.rank() == 0 and (step % 50 == 0 or step == args.steps - 1):
            print("step %d: loss %.4f" % (step, avg), flush=True)
    # the replicas stayed identical (same averaged update everywhere)
    w = model.inp.weight.detach()
    spread = float((hvd.allreduce(w, average=True, name="w_mean") - w).abs().max())
    # neighbours of a word score higher than random words
    with torch.no_grad():
        c = torch.arange(0, args.vocab, device=dev)
        nb = (c * 7 + 1) % args.vocab
        rnd = (c * 13 + 101) % args.vocab
        s_nb = (model.inp(c) * model.out(nb)).sum(-1).mean()
        s_rnd = (model.inp(c) * model.out(rnd)).sum(-1).mean()
    if hvd.rank() == 0:
        print("loss %.4f -> %.4f; neighbour score %.3f vs random %.3f; replica spread %.2e; "
              "sparse gradients exchanged as %s" %
              (first, last, float(s_nb), float(s_rnd), spread,
               "dense all-reduce" if args.sparse_as_dense else "all-gather of (indices, values)"),
              flush=True)
    hvd.shutdown()


if __name__ == "__main__":
    main()
