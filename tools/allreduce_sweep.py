"""Dense all-reduce bandwidth sweep 1 KB – 1 GB: our P2P kernels vs NCCL
(BASELINE.json config 5).  Run under torchrun:

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port 29540 tools/allreduce_sweep.py --out gpurun_out/allreduce_sweep_8.json

Per size: device time (CUDA events, max over ranks) of 20 back-to-back calls;
algbw = bytes/time, busbw = algbw·2(N-1)/N.  Our path is zero-copy on symmetric
tensors: one-shot below `--oneshot-max`, two-shot in place above.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from parallax_b200 import collectives as hvd
from parallax_b200.parallel import nvops
from parallax_b200.parallel.symmetric import CH_USER

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
ap.add_argument("--max-bytes", type=int, default=1 << 30)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--blocks", type=int, default=32)
ap.add_argument("--oneshot-max", type=int, default=256 << 10)
args = ap.parse_args()

hvd.init(oneshot_bytes=args.oneshot_max)
st = hvd._st()
comm, W, rank, dev = st.comm, st.comm.world, st.comm.rank, st.comm.device
heap = st.fabric.heap


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t) * 1e-3


from parallax_b200.parallel import multicast
mcbuf = None
try:
    if multicast.supported(comm):
        mcbuf = multicast.MulticastBuffer(st.fabric, args.max_bytes)
        if rank == 0:
            print("NVLS multicast buffer ready (%d MiB)" % (mcbuf.nbytes >> 20), flush=True)
    elif rank == 0:
        print("NVLS multicast not supported on this box", flush=True)
except Exception as e:
    mcbuf = None
    if rank == 0:
        print("NVLS setup failed:", e, flush=True)
rows = []
size = 1024
buf = hvd.symmetric_empty(args.max_bytes // 2, torch.bfloat16)     # one symmetric segment
sbuf, npad_total = st.user_bufs[buf.data_ptr()]
while size <= args.max_bytes:
    n = size // 2
    x = buf[:n]
    x.normal_()
    y = torch.randn(n, device=dev).bfloat16()
    q = W * 8
    npad = (n + q - 1) // q * q
    t_nccl = timed(lambda: dist.all_reduce(y), args.iters)
    if size <= args.oneshot_max:
        out = torch.empty_like(x)
        t_ours = timed(lambda: nvops.allreduce_oneshot(heap, x, out, st.stage, n, torch.bfloat16,
                                                       1.0, CH_USER[0], max_blocks=8), args.iters)
        algo = "oneshot"
    else:
        t_ours = timed(lambda: nvops.allreduce_twoshot(heap, sbuf.c_ptrs(), npad, torch.bfloat16,
                                                       1.0, CH_USER, max_blocks=args.blocks),
                       args.iters)
        algo = "twoshot"
    t_bulk = None
    if size > args.oneshot_max:
        t_bulk = timed(lambda: nvops.allreduce_twoshot_bulk(heap, sbuf.c_ptrs(), npad,
                                                            torch.bfloat16, 1.0, CH_USER,
                                                            max_blocks=args.blocks), args.iters)
    t_nvls = None
    if mcbuf is not None and size >= 65536:
        mx = mcbuf.tensor(torch.bfloat16, npad)
        mx.normal_()
        if size == 65536:      # correctness check against NCCL once
            ref = mx.clone().float()
            dist.all_reduce(ref)
            torch.cuda.synchronize(); dist.barrier()
            mcbuf.allreduce_(npad, torch.bfloat16, 1.0, CH_USER, max_blocks=args.blocks)
            torch.cuda.synchronize()
            err = float((mx.float() - ref).abs().max())
            if rank == 0:
                print("NVLS correctness: max |err| = %.4f (ref max %.2f)" % (err, float(ref.abs().max())), flush=True)
        t_nvls = timed(lambda: mcbuf.allreduce_(npad, torch.bfloat16, 1.0, CH_USER,
                                                max_blocks=args.blocks), args.iters)
    f = 2.0 * (W - 1) / W
    row = {"bytes": size, "algo": algo, "ours_us": t_ours * 1e6, "nccl_us": t_nccl * 1e6,
           "nvls_us": None if t_nvls is None else t_nvls * 1e6,
           "tma_bulk_us": None if t_bulk is None else t_bulk * 1e6,
           "tma_bulk_busbw_GBs": None if t_bulk is None else size / t_bulk * f / 1e9,
           "nvls_busbw_GBs": None if t_nvls is None else size / t_nvls * f / 1e9,
           "ours_busbw_GBs": size / t_ours * f / 1e9, "nccl_busbw_GBs": size / t_nccl * f / 1e9,
           "speedup": t_nccl / t_ours, "n_gpus": W}
    rows.append(row)
    if rank == 0:
        print("%11d B %8s  ours %9.1f us (%7.1f GB/s bus)   nccl %9.1f us (%7.1f GB/s bus)  x%.2f"
              % (size, algo, row["ours_us"], row["ours_busbw_GBs"], row["nccl_us"],
                 row["nccl_busbw_GBs"], row["speedup"]) +
              ("" if t_nvls is None else "   nvls %9.1f us (%7.1f GB/s bus)" %
               (row["nvls_us"], row["nvls_busbw_GBs"])) +
              ("" if t_bulk is None else "   tma-bulk %9.1f us (%7.1f GB/s bus)" %
               (row["tma_bulk_us"], row["tma_bulk_busbw_GBs"])), flush=True)
    size *= 4
if rank == 0 and args.out:
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(rows, fh, indent=1)
hvd.shutdown()
comm.shutdown()
