"""Helpers for single-GPU tests that simulate a multi-rank world in one
process (every "rank" gets its own streams and symmetric segments on cuda:0;
the kernels address peers through plain pointers, exactly as over IPC)."""
import torch


class FakeComm(object):
    def __init__(self, rank, world, device):
        self.rank, self.world, self.local_rank = rank, world, rank
        self.device = torch.device(device)
        self.distributed = False
        self.is_cuda = True

    def barrier(self):
        pass

    def all_gather_object(self, o):
        return [o]

    def broadcast_object(self, o, src=0):
        return o


def make_world(world, device="cuda:0", options=None):
    from parallax_b200.parallel.symmetric import LocalWorld
    from parallax_b200.parallel.nvlink_backend import NVFabric
    lw = LocalWorld(world)
    fabrics = []
    for r in range(world):
        comm = FakeComm(r, world, device)
        opts = {"comm_blocks": 4}
        opts.update(options or {})
        fabrics.append(NVFabric(comm, exchange=lw.exchange_for(r), options=opts))
    # Upload every rank's pad-pointer table NOW: a lazy (blocking) H2D copy issued
    # for rank k while rank 0's barrier kernel is already spinning on the same GPU
    # can stall behind it.  Nothing host-blocking may happen between the launches
    # of the simulated ranks.
    import torch
    for f in fabrics:
        f.heap.pads_dev()
    torch.cuda.synchronize()
    return fabrics
