"""Hypothesis sharding across ranks: the partition the C++ side uses
(csrc/ransac.cu fit_view) restated for the Python plumbing, plus the torch.distributed exchange
callback that m3d_ctx_set_exchange can use instead of the library's own NCCL communicator."""
import ctypes as C

import numpy as np


def shard_rows(rows, rank, world):
    """rank r scores rows [lo, hi) of a wave of `rows` hypotheses; S = padded shard length."""
    S = (rows + world - 1) // world
    lo = min(rows, rank * S)
    hi = min(rows, lo + S)
    return lo, hi, S


def gather_counts(packed, world, all_gather):
    """all-gather one rank's packed uint32 counts (bit31 = MinimalFit failed); rank-major = row order"""
    import torch
    t = torch.from_numpy(packed.astype(np.int32, copy=True))
    outs = [torch.empty_like(t) for _ in range(world)]
    all_gather(t, outs)
    return torch.cat(outs).numpy().astype(np.uint32)


def torch_exchange(group=None, device=None):
    """Exchange callback for Context.set_exchange: all-gathers `nbytes` per rank through
    torch.distributed (NCCL on CUDA tensors when on_device, gloo on host memory otherwise)."""
    import torch
    import torch.distributed as dist

    def fn(send_ptr, recv_ptr, nbytes, on_device):
        world = dist.get_world_size(group)
        if on_device:
            raise NotImplementedError("device exchange goes through m3d_ctx_init_nccl")
        src = (C.c_ubyte * nbytes).from_address(send_ptr)
        t = torch.frombuffer(src, dtype=torch.uint8).clone()
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t, group=group)
        dst = (C.c_ubyte * (nbytes * world)).from_address(recv_ptr)
        torch.frombuffer(dst, dtype=torch.uint8).copy_(torch.cat(outs))
        return 0

    return fn
