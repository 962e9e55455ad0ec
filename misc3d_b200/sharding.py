"""Hypothesis sharding across ranks: the partition the C++ side uses
(csrc/ransac.cu fit_view) restated for the Python plumbing, plus the torch.distributed exchange
callback that m3d_ctx_set_exchange can use instead of the library's own NCCL communicator."""
import ctypes as C

import numpy as np


SHARD_BLOCK = 256  # csrc/scan.h kShardBlock


def shard_rows(rows, rank, world):
    """The wave rows rank `rank` scores, in its local order, and the padded per-rank stride S.
    Rows are dealt out in cyclic blocks of SHARD_BLOCK (block b -> rank b % world): every rank owns rows near
    the start of the wave, so it can start its GPU after drawing a fraction of the (sequential) sample table."""
    if world <= 1:
        return np.arange(rows, dtype=np.uint32), rows
    blocks = (rows + SHARD_BLOCK - 1) // SHARD_BLOCK
    S = ((blocks + world - 1) // world) * SHARD_BLOCK
    mine = [np.arange(b * SHARD_BLOCK, min(rows, (b + 1) * SHARD_BLOCK), dtype=np.uint32)
            for b in range(rank, blocks, world)]
    return (np.concatenate(mine) if mine else np.empty(0, np.uint32)), S


def gathered_to_wave_order(allc, rows, world):
    """rank-major all-gathered buffer (world x S) -> counts in wave-row order"""
    if world <= 1:
        return allc[:rows]
    S = len(allc) // world
    g = np.arange(rows)
    rank = (g // SHARD_BLOCK) % world
    local = (g // (SHARD_BLOCK * world)) * SHARD_BLOCK + g % SHARD_BLOCK
    return allc[rank * S + local]


def gather_counts(packed, world, all_gather):
    """all-gather one rank's packed uint32 counts (bit31 = MinimalFit failed); returns the rank-major buffer"""
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
