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


NO_ROW = 0xFFFFFFFF
TIED_CAP = 8  # csrc/loop_kernels.cuh kTiedCap
REC_WORDS = 16  # BestRec = 64 bytes


def best_record(packed, rows_of_rank, n_points):
    """csrc/loop_kernels.cuh wave_best_kernel restated: the 16-word record one rank derives from its shard's packed
    counts (bit31 = MinimalFit failed) -- max count, first row reaching it, number of ties, valid rows, first row
    with fitness 1, the first TIED_CAP tied rows (wave row numbers, ascending)."""
    packed = np.asarray(packed, dtype=np.uint32)[: len(rows_of_rank)]
    rows_of_rank = np.asarray(rows_of_rank, dtype=np.uint32)
    valid = (packed >> 31) == 0
    cnt = np.where(valid, packed, 0).astype(np.uint32)
    rec = np.full(REC_WORDS, NO_ROW, dtype=np.uint32)
    rec[[0, 2, 3, 5, 6, 7]] = 0
    rec[3] = int(valid.sum())
    full = rows_of_rank[valid & (cnt == n_points) & (cnt > 0)]
    if len(full):
        rec[4] = full.min()
    best = int(cnt.max()) if len(cnt) else 0
    if best:
        tied = np.sort(rows_of_rank[cnt == best])
        rec[0], rec[1], rec[2] = best, tied[0], len(tied)
        rec[8: 8 + min(len(tied), TIED_CAP)] = tied[:TIED_CAP]
    return rec


def merge_records(recs):
    """best_merge_kernel restated: the record of the whole wave from the ranks' records (rank-major (world, 16))"""
    recs = np.asarray(recs, dtype=np.uint32).reshape(-1, REC_WORDS)
    m = np.full(REC_WORDS, NO_ROW, dtype=np.uint32)
    m[[0, 2, 3, 5, 6, 7]] = 0
    m[0] = recs[:, 0].max()
    m[3] = recs[:, 3].sum()
    m[4] = recs[:, 4].min()
    m[5] = np.bitwise_or.reduce(recs[:, 5])
    if m[0]:
        top = recs[recs[:, 0] == m[0]]
        m[1] = top[:, 1].min()
        m[2] = top[:, 2].sum()
        tied = np.sort(top[:, 8:].ravel())
        tied = tied[tied != NO_ROW][:TIED_CAP]
        m[8: 8 + len(tied)] = tied
        if (top[:, 2] > TIED_CAP).any() and m[2] <= TIED_CAP:
            m[2] = TIED_CAP + 1
    return m


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


def upload_slice(n_points, rank, world):
    """the part of a replicated n x 3 float64 cloud rank `rank` copies to its GPU in a multi-rank host-buffer fit
    (csrc/ransac.cu ChunkPlan::issue_copies): (offset, length, slice size) in doubles.  All slices have the same size S
    (the all-gather's count); the last ones may be shorter or empty."""
    tot = 3 * int(n_points)
    s = (tot + world - 1) // world
    off = s * rank
    return off, (min(s, tot - off) if off < tot else 0), s
