"""torch.distributed plumbing for the two places the hot path shards: the multistart
guesses of Optimize.optimal and the point ranges of prediction + implausibility.  One process
per GPU (torchrun); NCCL when the process group is NCCL (buffers on the rank's GPU), gloo otherwise.
Without an initialised process group everything degenerates to rank 0 of 1.

Multi-rank contract: every rank runs the same script on the same files.  The global
NumPy RNG is part of the reference's observable behaviour (data shuffle, multistart guesses, Latin
hypercubes): ``sync_numpy_rng`` makes rank 0's generator state the state of every rank right before
each such draw, so an unseeded run still gives all ranks the same training set, guesses and designs.
Files are written by rank 0 only (``is_writer``) and followed by a barrier; designs that the reference
writes and re-reads inside its loops are handed over in memory."""
import sys

import numpy as np


def rank_world():
    # a process group can only exist if the caller has imported torch.distributed already: do not pay the
    # torch import (about a second) in single-process runs
    dist = sys.modules.get("torch.distributed")
    if dist is not None and dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def is_writer():
    """True on the one rank that writes result / checkpoint files."""
    return rank_world()[0] == 0


def barrier():
    if rank_world()[1] > 1:
        import torch.distributed as dist
        dist.barrier()


def block(n, rank, world):
    """Contiguous block partition of range(n): rank g owns [g*n/G, (g+1)*n/G)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def block_aligned(n, rank, world, align=128):
    """Contiguous partition of range(n) into near-equal ranges whose interior boundaries are multiples of
    `align` (prediction tiles are 128 points wide): rank g owns [lo, hi)."""
    tiles = (n + align - 1) // align
    lo = min(n, ((tiles * rank) // world) * align)
    hi = min(n, ((tiles * (rank + 1)) // world) * align)
    return lo, hi


def _device():
    import torch
    import torch.distributed as dist
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def sync_numpy_rng():
    """Give every rank rank 0's global NumPy generator state (no-op without a process group)."""
    rank, world = rank_world()
    if world == 1:
        return
    import torch.distributed as dist
    box = [np.random.get_state() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    if rank != 0:
        np.random.set_state(box[0])


def all_reduce(arr, op="sum"):
    """All-reduce a NumPy array (sum / min / max) over the ranks; returns a NumPy array."""
    rank, world = rank_world()
    if world == 1:
        return arr
    import torch
    import torch.distributed as dist
    ops = {"sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX}
    a = np.ascontiguousarray(arr)
    as_int = a.dtype.kind in "ui"
    t = torch.from_numpy(a.astype(np.int64) if as_int else a.astype(np.float64)).to(_device())
    dist.all_reduce(t, op=ops[op])
    out = t.cpu().numpy()
    return out.astype(a.dtype) if as_int else out


def gather_blocks(table, n, bounds=block):
    """Every rank filled rows bounds(n, rank, world) of `table` [n, ...]; return the full table on all
    ranks: one all_gather of the (equal-size padded) blocks, rows copied through bit for bit -- a NaN stays
    a NaN (the caller's rule for a NaN objective applies exactly as on one rank)."""
    rank, world = rank_world()
    if world == 1:
        return table
    import torch
    import torch.distributed as dist
    table = np.asarray(table)
    spans = [bounds(n, g, world) for g in range(world)]
    width = max(hi - lo for lo, hi in spans)
    lo, hi = spans[rank]
    mine = np.zeros((width,) + table.shape[1:], dtype=np.float64)
    mine[:hi - lo] = table[lo:hi]
    t = torch.from_numpy(mine).to(_device())
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    out = np.empty(table.shape, dtype=np.float64)
    for (glo, ghi), part in zip(spans, parts):
        out[glo:ghi] = part.cpu().numpy()[:ghi - glo]
    return out
