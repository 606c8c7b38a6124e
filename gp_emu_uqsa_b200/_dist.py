"""torch.distributed plumbing for the two places the hot path shards: the multistart
guesses of Optimize.optimal and the point/cell ranges of prediction + implausibility.  One process
per GPU (torchrun); NCCL when the process group is NCCL (buffers on the rank's GPU), gloo otherwise.
Without an initialised process group everything degenerates to rank 0 of 1."""
import sys

import numpy as np


def rank_world():
    # a process group can only exist if the caller has imported torch.distributed already: do not pay the
    # torch import (about a second) in single-process runs
    dist = sys.modules.get("torch.distributed")
    if dist is not None and dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def block(n, rank, world):
    """Contiguous block partition of range(n): rank g owns [g*n/G, (g+1)*n/G)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def _device():
    import torch
    import torch.distributed as dist
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


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


def gather_blocks(table, n):
    """Every rank filled rows block(n, rank, world) of `table` [n, ...]; return the full table on all
    ranks.  The blocks are disjoint, so a sum over zero-filled copies is the concatenation."""
    rank, world = rank_world()
    if world == 1:
        return table
    lo, hi = block(n, rank, world)
    mine = np.zeros_like(table, dtype=np.float64)
    mine[lo:hi] = np.nan_to_num(np.asarray(table[lo:hi], dtype=np.float64), nan=0.0)
    return all_reduce(mine, "sum")
