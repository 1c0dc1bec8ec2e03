"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for the rendezvous only.

The hot path shards without a data-path collective: 8x8 blocks, frames and macroblock rows are independent.
This module holds the pure partitioning rules plus the two collectives the harness needs: a max-reduce for timing and the
optional gather of per-rank results (motion vectors / frames) on rank 0.  Works with the `nccl` backend on GPUs and with
`gloo` on CPU (tests/test_shard_cpu.py runs it at world_size 2)."""
import numpy as np


def split_range(n, rank, world):
    """Contiguous, balanced [lo, hi) of n units for `rank` (first n % world ranks get one extra)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def mb_row_range(mb_h, rank, world):
    """Macroblock rows of a picture owned by `rank` (motion search: each rank needs its rows of `cur` and those rows
    +-range of `ref`)."""
    return split_range(mb_h, rank, world)


def frames_for_rank(n_frames, rank, world):
    """Round-robin frame assignment (stream i -> GPU i mod world), the config-5 mapping."""
    return list(range(rank, n_frames, world))


def max_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_rows(local, lo, hi, total_rows, dst=0):
    """Optional result gather: every rank contributes rows [lo, hi) of a (total_rows, ...) array; rank `dst` returns the
    assembled array, the others None."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    parts = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object((lo, hi, np.ascontiguousarray(local)), parts, dst=dst)
    if dist.get_rank() != dst:
        return None
    out = np.zeros((total_rows,) + local.shape[1:], dtype=local.dtype)
    for (a, b, piece) in parts:
        out[a:b] = piece
    return out
