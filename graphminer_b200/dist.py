"""Process-per-GPU sharding helpers (torch.distributed is plumbing: rendezvous + one all-reduce).

The hot path shards by DFS root with no data-path collective: rank r owns the
contiguous source-vertex range [bounds[r], bounds[r+1]) and the per-rank 64-bit counts are summed by
a single all-reduce -- NCCL over NVLink on GPUs (replaces the host loop `total += h_counts[i]`,
src/triangle/multigpu.cu:84, and MPI_Allreduce, src/triangle/dist_gpu.cpp:30), gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np

from . import capi


def shard_bounds(rowptr, colidx, world_size: int, balance: bool = True):
    """n+1 source-range boundaries; balance=False reproduces graph_partition.cc:84-86 exactly."""
    return [int(x) for x in capi.host_shard_bounds(rowptr, colidx, world_size, balance)]


def allreduce_counts(counts, device="cpu"):
    """Sum a list of non-negative 64-bit counts over all ranks (exact: int64 two's-complement add)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(np.array(counts, dtype=np.uint64).view(np.int64), dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return [int(x) for x in t.cpu().numpy().view(np.uint64)]


def sharded_count(count_range, rowptr, colidx, ncounts=1, finish=None, balance=True, device="cpu"):
    """Run `count_range(begin, end) -> int | list[int]` on this rank's shard and reduce.

    `finish` (e.g. the motif formula fix-up, which divides and so must run AFTER the reduction) is
    applied to the reduced list on every rank."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    b = shard_bounds(rowptr, colidx, world, balance)
    local = count_range(b[rank], b[rank + 1])
    local = list(local) if isinstance(local, (list, tuple)) else [local]
    assert len(local) == ncounts
    total = allreduce_counts(local, device)
    return finish(total) if finish else total
