"""Multi-GPU plumbing of the hot path (one process per GPU, torch.distributed).

Reads/pairs are independent units: rank r of W owns the contiguous global index range
[r*units_per_rank, (r+1)*units_per_rank) (weak scaling) or an even split of a fixed total (strong).
There is no data-path collective; the only exchange is ONE sum all-reduce of the packed int64 counter
block at the end of a pass -- exactly what Stats::merge (src/stats.cpp:877-955) and FilterResult::merge
(src/filterresult.cpp:38-89) do across the reference's worker threads.
"""


def shard_range(rank, world, total):
    """Even contiguous split of [0,total): first (total % world) ranks get one extra unit."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def weak_first_index(rank, units_per_rank):
    return rank * units_per_rank


def allreduce_counters(t):
    """In-place SUM all-reduce of an int64 counter tensor (CPU/gloo or CUDA/NCCL); no-op when not distributed."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t
