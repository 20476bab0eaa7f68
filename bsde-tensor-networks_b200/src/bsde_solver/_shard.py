"""Sample sharding of the data-parallel path: rank r of `world` owns a contiguous block of the N samples; the cores
are replicated and the per-core normal equations are all-reduced (libbsde_b200: ncclAllReduce inside every micro-step)."""


def shard_range(n_samples: int, rank: int, world: int):
    """[lo, hi) of the samples owned by `rank`; block sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside 0..{world - 1}")
    base, extra = divmod(int(n_samples), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)
