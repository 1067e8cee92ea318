"""Multi-GPU sharding of a batch of independent MPC instances: contiguous block split of the batch
index, GPU g gets instances [g*B/G, (g+1)*B/G).  No data-path collective exists inside the solve; the only exchanges
are the GP-model broadcast and the gather of [u | x | status] blocks on rank 0."""


def shard_range(B, rank, world):
    """Contiguous block [lo, hi) of rank `rank`; blocks differ by at most one instance."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(int(B), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_order(B, world):
    """Index ranges of every rank in the gathered (rank-major) result; concatenation is the original batch order."""
    return [shard_range(B, r, world) for r in range(world)]
