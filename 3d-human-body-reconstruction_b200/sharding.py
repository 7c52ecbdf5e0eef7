"""Batch sharding across the GPUs of one box: bodies are independent (the reference evaluates them
one at a time, lib/model2video.py:514-518), so rank r owns the contiguous slice
[r*B/G, (r+1)*B/G) and NO collective runs on the data path."""


def shard_bounds(total, world_size, rank):
    """Contiguous, balanced slice of `total` bodies for `rank` (first `total % world` ranks get one more)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size/rank")
    base, rem = divmod(total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks_ms(local_ms, dist=None, device=None):
    """Step time of the job = slowest rank (one all_reduce(MAX) of a scalar; harness only)."""
    if dist is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(local_ms)
    import torch
    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
