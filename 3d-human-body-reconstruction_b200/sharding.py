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


def forward_shard(dm, betas, pose, transl=None, world_size=1, rank=0, **kwargs):
    """This rank's part of a batched forward over `world_size` GPUs: the contiguous slice
    [lo, hi) of the bodies, evaluated by the rank's own DeviceModel (`dm`, one handle per GPU) through
    smplk_forward.  `betas` / `pose` / `transl` describe the WHOLE batch (numpy or torch, any device; a
    one-row betas is shared); only the slice is copied to the device.  Returns (lo, hi, verts, joints)
    with the outputs left on the rank's GPU -- results stay sharded, no collective."""
    import numpy as np
    import torch
    from .body_models import body_model_apply
    total = pose.shape[0]
    lo, hi = shard_bounds(total, world_size, rank)
    dev = torch.device("cuda", dm.device)

    def take(a, rows):
        if a is None:
            return None
        a = a[rows] if a.shape[0] != 1 else a
        t = torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)) if not torch.is_tensor(a) else a.float()
        return t.to(dev, non_blocking=True).contiguous()
    rows = slice(lo, hi)
    if hi == lo:
        return lo, hi, None, None
    with torch.cuda.device(dev):
        v, j, _, _ = body_model_apply(dm, take(betas, rows), take(pose, rows), transl=take(transl, rows), **kwargs)
    return lo, hi, v, j
