"""One process per GPU, trajectories sharded by index, no collective on the data path (SURVEY.md
section 8e).  torch.distributed is used for the three things around the path: agreeing on the shard
ranges, the max-over-ranks of the timings, and the optional host-side gather of per-trajectory rows to
rank 0 (the `north_star`'s "host-side gather of per-trajectory F/G")."""
import numpy as np
import torch
import torch.distributed as dist

from .synth import shard_range


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def my_shard(B):
    """[b0, b1) of the global batch owned by this rank"""
    r, w = world()
    return shard_range(B, r, w)


def max_over_ranks(value, device="cpu"):
    r, w = world()
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if w > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_rows(local, B):
    """local: this rank's rows [b1-b0, m] (numpy, float64).  Returns the full [B, m] array on rank 0
    (None elsewhere).  Rows are placed by trajectory index, so the result does not depend on the
    number of ranks."""
    r, w = world()
    if w == 1:
        return local
    m = local.shape[1]
    per = (B + w - 1) // w
    pad = np.zeros((per, m))
    pad[:local.shape[0]] = local
    mine = torch.from_numpy(pad)
    parts = [torch.empty_like(mine) for _ in range(w)] if r == 0 else None
    dist.gather(mine, parts, dst=0)
    if r != 0:
        return None
    out = np.empty((B, m))
    for q in range(w):
        b0, b1 = shard_range(B, q, w)
        out[b0:b1] = parts[q].numpy()[:b1 - b0]
    return out
