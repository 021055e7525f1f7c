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


def eval_and_gather_device(ev, X, B, dst=0):
    """Device-side gather: every rank evaluates its shard X (torch CUDA tensor
    [b1-b0, >= n]) into F rows and COMPACT G rows, the shards are gathered on GPU `dst` with one NCCL gather
    per array -- G crosses NVLink as compact rows, a third of the bytes -- and expanded there into rows in
    coordinate order by the device expansion kernel.  Returns (F [B, neF], G [B, neG]) as CUDA tensors on rank
    `dst`, (None, None) elsewhere; rows are placed by trajectory index."""
    r, w = world()
    nb = X.shape[0]
    per = (B + w - 1) // w
    Lc = ev.compact_len
    Fl = torch.zeros(per, ev.neF, dtype=torch.float64, device=X.device)
    Gl = torch.zeros(per, Lc, dtype=torch.float64, device=X.device)
    if nb:
        ev.eval_batch_device(X, Fl[:nb], Gl[:nb], compact_rows=True)
    if w == 1:
        G = torch.empty(B, ev.neG, dtype=torch.float64, device=X.device)
        ev.expand_compact_device(Gl[:B], G)
        return Fl[:B], G
    Fp = [torch.empty_like(Fl) for _ in range(w)] if r == dst else None
    Gp = [torch.empty_like(Gl) for _ in range(w)] if r == dst else None
    dist.gather(Fl, Fp, dst=dst)
    dist.gather(Gl, Gp, dst=dst)
    if r != dst:
        return None, None
    torch.cuda.current_stream().synchronize()  # the gathers ran on torch's stream, the expansion runs on the context's
    F = torch.empty(B, ev.neF, dtype=torch.float64, device=X.device)
    G = torch.empty(B, ev.neG, dtype=torch.float64, device=X.device)
    for q in range(w):
        b0, b1 = shard_range(B, q, w)
        if b1 > b0:
            F[b0:b1] = Fp[q][:b1 - b0]
            ev.expand_compact_device(Gp[q][:b1 - b0], G[b0:b1], sync=False)
    ev.synchronize()
    return F, G
