"""Synthetic batched decision vectors, SURVEY.md section 8d: trajectory b is
x0[i]*(1+0.05*u) + 0.01*u' with (u, u') drawn interleaved per i from
numpy.random.Generator(PCG64(seed0 + b)).uniform(-1, 1, 2n).  The seed is the GLOBAL trajectory
index, so a shard [b0, b1) of a batch is the same on any number of GPUs."""
import numpy as np

SEED_G7 = 20260000
SEED_S10 = 20270000


def perturb(x0, seed):
    r = np.random.Generator(np.random.PCG64(seed)).uniform(-1.0, 1.0, size=2 * x0.size)
    return x0 * (1.0 + 0.05 * r[0::2]) + 0.01 * r[1::2]


def batch(x0, seed0, b0, b1, out=None, ld=None):
    """rows b0..b1-1 of the synthetic batch, into `out` [b1-b0, ld] (allocated if None)"""
    n = x0.size
    if out is None:
        out = np.zeros((b1 - b0, ld or n))
    for b in range(b0, b1):
        out[b - b0, :n] = perturb(x0, seed0 + b)
    return out


def shard_range(B, rank, world):
    """contiguous block of trajectory indices owned by `rank`"""
    per = (B + world - 1) // world
    b0 = min(B, rank * per)
    return b0, min(B, b0 + per)
