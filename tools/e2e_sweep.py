#!/usr/bin/env python
"""Host-pointer path (bench.py's e2e) by expansion threads per rank and by path, under torchrun on N GPUs sharing one
host: the 65,536-trajectory batch sharded as in bench.py, every rank through tolcuda_eval_batch with pinned host
buffers.  Prints one JSON line per point on rank 0.  Not the bench contract -- bench.py is.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/e2e_sweep.py [--threads 2,3,4,6,8]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tol_b200 as T  # noqa: E402
from tol_b200.evaluator import padded_ld  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--threads", default="2,3,4,6,8")
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--chunks", default="32")
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = np.load(os.path.join(ROOT, "tests", "golden", "S10_tempest_ts200.npz"))
ev = T.Evaluator.from_golden(g, device=local)
b0, b1 = T.synth.shard_range(args.batch, rank, world)
B, ts = b1 - b0, int(g["ts"])
U = min(B, 1024)
Xh = torch.zeros(B, padded_ld(ev.n), dtype=torch.float64, pin_memory=True)
T.synth.batch(g["x"][0], T.synth.SEED_S10, b0, b0 + U, out=Xh.numpy()[:U])
for a in range(U, B, U):
    Xh[a:a + U] = Xh[:min(U, B - a)]
Fh = torch.empty(B, padded_ld(ev.neF), dtype=torch.float64, pin_memory=True)
Gh = torch.empty(B, padded_ld(ev.neG), dtype=torch.float64, pin_memory=True)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def run(full):
    ev.eval_batch_host(Xh.numpy(), Fh.numpy(), Gh.numpy(), full_copy=full)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ev.eval_batch_host(Xh.numpy(), Fh.numpy(), Gh.numpy(), full_copy=full)
    dt = time.perf_counter() - t0
    barrier()
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]) / args.steps


for mb in [int(v) for v in args.chunks.split(",")]:
    ev.set_option("chunk_mb", mb)
    for th in [int(v) for v in args.threads.split(",")]:
        ev.set_host_threads(th)
        s = run(False)
        if rank == 0:
            print(json.dumps({"gpus": world, "path": "compact", "chunk_mb": mb, "threads_per_rank": th, "ms_per_step": 1e3 * s,
                              "node_evals_per_s": args.batch * ts / s}), flush=True)
s = run(True)
if rank == 0:
    print(json.dumps({"gpus": world, "path": "full_g_copy", "ms_per_step": 1e3 * s, "node_evals_per_s": args.batch * ts / s}), flush=True)
ev.close()
if world > 1:
    dist.destroy_process_group()
