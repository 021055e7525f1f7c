#!/usr/bin/env python
"""Launch-shape sweep of the F+G kernel (device-resident, CUDA events on the launching stream, back-to-back
launches): trajectories per CTA, single-trajectory tail waves, and programmatic dependent launch, per batch size.
Every combination produces the same bits (tests); this only times them.  Not the bench contract -- bench.py is.

    python tools/sweep.py [--quick] > gpurun_out/sweep.jsonl"""
import argparse
import itertools
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tol_b200 as T  # noqa: E402
from tol_b200.evaluator import padded_ld  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--cases", default="S10_tempest_ts200:8192,S10_tempest_ts200:16384,S10_tempest_ts200:65536,"
                                   "S10_tempest_ts200:4096,G7_skywalker_ts100:4096,G7_skywalker_ts100:65536,S10_tempest_ts100:65536")
ap.add_argument("--pers", default="0,1,2,3")
ap.add_argument("--tails", default="0,2,4")
ap.add_argument("--overlaps", default="0,1,2")
ap.add_argument("--ms", type=float, default=60.0, help="timed milliseconds per point (approx.)")
args = ap.parse_args()

PEAK = 6544.0
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
st = torch.cuda.Stream()
for case in args.cases.split(","):
    name, B = case.split(":")
    B = int(B)
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    ev = T.Evaluator.from_golden(g)
    ev.set_stream(st.cuda_stream)
    n, neF, neG, ts = ev.n, ev.neF, ev.neG, int(g["ts"])
    ldx, ldF, ldG = padded_ld(n), padded_ld(neF), padded_ld(neG)
    seed0 = T.synth.SEED_S10 if str(g["mission"]) == "S10" else T.synth.SEED_G7
    U = min(B, 512)
    Xu = torch.zeros(U, ldx, dtype=torch.float64)
    T.synth.batch(g["x"][0], seed0, 0, U, out=Xu.numpy())
    Xd = Xu.cuda()[torch.arange(B, device="cuda") % U].contiguous()
    outs = [(torch.empty(B, ldF, dtype=torch.float64, device="cuda"), torch.empty(B, ldG, dtype=torch.float64, device="cuda"))
            for _ in range(2)]
    by = 8.0 * B * (n + neF + neG)
    est_ms = by / (PEAK * 1e6)
    K = max(10, int(args.ms / est_ms))
    torch.cuda.synchronize()
    for per, tail, ov in itertools.product([int(v) for v in args.pers.split(",")], [int(v) for v in args.tails.split(",")],
                                           [int(v) for v in args.overlaps.split(",")]):
        if per in (0, 1) and tail != int(args.tails.split(",")[0]):
            continue  # the tail only exists with runs (per 0 = the library's own rule, reported once)
        ev.set_option("per", per)
        if per:
            ev.set_option("tail_x4", tail)

        def go(k):
            for i in range(k):
                F, G = outs[i & 1] if ov == 2 else outs[0]
                ev.eval_batch_device(Xd, F, G, sync=False, overlap=ov)
        with torch.cuda.stream(st):
            go(3)
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(2 if args.quick else 3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                go(K)
                e1.record(st)
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / K)
        print(json.dumps({"case": name, "B": B, "per": per, "tail_x4": tail if per else None, "overlap": ov, "K": K, "ms": round(best, 5),
                          "GBps": round(by / best / 1e6, 1), "frac": round(by / best / 1e6 / PEAK, 4)}), flush=True)
    ev.close()
    del Xd, outs
    torch.cuda.empty_cache()
