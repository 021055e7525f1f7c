"""Host half of the host-pointer batch path in isolation: tolcuda_expand_compact_g on pinned buffers for a
range of thread counts (GB/s of G rows written), then the whole tolcuda_eval_batch(HOST_PTRS) call for a
range of chunk sizes.  Run on the GPU box:  python tools/expandbw.py [B]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import tol_b200 as T  # noqa: E402
from tol_b200.evaluator import compact_len, expand_compact_g, padded_ld  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
g = np.load(os.path.join(ROOT, "tests", "golden", "S10_tempest_ts200.npz"))
m, ts = "S10", 200
n, neF, neG = T.problem_dims(m, ts)
Lc = compact_len(m, ts)
Gc = torch.randn(B, padded_ld(Lc), dtype=torch.float64).pin_memory()
G = torch.empty(B, padded_ld(neG), dtype=torch.float64).pin_memory()
for th in (1, 4, 16):
    expand_compact_g(m, ts, Gc.numpy(), G.numpy(), threads=th)
    t0 = time.perf_counter()
    for _ in range(3):
        expand_compact_g(m, ts, Gc.numpy(), G.numpy(), threads=th)
    dt = (time.perf_counter() - t0) / 3
    print("expand %2d threads: %7.2f ms  %6.1f GB/s written, %6.1f GB/s read+written" % (
        th, dt * 1e3, 8 * neG * B / dt / 1e9, 8 * (neG + Lc) * B / dt / 1e9), flush=True)

ev = T.Evaluator.from_golden(g)
X = torch.zeros(B, padded_ld(n), dtype=torch.float64).pin_memory()
T.synth.batch(g["x"][0], T.synth.SEED_S10, 0, B, out=X.numpy())
F = torch.empty(B, padded_ld(neF), dtype=torch.float64).pin_memory()
for mb in (1, 2, 4, 8, 16, 32, 64):
    ev.set_option("chunk_mb", mb)
    for th in (4, 8, 16):
        ev.set_host_threads(th)
        for full in (False,):
            ev.eval_batch_host(X.numpy(), F.numpy(), G.numpy(), full_copy=full)
            t0 = time.perf_counter()
            for _ in range(3):
                ev.eval_batch_host(X.numpy(), F.numpy(), G.numpy(), full_copy=full)
            dt = (time.perf_counter() - t0) / 3
            print("e2e chunk %3d MB, %2d threads: %7.2f ms  %.3e node-evals/s" % (mb, th, dt * 1e3, B * ts / dt), flush=True)
ev.close()
