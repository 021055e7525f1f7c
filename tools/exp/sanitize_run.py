"""Small evaluation of every kernel flavour (plain / compact / summary, kernel A with run lengths 1, 2, 3, kernel B,
odd leading dimensions, device expansion) with cross-checks between the paths; also the command to put under
compute-sanitizer where that tool is available:  compute-sanitizer --tool memcheck python tools/exp/sanitize_run.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import tol_b200 as T  # noqa: E402

for name in ("S10_tempest_ts200", "G7_skywalker_ts100", "S10_tempest_ts100_wind3", "S10_tempest_ts1", "S10_tempesteric_ts33",
             "G7_tempestwences_ts45_gains", "S10_skywalker_ts7_gains"):  # odd ts: record slots shifted by one double
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    for kernel, per in ((0, None), (0, 1), (0, 3), (2, None)):
        ev = T.Evaluator.from_golden(g, options={"kernel": kernel, "per": per or 0})
        B = 19
        X = T.synth.batch(g["x"][0], 5, 0, B)
        F, G = ev.eval_batch_host(X, full_copy=True)
        F2, G2 = ev.eval_batch_host(X)
        S, _, _ = ev.summary_host(X)
        Xd = torch.from_numpy(X).cuda()
        Fd = torch.empty(B, ev.neF + 1, dtype=torch.float64, device="cuda")  # odd leading dimension: lane-copy path
        Gd = torch.empty(B, ev.neG + 1, dtype=torch.float64, device="cuda")
        ev.eval_batch_device(Xd, Fd, Gd)
        Gc = torch.empty(B, ev.compact_len, dtype=torch.float64, device="cuda")
        ev.eval_batch_device(Xd, Fd, Gc, compact_rows=True)
        ev.expand_compact_device(Gc, Gd)
        D = torch.rand(B, ev.n, dtype=torch.float64, device="cuda")
        Lam = torch.rand(B, ev.neF, dtype=torch.float64, device="cuda")
        Y, Z = torch.empty(B, ev.neF, dtype=torch.float64, device="cuda"), torch.empty(B, ev.n, dtype=torch.float64, device="cuda")
        ev.jac_vec(Xd, D, Y)
        ev.jac_tvec(Xd, Lam, Z)
        lhs, rhs = (Lam * Y).sum(1), (Z * D).sum(1)
        assert torch.allclose(lhs, rhs, rtol=1e-9, atol=1e-9)
        f1, g1 = ev.eval(X[0])
        assert np.array_equal(G, G2) and np.array_equal(f1, F[0]) and np.array_equal(Gd.cpu().numpy()[:, :ev.neG], G)
        # launch shapes with a single-trajectory tail, and programmatic dependent launches (both forms)
        ld = T.evaluator.padded_ld(ev.neG)
        outs = [(torch.empty(B, ev.neF, dtype=torch.float64, device="cuda"), torch.empty(B, ld, dtype=torch.float64, device="cuda"))
                for _ in range(2)]
        for p2, tail in ((2, 1), (2, 64), (3, 2)):
            ev.set_option("per", p2)
            ev.set_option("tail_x4", tail)
            for ov in (0, 1, 2):
                for i in range(4):
                    ev.eval_batch_device(Xd, outs[i & 1][0], outs[i & 1][1], sync=False, overlap=ov)
                ev.synchronize()
                for Fo, Go in outs:
                    assert np.array_equal(Fo.cpu().numpy(), F) and np.array_equal(Go.cpu().numpy()[:, :ev.neG], G)
        ev.close()
print("sanitize_run ok")
