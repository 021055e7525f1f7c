#!/usr/bin/env python
"""How close to bit-identical with the reference is the GPU path?  For every fixture: fraction of F and G
entries whose bits equal the reference's (zeros of either sign count as equal), worst absolute and relative
error.  Also: how often CUDA's sin/cos differ from glibc's (numpy) on the angles of the fixtures."""
import glob
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tol_b200 as T  # noqa: E402

out = {}
tot_eq = tot = 0
for f in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))):
    g = np.load(f)
    ev = T.Evaluator.from_golden(g)
    eqF = eqG = nF = nG = 0
    worst = 0.0
    for s in range(g["x"].shape[0]):
        F, G = ev.eval(g["x"][s])
        for isF, got, ref in ((True, F, g["F"][s]), (False, G, g["G"][s])):
            e = (got == ref)
            if isF:
                eqF += int(e.sum()); nF += e.size
            else:
                eqG += int(e.sum()); nG += e.size
            worst = max(worst, float(np.max(np.abs(got - ref) / (1e-14 + 1e-12 * np.abs(ref)))))
    ev.close()
    out[os.path.basename(f)[:-4]] = {"F_equal": round(eqF / nF, 5), "G_equal": round(eqG / nG, 5), "worst_err_over_tol": round(worst, 4)}
    tot_eq += eqF + eqG
    tot += nF + nG
out["_all"] = tot_eq / tot
g = np.load(os.path.join(ROOT, "tests", "golden", "S10_tempest_ts200.npz"))
ang = g["x"][:, 1:].reshape(-1, 11)[:, 4:7].ravel()
ang = np.concatenate([ang, np.random.default_rng(0).uniform(-7, 7, 2_000_000)])
t = torch.from_numpy(ang).cuda()
out["sin_mismatch_vs_glibc"] = float((torch.sin(t).cpu().numpy() != np.sin(ang)).mean())
out["cos_mismatch_vs_glibc"] = float((torch.cos(t).cpu().numpy() != np.cos(ang)).mean())
print(json.dumps(out, indent=1))
