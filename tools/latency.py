#!/usr/bin/env python
"""Single-trajectory drop-in latency (BASELINE.json configs[0..1]): microseconds per DEFINEGusrfg_ call
through libtolcuda on cuda:0 versus the reference callback on one host core (oracle/_ref: as shipped with
its four file dumps in a scratch directory, and with the dumps sent to /dev/null)."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tol_b200 as T  # noqa: E402

out = {}
for name in ("G7_skywalker_ts100", "S10_tempest_ts100", "S10_tempest_ts200"):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    ev = T.Evaluator.from_golden(g)
    x = g["x"][1]
    res = {}
    for label, nf, ng in (("FG", 1, 1), ("F", 1, 0), ("G", 0, 1)):
        for _ in range(50):
            ev.usrfun(x, nf, ng)
        n = 2000
        t0 = time.perf_counter()
        for _ in range(n):
            ev.usrfun(x, nf, ng)
        res["gpu_us_" + label] = 1e6 * (time.perf_counter() - t0) / n
    ev.close()
    try:
        import refclient as R
        if R.available():
            ts = int(g["ts"])
            cwd = os.getcwd()
            os.chdir(tempfile.mkdtemp(prefix="tolref_cwd_"))
            for label, null_io in (("ref_us_as_shipped_with_dumps", False), ("ref_us_dumps_to_devnull", True)):
                p = R.RefProblem(str(g["mission"]), str(g["aircraft"]), tuple(g["enu"]), tuple(g["goal_enu"]),
                                 ts=None if ts == 100 else ts, null_io=null_io)
                for _ in range(3):
                    p.eval(x, full_callback=True)
                n = 30
                t0 = time.perf_counter()
                for _ in range(n):
                    p.eval(x, full_callback=True)
                res[label] = 1e6 * (time.perf_counter() - t0) / n
                p.close()
            os.chdir(cwd)
    except Exception as exc:  # reference arm is optional
        res["ref_error"] = str(exc)
    out[name] = res
print(json.dumps(out, indent=1))
