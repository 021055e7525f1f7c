#!/usr/bin/env python
"""Write bench.py's `roofline.traffic` record from an `ncu --set full` capture: dram__bytes_read.sum +
dram__bytes_write.sum of the captured launch, its algorithmic bytes, their ratio, and the SASS hash of the kernel in the
library the capture was taken with -- bench.py refuses the record when its own library's kernel hashes differently.

    python tools/traffic_record.py <workload key> <report.ncu-rep | raw-page.csv> <captured B> [--lib tol_b200/libtolcuda.so]

The second argument is the report itself or its raw page saved as CSV (`ncu -i rep --page raw --csv`)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from tol_b200.sass import bench_kernel_pattern, kernel_sass_hash  # noqa: E402

key, rep, B = sys.argv[1], sys.argv[2], int(sys.argv[3])
lib = sys.argv[sys.argv.index("--lib") + 1] if "--lib" in sys.argv else None
fixture = {"S10_tempest_ts200_B65536": "S10_tempest_ts200", "G7_skywalker_ts100_B4096": "G7_skywalker_ts100"}[key]
g = np.load(os.path.join(ROOT, "tests", "golden", fixture + ".npz"))
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, first = rows[0], rows[1], rows[2]
d, u = dict(zip(hdr, first)), dict(zip(hdr, units))


def to_bytes(name):
    v = float(d[name].replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u[name]]


rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
alg = 8.0 * (int(g["n"]) + int(g["neF"]) + int(g["neG"])) * B
h = kernel_sass_hash(bench_kernel_pattern(str(g["mission"]), int(g["wind_model"]), int(g["ts"])), lib)
assert h, "cuobjdump or the kernel not found"
path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
rec = json.load(open(path)) if os.path.exists(path) else {}
rec[key] = {"captured_B": B, "kernel": d.get("Kernel Name"), "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
            "algorithmic_bytes": alg, "ratio": round((rd + wr) / alg, 4), "duration_under_ncu_ms": float(d["gpu__time_duration.sum"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(u["gpu__time_duration.sum"], 1.0),
            "fp64_pipe_pct": float(d["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]),
            "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
            "dram_cycles_active_pct": float(d["dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]),
            "warp_instructions": float(d["smsp__inst_executed.sum"]),
            "source": os.path.basename(rep), "kernel_sass_sha256_16": h}
json.dump(rec, open(path, "w"), indent=1)
print(json.dumps(rec[key], indent=1))
