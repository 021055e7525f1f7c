#!/usr/bin/env python
"""Per-instruction stall reasons from an .ncu-rep: ncu_stalls.py rep [N]"""
import csv, io, subprocess, sys
rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 14
M = ["long_scoreboard", "short_scoreboard", "wait", "no_instructions", "mio_throttle", "barrier", "lg_throttle"]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--metrics",
                      ",".join("smsp__pcsamp_warps_issue_stalled_" + m for m in M) + ",smsp__pcsamp_sample_count,inst_executed"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, rows = rows[1], rows[2:]
tot = sum(int(r[2]) for r in rows)
print("samples", tot, "instructions", len(rows), "cols", hdr[2:])
for col in range(4, len(hdr)):
    t = sum(int(r[col] or 0) for r in rows)
    if t < 0.04 * tot:
        continue
    print("== %s: %d (%.1f%%)" % (hdr[col], t, 100.0 * t / tot))
    top = sorted(enumerate(rows), key=lambda x: -int(x[1][col] or 0))[:n]
    for i, r in sorted(top):
        prev = rows[i - 1][1].strip()[:46] if i else ""
        print("  %5d %5d  %-58s | prev: %s" % (i, int(r[col] or 0), r[1].strip()[:58], prev))
