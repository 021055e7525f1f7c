#!/usr/bin/env python
"""Top stall-sample SASS instructions of an .ncu-rep (source page).  Usage: ncu_sass.py rep [N]"""
import collections
import csv
import io
import subprocess
import sys

rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))[2:]
tot = sum(int(r[4]) for r in rows)
print("samples", tot, "sass instructions", len(rows))
top = sorted(enumerate(rows), key=lambda t: -int(t[1][4]))[:n]
for i, r in sorted(top):
    print("%5d %5.2f%% exec=%8s %s" % (i, 100 * int(r[4]) / tot, r[5], r[1].strip()[:100]))
agg = collections.Counter()
ex = collections.Counter()
for r in rows:
    t = r[1].strip().split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    agg[op] += int(r[4])
    ex[op] += int(r[5])
print("samples by opcode:", [(k, round(100 * v / tot, 1)) for k, v in agg.most_common(16)])
te = sum(ex.values())
print("executed by opcode:", [(k, round(100 * v / te, 1)) for k, v in ex.most_common(16)])
