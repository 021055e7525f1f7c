"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/results/<case>_s<k>.{json,txt}: the two result files the
UNMODIFIED reference writes (problem::writeJSON src/problem.cpp:1247-1365, problem::writeTXT :1371-1418) for
states taken from the committed fixtures tests/golden/<case>.npz (x = sample k, F[0] = its objective).

    python oracle/gen_results_golden.py        # needs /root/reference (build container only)

tests/test_host_cpu.py compares tolcuda_write_results_json / _txt with these files byte for byte."""
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refclient as R  # noqa: E402
from gen_golden import CASES  # noqa: E402

GOLD = os.path.join(HERE, "..", "tests", "golden")
OUT = os.path.join(GOLD, "results")
# fixture, sample: shipped sizes at the reference's own initial guess, small perturbed cases with edited
# gains, and ts = 1, 2 where jsoncpp keeps short arrays on one line
PICK = [("S10_tempest_ts100", 0), ("G7_skywalker_ts100", 1), ("G7_tempestwill_ts7_gains", 2),
        ("S10_skywalker_ts7_gains", 3), ("S10_tempest_ts1", 1), ("G7_skywalker_ts2", 0), ("G7_skywalker_ts2", 1)]


def main():
    os.makedirs(OUT, exist_ok=True)
    cases = {c[0]: c for c in CASES}
    for name, s in PICK:
        _, mission, ac, enu, goal, ts, gains, _, _ = cases[name]
        g = np.load(os.path.join(GOLD, name + ".npz"))
        p = R.RefProblem(mission, ac, enu, goal, ts=ts, gains=gains)
        x, F0 = g["x"][s], float(g["F"][s, 0])
        tmp = tempfile.mkdtemp()
        p.write_json(x, F0, os.path.join(OUT, "%s_s%d.json" % (name, s)))
        shutil.move(p.write_txt(x, F0, tmp), os.path.join(OUT, "%s_s%d.txt" % (name, s)))
        shutil.rmtree(tmp)
        p.close()
        print("wrote", name, s)


if __name__ == "__main__":
    main()
