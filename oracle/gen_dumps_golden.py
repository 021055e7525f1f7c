"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/dumps/<case>_s<k>/{Xoutput,Woutput,Foutput,Goutput}.txt: the
four files the UNMODIFIED reference callback DEFINEGusrfg_ rewrites on every call (src/DefineFG.cpp:16-46,
src/problem.cpp:740-756), for states taken from the committed fixtures tests/golden/<case>.npz.

    python oracle/gen_dumps_golden.py        # needs /root/reference (build container only)

tests/test_host_cpu.py compares tolcuda_write_dump / tolcuda_write_wind_dump with these files byte for byte (the 11
Goutput lines of S10 that print uninitialised memory are masked there)."""
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
OUT = os.path.join(GOLD, "dumps")
# small cases (the files are text): both formulations, the reference's own x0 and perturbed states
PICK = [("S10_skywalker_ts7_gains", 0), ("S10_skywalker_ts7_gains", 3), ("G7_tempestwill_ts7_gains", 2),
        ("G7_skywalker_ts2", 1), ("S10_tempest_ts1", 0)]


def main():
    cases = {c[0]: c for c in CASES}
    for name, s in PICK:
        _, mission, ac, enu, goal, ts, gains, _, _ = cases[name]
        g = np.load(os.path.join(GOLD, name + ".npz"))
        tmp, cwd = tempfile.mkdtemp(), os.getcwd()
        os.chdir(tmp)  # the reference writes into its working directory (Ioutput.txt already in the constructor)
        try:
            p = R.RefProblem(mission, ac, enu, goal, ts=ts, gains=gains, null_io=False)
            F, G = p.eval(g["x"][s], full_callback=True)
        finally:
            R.lib().tolref_set_null_io(1)
            os.chdir(cwd)
        dst = os.path.join(OUT, "%s_s%d" % (name, s))
        os.makedirs(dst, exist_ok=True)
        for f in ("Xoutput.txt", "Woutput.txt", "Foutput.txt", "Goutput.txt"):
            shutil.copy(os.path.join(tmp, f), os.path.join(dst, f))
        shutil.rmtree(tmp)
        p.close()
        assert np.array_equal(F, g["F"][s])
        print("wrote", name, s, sorted(os.listdir(dst)))


if __name__ == "__main__":
    main()
