import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN = sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))

# north_star tolerance: F/G within 1e-12 relative, 1e-14 absolute of the reference
RTOL, ATOL = 1e-12, 1e-14


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def assert_parity(got, ref, what=""):
    """|got - ref| <= ATOL + RTOL*|ref| elementwise; reports the worst offender"""
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    assert np.isfinite(got).all(), "%s: non-finite values at %s" % (what, np.where(~np.isfinite(got))[0][:8])
    err = np.abs(got - ref)
    tol = ATOL + RTOL * np.abs(ref)
    bad = err > tol
    if bad.any():
        i = np.unravel_index(np.argmax(err / tol), err.shape)
        raise AssertionError("%s: %d entries outside 1e-14+1e-12*|ref|; worst at %s: got %r ref %r (err %.3e)"
                             % (what, bad.sum(), i, got[i], ref[i], err[i]))


def port_from_golden(g):
    import portclient as P
    wm = int(g["wind_model"])
    p = P.PortProblem(str(g["mission"]), int(g["ts"]), g["ac"], g["gn"], g["goal_ned"], 1 if wm == 3 else wm)
    if wm == 3:
        p.set_wind_grid(g["grid_x"], g["grid_y"], g["grid_z"], g["grid_v"], g["grid_datum"], g["grid_spacing"])
    return p


@pytest.fixture(scope="session")
def oracle_built():
    """the plain-C oracle, built on demand (gcc only)"""
    import subprocess
    import portclient as P
    if not P.available():
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
    return P
