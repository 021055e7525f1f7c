"""TEST INFRASTRUCTURE ONLY: ctypes client for oracle/_ref/libtolref.so (the unmodified reference
sources compiled by oracle/Makefile).  Used by tests/, oracle/gen_golden.py and bench.py's
cpu_baseline / --impl reference legs.  The product package (tol_b200) never imports this."""
import ctypes as C
import os
import shutil
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libtolref.so")
REF_PARAMS = os.path.join(_HERE, "_ref", "params")

_lib = None


def available():
    return os.path.exists(REF_SO) and os.path.isdir(REF_PARAMS)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(REF_SO)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        L.tolref_create.restype = C.c_void_p
        L.tolref_create.argtypes = [C.c_char_p, C.c_char_p] + [C.c_double] * 7 + [C.c_char_p]
        L.tolref_destroy.argtypes = [C.c_void_p]
        L.tolref_dims.argtypes = [C.c_void_p, ip, ip, ip, ip, ip]
        L.tolref_pattern.argtypes = [C.c_void_p, ip, ip]
        L.tolref_x0.argtypes = [C.c_void_p, dp]
        L.tolref_bounds.argtypes = [C.c_void_p, dp, dp, dp, dp]
        L.tolref_params.argtypes = [C.c_void_p, dp, dp, dp, dp, dp, ip]
        L.tolref_set_wind_grid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp]
        L.tolref_get_wind.argtypes = [C.c_void_p, dp]
        L.tolref_chi_d.restype = C.c_double
        L.tolref_chi_d.argtypes = [C.c_void_p]
        L.tolref_eval.argtypes = [C.c_void_p, dp, C.c_int, dp, C.c_int, dp]
        L.tolref_usrfun.argtypes = [C.c_void_p, dp, C.c_int, dp, C.c_int, dp]
        for f in (L.tolref_eval_many, L.tolref_usrfun_many):
            f.argtypes = [C.c_void_p, C.c_int, dp, C.c_long, dp, C.c_long, dp, C.c_long]
        L.tolref_set_null_io.argtypes = [C.c_int]
        L.tolref_write_json.argtypes = [C.c_void_p, dp, C.c_double, C.c_char_p]
        L.tolref_write_txt.argtypes = [C.c_void_p, dp, C.c_double]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def make_root(ts=None, mission=None, gains=None, src=REF_PARAMS):
    """Return (root_path_with_trailing_slash, tmpdir_or_None).  ts != shipped value needs a private
    copy of the .param tree with line 2 of problems/<mission>/snopt.param edited (the reference
    reads ts only from that file, reference src/parameters.cpp:130-148); `gains` (kT,kp,kv,ka,kdt)
    likewise rewrites problems/<mission>/gains.param."""
    if ts is None and gains is None:
        return src.rstrip("/") + "/", None
    tmp = tempfile.mkdtemp(prefix="tolref_params_")
    for d in ("aircraft", "problems"):
        shutil.copytree(os.path.join(src, d), os.path.join(tmp, d))
    if ts is not None:
        p = os.path.join(tmp, "problems", mission, "snopt.param")
        os.chmod(p, 0o644)
        with open(p, "r", newline="") as fh:
            lines = fh.read().split("\n")
        lines[1] = "%d       // Number of time segments  " % ts
        with open(p, "w", newline="") as fh:
            fh.write("\n".join(lines))
    if gains is not None:
        p = os.path.join(tmp, "problems", mission, "gains.param")
        os.chmod(p, 0o644)
        with open(p, "w", newline="") as fh:
            fh.write("//edited gains\n" + "\n".join("%r   // gain" % float(g) for g in gains))
    return tmp + "/", tmp


class RefProblem:
    """One reference `problemG7` / `problemS10` object (reference src/tol.cpp:5-36)."""

    def __init__(self, mission, aircraft, enu=(0.0, 0.0, 70.0), goal=(0.0, 0.0, 0.0, 0.0), ts=None,
                 gains=None, null_io=True):
        L = lib()
        L.tolref_set_null_io(1 if null_io else 0)
        root, self._tmp = make_root(ts, mission, gains)
        self.mission, self.aircraft = mission, aircraft
        self.h = L.tolref_create(mission.encode(), aircraft.encode(), *[float(v) for v in enu],
                                 *[float(v) for v in goal], root.encode())
        if not self.h:
            raise RuntimeError("tolref_create failed for %s/%s" % (mission, aircraft))
        v = [C.c_int() for _ in range(5)]
        L.tolref_dims(self.h, *[C.byref(t) for t in v])
        self.n, self.neF, self.neG, self.ts, self.nb = [t.value for t in v]

    def close(self):
        if getattr(self, "h", None):
            lib().tolref_destroy(self.h)
            self.h = None
        if getattr(self, "_tmp", None):
            shutil.rmtree(self._tmp, ignore_errors=True)
            self._tmp = None

    __del__ = close

    def pattern(self):
        i = np.empty(self.neG, np.int32)
        j = np.empty(self.neG, np.int32)
        lib().tolref_pattern(self.h, _ip(i), _ip(j))
        return i, j

    def x0(self):
        x = np.empty(self.n)
        lib().tolref_x0(self.h, _dp(x))
        return x

    def bounds(self):
        xl, xu = np.empty(self.n), np.empty(self.n)
        fl, fu = np.empty(self.neF), np.empty(self.neF)
        lib().tolref_bounds(self.h, _dp(xl), _dp(xu), _dp(fl), _dp(fu))
        return xl, xu, fl, fu

    def params(self):
        ac, gn, lm, sn, goal = np.empty(15), np.empty(5), np.empty(8), np.empty(6), np.empty(4)
        wm = C.c_int()
        lib().tolref_params(self.h, _dp(ac), _dp(gn), _dp(lm), _dp(sn), _dp(goal), C.byref(wm))
        return dict(ac=ac, gn=gn, lm=lm, sn=sn, goal=goal, wind_model=wm.value)

    def set_wind_grid(self, gx, gy, gz, uvw, datum, spacing):
        """switch the reference to its wind model 3 on a synthetic wind cube; uvw[ne, nn, nu, 3]"""
        gx, gy, gz = (np.ascontiguousarray(a, dtype=np.float64) for a in (gx, gy, gz))
        uvw = np.ascontiguousarray(uvw, dtype=np.float64)
        assert uvw.shape == (gx.size, gy.size, gz.size, 3)
        d, sp = np.asarray(datum, float), np.asarray(spacing, float)
        lib().tolref_set_wind_grid(self.h, gx.size, gy.size, gz.size, _dp(gx), _dp(gy), _dp(gz), _dp(uvw),
                                   _dp(d), _dp(sp))

    def wind(self):
        out = np.empty((12, self.ts + 1))
        lib().tolref_get_wind(self.h, _dp(out))
        return out

    def chi_d(self):
        return lib().tolref_chi_d(self.h)

    def ub_mask(self):
        """Indices of G the reference leaves UNINITIALISED (S10 boundary rows' dt column:
        reference src/problemS10.cpp:397,414 returns `Gs` without ever assigning it)."""
        if self.mission != "S10":
            return np.zeros(0, np.int64)
        return self.neG - 3 * self.nb + 3 * np.arange(self.nb)

    def write_json(self, x, F0, path):
        """reference problem::writeJSON (src/problem.cpp:1247-1365) for the state (x, F[0])"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        lib().tolref_write_json(self.h, _dp(x), C.c_double(F0), str(path).encode())

    def write_txt(self, x, F0, directory):
        """reference problem::writeTXT (src/problem.cpp:1371-1418): always writes snopt_output.txt into
        the current directory, so run it from `directory`; returns the file's path"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        L = lib()
        cwd = os.getcwd()
        os.chdir(directory)
        try:
            L.tolref_set_null_io(0)
            L.tolref_write_txt(self.h, _dp(x), C.c_double(F0))
        finally:
            L.tolref_set_null_io(1)
            os.chdir(cwd)
        return os.path.join(directory, "snopt_output.txt")

    def eval(self, x, full_callback=False):
        x = np.ascontiguousarray(x, dtype=np.float64)
        F = np.empty(self.neF)
        G = np.empty(self.neG)
        fn = lib().tolref_usrfun if full_callback else lib().tolref_eval
        fn(self.h, _dp(x), 1, _dp(F), 1, _dp(G))
        return F, G

    def eval_many(self, X, F, G, full_callback=False):
        fn = lib().tolref_usrfun_many if full_callback else lib().tolref_eval_many
        fn(self.h, X.shape[0], _dp(X), X.strides[0] // 8, _dp(F), F.strides[0] // 8, _dp(G),
           G.strides[0] // 8)
