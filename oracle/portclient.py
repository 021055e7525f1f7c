"""TEST INFRASTRUCTURE ONLY: ctypes client for oracle/libtoloracle.so, the plain-C restatement of the
reference's user-function path (oracle/fg_oracle.c).  Used by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs only; the product package never imports it."""
import ctypes as C
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "libtoloracle.so")

G7, S10 = 7, 10


class ToloProblem(C.Structure):
    _fields_ = [(k, C.c_int) for k in
                ("formulation", "ts", "numinp", "numstates", "numbounds", "wind_model")] + \
               [(k, C.c_double) for k in
                ("mm", "SS", "ee", "AR", "Cd0", "kT", "kp", "kv", "kdt", "xg", "yg", "rg", "chi_d")] + \
               [(k, C.c_int) for k in ("grid_ne", "grid_nn", "grid_nu")] + \
               [(k, C.POINTER(C.c_double)) for k in ("grid_x", "grid_y", "grid_z", "grid_v")] + \
               [("datum", C.c_double * 3), ("spacing", C.c_double * 3)]


_lib = None


def available():
    return os.path.exists(PORT_SO)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(PORT_SO)
        dp, ip, pp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(ToloProblem)
        L.tolo_dims.argtypes = [pp, ip, ip, ip]
        L.tolo_pattern.argtypes = [pp, ip, ip]
        L.tolo_pattern.restype = C.c_int
        L.tolo_eval.argtypes = [pp, dp, C.c_int, dp, C.c_int, dp]
        L.tolo_eval_many.argtypes = [pp, C.c_int, dp, C.c_long, dp, C.c_long, dp, C.c_long]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class PortProblem:
    """mission 'G7'|'S10'; ac = the reference's 15 aircraft values (only mm,SS,ee,AR,Cd0 = indices
    0,2,3,4,5 are on the path), gn = kT,kp,kv,ka,kdt, goal = xg,yg,zg,rg in NED."""

    def __init__(self, mission, ts, ac, gn, goal, wind_model=1, chi_d=None):
        form = {"G7": G7, "S10": S10}[mission]
        nb = 12 if form == G7 else 11
        if chi_d is None:
            # reference src/problemG7.cpp:524 with xi = yi = 0 (src/problem.cpp:111-112).  libm's atan2 (what the
            # reference calls), NOT numpy's: np.arctan2 is a vectorised routine of its own and differs from glibc
            # in the last bit for ~8 % of arguments
            chi_d = math.atan2(float(goal[1]) - 0.0, float(goal[0]) - 0.0) if form == G7 else 0.0
        self.p = ToloProblem(form, int(ts), 11, 8, nb, int(wind_model), ac[0], ac[2], ac[3], ac[4],
                             ac[5], gn[0], gn[1], gn[2], gn[4], goal[0], goal[1], goal[3], chi_d)
        self.mission, self.ts, self.nb = mission, int(ts), nb
        self._grid = None
        n, neF, neG = C.c_int(), C.c_int(), C.c_int()
        lib().tolo_dims(C.byref(self.p), C.byref(n), C.byref(neF), C.byref(neG))
        self.n, self.neF, self.neG = n.value, neF.value, neG.value

    def set_wind_grid(self, gx, gy, gz, v, datum, spacing):
        """wind model 3 on a wind cube: v[ne, nn, nu] (the reference interpolates the v component only)"""
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (gx, gy, gz, v)]
        assert arrs[3].shape == (arrs[0].size, arrs[1].size, arrs[2].size)
        self._grid = arrs  # keep alive
        self.p.wind_model = 3
        self.p.grid_ne, self.p.grid_nn, self.p.grid_nu = arrs[0].size, arrs[1].size, arrs[2].size
        self.p.grid_x, self.p.grid_y, self.p.grid_z, self.p.grid_v = (_dp(a) for a in arrs)
        self.p.datum[:] = [float(t) for t in datum]
        self.p.spacing[:] = [float(t) for t in spacing]

    def pattern(self):
        i = np.empty(self.neG, np.int32)
        j = np.empty(self.neG, np.int32)
        k = lib().tolo_pattern(C.byref(self.p), _ip(i), _ip(j))
        assert k == self.neG
        return i, j

    def ub_mask(self):
        if self.mission != "S10":
            return np.zeros(0, np.int64)
        return self.neG - 3 * self.nb + 3 * np.arange(self.nb)

    def eval(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        F, G = np.empty(self.neF), np.empty(self.neG)
        lib().tolo_eval(C.byref(self.p), _dp(x), 1, _dp(F), 1, _dp(G))
        return F, G

    def eval_many(self, X, F, G):
        lib().tolo_eval_many(C.byref(self.p), X.shape[0], _dp(X), X.strides[0] // 8, _dp(F),
                             F.strides[0] // 8, _dp(G), G.strides[0] // 8)
