"""ctypes wrapper of oracle/c/libdgref.so (test infrastructure / CPU baseline only)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "c", "libdgref.so")


class ref_params(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("R_d", "cp_d", "cv_d", "T_0", "MSLP", "grav", "Omega")] + \
               [(n, C.c_int32) for n in ("naux", "a_Phi", "a_gradPhi", "a_ref_rho", "a_ref_p",
                                         "a_theta_v", "a_T", "subtract_off", "gravity", "coriolis",
                                         "nf_first")] + [("bc_kind", C.c_int32 * 6)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise RuntimeError(f"{_SO} missing: run `make -C oracle/c` (done by __graft_entry__.build())")
        _lib = C.CDLL(_SO)
        _lib.ref_num_threads.restype = C.c_int
        _lib.ref_set_num_threads.argtypes = [C.c_int]
    return _lib


def use_all_cores():
    """All host cores this process may run on, whatever OMP_NUM_THREADS says (torchrun sets it to 1)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().ref_set_num_threads(n)
    return n


def params_from_model(model, nf="rusanov"):
    ps = model.ps
    P = ref_params()
    P.R_d, P.cp_d, P.cv_d, P.T_0 = float(ps.R_d), float(ps.cp_d), float(ps.cv_d), float(ps.T_0)
    P.MSLP, P.grav, P.Omega = float(ps.MSLP), float(ps.grav), float(ps.Omega)
    P.naux = model.A
    P.a_Phi = -1 if model.a_Φ is None else model.a_Φ
    P.a_gradPhi = -1 if model.a_gradΦ is None else model.a_gradΦ.start
    P.a_ref_rho = -1 if model.a_ref is None else model.a_ref["ρ"]
    P.a_ref_p = -1 if model.a_ref is None else model.a_ref["p"]
    P.a_theta_v, P.a_T = model.a_θv, model.a_T
    P.subtract_off = int(model.subtract_off)
    P.gravity = int("gravity" in model.sources)
    P.coriolis = int("coriolis" in model.sources)
    P.nf_first = {"rusanov": 0, "central": 1}[nf]
    for i, b in enumerate(model.bcs):
        P.bc_kind[i] = 1 if b == "freeslip" else 2
    return P


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class CRefDG:
    """Reference-schedule tendency / LSRK steps on one rank's oracle grid (float64, N = 4)."""

    def __init__(self, model, grid, nf="rusanov"):
        assert grid.FT == np.float64 and grid.N == (4, 4, 4)
        self.P = params_from_model(model, nf)
        self.g = grid
        self.D = np.ascontiguousarray(grid.D[0])
        self.elems = np.arange(1, grid.nreal + 1, dtype=np.int64)

    def tendency(self, dQ, Q, aux, alpha=1.0, beta=0.0):
        g = self.g
        lib().ref_tendency(C.byref(self.P), _p(dQ), _p(Q), _p(aux), _p(g.vgeo), _p(g.sgeo),
                           _p(g.vmapM), _p(g.vmapP), _p(g.elemtobndy), _p(self.D), _p(self.elems),
                           C.c_int64(g.nreal), C.c_double(alpha), C.c_double(beta))

    def lsrk_steps(self, Q, dQ, aux, dt, rka, rkb, nsteps):
        g = self.g
        a = np.ascontiguousarray(rka, dtype=np.float64)
        b = np.ascontiguousarray(rkb, dtype=np.float64)
        lib().ref_lsrk_steps(C.byref(self.P), _p(Q), _p(dQ), _p(aux), _p(g.vgeo), _p(g.sgeo),
                             _p(g.vmapM), _p(g.vmapP), _p(g.elemtobndy), _p(self.D), _p(self.elems),
                             C.c_int64(g.nreal), C.c_double(dt), C.c_int(len(a)), _p(a), _p(b),
                             C.c_int64(nsteps))
