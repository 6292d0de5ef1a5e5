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
                                         "nf_first")] + [("bc_kind", C.c_int32 * 6)] + \
               [(n, C.c_int32) for n in ("second_order", "turbulence", "with_divergence",
                                         "horizontal_diffusion", "a_Delta", "ngradflux", "held_suarez",
                                         "sponge")] + \
               [(n, C.c_double) for n in ("turb_param", "inv_Pr_turb", "day", "sponge_z_max",
                                          "sponge_z_sponge", "sponge_alpha_max", "sponge_gamma")] + \
               [("sponge_u", C.c_double * 3)]


_lib = None
_lib_ld = None
_SO_LD = os.path.join(_HERE, "c", "libdgref_ld.so")


def lib_ld():
    """The same C code built in x87 extended precision (`real` = long double; 64-bit mantissa)."""
    global _lib_ld
    if _lib_ld is None:
        if not os.path.exists(_SO_LD):
            raise RuntimeError(f"{_SO_LD} missing: run `make -C oracle/c`")
        _lib_ld = C.CDLL(_SO_LD)
        _lib_ld.ref_real_bytes.restype = C.c_int
        assert _lib_ld.ref_real_bytes() == np.dtype(np.longdouble).itemsize == 16, "x86-64 long double expected"
        _lib_ld.ref_set_num_threads.argtypes = [C.c_int]
    return _lib_ld


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


def params_from_model(model, nf="rusanov", second_order=False, diffusion_direction="every"):
    """ref_params of an oracle DryAtmosModel.  `second_order`: run the gradient pass and the viscous
    fluxes (the reference always does; False = the GPU arm's skip_zero_viscosity for nu = 0)."""
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
    P.second_order = int(second_order)
    k = model.turbulence
    P.turbulence = {"constant_kinematic": 0, "constant_dynamic": 1, "smagorinsky": 2}[k[0]]
    P.turb_param = float(k[1])
    P.with_divergence = int(bool(k[2])) if len(k) > 2 else 0
    P.horizontal_diffusion = int(diffusion_direction == "horizontal")
    P.a_Delta = -1 if model.a_Δ is None else model.a_Δ
    P.ngradflux = model.GF
    P.inv_Pr_turb, P.day = float(ps.inv_Pr_turb), float(ps.day)
    P.held_suarez = int("held_suarez" in model.sources)
    for s in model.sources:
        if isinstance(s, tuple) and s[0] == "rayleigh_sponge":
            P.sponge = 1
            P.sponge_z_max, P.sponge_z_sponge, P.sponge_alpha_max = float(s[1]), float(s[2]), float(s[3])
            for i in range(3):
                P.sponge_u[i] = float(s[4][i])
            P.sponge_gamma = float(s[5])
    assert getattr(model, "hyperdiffusion", None) is None and not getattr(model, "NT", 0), \
        "the C twin restates neither hyperdiffusion nor tracers"
    return P


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class CRefDG:
    """Reference-schedule tendency / LSRK steps on one rank's oracle grid (float64, N = 4)."""

    def __init__(self, model, grid, nf="rusanov", second_order=False, diffusion_direction="every"):
        assert grid.FT == np.float64 and grid.N == (4, 4, 4)
        self._init(params_from_model(model, nf, second_order, diffusion_direction), grid, grid.D[0], grid.nelem)

    def _init(self, P, g, D, nelem):
        self.P, self.g = P, g
        self.D = np.ascontiguousarray(D, dtype=np.float64)
        self.elems = np.arange(1, g.nreal + 1, dtype=np.int64)
        # state_gradient_flux of the twin (written by the gradient pass when P.second_order)
        self.gradflux = np.zeros((nelem if P.second_order else 1, max(int(P.ngradflux), 1), 125))
        for name in ("vgeo", "sgeo", "vmapM", "vmapP", "elemtobndy"):
            a = getattr(g, name)
            assert a.flags["C_CONTIGUOUS"] and a.dtype == (np.float64 if name in ("vgeo", "sgeo") else np.int64), name

    @classmethod
    def from_arrays(cls, P, vgeo, sgeo, vmapM, vmapP, elemtobndy, D, nreal):
        """Twin over raw arrays in the reference layout (vgeo [nelem][25][Np], sgeo [nelem][6][Nfp][5],
        1-based Int64 vmaps, elemtobndy [nelem][6], row-major D) with a ready ``ref_params``."""
        import types
        g = types.SimpleNamespace(vgeo=vgeo, sgeo=sgeo, vmapM=vmapM, vmapP=vmapP, elemtobndy=elemtobndy,
                                  nreal=int(nreal))
        self = cls.__new__(cls)
        self._init(P, g, D, vgeo.shape[0])
        return self

    def tendency_extended(self, Q, aux):
        """One evaluation (alpha = 1, beta = 0) of the same schedule in extended precision on the same
        Float64 inputs; returns the tendency as np.longdouble (nelem, 5, Np).  The yardstick for
        ill-conditioned states: how far is a Float64 evaluation from the exactly rounded one?"""
        L = lib_ld()
        try:
            L.ref_set_num_threads(len(os.sched_getaffinity(0)))
        except AttributeError:
            pass
        g = self.g
        ld = lambda a: np.ascontiguousarray(a, dtype=np.longdouble)
        if not hasattr(self, "_ld"):
            self._ld = dict(vgeo=ld(g.vgeo), sgeo=ld(g.sgeo), D=ld(self.D))
        x = self._ld
        Ql, al, gl = ld(Q), ld(aux), ld(self.gradflux)
        dQ = np.zeros_like(Ql)
        L.ref_tendency(C.byref(self.P), _p(dQ), _p(Ql), _p(al), _p(gl), _p(x["vgeo"]), _p(x["sgeo"]),
                       _p(g.vmapM), _p(g.vmapP), _p(g.elemtobndy), _p(x["D"]), _p(self.elems),
                       C.c_int64(g.nreal), C.c_longdouble(1.0), C.c_longdouble(0.0))
        return dQ

    def tendency(self, dQ, Q, aux, alpha=1.0, beta=0.0):
        g = self.g
        lib().ref_tendency(C.byref(self.P), _p(dQ), _p(Q), _p(aux), _p(self.gradflux), _p(g.vgeo), _p(g.sgeo),
                           _p(g.vmapM), _p(g.vmapP), _p(g.elemtobndy), _p(self.D), _p(self.elems),
                           C.c_int64(g.nreal), C.c_double(alpha), C.c_double(beta))

    def lsrk_steps(self, Q, dQ, aux, dt, rka, rkb, nsteps):
        g = self.g
        a = np.ascontiguousarray(rka, dtype=np.float64)
        b = np.ascontiguousarray(rkb, dtype=np.float64)
        lib().ref_lsrk_steps(C.byref(self.P), _p(Q), _p(dQ), _p(aux), _p(self.gradflux), _p(g.vgeo), _p(g.sgeo),
                             _p(g.vmapM), _p(g.vmapP), _p(g.elemtobndy), _p(self.D), _p(self.elems),
                             C.c_int64(g.nreal), C.c_double(dt), C.c_int(len(a)), _p(a), _p(b),
                             C.c_int64(nsteps))


# ---------------------------------------------------------------------------------------
# ocean HydrostaticBoussinesqModel twin (oracle/c/hb_ref.c)
# ---------------------------------------------------------------------------------------
_SO_HB = os.path.join(_HERE, "c", "libhbref.so")
_lib_hb = None


class hb_params(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("grav", "rho0", "ch", "cz", "alphaT", "nuh", "nuz", "kappah", "kappaz",
                                          "kappac", "f0", "beta", "Ly", "tau0", "lambda_r", "thetaE")] + \
               [("bc_vel", C.c_int32 * 6), ("bc_temp", C.c_int32 * 6), ("nf_first", C.c_int32), ("nvert", C.c_int32)]


HB_VEL = {"noslip": 1, "freeslip": 2, "penetrable_freeslip": 3, "kinematic_stress": 4}
HB_TEMP = {"insulating": 1, "temperature_flux": 2}


def lib_hb():
    global _lib_hb
    if _lib_hb is None:
        if not os.path.exists(_SO_HB):
            raise RuntimeError(f"{_SO_HB} missing: run `make -C oracle/c` (done by __graft_entry__.build())")
        _lib_hb = C.CDLL(_SO_HB)
        _lib_hb.hb_set_num_threads.argtypes = [C.c_int]
    return _lib_hb


def hb_params_from(grav, rho0, ch, cz, alphaT, nuh, nuz, kappah, kappaz, kappac, f0, beta, Ly, tau0, lambda_r,
                   thetaE, bcs, nvert, nf="rusanov"):
    """hb_params from plain numbers; ``bcs``: per boundary tag a (velocity, temperature) pair of names."""
    P = hb_params()
    for k, v in dict(grav=grav, rho0=rho0, ch=ch, cz=cz, alphaT=alphaT, nuh=nuh, nuz=nuz, kappah=kappah,
                     kappaz=kappaz, kappac=kappac, f0=f0, beta=beta, Ly=Ly, tau0=tau0, lambda_r=lambda_r,
                     thetaE=thetaE).items():
        setattr(P, k, float(v))
    for i, (vel, temp) in enumerate(bcs):
        P.bc_vel[i], P.bc_temp[i] = HB_VEL[vel], HB_TEMP[temp]
    P.nf_first = {"rusanov": 0, "central": 1}[nf]
    P.nvert = int(nvert)
    return P


def hb_params_from_model(model, nvert, nf="rusanov"):
    """hb_params of an oracle HBModel over an OceanGyre problem."""
    pr = model.problem
    return hb_params_from(model.grav, model.rho0, model.ch, model.cz, model.alphaT, model.nuh, model.nuz,
                          model.kappah, model.kappaz, model.kappac, model.f0, model.beta, pr.Ly, pr.tau0,
                          pr.lambda_r, pr.thetaE, model.bcs, nvert, nf)


class CRefHB:
    """Reference-schedule HBModel tendency / LSRK steps on one rank (float64, N = 4, no ghost elements)."""

    def __init__(self, P, vgeo, sgeo, vmapM, vmapP, elemtobndy, D, Imat, Fc, Fe, nreal):
        self.P, self.nreal = P, int(nreal)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)
        self.vgeo, self.sgeo, self.D, self.Imat = f64(vgeo), f64(sgeo), f64(D), f64(Imat)
        eye = np.eye(5)
        self.Fc, self.Fe = f64(eye if Fc is None else Fc), f64(eye if Fe is None else Fe)
        self.vmapM, self.vmapP, self.bnd = i64(vmapM), i64(vmapP), i64(elemtobndy)
        assert self.vgeo.shape[0] == self.nreal, "the ocean twin runs one rank without ghost elements"
        assert self.nreal % P.nvert == 0
        self.gradflux = np.zeros((self.nreal, 10, 125))

    @classmethod
    def from_grid(cls, model, g, nf="rusanov"):
        return cls(hb_params_from_model(model, g.topology.stacksize, nf), g.vgeo, g.sgeo, g.vmapM, g.vmapP,
                   g.elemtobndy, g.D[0], g.Imat[2], model.vert_filter, model.exp_filter, g.nreal)

    def _args(self):
        return (_p(self.vgeo), _p(self.sgeo), _p(self.vmapM), _p(self.vmapP), _p(self.bnd), _p(self.D),
                _p(self.Imat), _p(self.Fc), _p(self.Fe), C.c_int64(self.nreal))

    def tendency(self, dQ, Q, aux, alpha=1.0, beta=0.0):
        """In place: Q is filtered, aux gets w / pkin / wz0, self.gradflux the gradient flux."""
        lib_hb().hb_ref_tendency(C.byref(self.P), _p(dQ), _p(Q), _p(aux), _p(self.gradflux), *self._args(),
                                 C.c_double(alpha), C.c_double(beta))

    def lsrk_steps(self, Q, dQ, aux, dt, rka, rkb, nsteps):
        a = np.ascontiguousarray(rka, dtype=np.float64)
        b = np.ascontiguousarray(rkb, dtype=np.float64)
        lib_hb().hb_ref_lsrk_steps(C.byref(self.P), _p(Q), _p(dQ), _p(aux), _p(self.gradflux), *self._args(),
                                   C.c_double(dt), C.c_int(len(a)), _p(a), _p(b), C.c_int64(nsteps))


def use_all_cores_hb():
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib_hb().hb_set_num_threads(n)
    return n
