"""Host-side mirror of the reference's DGModel / MPIStateArray / LSRK interface over libcmdg.

* ``MPIStateArray``  <- src/Arrays/MPIStateArrays.jl:46-172 (``data`` is the reference's
  ``Np x nstate x nelem`` array: a torch tensor of shape ``(nelem, nstate, Np)`` has the same bytes)
* ``DiscontinuousSpectralElementGrid`` <- src/Numerics/Mesh/Grids.jl:170-265 (device arrays only)
* ``DGModel``        <- src/Numerics/DGMethods/DGModel.jl:3-65, callable as ``:85-427``
* ``LSRK54CarpenterKennedy``, ``LSRK144NiegemannDiehlBusch``, ``solve``, ``dostep``
  <- src/Numerics/ODESolvers/LowStorageRungeKuttaMethod.jl, ODESolvers.jl:110-158

PyTorch is used for device memory and streams only; all arithmetic on the path is done by the
hand-written CUDA kernels in csrc/ through the C ABI.
"""
import ctypes as C
from fractions import Fraction as Fr

import numpy as np
import torch

from . import _lib
from . import balance_laws as bl


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class DiscontinuousSpectralElementGrid:
    """Device-resident grid arrays in the reference layout (1-based Int64 index arrays)."""

    def __init__(self, N, vgeo, sgeo, vmapM, vmapP, elemtobndy, D, nrealelem,
                 interiorelems=None, exteriorelems=None, vmapsend=None, vmaprecv=None,
                 nabrtorank=(), nabrtovmapsend=(), nabrtovmaprecv=(), nvertelem=0,
                 device="cuda", Imat=None, xi=None):
        dev = torch.device(device)
        self.N = int(N)
        self.Nq = self.N + 1
        self.Np = self.Nq ** 3
        self.Nfp = self.Nq ** 2
        tt = lambda a, dt=None: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev)
        self.vgeo = tt(vgeo)
        self.sgeo = tt(sgeo)
        self.FT = self.vgeo.dtype
        self.nelem = self.vgeo.shape[0]
        self.nrealelem = int(nrealelem)
        assert self.vgeo.shape == (self.nelem, 25, self.Np)
        assert self.sgeo.shape == (self.nelem, 6, self.Nfp, 5)
        self.vmapM = tt(vmapM, torch.int64)
        self.vmapP = tt(vmapP, torch.int64)
        self.elemtobndy = tt(elemtobndy, torch.int64)
        # D in Julia (column-major) memory order: element [a, b] at offset a + Nq*b
        self.D = tt(np.asarray(D).T.copy()).to(self.FT)
        if interiorelems is None:
            interiorelems = np.arange(1, self.nrealelem + 1)
            exteriorelems = np.zeros(0, dtype=np.int64)
        self.interiorelems = tt(np.asarray(interiorelems, dtype=np.int64), torch.int64)
        self.exteriorelems = tt(np.asarray(exteriorelems, dtype=np.int64), torch.int64)
        z = np.zeros(0, dtype=np.int64)
        self.vmapsend = tt(z if vmapsend is None else vmapsend, torch.int64)
        self.vmaprecv = tt(z if vmaprecv is None else vmaprecv, torch.int64)
        self.nabrtorank = [int(r) for r in nabrtorank]
        self.nabrtovmapsend = [(int(a), int(b)) for a, b in nabrtovmapsend]
        self.nabrtovmaprecv = [(int(a), int(b)) for a, b in nabrtovmaprecv]
        self.nvertelem = int(nvertelem)
        self.device = dev
        # grid.Imat[end] (Julia memory order) and the 1-D reference points, used by the ocean model
        self.Imat = None if Imat is None else tt(np.asarray(Imat).T.copy()).to(self.FT)
        self.xi = None if xi is None else np.asarray(xi, dtype=np.float64)


class MPIStateArray:
    def __init__(self, grid, nstate, data=None):
        self.grid = grid
        self.nstate = nstate
        if data is None:
            self.data = torch.zeros((grid.nelem, nstate, grid.Np), dtype=grid.FT, device=grid.device)
        else:
            # own copy (on a CPU device torch would otherwise alias the caller's NumPy buffer)
            self.data = torch.as_tensor(np.array(data, order="C", copy=True)).to(grid.device).to(grid.FT)
            assert self.data.shape == (grid.nelem, nstate, grid.Np)

    @property
    def realdata(self):
        return self.data[:self.grid.nrealelem]

    def similar(self):
        return MPIStateArray(self.grid, self.nstate)

    def weights(self):
        return self.grid.vgeo[:, 9, :]  # vgeo[:, _M, :] (create_states.jl:16-17)


def norm(Q, weighted=True):
    """sqrt(sum M Q^2) over real elements (MPIStateArrays.jl:583-604), rank-local part."""
    d = Q.realdata.double()
    w = Q.weights()[:Q.grid.nrealelem, None, :].double() if weighted else 1.0
    return float(torch.sqrt((d * d * w).sum()))


def euclidean_distance(A, B):
    d = (A.realdata - B.realdata).double()
    w = A.weights()[:A.grid.nrealelem, None, :].double()
    return float(torch.sqrt((d * d * w).sum()))


_NF1 = {bl.RusanovNumericalFlux: _lib.NF_RUSANOV, bl.CentralNumericalFluxFirstOrder: _lib.NF_CENTRAL,
        bl.RoeNumericalFlux: _lib.NF_ROE}


class DGModel:
    """``DGModel(balance_law, grid, nf1, nf2, nfgrad; state_auxiliary, ...)`` over libcmdg."""

    def __init__(self, balance_law, grid, numerical_flux_first_order,
                 numerical_flux_second_order, numerical_flux_gradient,
                 state_auxiliary=None, state_gradient_flux=None,
                 direction=None, diffusion_direction=None,
                 skip_zero_viscosity=False, write_aux_diagnostics=True, modeldata=None):
        if isinstance(balance_law, bl.HydrostaticBoussinesqModel):
            self._init_ocean(balance_law, grid, numerical_flux_first_order,
                             numerical_flux_second_order, numerical_flux_gradient,
                             state_auxiliary, state_gradient_flux, modeldata)
            return
        if not isinstance(balance_law, bl.AtmosModel):
            raise bl.UnsupportedModelError(
                f"balance law {type(balance_law).__name__} is not compiled into libcmdg")
        balance_law.validate()
        if type(numerical_flux_first_order) not in _NF1:
            raise bl.UnsupportedModelError(
                f"numerical flux {type(numerical_flux_first_order).__name__} is not supported")
        if not isinstance(numerical_flux_second_order, bl.CentralNumericalFluxSecondOrder) or \
                not isinstance(numerical_flux_gradient, bl.CentralNumericalFluxGradient):
            raise bl.UnsupportedModelError("second-order/gradient fluxes must be Central")
        if direction is not None and not isinstance(direction, bl.EveryDirection):
            raise bl.UnsupportedModelError("only direction = EveryDirection() is supported")
        self.balance_law = balance_law
        self.grid = grid
        self.numerical_flux_first_order = numerical_flux_first_order
        self.numerical_flux_second_order = numerical_flux_second_order
        self.numerical_flux_gradient = numerical_flux_gradient
        self.diffusion_direction = diffusion_direction or bl.EveryDirection()
        m = balance_law
        A, GF = m.number_states("Auxiliary"), m.number_states("GradientFlux")
        self.state_auxiliary = state_auxiliary or MPIStateArray(grid, A)
        assert self.state_auxiliary.nstate == A, "state_auxiliary has the wrong number of columns"
        self.state_gradient_flux = state_gradient_flux or MPIStateArray(grid, GF)
        L = _lib.lib()
        d = _lib.cmdg_desc()
        d.struct_bytes = C.sizeof(_lib.cmdg_desc)
        d.float_bytes = 8 if grid.FT == torch.float64 else 4
        d.dim, d.N = 3, grid.N
        d.nelem, d.nrealelem, d.nvertelem = grid.nelem, grid.nrealelem, grid.nvertelem
        d.model = _lib.MODEL_ATMOS_DRY
        d.nf_first = _NF1[type(numerical_flux_first_order)]
        d.nf_second = d.nf_gradient = _lib.NF_CENTRAL
        o = m.orientation
        d.orientation = (_lib.ORIENT_NONE if isinstance(o, bl.NoOrientation) else
                         _lib.ORIENT_FLAT if isinstance(o, bl.FlatOrientation) else _lib.ORIENT_SPHERICAL)
        if isinstance(m.ref_state, bl.HydrostaticState):
            d.ref_state, d.subtract_off = _lib.REF_HYDROSTATIC, int(m.ref_state.subtract_off)
        else:
            d.ref_state, d.subtract_off = _lib.REF_NONE, 0
        t = m.turbulence
        if isinstance(t, bl.ConstantKinematicViscosity):
            d.turbulence, d.turb_param, d.turb_with_divergence = _lib.TURB_CONSTANT_KINEMATIC, t.ν, int(t.with_divergence)
        elif isinstance(t, bl.ConstantDynamicViscosity):
            d.turbulence, d.turb_param, d.turb_with_divergence = _lib.TURB_CONSTANT_DYNAMIC, t.ρν, int(t.with_divergence)
        elif isinstance(t, bl.SmagorinskyLilly):
            d.turbulence, d.turb_param = _lib.TURB_SMAGORINSKY, t.C_smag
        else:
            raise bl.UnsupportedModelError(f"turbulence closure {type(t).__name__} is not supported")
        d.sources = 0
        for s in m.source:
            d.sources |= {bl.Gravity: _lib.SRC_GRAVITY, bl.Coriolis: _lib.SRC_CORIOLIS,
                          bl.HeldSuarezForcing: _lib.SRC_HELD_SUAREZ,
                          bl.RayleighSponge: _lib.SRC_RAYLEIGH_SPONGE}[type(s)]
            if isinstance(s, bl.RayleighSponge):
                d.sponge_z_max, d.sponge_z_sponge = s.z_max, s.z_sponge
                d.sponge_alpha_max, d.sponge_gamma = s.α_max, s.γ
                for i in range(3):
                    d.sponge_u_relax[i] = s.u_relaxation[i]
        d.diffusion_direction = (_lib.DIR_HORIZONTAL if isinstance(self.diffusion_direction, bl.HorizontalDirection)
                                 else _lib.DIR_EVERY)
        if isinstance(m.hyperdiffusion, bl.DryBiharmonic):
            d.hyperdiffusion, d.hyper_tau = _lib.HYPER_DRY_BIHARMONIC, float(m.hyperdiffusion.τ_timescale)
        else:
            d.hyperdiffusion = _lib.HYPER_NONE
        nt = len(m.tracers.δ_χ) if isinstance(m.tracers, bl.NTracers) else 0
        if nt and isinstance(numerical_flux_first_order, bl.RoeNumericalFlux):
            raise bl.UnsupportedModelError("NTracers: Rusanov / Central first-order fluxes only")
        d.ntracers = nt
        for i in range(nt):
            d.tracer_delta_chi[i] = m.tracers.δ_χ[i]
        d.skip_zero_viscosity = int(skip_zero_viscosity)
        d.write_aux_diagnostics = int(write_aux_diagnostics)
        d.nbc = len(m.boundaryconditions)
        for i, bc in enumerate(m.boundaryconditions):
            d.bc_kind[i] = _lib.BC_FREESLIP if isinstance(bc.momentum.drag, bl.FreeSlip) else _lib.BC_NOSLIP
        d.nstate, d.naux = m.number_states("Prognostic"), A
        d.ngrad, d.ngradflux = m.number_states("Gradient"), GF
        p = m.param_set
        d.R_d, d.cp_d, d.cv_d, d.T_0 = p.R_d, p.cp_d, p.cv_d, p.T_0
        d.MSLP, d.grav, d.Omega, d.inv_Pr_turb = p.MSLP, p.grav, p.Omega, p.inv_Pr_turb
        d.day = p.day
        self._desc = d
        self._h = C.c_void_p()
        _lib.check(L.cmdg_create(C.byref(d), C.byref(self._h)))
        g = grid
        nn = len(g.nabrtorank)
        ranks = (C.c_int32 * max(nn, 1))(*g.nabrtorank)
        sr = (C.c_int64 * max(2 * nn, 1))(*[v for ab in g.nabrtovmapsend for v in ab])
        rr = (C.c_int64 * max(2 * nn, 1))(*[v for ab in g.nabrtovmaprecv for v in ab])
        _lib.check(L.cmdg_bind_grid(
            self._h, _ptr(g.vgeo), _ptr(g.sgeo), _ptr(g.vmapM), _ptr(g.vmapP), _ptr(g.elemtobndy),
            _ptr(g.D), _ptr(g.interiorelems), g.interiorelems.numel(), _ptr(g.exteriorelems),
            g.exteriorelems.numel(), _ptr(g.vmapsend), g.vmapsend.numel(), _ptr(g.vmaprecv),
            g.vmaprecv.numel(), ranks, sr, rr, nn), self._h)
        _lib.check(L.cmdg_bind_state(self._h, _ptr(self.state_auxiliary.data),
                                     _ptr(self.state_gradient_flux.data)), self._h)

    def _init_ocean(self, m, grid, nf1, nf2, nfg, state_auxiliary, state_gradient_flux, modeldata):
        m.validate()
        if type(nf1) not in (bl.RusanovNumericalFlux, bl.CentralNumericalFluxFirstOrder):
            raise bl.UnsupportedModelError("HBModel supports Rusanov / Central first-order fluxes")
        if not isinstance(nf2, bl.CentralNumericalFluxSecondOrder) or \
                not isinstance(nfg, bl.CentralNumericalFluxGradient):
            raise bl.UnsupportedModelError("second-order/gradient fluxes must be Central")
        if modeldata is None or "vert_filter" not in modeldata or "exp_filter" not in modeldata:
            raise ValueError("HBModel needs modeldata = dict(vert_filter=..., exp_filter=...)")
        if grid.Imat is None:
            raise ValueError("the grid has no Imat (stack-integral operator)")
        self.balance_law, self.grid = m, grid
        self.numerical_flux_first_order = nf1
        self.diffusion_direction = bl.EveryDirection()
        self.modeldata = modeldata
        self.state_auxiliary = state_auxiliary or MPIStateArray(grid, 8)
        self.state_gradient_flux = state_gradient_flux or MPIStateArray(grid, 10)
        L = _lib.lib()
        d = _lib.cmdg_desc()
        d.struct_bytes = C.sizeof(_lib.cmdg_desc)
        d.float_bytes = 8 if grid.FT == torch.float64 else 4
        d.dim, d.N = 3, grid.N
        d.nelem, d.nrealelem, d.nvertelem = grid.nelem, grid.nrealelem, grid.nvertelem
        d.model = _lib.MODEL_HB
        d.nf_first = _NF1[type(nf1)]
        d.nf_second = d.nf_gradient = _lib.NF_CENTRAL
        d.nstate, d.naux, d.ngrad, d.ngradflux = 4, 8, 5, 10
        self._desc = d
        self._h = C.c_void_p()
        _lib.check(L.cmdg_create(C.byref(d), C.byref(self._h)))
        o = _lib.cmdg_ocean_desc()
        o.struct_bytes = C.sizeof(_lib.cmdg_ocean_desc)
        bcs = m.problem.boundary_conditions
        o.nbc = len(bcs)
        for i, bc in enumerate(bcs):
            o.bc_velocity[i], o.bc_temperature[i] = bl.ocean_bc_codes(bc)
        p = m.problem
        o.grav, o.rho0, o.ch, o.cz, o.alphaT = m.param_set.grav, m.ρₒ, m.cʰ, m.cᶻ, m.αᵀ
        o.nuh, o.nuz, o.kappah, o.kappaz, o.kappac = m.νʰ, m.νᶻ, m.κʰ, m.κᶻ, m.κᶜ
        o.f0, o.beta = m.fₒ, m.β
        o.Lx, o.Ly, o.H, o.tau0, o.lambda_r, o.thetaE = p.Lˣ, p.Lʸ, p.H, p.τₒ, p.λʳ, p.θᴱ
        _lib.check(L.cmdg_set_ocean_model(self._h, C.byref(o)), self._h)
        g = grid
        nn = len(g.nabrtorank)
        ranks = (C.c_int32 * max(nn, 1))(*g.nabrtorank)
        sr = (C.c_int64 * max(2 * nn, 1))(*[v for ab in g.nabrtovmapsend for v in ab])
        rr = (C.c_int64 * max(2 * nn, 1))(*[v for ab in g.nabrtovmaprecv for v in ab])
        _lib.check(L.cmdg_bind_grid(
            self._h, _ptr(g.vgeo), _ptr(g.sgeo), _ptr(g.vmapM), _ptr(g.vmapP), _ptr(g.elemtobndy),
            _ptr(g.D), _ptr(g.interiorelems), g.interiorelems.numel(), _ptr(g.exteriorelems),
            g.exteriorelems.numel(), _ptr(g.vmapsend), g.vmapsend.numel(), _ptr(g.vmaprecv),
            g.vmaprecv.numel(), ranks, sr, rr, nn), self._h)
        # filter matrices / Imat as device arrays in Julia (column-major) memory order
        dev = grid.device
        self._Fc = torch.as_tensor(np.ascontiguousarray(modeldata["vert_filter"].filter_matrix.T)).to(dev).to(grid.FT)
        self._Fe = torch.as_tensor(np.ascontiguousarray(modeldata["exp_filter"].filter_matrix.T)).to(dev).to(grid.FT)
        _lib.check(L.cmdg_bind_ocean_operators(self._h, _ptr(self._Fc), _ptr(self._Fe), _ptr(g.Imat)), self._h)
        _lib.check(L.cmdg_bind_state(self._h, _ptr(self.state_auxiliary.data),
                                     _ptr(self.state_gradient_flux.data)), self._h)

    # -- (dg::DGModel)(tendency, Q, param, t, alpha, beta) / (...; increment) ------------
    def __call__(self, tendency, Q, param=None, t=0.0, α=1.0, β=0.0, increment=None):
        if increment is not None:
            α, β = 1.0, (1.0 if increment else 0.0)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().cmdg_tendency(self._h, _ptr(tendency.data), _ptr(Q.data), float(t),
                                            float(α), float(β), C.c_void_p(st)), self._h)
        # the reference returns after checked_wait (DGModel.jl:426)
        torch.cuda.current_stream().synchronize()

    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        buf = C.create_string_buffer(unique_id, 128)
        _lib.check(_lib.lib().cmdg_comm_init(self._h, buf, rank, nranks), self._h)

    def kernel_launches(self):
        return int(_lib.lib().cmdg_kernel_launches(self._h))

    # begin_ghost_exchange! + end_ghost_exchange! (MPIStateArrays.jl:411-483)
    def ghost_exchange(self, arr):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        L = _lib.lib()
        _lib.check(L.cmdg_exchange_begin(self._h, _ptr(arr.data), arr.nstate, st), self._h)
        _lib.check(L.cmdg_exchange_end(self._h, _ptr(arr.data), arr.nstate, st), self._h)
        torch.cuda.current_stream().synchronize()

    # -- Filters.apply!(Q, target, grid, filter; state_auxiliary, direction) ------------------
    def _filter_args(self, target, filter, direction):
        if isinstance(target, bl.AtmosFilterPerturbations):
            kind, mask = _lib.FILTER_ATMOS_PERTURBATIONS, 0x1f
        elif isinstance(target, bl.FilterIndices):
            kind, mask = _lib.FILTER_INDICES, sum(1 << (i - 1) for i in target.I)
        else:
            raise bl.UnsupportedModelError(f"filter target {type(target).__name__} is not supported")
        if not hasattr(filter, "filter_matrix"):
            raise bl.UnsupportedModelError(f"filter {type(filter).__name__} is not supported")
        d = {bl.EveryDirection: _lib.DIR_EVERY, bl.HorizontalDirection: _lib.DIR_HORIZONTAL,
             bl.VerticalDirection: _lib.DIR_VERTICAL}[type(direction or bl.EveryDirection())]
        # device copy in Julia (column-major) memory order, as filter.filter_matrices[i] is
        W = torch.as_tensor(np.ascontiguousarray(np.asarray(filter.filter_matrix).T)).to(self.grid.device).to(self.grid.FT)
        return kind, mask, W, d

    def apply_filter(self, Q, target, filter, direction=None):
        kind, mask, W, d = self._filter_args(target, filter, direction)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.lib().cmdg_filter_apply(self._h, _ptr(Q.data), Q.nstate, kind, mask, _ptr(W),
                                                _ptr(W), d, st), self._h)
        torch.cuda.current_stream().synchronize()

    def set_step_filter(self, target, filter=None, direction=None):
        """Per-step `cbfilter` callback of the GCM drivers, run inside cmdg_lsrk_steps
        (``target=None`` removes it)."""
        if target is None:
            _lib.check(_lib.lib().cmdg_set_step_filter(self._h, -1, 0, None, None, 0), self._h)
            return
        kind, mask, W, d = self._filter_args(target, filter, direction)
        _lib.check(_lib.lib().cmdg_set_step_filter(self._h, kind, mask, _ptr(W), _ptr(W), d), self._h)

    # -- courant(local_courant, dg, m, Q, dt, simtime, direction) -----------------------------
    def courant(self, local_courant, Q, Δt, direction=None):
        """``local_courant`` in {"advective", "nondiffusive", "diffusive"} (the functions of
        src/Atmos/Model/courant.jl); returns the rank-local maximum."""
        kind = {"advective": _lib.COURANT_ADVECTIVE, "nondiffusive": _lib.COURANT_NONDIFFUSIVE,
                "diffusive": _lib.COURANT_DIFFUSIVE}[local_courant]
        d = {bl.EveryDirection: _lib.DIR_EVERY, bl.HorizontalDirection: _lib.DIR_HORIZONTAL,
             bl.VerticalDirection: _lib.DIR_VERTICAL}[type(direction or bl.EveryDirection())]
        out = C.c_double(0.0)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.lib().cmdg_courant(self._h, _ptr(Q.data), _ptr(self.grid.vgeo), float(Δt), kind, d,
                                           C.byref(out), st), self._h)
        return out.value

    # -- check_for_crashes (MPIStateArrays.jl:910-935): NaN/Inf on this rank, and on any rank ------------
    def check_for_crashes(self, Q, raise_on_failure=True):
        """Returns (local_bad, any_bad).  With ``raise_on_failure`` a rank whose own state is not finite raises
        ``FloatingPointError``; the other ranks raise ``ErrorOnRemoteNode`` as the reference does."""
        lb, ab = C.c_int32(0), C.c_int32(0)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.lib().cmdg_check_for_crashes(self._h, _ptr(Q.data), Q.nstate, C.byref(lb), C.byref(ab), st),
                   self._h)
        if raise_on_failure and lb.value:
            raise FloatingPointError("non-finite values in the prognostic state on this rank")
        if raise_on_failure and ab.value:
            raise ErrorOnRemoteNode("another rank reported non-finite values")
        return bool(lb.value), bool(ab.value)

    def set_timing(self, enable=True):
        _lib.check(_lib.lib().cmdg_set_timing(self._h, int(enable)), self._h)

    def last_kernel_ms(self):
        n = C.c_int64(0)
        ms = _lib.lib().cmdg_last_kernel_ms(self._h, C.byref(n))
        return float(ms), int(n.value)

    def kernel_class_ms(self):
        """Device ms and launch count per kernel class of the last fused-stepper call."""
        out = {}
        for i, name in enumerate(("tendency", "gradient", "hyper_divergence", "hyper_flux", "tracer_gradient",
                                  "tracer_tendency", "hb_filter", "hb_column")):
            n = C.c_int64(0)
            ms = _lib.lib().cmdg_kernel_class_ms(self._h, i, C.byref(n))
            out[name] = (float(ms), int(n.value))
        return out

    def sync(self):
        _lib.check(_lib.lib().cmdg_sync(self._h), self._h)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().cmdg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ErrorOnRemoteNode(RuntimeError):
    """``ErrorOnRemoteNode`` (src/Arrays/MPIStateArrays.jl:17)."""


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _lib.check(_lib.lib().cmdg_comm_unique_id(buf))
    return buf.raw


# ---------------------------------------------------------------------------------------
# Low-storage RK
# ---------------------------------------------------------------------------------------
_LSRK54 = (
    (Fr(0), Fr(-567301805773, 1357537059087), Fr(-2404267990393, 2016746695238),
     Fr(-3550918686646, 2091501179385), Fr(-1275806237668, 842570457699)),
    (Fr(1432997174477, 9575080441755), Fr(5161836677717, 13612068292357),
     Fr(1720146321549, 2090206949498), Fr(3134564353537, 4481467310338),
     Fr(2277821191437, 14882151754819)),
    (Fr(0), Fr(1432997174477, 9575080441755), Fr(2526269341429, 6820363962896),
     Fr(2006345519317, 3224310063776), Fr(2802321613138, 2924317926251)),
)
_LSRK144 = (
    (0.0, -0.7188012108672410, -0.7785331173421570, -0.0053282796654044, -0.8552979934029281,
     -3.9564138245774565, -1.5780575380587385, -2.0837094552574054, -0.7483334182761610,
     -0.7032861106563359, 0.0013917096117681, -0.0932075369637460, -0.9514200470875948,
     -7.1151571693922548),
    (0.0367762454319673, 0.3136296607553959, 0.1531848691869027, 0.0030097086818182,
     0.3326293790646110, 0.2440251405350864, 0.3718879239592277, 0.6204126221582444,
     0.1524043173028741, 0.0760894927419266, 0.0077604214040978, 0.0024647284755382,
     0.0780348340049386, 5.5059777270269628),
    (0.0, 0.0367762454319673, 0.1249685262725025, 0.2446177702277698, 0.2476149531070420,
     0.2969311120382472, 0.3978149645802642, 0.5270854589440328, 0.6981269994175695,
     0.8190890835352128, 0.8527059887098624, 0.8604711817462826, 0.8627060376969976,
     0.8734213127600976),
)


class LowStorageRungeKutta2N:
    def __init__(self, rhs, RKA, RKB, RKC, Q, dt=0.0, t0=0.0):
        FT = np.float64 if Q.data.dtype == torch.float64 else np.float32
        conv = lambda xs: tuple(float(FT(x.numerator / x.denominator if isinstance(x, Fr) else x)) for x in xs)
        self.rhs = rhs
        self.RKA, self.RKB, self.RKC = conv(RKA), conv(RKB), conv(RKC)
        self.dt, self.t, self.steps = float(dt), float(t0), 0
        self.dQ = Q.similar()
        ns = len(self.RKA)
        arr = C.c_double * ns
        self._a, self._b, self._c = arr(*self.RKA), arr(*self.RKB), arr(*self.RKC)

    # dostep! with separate tendency and update kernels, exactly the reference's call sequence
    def dostep_unfused(self, Q, time):
        L = _lib.lib()
        ns = len(self.RKA)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for s in range(ns):
            self.rhs(self.dQ, Q, None, time + self.RKC[s] * self.dt, increment=True)
            _lib.check(L.cmdg_lsrk_update(self.rhs._h, _ptr(self.dQ.data), _ptr(Q.data),
                                          self.RKA[(s + 1) % ns], self.RKB[s], self.dt, st), self.rhs._h)
        torch.cuda.current_stream().synchronize()

    # dostep!: one fused kernel per stage (cmdg_lsrk_steps)
    def dostep(self, Q, time, nsteps=1):
        L = _lib.lib()
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(L.cmdg_lsrk_steps(self.rhs._h, _ptr(Q.data), _ptr(self.dQ.data), float(time),
                                     self.dt, len(self.RKA), self._a, self._b, self._c,
                                     int(nsteps), st), self.rhs._h)

    # the same through HOST buffers (cmdg_lsrk_steps_host): realview(Q) as a pinned CPU tensor
    def dostep_host(self, Q_host, time, nsteps=1):
        assert Q_host.device.type == "cpu" and Q_host.is_contiguous()
        _lib.check(_lib.lib().cmdg_lsrk_steps_host(
            self.rhs._h, C.c_void_p(Q_host.data_ptr()), float(time), self.dt, len(self.RKA),
            self._a, self._b, self._c, int(nsteps)), self.rhs._h)

    def general_dostep(self, Q, timeend, adjustfinalstep=True, fused=True):
        time, dt = self.t, self.dt
        final = False
        if adjustfinalstep and time + dt > timeend:
            orig, self.dt, final = dt, timeend - time, True
        (self.dostep if fused else self.dostep_unfused)(Q, time)
        if not final:
            self.t = time + dt
        else:
            self.dt, self.t = orig, timeend
        return self.t


def LSRK54CarpenterKennedy(rhs, Q, dt=0.0, t0=0.0):
    return LowStorageRungeKutta2N(rhs, *_LSRK54, Q, dt=dt, t0=t0)


def LSRK144NiegemannDiehlBusch(rhs, Q, dt=0.0, t0=0.0):
    return LowStorageRungeKutta2N(rhs, *_LSRK144, Q, dt=dt, t0=t0)


def solve(Q, solver, timeend=float("inf"), numberofsteps=0, adjustfinalstep=True, fused=True,
          callbacks=()):
    """``solve!`` (ODESolvers.jl:110-158).  Without callbacks and with a fixed number of steps the
    whole loop runs inside one cmdg_lsrk_steps call."""
    assert np.isfinite(timeend) or numberofsteps > 0
    if fused and not callbacks and numberofsteps > 0 and not np.isfinite(timeend):
        solver.dostep(Q, solver.t, nsteps=numberofsteps)
        solver.t += numberofsteps * solver.dt
        solver.steps = numberofsteps
        torch.cuda.current_stream().synchronize()
        return solver.t
    step, time = 0, solver.t
    while time < timeend:
        step += 1
        solver.steps = step
        time = solver.general_dostep(Q, timeend, adjustfinalstep, fused)
        for cb in callbacks:
            cb(solver, Q, time)
        if step == numberofsteps:
            break
    torch.cuda.current_stream().synchronize()
    return solver.t
