"""Host-side mirror of the reference's model/numerical-flux type tags for the hot path.

These records carry exactly what the reference encodes in Julia types and what libcmdg's
``cmdg_desc`` needs (include/cmdg.h):

* ``AtmosModel`` / ``AtmosPhysics`` / ``AtmosProblem``  <- src/Atmos/Model/AtmosModel.jl:110-377,
  src/Atmos/Model/problem.jl:13-37
* ``NoOrientation, FlatOrientation, SphericalOrientation`` <- src/Common/Orientations/Orientations.jl
* ``NoReferenceState, HydrostaticState, DecayingTemperatureProfile`` <- src/Atmos/Model/ref_state.jl:22-64
* ``ConstantDynamicViscosity, ConstantKinematicViscosity, SmagorinskyLilly``
  <- src/Common/TurbulenceClosures/TurbulenceClosures.jl:287-420
* ``Gravity, Coriolis`` <- src/Atmos/Model/tendencies_momentum.jl:62-92
* ``AtmosBC, Impenetrable, FreeSlip, NoSlip, Insulating`` <- src/Atmos/Model/boundaryconditions.jl:34-55
* ``RusanovNumericalFlux, CentralNumericalFluxFirstOrder, RoeNumericalFlux,
  CentralNumericalFluxSecondOrder, CentralNumericalFluxGradient``
  <- src/Numerics/DGMethods/NumericalFluxes.jl
* ``EarthParameterSet`` <- CLIMAParameters.jl 0.2.0 (un-vendored; values restated)

Anything outside the supported set raises ``UnsupportedModelError`` -- there is no fallback.
"""
from dataclasses import dataclass, field
from typing import Optional, Tuple


class UnsupportedModelError(NotImplementedError):
    pass


@dataclass(frozen=True)
class EarthParameterSet:
    gas_constant: float = 8.3144598
    molmass_dryair: float = 28.97e-3
    kappa_d: float = 2 / 7
    T_0: float = 273.16
    MSLP: float = 1.01325e5
    grav: float = 9.81
    Omega: float = 7.2921159e-5
    planet_radius: float = 6.371e6
    inv_Pr_turb: float = 3.0
    C_smag: float = 0.21
    T_surf_ref: float = 290.0
    T_min_ref: float = 220.0
    day: float = 86400.0

    @property
    def R_d(self):
        return self.gas_constant / self.molmass_dryair

    @property
    def cp_d(self):
        return self.R_d / self.kappa_d

    @property
    def cv_d(self):
        return self.cp_d - self.R_d


# --- orientations ----------------------------------------------------------
class NoOrientation:
    pass


class FlatOrientation:
    pass


class SphericalOrientation:
    pass


# --- reference states ------------------------------------------------------
class NoReferenceState:
    pass


@dataclass(frozen=True)
class DecayingTemperatureProfile:
    T_virt_surf: float = 290.0
    T_min_ref: float = 220.0
    H_t: float = 8.0e3


@dataclass(frozen=True)
class DryAdiabaticProfile:
    """``DryAdiabaticProfile{FT}(param_set, T_surface, T_min_ref)`` (TemperatureProfiles.jl:43-98)."""
    T_surface: float = 290.0
    T_min_ref: float = 220.0


@dataclass(frozen=True)
class HydrostaticState:
    virtual_temperature_profile: object
    relative_humidity: float = 0.0
    subtract_off: bool = True


# --- turbulence closures ---------------------------------------------------
@dataclass(frozen=True)
class ConstantDynamicViscosity:
    ρν: float = 0.0
    with_divergence: bool = False


@dataclass(frozen=True)
class ConstantKinematicViscosity:
    ν: float = 0.0
    with_divergence: bool = False


@dataclass(frozen=True)
class SmagorinskyLilly:
    C_smag: float = 0.21


# --- hyperdiffusion (src/Common/TurbulenceClosures/TurbulenceClosures.jl:780-848) ----------
class NoHyperDiffusion:
    pass


@dataclass(frozen=True)
class DryBiharmonic:
    """``DryBiharmonic{FT}(tau_timescale)``: fourth-order horizontal hyperdiffusion of u_h and h_tot
    with nu4 = (Delta_h / 2)^4 / 2 / tau.  Needs ``diffusion_direction = HorizontalDirection()``."""
    τ_timescale: float


# --- tracers (src/Atmos/Model/tracers.jl) -------------------------------------
class NoTracers:
    pass


@dataclass(frozen=True)
class NTracers:
    """``NTracers{N, FT}(delta_chi)``: N passive tracers with diffusivity ratios delta_chi (N <= 4)."""
    δ_χ: Tuple

    def __post_init__(self):
        object.__setattr__(self, "δ_χ", tuple(float(x) for x in self.δ_χ))


# --- sources ---------------------------------------------------------------
class Gravity:
    pass


class Coriolis:
    pass


class HeldSuarezForcing:
    """experiments/AtmosGCM/heldsuarez.jl:112-172 (tutorials/Atmos/heldsuarez.jl:45-118)."""


HeldSuarezForcingTutorial = HeldSuarezForcing


@dataclass(frozen=True)
class RayleighSponge:
    """src/Atmos/Model/tendencies_momentum.jl:104-137."""
    z_max: float
    z_sponge: float
    α_max: float
    u_relaxation: Tuple = (0.0, 0.0, 0.0)
    γ: float = 2.0


# --- boundary conditions ---------------------------------------------------
class FreeSlip:
    pass


class NoSlip:
    pass


@dataclass(frozen=True)
class Impenetrable:
    drag: object = field(default_factory=FreeSlip)


class Insulating:
    pass


@dataclass(frozen=True)
class AtmosBC:
    momentum: Impenetrable = field(default_factory=Impenetrable)
    energy: object = field(default_factory=Insulating)


# --- moisture --------------------------------------------------------------
class DryModel:
    pass


# --- numerical fluxes ------------------------------------------------------
class RusanovNumericalFlux:
    pass


class CentralNumericalFluxFirstOrder:
    pass


class RoeNumericalFlux:
    pass


class CentralNumericalFluxSecondOrder:
    pass


class CentralNumericalFluxGradient:
    pass


# --- directions ------------------------------------------------------------
class EveryDirection:
    pass


class HorizontalDirection:
    pass


class VerticalDirection:
    pass


@dataclass
class AtmosModel:
    """Dry, compressible, total-energy AtmosModel (the subset libcmdg compiles in)."""
    param_set: EarthParameterSet = field(default_factory=EarthParameterSet)
    orientation: object = field(default_factory=NoOrientation)
    ref_state: object = field(default_factory=NoReferenceState)
    turbulence: object = field(default_factory=ConstantDynamicViscosity)
    moisture: object = field(default_factory=DryModel)
    source: Tuple = ()
    boundaryconditions: Tuple = ()
    # anything else the reference's AtmosModel can hold is unsupported here
    hyperdiffusion: Optional[object] = None
    precipitation: Optional[object] = None
    radiation: Optional[object] = None
    tracers: Optional[object] = None
    turbconv: Optional[object] = None

    def number_states(self, kind):
        smag = isinstance(self.turbulence, SmagorinskyLilly)
        hyp = isinstance(self.hyperdiffusion, DryBiharmonic)
        nt = len(self.tracers.δ_χ) if isinstance(self.tracers, NTracers) else 0
        if kind == "Prognostic":
            return 5 + nt
        if kind == "Gradient":
            return (5 if smag else 4) + (4 if hyp else 0) + nt
        if kind == "GradientLaplacian":
            return 4 if hyp else 0
        if kind == "Hyperdiffusive":
            return 12 if hyp else 0
        if kind == "GradientFlux":
            return (10 if smag else 9) + 3 * nt
        if kind == "Auxiliary":
            c = 3
            if not isinstance(self.orientation, NoOrientation):
                c += 4
            if isinstance(self.ref_state, HydrostaticState):
                c += 7
            if smag:
                c += 1
            if hyp:
                c += 1
            return c + 2 + nt
        raise KeyError(kind)

    def validate(self):
        if self.hyperdiffusion is not None and \
                not isinstance(self.hyperdiffusion, (NoHyperDiffusion, DryBiharmonic)):
            raise UnsupportedModelError(
                f"hyperdiffusion model {type(self.hyperdiffusion).__name__} is not supported by libcmdg "
                "(NoHyperDiffusion / DryBiharmonic only)")
        if isinstance(self.hyperdiffusion, DryBiharmonic) and isinstance(self.orientation, NoOrientation):
            raise UnsupportedModelError("DryBiharmonic needs an orientation")
        if self.tracers is not None and not isinstance(self.tracers, (NoTracers, NTracers)):
            raise UnsupportedModelError(f"tracer model {type(self.tracers).__name__} is not supported")
        if isinstance(self.tracers, NTracers):
            if not 1 <= len(self.tracers.δ_χ) <= 4:
                raise UnsupportedModelError("NTracers: 1..4 tracers are compiled into libcmdg")
            if isinstance(self.hyperdiffusion, DryBiharmonic):
                raise UnsupportedModelError("NTracers with DryBiharmonic is not supported")
        for name in ("precipitation", "radiation", "turbconv"):
            if getattr(self, name) is not None:
                raise UnsupportedModelError(
                    f"AtmosModel.{name} = {getattr(self, name)!r} is not supported by libcmdg "
                    "(dry compressible Euler / Navier-Stokes subset only); no fallback exists")
        if not isinstance(self.moisture, DryModel):
            raise UnsupportedModelError("only DryModel moisture is supported")
        for s in self.source:
            if not isinstance(s, (Gravity, Coriolis, HeldSuarezForcing, RayleighSponge)):
                raise UnsupportedModelError(f"unsupported source {type(s).__name__}")
        if sum(isinstance(s, RayleighSponge) for s in self.source) > 1:
            raise UnsupportedModelError("at most one RayleighSponge source is supported")
        if any(isinstance(s, (HeldSuarezForcing, RayleighSponge)) for s in self.source) and \
                isinstance(self.orientation, NoOrientation):
            raise UnsupportedModelError("HeldSuarezForcing / RayleighSponge need an orientation")
        for bc in self.boundaryconditions:
            if not (isinstance(bc, AtmosBC) and isinstance(bc.momentum, Impenetrable)
                    and isinstance(bc.momentum.drag, (FreeSlip, NoSlip))
                    and isinstance(bc.energy, Insulating)):
                raise UnsupportedModelError(f"unsupported boundary condition {bc!r}")


# ---------------------------------------------------------------------------------------
# Ocean: HydrostaticBoussinesqModel (src/Ocean/HydrostaticBoussinesq), OceanBC (src/Ocean/OceanBC.jl),
# OceanGyre (src/Ocean/OceanProblems/ocean_gyre.jl), spectral filters (src/Numerics/Mesh/Filters.jl)
# ---------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Penetrable:
    drag: object = field(default_factory=FreeSlip)


class KinematicStress:
    pass


class TemperatureFlux:
    pass


@dataclass(frozen=True)
class OceanBC:
    velocity: object = field(default_factory=lambda: Impenetrable(NoSlip()))
    temperature: object = field(default_factory=Insulating)


@dataclass(frozen=True)
class OceanGyre:
    Lˣ: float
    Lʸ: float
    H: float
    τₒ: float = 1e-1
    λʳ: float = 4 / 86400
    θᴱ: float = 10.0
    boundary_conditions: Tuple = (OceanBC(Impenetrable(NoSlip()), Insulating()),
                                  OceanBC(Impenetrable(NoSlip()), Insulating()),
                                  OceanBC(Penetrable(KinematicStress()), TemperatureFlux()))


@dataclass(frozen=True)
class HomogeneousBox:
    """``HomogeneousBox`` (src/Ocean/OceanProblems/homogeneous_box.jl:15-66): jet-like wind stress, constant
    temperature; no temperature relaxation (lambda_r = theta_E = 0 in the descriptor)."""
    Lˣ: float
    Lʸ: float
    H: float
    τₒ: float = 1e-1
    boundary_conditions: Tuple = (OceanBC(Impenetrable(NoSlip()), Insulating()),
                                  OceanBC(Impenetrable(NoSlip()), Insulating()),
                                  OceanBC(Penetrable(KinematicStress()), Insulating()))
    λʳ: float = 0.0
    θᴱ: float = 0.0


@dataclass
class HydrostaticBoussinesqModel:
    problem: OceanGyre
    param_set: EarthParameterSet = field(default_factory=EarthParameterSet)
    ρₒ: float = 1000.0
    cʰ: float = 0.0
    cᶻ: float = 0.0
    αᵀ: float = 2e-4
    νʰ: float = 5e3
    νᶻ: float = 5e-3
    κʰ: float = 1e3
    κᶻ: float = 1e-4
    κᶜ: float = 1e-1
    fₒ: float = 1e-4
    β: float = 1e-11
    momentum_advection: Optional[object] = None   # only `nothing` (the default) is supported
    coupling: Optional[object] = None             # Uncoupled
    forcing: Optional[object] = None              # no forcing

    def number_states(self, kind):
        return {"Prognostic": 4, "Auxiliary": 8, "Gradient": 5, "GradientFlux": 10}[kind]

    def validate(self):
        for name in ("momentum_advection", "coupling", "forcing"):
            if getattr(self, name) is not None:
                raise UnsupportedModelError(f"HBModel.{name} is not supported by libcmdg")
        if not isinstance(self.problem, (OceanGyre, HomogeneousBox)):
            raise UnsupportedModelError("only OceanGyre / HomogeneousBox problems are supported")
        for bc in self.problem.boundary_conditions:
            ocean_bc_codes(bc)


HBModel = HydrostaticBoussinesqModel


def ocean_bc_codes(bc):
    v, t = bc.velocity, bc.temperature
    if isinstance(v, Impenetrable) and isinstance(v.drag, NoSlip):
        vc = 1
    elif isinstance(v, Impenetrable) and isinstance(v.drag, FreeSlip):
        vc = 2
    elif isinstance(v, Penetrable) and isinstance(v.drag, FreeSlip):
        vc = 3
    elif isinstance(v, Penetrable) and isinstance(v.drag, KinematicStress):
        vc = 4
    else:
        raise UnsupportedModelError(f"unsupported ocean velocity BC {v!r}")
    if isinstance(t, Insulating):
        tc = 1
    elif isinstance(t, TemperatureFlux):
        tc = 2
    else:
        raise UnsupportedModelError(f"unsupported ocean temperature BC {t!r}")
    return vc, tc


def spectral_filter_matrix(r, Nc, σ):
    """Filters.jl:114-131."""
    import numpy as np
    r = np.asarray(r, dtype=np.float64)
    N = len(r) - 1
    V = np.stack([np.polynomial.legendre.legval(r, [0] * n + [1]) * np.sqrt((2 * n + 1) / 2)
                  for n in range(N + 1)], axis=1)
    Σ = np.ones(N + 1)
    for n in range(Nc, N + 1):
        Σ[n] = σ((n - Nc) / (N - Nc))
    return (V * Σ[None, :]) @ np.linalg.inv(V)


class FilterIndices:
    """``FilterIndices(I...)`` (Filters.jl:60-98), 1-based state indices as in Julia."""

    def __init__(self, *I):
        self.I = tuple(int(i) for i in I)


class AtmosFilterPerturbations:
    """``AtmosFilterPerturbations(atmos)`` (src/Atmos/Model/filters.jl:4-48)."""

    def __init__(self, atmos):
        self.atmos = atmos


class CutoffFilter:
    """``CutoffFilter(grid, Nc)`` (Filters.jl:275-314); only the vertical matrix is used here."""

    def __init__(self, grid, Nc):
        self.filter_matrix = spectral_filter_matrix(grid.xi, Nc, lambda η: 0.0)


class ExponentialFilter:
    """``ExponentialFilter(grid, Nc, s)`` (Filters.jl:172-229)."""

    def __init__(self, grid, Nc=0, s=32, α=None):
        import numpy as np
        α = -np.log(np.finfo(np.float64).eps) if α is None else α
        assert s % 2 == 0
        self.filter_matrix = spectral_filter_matrix(grid.xi, Nc, lambda η: np.exp(-α * η ** s))
