"""Dry AtmosModel balance law (test infrastructure -- see oracle/__init__.py).

Vectorised NumPy restatement of the dry/compressible subset of
``src/Atmos/Model`` that the hot path inlines into its kernels:

* state layouts (``AtmosModel.jl:387-497``): prognostic ``rho, rhou[3], rhoe``;
  auxiliary ``coord[3], [Phi, gradPhi[3]], [ref rho,p,T,rhoe,rhoq_tot,rhoq_liq,rhoq_ice],
  [Delta], theta_v, air_T``; gradient ``u[3], h_tot, [theta_v]``; gradient flux
  ``grad h_tot[3], S[6] (11,21,31,22,32,33), [N2]``
* first-order fluxes ``tendencies_mass.jl:7-9``, ``tendencies_momentum.jl:13-30``,
  ``tendencies_energy.jl:7-21`` summed by ``BalanceLaws/kernels.jl:31-53``
* sources Gravity / Coriolis ``tendencies_momentum.jl:66-92``, RayleighSponge ``:104-137``,
  HeldSuarezForcing ``experiments/AtmosGCM/heldsuarez.jl:112-172`` (= the tutorial's
  ``HeldSuarezForcingTutorial``, ``tutorials/Atmos/heldsuarez.jl:45-118``) with
  ``latitude``/``projection_tangential`` of ``src/Common/Orientations/Orientations.jl:73-99,178-179``
* ``wavespeed`` ``AtmosModel.jl:776-796``
* thermodynamic state ``thermo_states.jl:67-77``, ``moisture.jl:32-69``
* gradient argument / flux ``AtmosModel.jl:622-744``, ``energy.jl:20-59``,
  ``TurbulenceClosures.jl:351-362,446-470``
* second-order fluxes ``tendencies_momentum.jl:36-43``, ``tendencies_energy.jl:27-59``,
  turbulence tensors ``TurbulenceClosures.jl:364-404,472-499``
* wall boundary state ``boundaryconditions.jl:60-111``, ``bc_momentum.jl:24-49``,
  ``bc_energy.jl:10-17``
* Roe flux ``AtmosModel.jl:967-1078``

Thermodynamics.jl 0.5.1 and CLIMAParameters.jl 0.2.0 are un-vendored
dependencies (``Manifest.toml:985-989,82-86``); their dry-air formulas and
constants are restated in ``Params`` / ``thermo_*`` below and pinned end to
end by the isentropic-vortex golden errors (tests/test_oracle_golden.py).
**Parity unpinned**: ``MSLP``, ``Omega``, ``planet_radius``, ``grav`` enter only
configurations for which the reference has no tight golden value.
"""
import numpy as np


class Params:
    """CLIMAParameters.Planet values used on the path (v0.2.0)."""

    def __init__(self, FT=np.float64):
        FT = np.dtype(FT).type
        gas_constant = 8.3144598
        molmass_dryair = 28.97e-3
        self.FT = FT
        self.R_d = FT(gas_constant / molmass_dryair)
        self.kappa_d = FT(2 / 7)
        self.cp_d = FT((gas_constant / molmass_dryair) / (2 / 7))
        self.cv_d = FT((gas_constant / molmass_dryair) / (2 / 7) - gas_constant / molmass_dryair)
        self.T_0 = FT(273.16)
        self.MSLP = FT(1.01325e5)
        self.grav = FT(9.81)
        self.Omega = FT(7.2921159e-5)
        self.planet_radius = FT(6.371e6)
        self.inv_Pr_turb = FT(3)
        self.C_smag = FT(0.21)
        self.T_surf_ref = FT(290)
        self.T_min_ref = FT(220)
        self.day = FT(86400)


# --- Thermodynamics.jl (dry air) -------------------------------------------
def thermo_internal_energy(ρ, ρe, ρu, e_pot):
    ρinv = 1 / ρ
    ρe_kin = ρinv * (ρu[0] ** 2 + ρu[1] ** 2 + ρu[2] ** 2) / 2
    ρe_pot = ρ * e_pot
    return ρinv * (ρe - ρe_kin - ρe_pot)


def air_temperature(ps, e_int):
    return ps.T_0 + e_int / ps.cv_d


def air_pressure(ps, T, ρ):
    return ps.R_d * ρ * T


def soundspeed_air(ps, T):
    return np.sqrt(ps.cp_d / ps.cv_d * ps.R_d * T)


def air_density(ps, T, p):
    return p / (ps.R_d * T)


def total_energy(ps, e_kin, e_pot, T):
    return e_kin + e_pot + ps.cv_d * (T - ps.T_0)


def exner_given_pressure(ps, p):
    return (p / ps.MSLP) ** (ps.R_d / ps.cp_d)


class DryAtmosModel:
    """Configuration record + pointwise physics of the dry AtmosModel."""

    S = 5
    iρ, iρu, iρe = 0, slice(1, 4), 4

    def __init__(self, FT=np.float64, orientation="none", ref_state=None,
                 turbulence=("constant_dynamic", 0.0, False), sources=(),
                 bcs=(), params=None, hyperdiffusion=None, tracers=None):
        self.FT = np.dtype(FT).type
        self.ps = params or Params(FT)
        self.orientation = orientation            # none | flat | spherical
        self.ref_state = ref_state                # None or dict(T_surf,T_min,H_t,subtract_off)
        self.turbulence = turbulence
        # entries: "gravity" | "coriolis" | "held_suarez" |
        # ("rayleigh_sponge", z_max, z_sponge, alpha_max, (u1,u2,u3), gamma), summed in tuple order
        self.sources = tuple(sources)
        self.bcs = tuple(bcs)                     # per boundary tag: "freeslip" | "noslip"
        # auxiliary layout
        c = 3
        self.a_coord = slice(0, 3)
        self.a_Φ = self.a_gradΦ = None
        if orientation != "none":
            self.a_Φ = c
            self.a_gradΦ = slice(c + 1, c + 4)
            c += 4
        self.a_ref = None
        if ref_state is not None:
            self.a_ref = dict(ρ=c, p=c + 1, T=c + 2, ρe=c + 3, ρq_tot=c + 4, ρq_liq=c + 5, ρq_ice=c + 6)
            c += 7
        self.a_Δ = None
        if turbulence[0] == "smagorinsky":
            self.a_Δ = c
            c += 1
        # DryBiharmonic hyperdiffusion (TurbulenceClosures.jl:793-848): ("dry_biharmonic", tau)
        self.hyperdiffusion = hyperdiffusion
        self.a_Δh = None
        if hyperdiffusion is not None:
            assert hyperdiffusion[0] == "dry_biharmonic"
            self.a_Δh = c
            c += 1
        self.a_θv = c
        self.a_T = c + 1
        c += 2
        # NTracers{N}(delta_chi) (src/Atmos/Model/tracers.jl:113-131): N passive tracers rho*chi after
        # rho*e; aux.tracers.delta_chi after aux.moisture; chi in the gradient variables and grad chi
        # (3 x N, column-major: d + 3 i) in the gradient flux after everything else
        self.tracers = None if tracers is None else tuple(float(x) for x in tracers)
        self.NT = 0 if tracers is None else len(self.tracers)
        self.S = 5 + self.NT
        self.a_δχ = None
        if self.NT:
            self.a_δχ = slice(c, c + self.NT)
            c += self.NT
        self.A = c
        # gradient / gradient-flux layout
        self.smag = turbulence[0] == "smagorinsky"
        self.G = 5 if self.smag else 4
        self.GF = 10 if self.smag else 9
        # Gradient vars: u, h_tot, [theta_v], [hyperdiffusion: u_h (3), h_tot]; the last four are the
        # GradientLaplacian variables (hypervisc_indexmap), Hyperdiffusive = nu grad^3 u_h (3x3
        # column major), nu grad^3 h_tot (3)
        self.hyper_G = None
        self.ngradlap = self.nhyper = 0
        if hyperdiffusion is not None:
            self.hyper_G = self.G
            self.G += 4
            self.ngradlap, self.nhyper = 4, 12
        self.G_χ, self.GF_χ = self.G, self.GF
        self.G += self.NT
        self.GF += 3 * self.NT
        self.subtract_off = bool(ref_state is not None and ref_state.get("subtract_off", True))

    # ------------------------------------------------------------------
    def viscous(self):
        """True when second-order fluxes can be non-zero."""
        return self.hyperdiffusion is not None or \
            not (self.turbulence[0].startswith("constant") and self.turbulence[1] == 0)

    def Φ(self, aux):
        if self.a_Φ is None:
            return -np.zeros_like(aux[0])
        return aux[self.a_Φ]

    def thermo(self, Q, aux):
        """(T, p) of ``recover_thermo_state`` for the dry model."""
        e_int = thermo_internal_energy(Q[0], Q[4], Q[1:4], self.Φ(aux))
        T = air_temperature(self.ps, e_int)
        return T, air_pressure(self.ps, T, Q[0])

    def flux_first_order(self, Q, aux):
        """F[d, s, ...] (3, S, ...)."""
        ρ, ρu, ρe = Q[0], Q[1:4], Q[4]
        T, p = self.thermo(Q, aux)
        F = np.zeros((3, self.S) + Q.shape[1:], dtype=Q.dtype)
        u = ρu / ρ
        F[:, 0] = ρu
        for d in range(3):
            for c in range(3):
                F[d, 1 + c] = ρu[d] * u[c]
        pp = p - aux[self.a_ref["p"]] if self.subtract_off else p
        for d in range(3):
            F[d, 1 + d] = F[d, 1 + d] + pp
        F[:, 4] = u * ρe + u * p
        for i in range(self.NT):
            # flux(::Tracers, ::Advect) (tendencies_tracers.jl:7-11): rho*chi_i u
            F[:, 5 + i] = Q[5 + i] * u
        return F

    def source(self, Q, aux):
        S = np.zeros_like(Q)
        srcs = []
        if "gravity" in self.sources:
            gΦ = aux[self.a_gradΦ]
            if self.subtract_off:
                srcs.append(-(Q[0] - aux[self.a_ref["ρ"]]) * gΦ)
            else:
                srcs.append(-Q[0] * gΦ)
        if "coriolis" in self.sources:
            w = 2 * self.ps.Omega
            ρu = Q[1:4]
            # -(0,0,2Ω) x ρu
            cx = np.stack([0 * ρu[2] - w * ρu[1], w * ρu[0] - 0 * ρu[2], 0 * ρu[1] - 0 * ρu[0]])
            srcs.append(-cx)
        FT = self.FT
        for s in self.sources:
            if s == "held_suarez":
                k_v, k_T, T_equil = self.held_suarez_coefficients(Q, aux)
                ρu = Q[1:4]
                n̂ = aux[self.a_gradΦ] / self.ps.grav
                proj_n = n̂ * (n̂[0] * ρu[0] + n̂[1] * ρu[1] + n̂[2] * ρu[2])
                srcs.append(-k_v * (ρu - proj_n))
                T, _ = self.thermo(Q, aux)
                S[4] = -k_T * Q[0] * self.ps.cv_d * (T - T_equil)
            elif isinstance(s, tuple) and s[0] == "rayleigh_sponge":
                _, z_max, z_sponge, α_max, u_relax, γ = s
                z = self.Φ(aux) / self.ps.grav
                r = (z - FT(z_sponge)) / (FT(z_max) - FT(z_sponge))
                with np.errstate(invalid="ignore"):
                    β = FT(α_max) * np.sin(np.pi * (r / 2)) ** FT(γ)
                β = np.where(z >= FT(z_sponge), β, 0 * β)
                ur = np.asarray(u_relax, dtype=Q.dtype).reshape((3,) + (1,) * (Q.ndim - 1))
                srcs.append(-β * (Q[1:4] - Q[0] * ur))
        if srcs:
            tot = srcs[0]
            for s in srcs[1:]:
                tot = tot + s
            S[1:4] = tot
        return S

    def held_suarez_coefficients(self, Q, aux):
        """(k_v, k_T, T_equil) of ``held_suarez_forcing_coefficients``
        (experiments/AtmosGCM/heldsuarez.jl:116-153)."""
        ps, FT = self.ps, self.FT
        k_a = FT(1 / (40 * ps.day))
        k_f = FT(1 / ps.day)
        k_s = FT(1 / (4 * ps.day))
        ΔT_y, Δθ_z, T_equator, T_min, σ_b = FT(60), FT(10), FT(315), FT(200), FT(7 / 10)
        x = aux[self.a_coord]
        φ = np.arcsin(x[2] / np.sqrt(x[0] ** 2 + x[1] ** 2 + x[2] ** 2))
        _, p = self.thermo(Q, aux)
        σ = p / ps.MSLP
        exner_p = σ ** (ps.R_d / ps.cp_d)
        Δσ = (σ - σ_b) / (1 - σ_b)
        height_factor = np.maximum(0, Δσ)
        T_equil = (T_equator - ΔT_y * np.sin(φ) ** 2 - Δθ_z * np.log(σ) * np.cos(φ) ** 2) * exner_p
        T_equil = np.maximum(T_min, T_equil)
        k_T = k_a + (k_s - k_a) * height_factor * np.cos(φ) ** 4
        k_v = k_f * height_factor
        return k_v, k_T, T_equil

    def wavespeed(self, n, Q, aux):
        u = (1 / Q[0]) * Q[1:4]
        uN = np.abs(n[0] * u[0] + n[1] * u[1] + n[2] * u[2])
        T, _ = self.thermo(Q, aux)
        ws = uN + soundspeed_air(self.ps, T)
        if not self.NT:
            return ws
        # wavespeed_tracers! (tracers.jl:165-180): the tracers travel with |u . n| only
        return np.stack([ws] * 5 + [uN] * self.NT)

    # --- auxiliary update (moisture.jl:58-69) --------------------------
    def nodal_update_aux(self, Q, aux):
        T, p = self.thermo(Q, aux)
        aux[self.a_θv] = T / exner_given_pressure(self.ps, p)
        aux[self.a_T] = T

    # --- boundary state (first-order / gradient fluxes) ----------------
    def boundary_state(self, kind, bctag, n, Qm, auxm):
        """Ghost state for wall faces; ``kind`` in {"first", "gradient"}."""
        Qp = Qm.copy()
        auxp = auxm.copy()
        bc = self.bcs[bctag - 1]
        ρun = Qm[1] * n[0] + Qm[2] * n[1] + Qm[3] * n[2]
        if bc == "freeslip":
            fac = 2 if kind == "first" else 1
            Qp[1:4] = Qm[1:4] - fac * ρun * n
        elif bc == "noslip":
            Qp[1:4] = -Qm[1:4] if kind == "first" else 0 * Qm[1:4]
        else:
            raise ValueError(f"unsupported boundary condition {bc!r}")
        self.nodal_update_aux(Qp, auxp)
        return Qp, auxp

    # --- gradient pass -------------------------------------------------
    def gradient_argument(self, Q, aux):
        ρinv = 1 / Q[0]
        G = np.zeros((self.G,) + Q.shape[1:], dtype=Q.dtype)
        G[0:3] = ρinv * Q[1:4]
        T, _ = self.thermo(Q, aux)
        e_tot = Q[4] * (1 / Q[0])
        G[3] = e_tot + self.ps.R_d * T
        if self.smag:
            G[4] = aux[self.a_θv]
        if self.hyper_G is not None:
            # compute_gradient_argument!(::DryBiharmonic, ...): u_h = (I - k k') u, h_tot
            k = aux[self.a_gradΦ] / self.ps.grav
            u = G[0:3]
            ku = k[0] * u[0] + k[1] * u[1] + k[2] * u[2]
            G[self.hyper_G:self.hyper_G + 3] = u - k * ku
            G[self.hyper_G + 3] = G[3]
        for i in range(self.NT):
            G[self.G_χ + i] = Q[5 + i] * ρinv
        return G

    def transform_post_gradient_laplacian(self, gradlap, Q, aux):
        """``transform_post_gradient_laplacian!(::DryBiharmonic, ...)``: gradlap[d, g] ->
        hyperdiffusive (12, ...): nu4 grad(lap u_h) at d + 3 c, nu4 grad(lap h_tot) at 9 + d."""
        τ = self.FT(self.hyperdiffusion[1])
        ν4 = (aux[self.a_Δh] / 2) ** 4 / 2 / τ
        H = np.zeros((12,) + Q.shape[1:], dtype=Q.dtype)
        for c in range(3):
            for d in range(3):
                H[d + 3 * c] = ν4 * gradlap[d, c]
        for d in range(3):
            H[9 + d] = ν4 * gradlap[d, 3]
        return H

    def flux_hyperdiffusive(self, Q, H):
        assert self.NT == 0, "DryBiharmonic with tracers is not restated"
        """HyperdiffViscousFlux / HyperdiffEnthalpyFlux (tendencies_momentum.jl:51-54,
        tendencies_energy.jl:40-48): F[d, 1 + c] = rho H[d, c]; F[d, 4] = H[d, :] . rhou + H_h[d] rho."""
        F = np.zeros((3, 5) + Q.shape[1:], dtype=Q.dtype)
        ρ, ρu = Q[0], Q[1:4]
        for d in range(3):
            for c in range(3):
                F[d, 1 + c] = ρ * H[d + 3 * c]
            F[d, 4] = (H[d] * ρu[0] + H[d + 3] * ρu[1] + H[d + 6] * ρu[2]) + H[9 + d] * ρ
        return F

    def gradient_flux(self, gradG, Q, aux):
        """gradG[d, g, ...] -> GF (linear in gradG)."""
        GF = np.zeros((self.GF,) + Q.shape[1:], dtype=Q.dtype)
        GF[0:3] = gradG[:, 3]
        du = gradG[:, 0:3]  # du[d, c] = d u_c / d x_d
        GF[3] = du[0, 0]
        GF[4] = (du[1, 0] + du[0, 1]) / 2
        GF[5] = (du[2, 0] + du[0, 2]) / 2
        GF[6] = du[1, 1]
        GF[7] = (du[2, 1] + du[1, 2]) / 2
        GF[8] = du[2, 2]
        if self.smag:
            gΦ = aux[self.a_gradΦ]
            gθ = gradG[:, 4]
            GF[9] = (gθ[0] * gΦ[0] + gθ[1] * gΦ[1] + gθ[2] * gΦ[2]) / aux[self.a_θv]
        for i in range(self.NT):
            for d in range(3):
                GF[self.GF_χ + d + 3 * i] = gradG[d, self.G_χ + i]
        return GF

    def turbulence_tensors(self, Q, GF, aux):
        """(D_t[3] or scalar, tau[3,3]); tau[i, j]."""
        ps = self.ps
        S6 = GF[3:9]
        Sm = [[S6[0], S6[1], S6[2]], [S6[1], S6[3], S6[4]], [S6[2], S6[4], S6[5]]]
        kind = self.turbulence[0]
        if kind in ("constant_kinematic", "constant_dynamic"):
            ν = self.FT(self.turbulence[1]) if kind == "constant_kinematic" else self.FT(self.turbulence[1]) / Q[0]
            D_t = ν * ps.inv_Pr_turb
            τ = [[(-2 * ν) * Sm[i][j] for j in range(3)] for i in range(3)]
            if self.turbulence[2]:
                tr = S6[0] + S6[3] + S6[5]
                for i in range(3):
                    τ[i][i] = τ[i][i] + (2 * ν / 3) * tr
            return [D_t, D_t, D_t], τ
        if kind == "smagorinsky":
            FT = self.FT
            norm2 = (S6[0] ** 2 + 2 * S6[1] ** 2 + 2 * S6[2] ** 2 + S6[3] ** 2
                     + 2 * S6[4] ** 2 + S6[5] ** 2)
            normS = np.sqrt(2 * norm2)
            k = aux[self.a_gradΦ] / ps.grav
            Ri = GF[9] / (normS ** 2 + np.spacing(normS))
            f_b2 = np.sqrt(np.clip(FT(1) - Ri * ps.inv_Pr_turb, FT(0), FT(1)))
            ν0 = normS * (FT(self.turbulence[1]) * aux[self.a_Δ]) ** 2 + FT(1e-5)
            dotνk = ν0 * k[0] + ν0 * k[1] + ν0 * k[2]
            ν_v = k * dotνk
            ν_h = ν0 - ν_v
            ν = ν_h + ν_v * f_b2
            D_t = [ν[i] * ps.inv_Pr_turb for i in range(3)]
            τ = [[-2 * ν[i] * Sm[i][j] for j in range(3)] for i in range(3)]
            return D_t, τ
        raise ValueError(kind)

    def flux_second_order(self, Q, GF, aux):
        F = np.zeros((3, self.S) + Q.shape[1:], dtype=Q.dtype)
        if self.GF == 0:
            return F
        D_t, τ = self.turbulence_tensors(Q, GF, aux)
        ρ, ρu = Q[0], Q[1:4]
        for i in range(3):
            for j in range(3):
                F[i, 1 + j] = τ[i][j] * ρ
        for i in range(3):
            visc = τ[i][0] * ρu[0] + τ[i][1] * ρu[1] + τ[i][2] * ρu[2]
            d_h = (-D_t[i]) * GF[i]
            F[i, 4] = visc + d_h * ρ
        for t in range(self.NT):
            # flux(::Tracers, ::Diffusion) (tendencies_tracers.jl:17-22): d_chi = (-D_t) delta_chi' .* grad chi
            δ = aux[self.a_δχ][t]
            for i in range(3):
                F[i, 5 + t] = (((-D_t[i]) * δ) * GF[self.GF_χ + i + 3 * t]) * ρ
        return F

    # --- Roe flux (dry) ------------------------------------------------
    def roe_dissipation(self, n, Qm, auxm, Qp, auxp):
        ps = self.ps
        Φ = self.Φ(auxm)
        ρm, ρum, ρem = Qm[0], Qm[1:4], Qm[4]
        Tm, pm = self.thermo(Qm, auxm)
        um = ρum / ρm
        hm = ρem / ρm + ps.R_d * Tm
        cm = soundspeed_air(ps, Tm)
        ρp, ρup, ρep = Qp[0], Qp[1:4], Qp[4]
        Tp, pp = self.thermo(Qp, auxp)
        up = ρup / ρp
        hp = ρep / ρp + ps.R_d * Tp
        cp = soundspeed_air(ps, Tp)

        def ravg(a, b):
            return (np.sqrt(ρm) * a + np.sqrt(ρp) * b) / (np.sqrt(ρm) + np.sqrt(ρp))

        ρt = np.sqrt(ρm * ρp)
        ut = ravg(um, up)
        ht = ravg(hm, hp)
        ct = np.sqrt(ravg(cm ** 2, cp ** 2))
        utn = ut[0] * n[0] + ut[1] * n[1] + ut[2] * n[2]
        Δρ = ρp - ρm
        Δp = pp - pm
        Δu = up - um
        Δun = Δu[0] * n[0] + Δu[1] * n[1] + Δu[2] * n[2]
        w1 = np.abs(utn - ct) * (Δp - ρt * ct * Δun) / (2 * ct ** 2)
        w2 = np.abs(utn + ct) * (Δp + ρt * ct * Δun) / (2 * ct ** 2)
        w3 = np.abs(utn) * (Δρ - Δp / ct ** 2)
        w4 = np.abs(utn) * ρt
        D = np.zeros_like(Qm)
        D[0] = (w1 + w2 + w3) / 2
        D[1:4] = (w1 * (ut - ct * n) + w2 * (ut + ct * n) + w3 * ut + w4 * (Δu - Δun * n)) / 2
        utut = ut[0] ** 2 + ut[1] ** 2 + ut[2] ** 2
        utΔu = ut[0] * Δu[0] + ut[1] * Δu[1] + ut[2] * Δu[2]
        D[4] = (w1 * (ht - ct * utn) + w2 * (ht + ct * utn)
                + w3 * (utut / 2 + Φ - ps.T_0 * ps.cv_d) + w4 * (utΔu - utn * Δun)) / 2
        return D


# --- initial conditions -----------------------------------------------------
class IsentropicVortexSetup:
    """``test/Numerics/DGMethods/Euler/isentropicvortex_setup.jl:3-66``."""

    def __init__(self, ps, FT=np.float64):
        FT = np.dtype(FT).type
        self.ps = ps
        self.p_inf = FT(10 ** 5)
        self.T_inf = FT(300)
        self.ρ_inf = air_density(ps, self.T_inf, self.p_inf)
        self.translation_speed = FT(150)
        self.translation_angle = FT(np.pi / 4)
        self.vortex_speed = FT(50)
        self.vortex_radius = FT(1) / FT(200)
        self.domain_halflength = FT(1) / FT(20)

    def __call__(self, x1, x2, x3, t):
        ps = self.ps
        FT = x1.dtype.type
        α = self.translation_angle
        u_inf = [self.translation_speed * np.cos(α), self.translation_speed * np.sin(α), FT(0)]
        L = self.domain_halflength
        x = [x1 - u_inf[0] * t, x2 - u_inf[1] * t, x3 - u_inf[2] * t]
        x = [xi - np.floor((xi + L) / (2 * L)) * (2 * L) for xi in x]
        R = self.vortex_radius
        r = np.sqrt(x[0] ** 2 + x[1] ** 2)
        δu_x = -self.vortex_speed * x[1] / R * np.exp(-(r / R) ** 2 / 2)
        δu_y = self.vortex_speed * x[0] / R * np.exp(-(r / R) ** 2 / 2)
        u = [u_inf[0] + δu_x, u_inf[1] + δu_y, u_inf[2] + 0 * δu_x]
        κ = ps.kappa_d
        T = self.T_inf * (1 - κ * self.vortex_speed ** 2 / 2 * self.ρ_inf / self.p_inf
                          * np.exp(-(r / R) ** 2))
        p = self.p_inf * (T / self.T_inf) ** (FT(1) / κ)
        ρ = air_density(ps, T, p)
        e_kin = (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]) / 2
        Q = np.stack([ρ, ρ * u[0], ρ * u[1], ρ * u[2],
                      ρ * total_energy(ps, e_kin, FT(0), T)])
        return Q.astype(FT)


def decaying_temperature_profile(ps, z, T_virt_surf, T_min_ref, H_t):
    """``src/Atmos/TemperatureProfiles/TemperatureProfiles.jl:133-156``."""
    H_sfc = ps.R_d * T_virt_surf / ps.grav
    zp = z / H_t
    tanh_zp = np.tanh(zp)
    ΔTv = T_virt_surf - T_min_ref
    Tv = T_virt_surf - ΔTv * tanh_zp
    ΔTvp = ΔTv / T_virt_surf
    p = -H_t * (zp + ΔTvp * (np.log(1 - ΔTvp * tanh_zp) - np.log(1 + tanh_zp) + zp))
    p = p / (H_sfc * (1 - ΔTvp ** 2))
    p = ps.MSLP * np.exp(p)
    return Tv, p


def dry_adiabatic_profile(ps, z, T_surface, T_min_ref):
    """``DryAdiabaticProfile`` (src/Atmos/TemperatureProfiles/TemperatureProfiles.jl:76-98)."""
    Γ = ps.grav / ps.cp_d
    T = np.maximum(T_surface - Γ * z, T_min_ref)
    p = ps.MSLP * (T / T_surface) ** (ps.grav / (ps.R_d * Γ))
    z_top = (T_surface - T_min_ref) / Γ
    H_min = ps.R_d * T_min_ref / ps.grav if T_min_ref != 0 else np.inf
    with np.errstate(invalid="ignore", divide="ignore"):
        p = np.where(T == T_min_ref, p * np.exp(-(z - z_top) / H_min), p)
    return T, p


def reference_profile(ps, rs, z, FT):
    """(T_v, p) of the reference state's temperature profile; ``rs["profile"]`` selects it
    (default: DecayingTemperatureProfile)."""
    if rs.get("profile", "decaying") == "dry_adiabatic":
        return dry_adiabatic_profile(ps, z, FT(rs["T_surf"]), FT(rs["T_min"]))
    return decaying_temperature_profile(ps, z, FT(rs["T_surf"]), FT(rs["T_min"]), FT(rs["H_t"]))


def init_risingbubble(model, aux, xc=5000.0, zc=2000.0, rc=2000.0, θamplitude=2.0):
    """``init_risingbubble!`` (tutorials/Atmos/risingbubble.jl:106-176) without the tracers."""
    ps = model.ps
    FT = aux.dtype.type
    x, z = aux[0], aux[2]
    r = np.sqrt((x - FT(xc)) ** 2 + (z - FT(zc)) ** 2)
    θ_ref = FT(model.ref_state["T_surf"])
    Δθ = np.where(r <= rc, FT(θamplitude) * (1.0 - r / FT(rc)), FT(0))
    θ = θ_ref + Δθ
    π_exner = FT(1) - ps.grav / (ps.cp_d * θ) * z
    ρ = ps.MSLP / (ps.R_d * θ) * π_exner ** (ps.cv_d / ps.R_d)
    T = θ * π_exner
    ρe = ρ * total_energy(ps, FT(0), aux[model.a_Φ], T)
    zero = np.zeros_like(ρ)
    return np.stack([ρ, zero, zero, zero, ρe]).astype(FT)


def init_baroclinic_wave(model, aux):
    """Dry branch of ``experiments/TestCase/baroclinic_wave.jl:31-163``.

    ``aux``: (A, ...) auxiliary array with coord / orientation filled.
    """
    ps = model.ps
    FT = aux.dtype.type
    grav, R_d, Ω, a, p_0 = ps.grav, ps.R_d, ps.Omega, ps.planet_radius, ps.MSLP
    k = FT(3)
    T_E, T_P = FT(310), FT(240)
    T_0 = FT(0.5) * (T_E + T_P)
    Γ = FT(0.005)
    A = 1 / Γ
    B = (T_0 - T_P) / T_0 / T_P
    C = FT(0.5) * (k + 2) * (T_E - T_P) / T_E / T_P
    b = FT(2)
    H = R_d * T_0 / grav
    z_t = FT(15e3)
    λ_c = FT(np.pi / 9)
    φ_c = FT(2 * np.pi / 9)
    d_0 = a / 6
    V_p = FT(1)
    coord = aux[model.a_coord]
    normc = np.sqrt(coord[0] ** 2 + coord[1] ** 2 + coord[2] ** 2)
    φ = np.arcsin(coord[2] / normc)
    λ = np.arctan2(coord[1], coord[0])
    z = aux[model.a_Φ] / grav
    γ = FT(1)
    τ_z_1 = np.exp(Γ * z / T_0)
    τ_z_2 = 1 - 2 * (z / b / H) ** 2
    τ_z_3 = np.exp(-(z / b / H) ** 2)
    τ_1 = 1 / T_0 * τ_z_1 + B * τ_z_2 * τ_z_3
    τ_2 = C * τ_z_2 * τ_z_3
    τ_int_1 = A * (τ_z_1 - 1) + B * z * τ_z_3
    τ_int_2 = C * z * τ_z_3
    cφz = np.cos(φ) * (1 + γ * z / a)
    I_T = cφz ** k - k / (k + 2) * cφz ** (k + 2)
    T_v = (τ_1 - τ_2 * I_T) ** (-1)
    p = p_0 * np.exp(-grav / R_d * (τ_int_1 - τ_int_2 * I_T))
    U = grav * k / a * τ_int_2 * T_v * (cφz ** (k - 1) - cφz ** (k + 1))
    u_ref = (-Ω * (a + γ * z) * np.cos(φ)
             + np.sqrt((Ω * (a + γ * z) * np.cos(φ)) ** 2 + (a + γ * z) * np.cos(φ) * U))
    F_z = 1 - 3 * (z / z_t) ** 2 + 2 * (z / z_t) ** 3
    F_z = np.where(z > z_t, FT(0), F_z)
    arg = np.sin(φ) * np.sin(φ_c) + np.cos(φ) * np.cos(φ_c) * np.cos(λ - λ_c)
    d = a * np.arccos(np.clip(arg, -1, 1))
    c3 = np.cos(np.pi * d / 2 / d_0) ** 3
    s1 = np.sin(np.pi * d / 2 / d_0)
    mask = (0 < d) & (d < d_0) & (d != FT(a * np.pi))
    with np.errstate(divide="ignore", invalid="ignore"):
        up = (-16 * V_p / 3 / np.sqrt(FT(3)) * F_z * c3 * s1
              * (-np.sin(φ_c) * np.cos(φ) + np.cos(φ_c) * np.sin(φ) * np.cos(λ - λ_c))
              / np.sin(d / a))
        vp = (16 * V_p / 3 / np.sqrt(FT(3)) * F_z * c3 * s1 * np.cos(φ_c) * np.sin(λ - λ_c)
              / np.sin(d / a))
    up = np.where(mask, up, FT(0))
    vp = np.where(mask, vp, FT(0))
    us = [u_ref + up, vp, 0 * up]
    slat, clat, slon, clon = np.sin(φ), np.cos(φ), np.sin(λ), np.cos(λ)
    u_cart = [-slon * us[0] - slat * clon * us[1] + clat * clon * us[2],
              clon * us[0] - slat * slon * us[1] + clat * slon * us[2],
              clat * us[1] + slat * us[2]]
    T = T_v
    ρ = air_density(ps, T, p)
    e_pot = aux[model.a_Φ]
    e_kin = FT(0.5) * (u_cart[0] ** 2 + u_cart[1] ** 2 + u_cart[2] ** 2)
    e_tot = total_energy(ps, e_kin, e_pot, T)
    return np.stack([ρ, ρ * u_cart[0], ρ * u_cart[1], ρ * u_cart[2], ρ * e_tot]).astype(FT)
