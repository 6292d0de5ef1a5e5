"""Harness-side initialisation of the AtmosModel auxiliary state and of the synthetic initial
conditions (torch, on the device; setup code, not the hot path).  In production these arrays
are produced by Julia before the first tendency call:

* ``init_state_auxiliary!``  <- src/Atmos/Model/AtmosModel.jl:880-925, Orientations.jl
  (Phi, grad Phi by the element-local derivative, DGModel_kernels.jl:3097-3232),
  ref_state.jl:70-264 (hydrostatic reference state whose density balances the *discrete*
  pressure gradient, obtained by running a DG operator over ``PressureGradientModel``)
* ``IsentropicVortexSetup``  <- test/Numerics/DGMethods/Euler/isentropicvortex_setup.jl:3-66
* ``init_baroclinic_wave!``  <- experiments/TestCase/baroclinic_wave.jl:31-163 (dry branch)
"""
import math

import torch

from . import balance_laws as bl
from .dgmodel import MPIStateArray

_X1, _X2, _X3, _M, _MI = 12, 13, 14, 9, 10


def aux_layout(m):
    c = 3
    lay = {"coord": 0}
    if not isinstance(m.orientation, bl.NoOrientation):
        lay["Φ"], lay["∇Φ"] = c, c + 1
        c += 4
    if isinstance(m.ref_state, bl.HydrostaticState):
        lay.update({"ref_ρ": c, "ref_p": c + 1, "ref_T": c + 2, "ref_ρe": c + 3})
        c += 7
    if isinstance(m.turbulence, bl.SmagorinskyLilly):
        lay["Δ"] = c
        c += 1
    if isinstance(m.hyperdiffusion, bl.DryBiharmonic):
        lay["Δh"] = c
        c += 1
    lay["θ_v"], lay["T"] = c, c + 1
    c += 2
    if isinstance(m.tracers, bl.NTracers):
        lay["δ_χ"] = c
        c += len(m.tracers.δ_χ)
    lay["A"] = c
    return lay


def _dT(D, x, axis):
    """sum_n D[n, a] x[.., n, ..] along reference axis (0 = last torch dim)."""
    ax = x.dim() - 1 - axis
    return torch.matmul(x.movedim(ax, -1), D).movedim(-1, ax)


def _d(D, x, axis):
    ax = x.dim() - 1 - axis
    return torch.matmul(x.movedim(ax, -1), D.T).movedim(-1, ax)


def local_gradient(grid, f):
    """Element-local strong-form gradient of f (nreal, Np) -> 3 x (nreal, Np)."""
    nr, Nq = f.shape[0], grid.Nq
    D = torch.as_tensor(grid.D_host, dtype=grid.FT, device=grid.device)
    fr = f.reshape(nr, Nq, Nq, Nq)
    G = [_d(D, fr, a).reshape(nr, grid.Np) for a in range(3)]
    vg = grid.vgeo[:nr]
    return [vg[:, 3 * d] * G[0] + vg[:, 3 * d + 1] * G[1] + vg[:, 3 * d + 2] * G[2] for d in range(3)]


def reference_pressure_gradient(grid, p):
    """DG gradient of a nodal field ``p`` (nelem, Np; ghosts valid at face nodes):
    ``PressureGradientModel`` with central fluxes (ref_state.jl:204-264)."""
    nr, Nq, Np = grid.nrealelem, grid.Nq, grid.Np
    D = torch.as_tensor(grid.D_host, dtype=grid.FT, device=grid.device)
    vg = grid.vgeo[:nr]
    M, MI = vg[:, _M], vg[:, _MI]
    pr = p[:nr]
    out = []
    for s in range(3):
        acc = 0
        for m in range(3):
            Ft = (M * vg[:, 3 * s + m] * (-pr)).reshape(nr, Nq, Nq, Nq)
            acc = acc + _dT(D, Ft, m).reshape(nr, Np)
        out.append(MI * acc)
    out = torch.stack(out, dim=1).contiguous()      # (nr, 3, Np)
    idm = grid.vmapM[:nr] - 1
    idp = grid.vmapP[:nr] - 1
    bnd = (grid.elemtobndy[:nr] != 0)[:, :, None]
    idp = torch.where(bnd, idm, idp)
    pf = p.reshape(-1)
    psum = -pf[idm] - pf[idp]                        # (nr, 6, Nfp)
    sg = grid.sgeo[:nr]
    lift = sg[..., 3] * sg[..., 4]
    e = torch.arange(nr, device=grid.device)[:, None, None]
    vid = idm - e * Np
    flat = out.reshape(-1)
    for s in range(3):
        contrib = lift * psum * (sg[..., s] / 2)
        flat.index_add_(0, ((e * 3 + s) * Np + vid).reshape(-1), -contrib.reshape(-1))
    return flat.reshape(nr, 3, Np)


def init_state_auxiliary(model, grid, exchange=None):
    """Fill ``state_auxiliary`` (real elements; ``exchange(aux)`` refreshes ghosts between the
    passes when the grid is partitioned)."""
    p = model.param_set
    lay = aux_layout(model)
    aux = MPIStateArray(grid, lay["A"])
    a = aux.data
    nr = grid.nrealelem
    x = [grid.vgeo[:nr, _X1], grid.vgeo[:nr, _X2], grid.vgeo[:nr, _X3]]
    ex = exchange or (lambda arr: None)
    if isinstance(model.orientation, bl.SphericalOrientation):
        a[:nr, lay["Φ"]] = p.grav * (torch.sqrt(x[0] ** 2 + x[1] ** 2 + x[2] ** 2) - p.planet_radius)
    elif isinstance(model.orientation, bl.FlatOrientation):
        a[:nr, lay["Φ"]] = p.grav * x[2]
    if "Φ" in lay:
        ex(aux)
        g = local_gradient(grid, a[:nr, lay["Φ"]])
        for d in range(3):
            a[:nr, lay["∇Φ"] + d] = g[d]
    if "ref_p" in lay:
        prof = model.ref_state.virtual_temperature_profile
        z = a[:nr, lay["Φ"]] / p.grav
        if isinstance(prof, bl.DryAdiabaticProfile):
            Γ = p.grav / p.cp_d
            Tv = torch.clamp(prof.T_surface - Γ * z, min=prof.T_min_ref)
            pr = p.MSLP * (Tv / prof.T_surface) ** (p.grav / (p.R_d * Γ))
            if prof.T_min_ref > 0:
                z_top = (prof.T_surface - prof.T_min_ref) / Γ
                H_min = p.R_d * prof.T_min_ref / p.grav
                pr = torch.where(Tv == prof.T_min_ref, pr * torch.exp(-(z - z_top) / H_min), pr)
            prof = None
    if "ref_p" in lay and prof is not None:
        H_sfc = p.R_d * prof.T_virt_surf / p.grav
        zp = z / prof.H_t
        th = torch.tanh(zp)
        dTv = prof.T_virt_surf - prof.T_min_ref
        Tv = prof.T_virt_surf - dTv * th
        dTvp = dTv / prof.T_virt_surf
        pr = -prof.H_t * (zp + dTvp * (torch.log(1 - dTvp * th) - torch.log(1 + th) + zp))
        pr = pr / (H_sfc * (1 - dTvp ** 2))
        pr = p.MSLP * torch.exp(pr)
    if "ref_p" in lay:
        a[:nr, lay["ref_p"]] = pr
        a[:nr, lay["ref_ρ"]] = pr / (Tv * p.R_d)
        ex(aux)
        gp = reference_pressure_gradient(grid, a[:, lay["ref_p"]].contiguous())
        gΦ = a[:nr, lay["∇Φ"]:lay["∇Φ"] + 3]
        k = gΦ / p.grav
        num = -(k * gp).sum(dim=1)
        den = (k * gΦ).sum(dim=1)
        ρ = num / den
        a[:nr, lay["ref_ρ"]] = ρ
        T = pr / (ρ * p.R_d)
        a[:nr, lay["ref_T"]] = T
        a[:nr, lay["ref_ρe"]] = ρ * (a[:nr, lay["Φ"]] + p.cv_d * (T - p.T_0))
    for d in range(3):
        a[:nr, d] = x[d]
    if "Δ" in lay:
        vg = grid.vgeo[:nr]
        det = (vg[:, 0] * (vg[:, 4] * vg[:, 8] - vg[:, 7] * vg[:, 5])
               - vg[:, 3] * (vg[:, 1] * vg[:, 8] - vg[:, 7] * vg[:, 2])
               + vg[:, 6] * (vg[:, 1] * vg[:, 5] - vg[:, 4] * vg[:, 2]))
        a[:nr, lay["Δ"]] = 2 / (torch.sign(det) * det.abs() ** (1.0 / 3.0) * max(1, grid.N))
    if "δ_χ" in lay:
        for i, δ in enumerate(model.tracers.δ_χ):      # atmos_init_aux!(::NTracers) (tracers.jl:133-140)
            a[:nr, lay["δ_χ"] + i] = δ
    if "Δh" in lay:
        # lengthscale_horizontal (src/Numerics/Mesh/Geometry.jl:129-152): mean of |J e1|, |J e2| times 2 / N
        vg = grid.vgeo[:nr].double()
        invJ = torch.stack([torch.stack([vg[:, 3 * j + i] for j in range(3)], dim=-1) for i in range(3)], dim=-2)
        e = torch.zeros(invJ.shape[:-1] + (2,), dtype=invJ.dtype, device=invJ.device)
        e[..., 0, 0] = 1
        e[..., 1, 1] = 1
        sol = torch.linalg.solve(invJ, e)
        Δ = (sol.pow(2).sum(dim=-2).sqrt() * 2 / max(1, grid.N)).mean(dim=-1)
        a[:nr, lay["Δh"]] = Δ.to(a.dtype)
    ex(aux)
    return aux


def isentropic_vortex(model, grid, t=0.0):
    """Prognostic state (nreal, 5, Np) of the isentropic vortex at time t."""
    p = model.param_set
    nr = grid.nrealelem
    FT = grid.FT
    p_inf, T_inf = 1e5, 300.0
    ρ_inf = p_inf / (p.R_d * T_inf)
    speed, α, vs, R, L = 150.0, math.pi / 4, 50.0, 1 / 200, 1 / 20
    u_inf = (speed * math.cos(α), speed * math.sin(α), 0.0)
    x = [grid.vgeo[:nr, _X1 + d] - u_inf[d] * t for d in range(3)]
    x = [xi - torch.floor((xi + L) / (2 * L)) * (2 * L) for xi in x]
    r = torch.sqrt(x[0] ** 2 + x[1] ** 2)
    ex2 = torch.exp(-(r / R) ** 2 / 2)
    u = [u_inf[0] - vs * x[1] / R * ex2, u_inf[1] + vs * x[0] / R * ex2, torch.zeros_like(r)]
    T = T_inf * (1 - p.kappa_d * vs ** 2 / 2 * ρ_inf / p_inf * torch.exp(-(r / R) ** 2))
    pr = p_inf * (T / T_inf) ** (1 / p.kappa_d)
    ρ = pr / (p.R_d * T)
    e_kin = (u[0] ** 2 + u[1] ** 2 + u[2] ** 2) / 2
    Q = torch.stack([ρ, ρ * u[0], ρ * u[1], ρ * u[2], ρ * (e_kin + p.cv_d * (T - p.T_0))], dim=1)
    return Q.to(FT)


def baroclinic_wave(model, grid, aux):
    """Dry baroclinic-wave initial state (nreal, 5, Np)."""
    p = model.param_set
    lay = aux_layout(model)
    nr = grid.nrealelem
    a = aux.data[:nr]
    grav, R_d, Ω, rad, p_0 = p.grav, p.R_d, p.Omega, p.planet_radius, p.MSLP
    k, T_E, T_P = 3.0, 310.0, 240.0
    T_0 = 0.5 * (T_E + T_P)
    Γ = 0.005
    A, B = 1 / Γ, (T_0 - T_P) / T_0 / T_P
    C = 0.5 * (k + 2) * (T_E - T_P) / T_E / T_P
    b, H = 2.0, R_d * T_0 / grav
    z_t, λ_c, φ_c, d_0, V_p = 15e3, math.pi / 9, 2 * math.pi / 9, rad / 6, 1.0
    c0, c1, c2 = a[:, 0], a[:, 1], a[:, 2]
    φ = torch.asin(c2 / torch.sqrt(c0 ** 2 + c1 ** 2 + c2 ** 2))
    λ = torch.atan2(c1, c0)
    z = a[:, lay["Φ"]] / grav
    τ_z_1 = torch.exp(Γ * z / T_0)
    τ_z_2 = 1 - 2 * (z / b / H) ** 2
    τ_z_3 = torch.exp(-(z / b / H) ** 2)
    τ_1 = 1 / T_0 * τ_z_1 + B * τ_z_2 * τ_z_3
    τ_2 = C * τ_z_2 * τ_z_3
    τ_int_1 = A * (τ_z_1 - 1) + B * z * τ_z_3
    τ_int_2 = C * z * τ_z_3
    cz = torch.cos(φ) * (1 + z / rad)
    I_T = cz ** k - k / (k + 2) * cz ** (k + 2)
    T_v = 1 / (τ_1 - τ_2 * I_T)
    pr = p_0 * torch.exp(-grav / R_d * (τ_int_1 - τ_int_2 * I_T))
    U = grav * k / rad * τ_int_2 * T_v * (cz ** (k - 1) - cz ** (k + 1))
    rc = (rad + z) * torch.cos(φ)
    u_ref = -Ω * rc + torch.sqrt((Ω * rc) ** 2 + rc * U)
    F_z = torch.where(z > z_t, torch.zeros_like(z), 1 - 3 * (z / z_t) ** 2 + 2 * (z / z_t) ** 3)
    arg = torch.sin(φ) * math.sin(φ_c) + torch.cos(φ) * math.cos(φ_c) * torch.cos(λ - λ_c)
    d = rad * torch.acos(arg.clamp(-1, 1))
    c3 = torch.cos(math.pi * d / 2 / d_0) ** 3
    s1 = torch.sin(math.pi * d / 2 / d_0)
    mask = (d > 0) & (d < d_0) & (d != rad * math.pi)
    sd = torch.where(mask, torch.sin(d / rad), torch.ones_like(d))
    f0 = 16 * V_p / 3 / math.sqrt(3.0)
    up = -f0 * F_z * c3 * s1 * (-math.sin(φ_c) * torch.cos(φ)
                                + math.cos(φ_c) * torch.sin(φ) * torch.cos(λ - λ_c)) / sd
    vp = f0 * F_z * c3 * s1 * math.cos(φ_c) * torch.sin(λ - λ_c) / sd
    zero = torch.zeros_like(d)
    us = [u_ref + torch.where(mask, up, zero), torch.where(mask, vp, zero), zero]
    sl, cl, sn, cn = torch.sin(φ), torch.cos(φ), torch.sin(λ), torch.cos(λ)
    uc = [-sn * us[0] - sl * cn * us[1] + cl * cn * us[2],
          cn * us[0] - sl * sn * us[1] + cl * sn * us[2], cl * us[1] + sl * us[2]]
    ρ = pr / (R_d * T_v)
    e_kin = 0.5 * (uc[0] ** 2 + uc[1] ** 2 + uc[2] ** 2)
    e_tot = e_kin + a[:, lay["Φ"]] + p.cv_d * (T_v - p.T_0)
    return torch.stack([ρ, ρ * uc[0], ρ * uc[1], ρ * uc[2], ρ * e_tot], dim=1).to(grid.FT)


def rising_bubble(model, grid, aux, xc=5000.0, zc=2000.0, rc=2000.0, θamplitude=2.0, wind=True):
    """``init_risingbubble!`` (tutorials/Atmos/risingbubble.jl:106-176): warm bubble on a neutrally stratified
    column at rest; with NTracers the tutorial's tracer layer (smoothed: the benchmark mesh does not resolve a
    50 m layer).  `wind` adds a smooth wind (w = 0 on the walls) so that advection, the Rusanov penalty and the
    Smagorinsky closure all have something to act on -- synthetic either way.  Returns (nreal, 5 + N, Np)."""
    p = model.param_set
    lay = aux_layout(model)
    nr = grid.nrealelem
    a = aux.data[:nr]
    x, z = a[:, 0], a[:, 2]
    r = torch.sqrt((x - xc) ** 2 + (z - zc) ** 2)
    θ_ref = model.ref_state.virtual_temperature_profile.T_surface
    θ = θ_ref + torch.where(r <= rc, θamplitude * (1.0 - r / rc), torch.zeros_like(r))
    π_exner = 1.0 - p.grav / (p.cp_d * θ) * z
    ρ = p.MSLP / (p.R_d * θ) * π_exner ** (p.cv_d / p.R_d)
    T = θ * π_exner
    zero = torch.zeros_like(ρ)
    Lx = float(x.max()) + 1e-30
    Lz = float(z.max()) + 1e-30
    u = [zero, zero, zero]
    if wind:
        u = [8 + 2 * torch.sin(2 * math.pi * z / Lz), zero,
             1.5 * torch.sin(2 * math.pi * x / Lx) * torch.sin(math.pi * z / Lz)]
    e_kin = 0.5 * (u[0] ** 2 + u[1] ** 2 + u[2] ** 2)
    cols = [ρ, ρ * u[0], ρ * u[1], ρ * u[2], ρ * (e_kin + a[:, lay["Φ"]] + p.cv_d * (T - p.T_0))]
    if isinstance(model.tracers, bl.NTracers):
        ρχ = torch.where((z > 0.04 * Lz) & (z <= 0.16 * Lz), 0.05 * (1 + 0.5 * torch.sin(2 * math.pi * x / Lx)), zero)
        cols += [ρχ / (i + 1) for i in range(len(model.tracers.δ_χ))]
    return torch.stack(cols, dim=1).to(grid.FT)


def ocean_gyre_state(problem, grid):
    """``ocean_init_state!(::HBModel, ::OceanGyre)`` (ocean_gyre.jl:38-66): state at rest,
    theta = (5 + 4 cos(pi y / Ly)) (1 + z / H).  Returns (Q (nreal,4,Np), aux MPIStateArray)."""
    nr = grid.nrealelem
    y, z = grid.vgeo[:nr, _X2], grid.vgeo[:nr, _X3]
    Q = torch.zeros((nr, 4, grid.Np), dtype=grid.FT, device=grid.device)
    Q[:, 3] = (5 + 4 * torch.cos(y * math.pi / problem.Lʸ)) * (1 + z / problem.H)
    aux = MPIStateArray(grid, 8)
    aux.data[:nr, 0] = y
    return Q, aux
