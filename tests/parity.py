"""Shared harness of the GPU parity tests: builds a case with the CPU oracle, hands the *same*
arrays (byte-identical layout) to libcmdg through the Python mirror of the DGModel interface,
and returns the differences."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import topologies as tp, grids as ogrids, atmos as oatmos, dgmodel as odg  # noqa: E402
from oracle import odesolvers as oode, mpistatearrays as omsa  # noqa: E402


def pkg():
    import __graft_entry__ as ge
    return ge.load_package()


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.sqrt(np.sum((a - b) ** 2) / max(np.sum(b ** 2), 1e-300)))


def device_grid(g, device="cuda"):
    """oracle Grid -> device grid with the reference's array layouts."""
    P = pkg()
    return P.DiscontinuousSpectralElementGrid(
        g.N[0], g.vgeo, g.sgeo, g.vmapM, g.vmapP, g.elemtobndy, g.D[0], g.nreal,
        interiorelems=g.interiorelems, exteriorelems=g.exteriorelems,
        vmapsend=g.vmapsend, vmaprecv=g.vmaprecv, nabrtorank=g.nabrtorank,
        nabrtovmapsend=g.nabrtovmapsend, nabrtovmaprecv=g.nabrtovmaprecv,
        nvertelem=g.topology.stacksize, device=device)


NF = {"rusanov": "RusanovNumericalFlux", "central": "CentralNumericalFluxFirstOrder",
      "roe": "RoeNumericalFlux"}


def device_model(om):
    """oracle DryAtmosModel -> package AtmosModel."""
    P = pkg()
    orient = {"none": P.NoOrientation, "flat": P.FlatOrientation, "spherical": P.SphericalOrientation}[om.orientation]()
    if om.ref_state is None:
        ref = P.NoReferenceState()
    else:
        if om.ref_state.get("profile", "decaying") == "dry_adiabatic":
            prof = P.DryAdiabaticProfile(om.ref_state["T_surf"], om.ref_state["T_min"])
        else:
            prof = P.DecayingTemperatureProfile(om.ref_state["T_surf"], om.ref_state["T_min"], om.ref_state["H_t"])
        ref = P.HydrostaticState(prof, subtract_off=om.ref_state.get("subtract_off", True))
    k = om.turbulence
    turb = {"constant_dynamic": lambda: P.ConstantDynamicViscosity(float(k[1]), bool(k[2])),
            "constant_kinematic": lambda: P.ConstantKinematicViscosity(float(k[1]), bool(k[2])),
            "smagorinsky": lambda: P.SmagorinskyLilly(float(k[1]))}[k[0]]()
    src = tuple(P.RayleighSponge(float(s[1]), float(s[2]), float(s[3]), tuple(float(x) for x in s[4]), float(s[5]))
                if isinstance(s, tuple) else
                {"gravity": P.Gravity, "coriolis": P.Coriolis, "held_suarez": P.HeldSuarezForcing}[s]()
                for s in om.sources)
    bcs = tuple(P.AtmosBC(P.Impenetrable(P.FreeSlip() if b == "freeslip" else P.NoSlip())) for b in om.bcs)
    hyp = None
    if getattr(om, "hyperdiffusion", None) is not None:
        hyp = P.DryBiharmonic(float(om.hyperdiffusion[1]))
    trc = P.NTracers(tuple(om.tracers)) if getattr(om, "tracers", None) else None
    return P.AtmosModel(orientation=orient, ref_state=ref, turbulence=turb, source=src,
                        boundaryconditions=bcs, hyperdiffusion=hyp, tracers=trc)


def make_device_dg(odgm, g, nf, diffusion_direction="every", skip_zero_viscosity=False):
    P = pkg()
    dgrid = device_grid(g)
    m = device_model(odgm.bl)
    aux = P.MPIStateArray(dgrid, odgm.bl.A, data=odgm.state_auxiliary[0].data)
    dd = P.HorizontalDirection() if diffusion_direction == "horizontal" else P.EveryDirection()
    dg = P.DGModel(m, dgrid, getattr(P, NF[nf])(), P.CentralNumericalFluxSecondOrder(),
                   P.CentralNumericalFluxGradient(), state_auxiliary=aux,
                   diffusion_direction=dd, skip_zero_viscosity=skip_zero_viscosity)
    return dg, dgrid


def compare_case(model, g, Q0, nf="rusanov", nsteps=1, dt=None, diffusion_direction="every",
                 skip_zero_viscosity=False, t=0.0):
    """Single-rank comparison of (i) one tendency evaluation with beta = 0 and with increment,
    (ii) ``nsteps`` LSRK54 steps, between the oracle and libcmdg.  ``Q0``: (S, nreal, Np)."""
    P = pkg()
    FT = g.FT
    odgm = odg.DGModel(model, [g], nf, diffusion_direction=diffusion_direction,
                       skip_zero_viscosity=skip_zero_viscosity)
    oQ = omsa.MPIStateArray.from_grid(g, 5)
    np.moveaxis(oQ.data[:g.nreal], 1, 0)[...] = Q0
    omsa.ghost_exchange([oQ])
    oQ0_data = oQ.data.copy()
    dg, dgrid = make_device_dg(odgm, g, nf, diffusion_direction, skip_zero_viscosity)
    dQ = P.MPIStateArray(dgrid, 5, data=oQ.data)
    # (i) tendency, beta = 0
    odQ = oQ.similar()
    odgm([odQ], [oQ], t, 1, 0)
    dT = P.MPIStateArray(dgrid, 5)
    dT.data.fill_(float("nan"))   # beta = 0 must not read the old tendency
    dg(dT, dQ, None, t, 1.0, 0.0)
    res = {"tendency_rel_l2": rel_l2(dT.realdata.cpu().numpy(), odQ.realdata)}
    if model.GF > 0 and not (skip_zero_viscosity and not model.viscous()):
        res["gradflux_rel_l2"] = rel_l2(dg.state_gradient_flux.realdata.cpu().numpy(),
                                        odgm.state_gradient_flux[0].realdata)
    res["aux_theta_T_rel_l2"] = rel_l2(
        dg.state_auxiliary.realdata[:, [model.a_θv, model.a_T]].cpu().numpy(),
        odgm.state_auxiliary[0].realdata[:, [model.a_θv, model.a_T]])
    # increment form with alpha != 1
    odgm([odQ], [oQ], t, 0.5, 2.0)
    dg(dT, dQ, None, t, 0.5, 2.0)
    res["tendency_inc_rel_l2"] = rel_l2(dT.realdata.cpu().numpy(), odQ.realdata)
    # (ii) time stepping
    if nsteps > 0:
        osol = oode.LSRK54CarpenterKennedy(odgm, [oQ], dt=dt, t0=t)
        oode.solve([oQ], osol, numberofsteps=nsteps)
        dsol = P.LSRK54CarpenterKennedy(dg, dQ, dt=dt, t0=t)
        dQ2 = P.MPIStateArray(dgrid, 5, data=dQ.data.cpu().numpy())
        P.solve(dQ, dsol, numberofsteps=nsteps)
        res["state_rel_l2"] = rel_l2(dQ.realdata.cpu().numpy(), oQ.realdata)
        # the un-fused call sequence (cmdg_tendency + cmdg_lsrk_update) must agree as well
        dsol2 = P.LSRK54CarpenterKennedy(dg, dQ2, dt=dt, t0=t)
        P.solve(dQ2, dsol2, numberofsteps=min(nsteps, 2), fused=False)
        if nsteps <= 2:
            res["state_unfused_rel_l2"] = rel_l2(dQ2.realdata.cpu().numpy(), oQ.realdata)
        res["dQ_after_step_max"] = float(dsol.dQ.realdata.abs().max())
        # the same steps through HOST buffers (cmdg_lsrk_steps_host; on one rank the Euler path pipelines the
        # upload with the first stage and the download with the last one): bit-identical to the resident path
        import torch
        Qh = torch.from_numpy(np.ascontiguousarray(oQ0_data[:g.nreal])).pin_memory()
        dsol3 = P.LSRK54CarpenterKennedy(dg, dQ, dt=dt, t0=t)
        k = min(nsteps, 3)
        for i in range(k):      # one call per step, as a host-side time loop does
            dsol3.dostep_host(Qh, t + i * dt, nsteps=1)
        Qd = P.MPIStateArray(dgrid, 5, data=oQ0_data)
        dsol4 = P.LSRK54CarpenterKennedy(dg, Qd, dt=dt, t0=t)
        P.solve(Qd, dsol4, numberofsteps=k)
        res["host_path_max_abs_diff"] = float((Qh - Qd.realdata.cpu()).abs().max())
        assert res["host_path_max_abs_diff"] == 0.0, res["host_path_max_abs_diff"]
    res["launches"] = dg.kernel_launches()
    dg.close()
    return res


def multi_rank_case(model, gs, Q0s, nf, dt, nsteps, rank, world, skip_zero_viscosity,
                    diffusion_direction="every", device=None):
    """One rank's share of a `world`-rank run (torch.distributed initialised by the caller when
    world > 1): the oracle emulates all ranks serially (Hilbert partition, ghost lists, halo
    exchange); this rank runs libcmdg on its partition with the NCCL halo exchange.  Returns the
    maxima over ranks of
      halo_exact        ghost face nodes after cmdg_exchange_begin/end == the oracle's (array_equal;
                        ghosts start as NaN, so only the exchange can have filled them),
      tendency_rel_l2   one cmdg_tendency (reference order: interior while the halo is in flight),
      state_rel_l2      `nsteps` fused LSRK54 steps (cmdg_lsrk_steps: exterior-first / overlapped).
    """
    import torch
    import torch.distributed as dist
    P = pkg()
    odgm = odg.DGModel(model, gs, nf, skip_zero_viscosity=skip_zero_viscosity,
                       diffusion_direction=diffusion_direction)
    oQ = []
    for g, q0 in zip(gs, Q0s):
        q = omsa.MPIStateArray.from_grid(g, 5)
        np.moveaxis(q.data[:g.nreal], 1, 0)[...] = q0
        oQ.append(q)
    g = gs[rank]
    device = device or f"cuda:{torch.cuda.current_device()}"
    dgrid = device_grid(g, device=device)
    m = device_model(model)
    aux = P.MPIStateArray(dgrid, model.A, data=odgm.state_auxiliary[rank].data)
    dd = P.HorizontalDirection() if diffusion_direction == "horizontal" else P.EveryDirection()
    dg = P.DGModel(m, dgrid, getattr(P, NF[nf])(), P.CentralNumericalFluxSecondOrder(),
                   P.CentralNumericalFluxGradient(), state_auxiliary=aux, diffusion_direction=dd,
                   skip_zero_viscosity=skip_zero_viscosity)
    if world > 1:
        uid = [P.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        dg.comm_init(uid[0], rank, world)
    # ghost elements of the device state start as NaN: only the exchange may fill them
    data = oQ[rank].data.copy()
    data[g.nreal:] = np.nan
    dQ = P.MPIStateArray(dgrid, 5, data=data)
    # halo exchange known answer: ghost face nodes must equal the oracle's after exchange
    omsa.ghost_exchange(oQ)
    halo_exact = True
    if world > 1:
        dg.ghost_exchange(dQ)
        e, n = np.divmod(g.vmaprecv - 1, g.Np)
        got = dQ.data.cpu().numpy()[e, :, n]
        halo_exact = bool(np.array_equal(got, oQ[rank].data[e, :, n]))
    # tendency
    odQ = [q.similar() for q in oQ]
    odgm(odQ, oQ, 0.0, 1, 0)
    dT = P.MPIStateArray(dgrid, 5)
    dT.data.fill_(float("nan"))
    dg(dT, dQ, None, 0.0, 1.0, 0.0)
    r1 = rel_l2(dT.realdata.cpu().numpy(), odQ[rank].realdata)
    # fused steps
    osol = oode.LSRK54CarpenterKennedy(odgm, oQ, dt=dt)
    oode.solve(oQ, osol, numberofsteps=nsteps)
    dsol = P.LSRK54CarpenterKennedy(dg, dQ, dt=dt)
    P.solve(dQ, dsol, numberofsteps=nsteps)
    r2 = rel_l2(dQ.realdata.cpu().numpy(), oQ[rank].realdata)
    res = torch.tensor([r1, r2, 0.0 if halo_exact else 1.0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(res, op=dist.ReduceOp.MAX)
    out = {"n_ranks": world, "halo_exact": float(res[2]) == 0.0, "tendency_rel_l2": float(res[0]),
           "state_rel_l2": float(res[1]), "nsteps": nsteps,
           "nreal_per_rank": [int(x.nreal) for x in gs], "ninterior_per_rank": [int(len(x.interiorelems)) for x in gs],
           "nghost_per_rank": [int(x.nelem - x.nreal) for x in gs],
           "nneighbours_per_rank": [int(len(x.nabrtorank)) for x in gs]}
    dg.close()
    return out


def vortex_setup(nelem=(5, 5, 1), FT=np.float64, csize=1):
    ps = oatmos.Params(FT)
    setup = oatmos.IsentropicVortexSetup(ps, FT)
    L = setup.domain_halflength
    br = tuple(np.linspace(-L, L, n + 1).astype(FT) for n in nelem)
    topos = tp.BrickTopology(csize, br, periodicity=(True, True, True))
    gs = [ogrids.Grid(t, 4, FT=FT) for t in topos]
    model = oatmos.DryAtmosModel(FT, orientation="none", ref_state=None,
                                 turbulence=("constant_dynamic", 0.0, False), sources=())
    elementsize = min(float(np.min(np.diff(b))) for b in br)
    dt = elementsize / float(oatmos.soundspeed_air(ps, setup.T_inf)) / 4 ** 2
    return model, gs, setup, dt


def vortex_case(nelem=(5, 5, 1), nf="rusanov", nsteps=1, FT=np.float64, skip_zero_viscosity=True):
    model, gs, setup, dt = vortex_setup(nelem, FT)
    g = gs[0]
    Q0 = setup(g.vgeo[:g.nreal, ogrids._x1], g.vgeo[:g.nreal, ogrids._x2],
               g.vgeo[:g.nreal, ogrids._x3], FT(0))
    res = compare_case(model, g, Q0, nf=nf, nsteps=nsteps, dt=dt,
                       skip_zero_viscosity=skip_zero_viscosity)
    if FT == np.float32:
        res.update(float32_truth_errors(model, g, Q0, nf, skip_zero_viscosity))
    return res


def float32_truth_errors(model, g, Q0, nf, skip_zero_viscosity, diffusion_direction="every"):
    """Float32 conditioning check: evaluate the oracle in Float64 on the *same Float32 inputs*
    (grid arrays and state upcast) and report the Float32 oracle's and libcmdg's errors
    against that truth.  The isentropic vortex lives on a 0.1 m box, so a Float32 tendency
    carries ~1e-5 of rounding whatever the evaluation order."""
    import copy
    P = pkg()
    g64 = copy.copy(g)
    g64.FT = np.float64
    g64.vgeo, g64.sgeo = g.vgeo.astype(np.float64), g.sgeo.astype(np.float64)
    g64.D = [d.astype(np.float64) for d in g.D]
    m64 = oatmos.DryAtmosModel(np.float64, orientation=model.orientation, ref_state=model.ref_state,
                               turbulence=model.turbulence, sources=model.sources, bcs=model.bcs,
                               hyperdiffusion=getattr(model, "hyperdiffusion", None))
    out = {}
    tend = {}
    aux32 = odg.DGModel(model, [g], nf, skip_zero_viscosity=skip_zero_viscosity,
                        diffusion_direction=diffusion_direction).state_auxiliary[0].data
    for name, gg, mm in (("truth", g64, m64), ("oracle32", g, model)):
        dgm = odg.DGModel(mm, [gg], nf, skip_zero_viscosity=skip_zero_viscosity,
                          diffusion_direction=diffusion_direction)
        if name == "truth":
            # same Float32 auxiliary state, upcast (the reference state is part of the input)
            dgm.state_auxiliary[0].data[...] = aux32.astype(np.float64)
        q = omsa.MPIStateArray.from_grid(gg, 5)
        np.moveaxis(q.data[:gg.nreal], 1, 0)[...] = Q0
        omsa.ghost_exchange([q])
        dq = q.similar()
        dgm([dq], [q], 0.0, 1, 0)
        tend[name] = dq.realdata.astype(np.float64)
        if name == "oracle32":
            dg, dgrid = make_device_dg(dgm, gg, nf, diffusion_direction, skip_zero_viscosity)
            dQ = P.MPIStateArray(dgrid, 5, data=q.data)
            dT = P.MPIStateArray(dgrid, 5)
            dg(dT, dQ, None, 0.0, 1.0, 0.0)
            tend["cuda32"] = dT.realdata.cpu().numpy().astype(np.float64)
            dg.close()
    out["oracle32_vs_truth"] = rel_l2(tend["oracle32"], tend["truth"])
    out["cuda32_vs_truth"] = rel_l2(tend["cuda32"], tend["truth"])
    return out


def gcm_setup(ne=3, nvert=2, FT=np.float64, turbulence=("constant_kinematic", 0.0, False),
              csize=1, domain_height=30e3, hyperdiffusion=None):
    """Dry baroclinic wave on the cubed sphere (experiments/TestCase/baroclinic_wave.jl with
    explicit stepping, no hyperdiffusion): spherical orientation, hydrostatic reference state
    (DecayingTemperatureProfile 290/220/8 km), Gravity + Coriolis, free-slip walls."""
    ps = oatmos.Params(FT)
    a = float(ps.planet_radius)
    Rrange = np.linspace(a, a + domain_height, nvert + 1)
    topos = tp.StackedCubedSphereTopology(csize, ne, Rrange, boundary=(1, 2))
    gs = [ogrids.Grid(t, 4, FT=FT, meshwarp=tp.equiangular_cubed_sphere_warp) for t in topos]
    model = oatmos.DryAtmosModel(
        FT, orientation="spherical",
        ref_state=dict(T_surf=290.0, T_min=220.0, H_t=8e3, subtract_off=True),
        turbulence=turbulence, sources=("gravity", "coriolis"), bcs=("freeslip", "freeslip"),
        hyperdiffusion=hyperdiffusion)
    return model, gs


def gcm_case(nf="rusanov", nsteps=1, dt=0.5, turbulence=("constant_kinematic", 0.0, False),
             diffusion_direction="every", skip_zero_viscosity=True, ne=3, nvert=2):
    model, gs = gcm_setup(ne, nvert, turbulence=turbulence)
    g = gs[0]
    odgm = odg.DGModel(model, [g], nf)
    aux = np.moveaxis(odgm.state_auxiliary[0].data[:g.nreal], 1, 0)
    Q0 = oatmos.init_baroclinic_wave(model, aux)
    return compare_case(model, g, Q0, nf=nf, nsteps=nsteps, dt=dt,
                        diffusion_direction=diffusion_direction,
                        skip_zero_viscosity=skip_zero_viscosity)


def heldsuarez_case(nf="rusanov", nsteps=1, dt=0.5, turbulence=("smagorinsky", 0.21),
                    diffusion_direction="horizontal", ne=3, nvert=3):
    """Held-Suarez dry GCM (tutorials/Atmos/heldsuarez.jl:160-201 without hyperdiffusion, explicit
    stepping): Smagorinsky + Gravity, Coriolis, HeldSuarezForcing, RayleighSponge(30 km, 12 km,
    1/900 s, 0, 2); state = reference state + smooth wind/temperature perturbation so that every
    forcing term (relaxation, boundary-layer friction, sponge) is active."""
    model, gs = gcm_setup(ne, nvert, turbulence=turbulence)
    model.sources = ("gravity", "coriolis", "held_suarez",
                     ("rayleigh_sponge", 30e3, 12e3, 1 / 60 / 15, (0.0, 0.0, 0.0), 2.0))
    g = gs[0]
    odgm = odg.DGModel(model, [g], nf, diffusion_direction=diffusion_direction)
    aux = np.moveaxis(odgm.state_auxiliary[0].data[:g.nreal], 1, 0)
    Q0 = oatmos.init_baroclinic_wave(model, aux)
    return compare_case(model, g, Q0, nf=nf, nsteps=nsteps, dt=dt,
                        diffusion_direction=diffusion_direction, skip_zero_viscosity=False)


def filter_case(direction="every", target="indices", nsteps=0, dt=0.5):
    """Filters.apply! parity: FilterIndices(1, 3, 5) / AtmosFilterPerturbations on a perturbed
    baroclinic-wave state (cubed sphere, ExponentialFilter(grid, 0, 10) as the GCM drivers use);
    with nsteps > 0 also the per-step filter inside the fused stepper (cbfilter callback)."""
    import types
    from oracle import filters as ofilters
    P = pkg()
    model, gs = gcm_setup(3, 2)
    g = gs[0]
    odgm = odg.DGModel(model, [g], "rusanov", skip_zero_viscosity=True)
    aux = np.moveaxis(odgm.state_auxiliary[0].data[:g.nreal], 1, 0)
    Q0 = oatmos.init_baroclinic_wave(model, aux)
    rng = np.random.default_rng(1)
    Q0 = Q0 * (1 + 1e-3 * rng.standard_normal(Q0.shape))
    oQ = omsa.MPIStateArray.from_grid(g, 5)
    np.moveaxis(oQ.data[:g.nreal], 1, 0)[...] = Q0
    W = ofilters.exponential_filter_matrix(g.xi[0], 0, 10)
    dg, dgrid = make_device_dg(odgm, g, "rusanov", skip_zero_viscosity=True)
    dQ = P.MPIStateArray(dgrid, 5, data=oQ.data)
    filt = types.SimpleNamespace(filter_matrix=W)
    if target == "indices":
        otarget, dtarget = ofilters.FilterIndices(0, 2, 4), P.FilterIndices(1, 3, 5)
    else:
        otarget, dtarget = ofilters.AtmosFilterPerturbations(model), P.AtmosFilterPerturbations(dg.balance_law)
    ddir = {"every": P.EveryDirection, "horizontal": P.HorizontalDirection, "vertical": P.VerticalDirection}[direction]()
    res = {}
    if nsteps == 0:
        before = oQ.data.copy()
        ofilters.apply(oQ.data, otarget, g, W, state_auxiliary=odgm.state_auxiliary[0].data, direction=direction)
        dg.apply_filter(dQ, dtarget, filt, ddir)
        got = dQ.realdata.cpu().numpy()
        res["filtered_rel_l2"] = rel_l2(got, oQ.realdata)
        # relative to what the filter changed (the perturbation), not to the O(1) state
        res["change_rel_l2"] = rel_l2(got - before[:g.nreal], oQ.realdata - before[:g.nreal])
    else:
        osol = oode.LSRK54CarpenterKennedy(odgm, [oQ], dt=dt, t0=0.0)
        for _ in range(nsteps):
            oode.solve([oQ], osol, numberofsteps=1)
            ofilters.apply(oQ.data, otarget, g, W, state_auxiliary=odgm.state_auxiliary[0].data, direction=direction)
            omsa.ghost_exchange([oQ])
        dg.set_step_filter(dtarget, filt, ddir)
        dsol = P.LSRK54CarpenterKennedy(dg, dQ, dt=dt, t0=0.0)
        P.solve(dQ, dsol, numberofsteps=nsteps)
        res["state_rel_l2"] = rel_l2(dQ.realdata.cpu().numpy(), oQ.realdata)
    res["launches"] = dg.kernel_launches()
    dg.close()
    return res


def courant_case():
    """courant(local_courant, dg, m, Q, dt, t, direction) for the three AtmosModel numbers and the
    three directions, Smagorinsky on the cubed sphere (gradient flux from one tendency call)."""
    from oracle import courant as ocourant
    P = pkg()
    model, gs = gcm_setup(3, 2, turbulence=("smagorinsky", 0.21))
    g = gs[0]
    odgm = odg.DGModel(model, [g], "rusanov", diffusion_direction="horizontal")
    aux = np.moveaxis(odgm.state_auxiliary[0].data[:g.nreal], 1, 0)
    Q0 = oatmos.init_baroclinic_wave(model, aux)
    oQ = omsa.MPIStateArray.from_grid(g, 5)
    np.moveaxis(oQ.data[:g.nreal], 1, 0)[...] = Q0
    omsa.ghost_exchange([oQ])
    odQ = oQ.similar()
    odgm([odQ], [oQ], 0.0, 1, 0)
    dg, dgrid = make_device_dg(odgm, g, "rusanov", "horizontal")
    dQ = P.MPIStateArray(dgrid, 5, data=oQ.data)
    dT = P.MPIStateArray(dgrid, 5)
    dg(dT, dQ, None, 0.0, 1.0, 0.0)
    out = {}
    dirs = {"every": P.EveryDirection, "horizontal": P.HorizontalDirection, "vertical": P.VerticalDirection}
    for kind in ("advective", "nondiffusive", "diffusive"):
        for dname, D in dirs.items():
            ref = ocourant.courant(model, g, oQ.data, odgm.state_auxiliary[0].data,
                                   odgm.state_gradient_flux[0].data, 0.5, kind, dname)
            got = dg.courant(kind, dQ, 0.5, D())
            out[(kind, dname)] = (got, ref)
    dg.close()
    return out


def box_setup(nelem=(3, 2, 3), FT=np.float64, turbulence=("smagorinsky", 0.21), csize=1,
              periodic_z=False, hyperdiffusion=None, periodic_xy=(True, True)):
    """LES-like box (tutorials/Atmos/risingbubble.jl without tracers): flat orientation,
    hydrostatic reference state, gravity, walls top/bottom, periodic horizontally."""
    br = (np.linspace(0, 1500, nelem[0] + 1), np.linspace(0, 1000, nelem[1] + 1),
          np.linspace(0, 1500, nelem[2] + 1))
    topos = tp.StackedBrickTopology(csize, br, periodicity=(periodic_xy[0], periodic_xy[1], periodic_z),
                                    boundary=((0, 0) if periodic_xy[0] else (1, 2),
                                              (0, 0) if periodic_xy[1] else (2, 1), (1, 2)))
    gs = [ogrids.Grid(t, 4, FT=FT) for t in topos]
    model = oatmos.DryAtmosModel(
        FT, orientation="flat",
        ref_state=dict(T_surf=300.0, T_min=220.0, H_t=8e3, subtract_off=True),
        turbulence=turbulence, sources=("gravity",), bcs=("freeslip", "noslip"),
        hyperdiffusion=hyperdiffusion)
    return model, gs


def bubble_state(model, g, aux):
    """Warm bubble + sheared wind on top of the reference state (smooth, non-trivial gradients)."""
    ps = model.ps
    FT = g.FT
    x, y, z = aux[0], aux[1], aux[2]
    r = np.sqrt((x - 750) ** 2 + (z - 500) ** 2 + (y - 500) ** 2) / 400
    dT = np.where(r < 1, 2.0 * np.cos(np.pi * r / 2) ** 2, 0.0)
    Tref = aux[model.a_ref["T"]]
    p = aux[model.a_ref["p"]]
    T = Tref + dT
    ρ = oatmos.air_density(ps, T, p)
    u = [5 + 3 * np.sin(2 * np.pi * z / 1500), 2 * np.cos(2 * np.pi * x / 1500), 1.5 * np.sin(2 * np.pi * y / 1000) * np.sin(np.pi * z / 1500)]
    e_kin = 0.5 * (u[0] ** 2 + u[1] ** 2 + u[2] ** 2)
    e_tot = oatmos.total_energy(ps, e_kin, aux[model.a_Φ], T)
    return np.stack([ρ, ρ * u[0], ρ * u[1], ρ * u[2], ρ * e_tot]).astype(FT)


def box_case(nf="rusanov", nsteps=1, dt=0.01, turbulence=("smagorinsky", 0.21),
             diffusion_direction="every", nelem=(3, 2, 3), FT=np.float64):
    model, gs = box_setup(nelem, FT=FT, turbulence=turbulence)
    g = gs[0]
    odgm = odg.DGModel(model, [g], nf, diffusion_direction=diffusion_direction)
    aux = np.moveaxis(odgm.state_auxiliary[0].data[:g.nreal], 1, 0)
    Q0 = bubble_state(model, g, aux)
    res = compare_case(model, g, Q0, nf=nf, nsteps=nsteps, dt=dt,
                       diffusion_direction=diffusion_direction)
    if FT == np.float32:
        res.update(float32_truth_errors(model, g, Q0, nf, False, diffusion_direction))
    return res


def risingbubble_case(nelem=(10, 1, 10), nsteps=2, dt=0.4, FT=np.float64, nf="rusanov", tracers=None,
                      turbulence=("smagorinsky", 0.21), skip_zero_viscosity=False):
    """BASELINE.json configs[0], tutorials/Atmos/risingbubble.jl: with `tracers=(1, 2, 3, 4)` as shipped
    (NTracers{4} injected in a layer, `init_risingbubble!`), with `tracers=None` minus the four passive
    tracers (SURVEY 8.0 note): 10 km x 500 m x 10 km box, periodic x / y, free-slip walls in z, N = 4,
    SmagorinskyLilly(C_smag = 0.21), HydrostaticState(DryAdiabaticProfile(300 K, 0 K)), Gravity,
    Rusanov, the tutorial's warm-bubble initial state, LSRK144NiegemannDiehlBusch (the tutorial's
    mesh is 20 x 1 x 20 elements of 500 m; `nelem` coarsens it for the CPU oracle)."""
    P = pkg()
    br = (np.linspace(0, 10000, nelem[0] + 1), np.linspace(0, 500, nelem[1] + 1),
          np.linspace(0, 10000, nelem[2] + 1))
    topos = tp.StackedBrickTopology(1, br, periodicity=(True, True, False), boundary=((0, 0), (0, 0), (1, 2)))
    g = ogrids.Grid(topos[0], 4, FT=FT)
    model = oatmos.DryAtmosModel(
        FT, orientation="flat",
        ref_state=dict(profile="dry_adiabatic", T_surf=300.0, T_min=0.0, H_t=0.0, subtract_off=True),
        turbulence=turbulence, sources=("gravity",), bcs=("freeslip", "freeslip"),
        tracers=tracers)
    S = model.S
    odgm = odg.DGModel(model, [g], nf, skip_zero_viscosity=skip_zero_viscosity)
    aux = np.moveaxis(odgm.state_auxiliary[0].data[:g.nreal], 1, 0)
    Q0 = oatmos.init_risingbubble(model, aux)
    if tracers:
        # the tutorial injects rho*chi = 0.05 for 500 < z <= 550 (a layer thinner than the coarsened test
        # mesh resolves: widened to 400 < z <= 1600 and tapered so that gradients are non-trivial), then
        # rho*chi, /2, /3, /4
        z, x = aux[2], aux[0]
        ρχ = np.where((z > 400) & (z <= 1600), 0.05 * (1 + 0.5 * np.sin(2 * np.pi * x / 10000)), 0.0).astype(FT)
        # the tutorial starts at rest, where the tracer tendency is a 1e-5 m^2/s diffusion only: add a smooth
        # wind (w = 0 on the walls) so that advection, the Rusanov penalty and the Smagorinsky D_t all act
        u = [8 + 2 * np.sin(2 * np.pi * z / 10000), 0 * z, 1.5 * np.sin(2 * np.pi * x / 10000) * np.sin(np.pi * z / 10000)]
        ρ = Q0[0]
        Q0[1], Q0[2], Q0[3] = ρ * u[0], ρ * u[1], ρ * u[2]
        Q0[4] = Q0[4] + ρ * (u[0] ** 2 + u[1] ** 2 + u[2] ** 2) / 2
        Q0 = np.concatenate([Q0, np.stack([ρχ / (i + 1) for i in range(len(tracers))])]).astype(FT)
    oQ = omsa.MPIStateArray.from_grid(g, S)
    np.moveaxis(oQ.data[:g.nreal], 1, 0)[...] = Q0
    omsa.ghost_exchange([oQ])
    dg, dgrid = make_device_dg(odgm, g, nf, skip_zero_viscosity=skip_zero_viscosity)
    dQ = P.MPIStateArray(dgrid, S, data=oQ.data)
    odQ = oQ.similar()
    odgm([odQ], [oQ], 0.0, 1, 0)
    dT = P.MPIStateArray(dgrid, S)
    dT.data.fill_(float("nan"))
    dg(dT, dQ, None, 0.0, 1.0, 0.0)
    got_t, got_gf = dT.realdata.cpu().numpy(), dg.state_gradient_flux.realdata.cpu().numpy()
    ngf = 10 if turbulence[0] == "smagorinsky" else 9
    second = not (skip_zero_viscosity and not model.viscous())
    res = {"tendency_rel_l2": rel_l2(got_t[:, :5], odQ.realdata[:, :5])}
    if second:
        res["gradflux_rel_l2"] = rel_l2(got_gf[:, :ngf], odgm.state_gradient_flux[0].realdata[:, :ngf])
    if tracers:
        res["tracer_tendency_rel_l2"] = [rel_l2(got_t[:, 5 + i], odQ.realdata[:, 5 + i]) for i in range(len(tracers))]
        if second:
            res["tracer_gradflux_rel_l2"] = rel_l2(got_gf[:, ngf:], odgm.state_gradient_flux[0].realdata[:, ngf:])
        # increment form
        odgm([odQ], [oQ], 0.0, 0.5, 2.0)
        dg(dT, dQ, None, 0.0, 0.5, 2.0)
        res["tendency_inc_rel_l2"] = rel_l2(dT.realdata.cpu().numpy(), odQ.realdata)
    Q_init = oQ.realdata.copy()
    osol = oode.LSRK144NiegemannDiehlBusch(odgm, [oQ], dt=dt, t0=0.0)
    oode.solve([oQ], osol, numberofsteps=nsteps)
    dsol = P.LSRK144NiegemannDiehlBusch(dg, dQ, dt=dt, t0=0.0)
    P.solve(dQ, dsol, numberofsteps=nsteps)
    got = dQ.realdata.cpu().numpy()
    res["state_rel_l2"] = rel_l2(got, oQ.realdata)
    # relative to what the steps changed (the state itself is dominated by the hydrostatic background)
    res["change_rel_l2"] = rel_l2(got - Q_init, oQ.realdata - Q_init)
    if tracers:
        res["tracer_state_rel_l2"] = rel_l2(got[:, 5:], oQ.realdata[:, 5:])
        res["tracer_change_rel_l2"] = rel_l2(got[:, 5:] - Q_init[:, 5:], oQ.realdata[:, 5:] - Q_init[:, 5:])
    res["launches"] = dg.kernel_launches()
    dg.close()
    return res


def balance_case(config="GCM", nf="roe", nsteps=10):
    """test/Atmos/Model/discrete_hydrostatic_balance.jl on the device: the reference state
    (subtract_off = false, Gravity) must stay put.  Returns the device's and the oracle's relative
    drift after `nsteps` LSRK54 steps and their mutual difference."""
    from tests.test_oracle_balance import _grid, HEIGHT  # noqa: F401
    P = pkg()
    ps = oatmos.Params(np.float64)
    g, orientation = _grid(config)
    T_surf = float(ps.T_surf_ref)
    model = oatmos.DryAtmosModel(np.float64, orientation=orientation,
                                 ref_state=dict(T_surf=T_surf, T_min=float(ps.T_min_ref),
                                                H_t=float(ps.R_d) * T_surf / float(ps.grav), subtract_off=False),
                                 turbulence=("constant_dynamic", 0.0, False), sources=("gravity",),
                                 bcs=("freeslip", "freeslip"))
    odgm = odg.DGModel(model, [g], nf, skip_zero_viscosity=True)
    aux = odgm.state_auxiliary[0].data
    oQ = omsa.MPIStateArray.from_grid(g, 5)
    oQ.data[:, 0] = aux[:, model.a_ref["ρ"]]
    oQ.data[:, 4] = aux[:, model.a_ref["ρe"]]
    Q0 = oQ.data.copy()
    vg = g.vgeo[:g.nreal]
    x = np.stack([vg[:, ogrids._x1], vg[:, ogrids._x2], vg[:, ogrids._x3]], axis=-1).reshape(g.nreal, 5, 5, 5, 3)
    dmin = min(np.linalg.norm(np.diff(x, axis=ax), axis=-1).min() for ax in (1, 2, 3))
    dt = 0.1 * dmin / float(oatmos.soundspeed_air(ps, np.float64(T_surf)))
    dg, dgrid = make_device_dg(odgm, g, nf, skip_zero_viscosity=True)
    dQ = P.MPIStateArray(dgrid, 5, data=oQ.data)
    osol = oode.LSRK54CarpenterKennedy(odgm, [oQ], dt=dt, t0=0.0)
    oode.solve([oQ], osol, numberofsteps=nsteps)
    dsol = P.LSRK54CarpenterKennedy(dg, dQ, dt=dt, t0=0.0)
    P.solve(dQ, dsol, numberofsteps=nsteps)
    got = dQ.realdata.cpu().numpy()
    M = vg[:, ogrids._M][:, None, :]
    nrm = np.sqrt(np.sum(M * Q0[:g.nreal] ** 2))
    res = {"device_drift": float(np.sqrt(np.sum(M * (got - Q0[:g.nreal]) ** 2)) / nrm),
           "oracle_drift": float(np.sqrt(np.sum(M * (oQ.realdata - Q0[:g.nreal]) ** 2)) / nrm),
           "state_rel_l2": rel_l2(got, oQ.realdata)}
    dg.close()
    return res


def hyperdiffusion_case(kind="sphere", nsteps=2, turbulence=("constant_kinematic", 0.0, False),
                        tau=None, nf="rusanov", FT=np.float64):
    """DryBiharmonic(tau) with diffusion_direction = HorizontalDirection(), as the GCM
    drivers configure it (experiments/TestCase/baroclinic_wave.jl:179, AtmosGCM/heldsuarez.jl:195).
    kind: "sphere" (cubed sphere, panel flips), "box" (flat box, periodic horizontally), "walled_box"
    (horizontal walls: the no-op boundary states of the divergence / higher-order fluxes).
    Besides the usual differences returns the hyperdiffusion's own share of the oracle tendency and
    the device-vs-oracle difference of that share alone."""
    # 8 h on the sphere as the drivers set it; 30 s on the 1.5 km box so that the term is of comparable
    # relative size there
    hyp = ("dry_biharmonic", tau or (8 * 3600.0 if kind == "sphere" else 30.0))
    if kind == "sphere":
        model, gs = gcm_setup(3, 2, FT=FT, turbulence=turbulence, hyperdiffusion=hyp)
        model0, _ = gcm_setup(3, 2, FT=FT, turbulence=turbulence)
        dt = 0.5
    else:
        pxy = (True, True) if kind == "box" else (False, False)
        model, gs = box_setup((3, 2, 3), FT=FT, turbulence=turbulence, hyperdiffusion=hyp, periodic_xy=pxy)
        model0, _ = box_setup((3, 2, 3), FT=FT, turbulence=turbulence, periodic_xy=pxy)
        dt = 0.01
    g = gs[0]
    odgm = odg.DGModel(model, [g], nf, diffusion_direction="horizontal")
    aux = np.moveaxis(odgm.state_auxiliary[0].data[:g.nreal], 1, 0)
    Q0 = oatmos.init_baroclinic_wave(model, aux) if kind == "sphere" else bubble_state(model, g, aux)
    res = compare_case(model, g, Q0, nf=nf, nsteps=nsteps, dt=dt, diffusion_direction="horizontal")
    if FT == np.float32:
        return res
    # the hyperdiffusion's own contribution: tendency(with) - tendency(without), oracle and device
    P = pkg()
    tend = {}
    for name, mm in (("with", model), ("without", model0)):
        o = odg.DGModel(mm, [g], nf, diffusion_direction="horizontal")
        q = omsa.MPIStateArray.from_grid(g, 5)
        np.moveaxis(q.data[:g.nreal], 1, 0)[...] = Q0
        omsa.ghost_exchange([q])
        dq = q.similar()
        o([dq], [q], 0.0, 1, 0)
        dg, dgrid = make_device_dg(o, g, nf, "horizontal")
        dQ = P.MPIStateArray(dgrid, 5, data=q.data)
        dT = P.MPIStateArray(dgrid, 5)
        dg(dT, dQ, None, 0.0, 1.0, 0.0)
        tend[name] = (dq.realdata.copy(), dT.realdata.cpu().numpy())
        dg.close()
    osh = tend["with"][0] - tend["without"][0]
    dsh = tend["with"][1] - tend["without"][1]
    res["hyper_share_of_tendency"] = rel_l2(tend["with"][0], tend["without"][0])
    res["hyper_share_rel_l2"] = rel_l2(dsh, osh)
    return res


# ---------------------------------------------------------------------------------------
# ocean HBModel
# ---------------------------------------------------------------------------------------
def ocean_setup(csize=1, nelem=(5, 5, 5)):
    from tests.test_oracle_ocean import gyre_setup
    return gyre_setup(csize, nelem)


def device_ocean_dg(omodel, g, odgm, nf="rusanov", rank=0):
    P = pkg()
    dgrid = P.DiscontinuousSpectralElementGrid(
        g.N[0], g.vgeo, g.sgeo, g.vmapM, g.vmapP, g.elemtobndy, g.D[0], g.nreal,
        interiorelems=g.interiorelems, exteriorelems=g.exteriorelems,
        vmapsend=g.vmapsend, vmaprecv=g.vmaprecv, nabrtorank=g.nabrtorank,
        nabrtovmapsend=g.nabrtovmapsend, nabrtovmaprecv=g.nabrtovmaprecv,
        nvertelem=g.topology.stacksize, Imat=g.Imat[2], xi=g.xi[2])
    pr = omodel.problem
    m = P.HBModel(P.OceanGyre(pr.Lx, pr.Ly, pr.H), cʰ=float(omodel.ch))
    aux = P.MPIStateArray(dgrid, 8, data=odgm.state_auxiliary[rank].data)
    md = dict(vert_filter=P.CutoffFilter(dgrid, 3), exp_filter=P.ExponentialFilter(dgrid, 1, 8))
    dg = P.DGModel(m, dgrid, getattr(P, NF[nf])(), P.CentralNumericalFluxSecondOrder(),
                   P.CentralNumericalFluxGradient(), state_auxiliary=aux, modeldata=md)
    return dg, dgrid


def ocean_case(nsteps=2, dt=120.0, nelem=(5, 5, 5), spinup=3, nf="rusanov"):
    """HBModel: one evaluation (filters + gradient pass + column integrals + tendency) and
    LSRK144 steps, oracle vs libcmdg; the state is first spun up by the oracle so that every
    term (advection by w, pressure, wind stress, convective adjustment) is active."""
    from oracle import ocean as oocean
    P = pkg()
    model, gs, prob = ocean_setup(1, nelem)
    g = gs[0]
    odgm = odg.DGModel(model, [g], nf)
    oQ = odg.init_ode_state(odgm, lambda x1, x2, x3, a, t: prob.init_state(x1, x2, x3), 0.0)
    osol = oode.LSRK144NiegemannDiehlBusch(odgm, oQ, dt=dt, t0=0.0)
    if spinup:
        oode.solve(oQ, osol, numberofsteps=spinup)
    dg, dgrid = device_ocean_dg(model, g, odgm, nf)
    dQ = P.MPIStateArray(dgrid, 4, data=oQ[0].data)
    odQ = [oQ[0].similar()]
    odgm(odQ, oQ, 0.0, 1, 0)
    dT = P.MPIStateArray(dgrid, 4)
    dT.data.fill_(float("nan"))
    dg(dT, dQ, None, 0.0, 1.0, 0.0)
    res = {
        "tendency_rel_l2": rel_l2(dT.realdata.cpu().numpy(), odQ[0].realdata),
        "filtered_state_rel_l2": rel_l2(dQ.realdata.cpu().numpy(), oQ[0].realdata),
        "gradflux_rel_l2": rel_l2(dg.state_gradient_flux.realdata.cpu().numpy(),
                                  odgm.state_gradient_flux[0].realdata),
        "aux_rel_l2": rel_l2(dg.state_auxiliary.realdata[:, 1:4].cpu().numpy(),
                             odgm.state_auxiliary[0].realdata[:, 1:4]),
    }
    osol2 = oode.LSRK144NiegemannDiehlBusch(odgm, oQ, dt=dt, t0=0.0)
    oode.solve(oQ, osol2, numberofsteps=nsteps)
    dsol = P.LSRK144NiegemannDiehlBusch(dg, dQ, dt=dt, t0=0.0)
    P.solve(dQ, dsol, numberofsteps=nsteps)
    res["state_rel_l2"] = rel_l2(dQ.realdata.cpu().numpy(), oQ[0].realdata)
    res["launches"] = dg.kernel_launches()
    dg.close()
    return res


def ocean_refvals_on_device():
    """test_ocean_gyre_short.jl run entirely through libcmdg: min/max/mean/std after 1 h."""
    from oracle import ocean as oocean
    P = pkg()
    model, gs, prob = ocean_setup(1, (5, 5, 5))
    g = gs[0]
    odgm = odg.DGModel(model, [g], "rusanov")
    oQ = odg.init_ode_state(odgm, lambda x1, x2, x3, a, t: prob.init_state(x1, x2, x3), 0.0)
    dg, dgrid = device_ocean_dg(model, g, odgm)
    dQ = P.MPIStateArray(dgrid, 4, data=oQ[0].data)
    sol = P.LSRK144NiegemannDiehlBusch(dg, dQ, dt=120.0, t0=0.0)
    P.solve(dQ, sol, numberofsteps=30)

    class Wrap:
        def __init__(self, t, nreal):
            self.realdata = t.cpu().numpy()[:nreal]
    out = {}
    for name, arr in (("Q", dQ), ("aux", dg.state_auxiliary)):
        w = Wrap(arr.data, g.nreal)
        for ivar in range(4):
            out[(name, ivar)] = oocean.statecheck(w, ivar)
    dg.close()
    return out


def ocean_windstress_on_device():
    """test/Ocean/HydrostaticBoussinesq/test_windstress_short.jl (explicit) run entirely through libcmdg:
    HomogeneousBox with (NoSlip, Insulating) sides, (FreeSlip, Insulating) bottom, (KinematicStress,
    Insulating) surface -- the ocean boundary-condition branches the gyre does not use -- 20 LSRK144
    steps of 180 s; returns min/max/mean/std per field and the oracle's final state difference."""
    from oracle import ocean as oocean
    from tests.test_oracle_ocean_windstress import HomogeneousBox
    P = pkg()
    Lx, Ly, H = 1e6, 1e6, 400.0
    br = (np.linspace(0, Lx, 6), np.linspace(0, Ly, 6), np.linspace(-H, 0, 6))
    topo = tp.StackedBrickTopology(1, br, periodicity=(False, False, False), boundary=((1, 1), (1, 1), (2, 3)))
    g = ogrids.Grid(topo[0], 4)
    prob = HomogeneousBox(Lx, Ly, H)
    xi = g.xi[2]
    bcs = (("noslip", "insulating"), ("freeslip", "insulating"), ("kinematic_stress", "insulating"))
    model = oocean.HBModel(prob, bcs=bcs, vert_filter=oocean.cutoff_filter_matrix(xi, 3),
                           exp_filter=oocean.exponential_filter_matrix(xi, 1, 8))
    odgm = odg.DGModel(model, [g], "rusanov")
    oQ = odg.init_ode_state(odgm, lambda x1, x2, x3, a, t: prob.init_state(x1, x2, x3), 0.0)
    dgrid = P.DiscontinuousSpectralElementGrid(
        g.N[0], g.vgeo, g.sgeo, g.vmapM, g.vmapP, g.elemtobndy, g.D[0], g.nreal,
        interiorelems=g.interiorelems, exteriorelems=g.exteriorelems,
        nvertelem=g.topology.stacksize, Imat=g.Imat[2], xi=g.xi[2])
    pbcs = (P.OceanBC(P.Impenetrable(P.NoSlip()), P.Insulating()),
            P.OceanBC(P.Impenetrable(P.FreeSlip()), P.Insulating()),
            P.OceanBC(P.Penetrable(P.KinematicStress()), P.Insulating()))
    m = P.HBModel(P.HomogeneousBox(Lx, Ly, H, boundary_conditions=pbcs), cʰ=float(model.ch))
    aux = P.MPIStateArray(dgrid, 8, data=odgm.state_auxiliary[0].data)
    md = dict(vert_filter=P.CutoffFilter(dgrid, 3), exp_filter=P.ExponentialFilter(dgrid, 1, 8))
    dg = P.DGModel(m, dgrid, P.RusanovNumericalFlux(), P.CentralNumericalFluxSecondOrder(),
                   P.CentralNumericalFluxGradient(), state_auxiliary=aux, modeldata=md)
    dQ = P.MPIStateArray(dgrid, 4, data=oQ[0].data)
    # one evaluation first (tendency parity with these boundary conditions), on a spun-up oracle state
    osol = oode.LSRK144NiegemannDiehlBusch(odgm, oQ, dt=180.0, t0=0.0)
    oode.solve(oQ, osol, numberofsteps=2)
    dQs = P.MPIStateArray(dgrid, 4, data=oQ[0].data)
    odQ = [oQ[0].similar()]
    odgm(odQ, oQ, 0.0, 1, 0)
    dT = P.MPIStateArray(dgrid, 4)
    dT.data.fill_(float("nan"))
    dg(dT, dQs, None, 0.0, 1.0, 0.0)
    res = {"tendency_rel_l2": rel_l2(dT.realdata.cpu().numpy(), odQ[0].realdata)}
    sol = P.LSRK144NiegemannDiehlBusch(dg, dQ, dt=180.0, t0=0.0)
    P.solve(dQ, sol, numberofsteps=20)

    class Wrap:
        def __init__(self, t, nreal):
            self.realdata = t.cpu().numpy()[:nreal]
    stats = {}
    for name, arr in (("Q", dQ), ("aux", dg.state_auxiliary)):
        w = Wrap(arr.data, g.nreal)
        for ivar in range(4):
            stats[(name, ivar)] = oocean.statecheck(w, ivar)
    res["stats"] = stats
    dg.close()
    return res


def ocean_case_float32(nsteps=2, dt=120.0, nelem=(5, 5, 5), spinup=3):
    """Float32 instantiation of the four HBModel kernels: the device runs in Float32 on Float32-rounded
    grid arrays and state; the oracle (Float64 only for the ocean) evaluates the same rounded inputs in
    Float64 -- i.e. it is the 'truth' of a Float32 run, so the differences are the Float32 rounding of
    the kernels themselves."""
    import copy
    P = pkg()
    model, gs, prob = ocean_setup(1, nelem)
    g64 = gs[0]
    # Float32-rounded geometry, seen by the oracle as Float64 values
    g = copy.copy(g64)
    g.vgeo = g64.vgeo.astype(np.float32).astype(np.float64)
    g.sgeo = g64.sgeo.astype(np.float32).astype(np.float64)
    odgm = odg.DGModel(model, [g], "rusanov")
    oQ = odg.init_ode_state(odgm, lambda x1, x2, x3, a, t: prob.init_state(x1, x2, x3), 0.0)
    osol = oode.LSRK144NiegemannDiehlBusch(odgm, oQ, dt=dt, t0=0.0)
    oode.solve(oQ, osol, numberofsteps=spinup)
    oQ[0].data[...] = oQ[0].data.astype(np.float32).astype(np.float64)
    odgm.state_auxiliary[0].data[...] = odgm.state_auxiliary[0].data.astype(np.float32).astype(np.float64)
    g32 = copy.copy(g)
    g32.vgeo, g32.sgeo = g.vgeo.astype(np.float32), g.sgeo.astype(np.float32)
    dg, dgrid = device_ocean_dg(model, g32, odgm)
    import torch
    assert dgrid.FT == torch.float32
    dQ = P.MPIStateArray(dgrid, 4, data=oQ[0].data)
    odQ = [oQ[0].similar()]
    odgm(odQ, oQ, 0.0, 1, 0)
    dT = P.MPIStateArray(dgrid, 4)
    dT.data.fill_(float("nan"))
    dg(dT, dQ, None, 0.0, 1.0, 0.0)
    got = dT.realdata.cpu().numpy().astype(np.float64)
    res = {"tendency_rel_l2": rel_l2(got, odQ[0].realdata),
           "tendency_per_state_rel_l2": [rel_l2(got[:, s], odQ[0].realdata[:, s]) for s in range(4)],
           "gradflux_rel_l2": rel_l2(dg.state_gradient_flux.realdata.cpu().numpy(), odgm.state_gradient_flux[0].realdata),
           "aux_rel_l2": rel_l2(dg.state_auxiliary.realdata[:, 1:4].cpu().numpy(), odgm.state_auxiliary[0].realdata[:, 1:4])}
    osol2 = oode.LSRK144NiegemannDiehlBusch(odgm, oQ, dt=dt, t0=0.0)
    oode.solve(oQ, osol2, numberofsteps=nsteps)
    dsol = P.LSRK144NiegemannDiehlBusch(dg, dQ, dt=dt, t0=0.0)
    P.solve(dQ, dsol, numberofsteps=nsteps)
    gotQ = dQ.realdata.cpu().numpy().astype(np.float64)
    res["state_rel_l2"] = rel_l2(gotQ, oQ[0].realdata)
    res["state_per_field_rel_l2"] = [rel_l2(gotQ[:, s], oQ[0].realdata[:, s]) for s in range(4)]
    res["finite"] = bool(np.isfinite(gotQ).all())
    dg.close()
    return res


def ocean_spindown_on_device(nsteps=720):
    """test/Ocean/HydrostaticBoussinesq/test_3D_spindown.jl (explicit) run entirely through libcmdg: SimpleBox
    spin-down (periodic in x and y, free-slip bottom, penetrable free-slip surface, insulating, c_h = 1, no buoyancy /
    diffusion / rotation), 5 x 5 x 8 elements, `nsteps` LSRK144 steps of 120 s in one cmdg_lsrk_steps call.  Returns the
    per-field statistics, the difference to the C twin of the oracle run on the same arrays, and the error against
    the analytic solution (the reference prints 0.0011289879366523504)."""
    from oracle import cref
    from tests.test_oracle_ocean_spindown import spindown_setup
    P = pkg()
    model, g, prob = spindown_setup()
    odgm = odg.DGModel(model, [g], "rusanov")
    oQ = odg.init_ode_state(odgm, lambda x1, x2, x3, a, t: prob.init_state(x1, x2, x3, 0.0, model), 0.0)
    dgrid = P.DiscontinuousSpectralElementGrid(
        g.N[0], g.vgeo, g.sgeo, g.vmapM, g.vmapP, g.elemtobndy, g.D[0], g.nreal,
        interiorelems=g.interiorelems, exteriorelems=g.exteriorelems,
        nvertelem=g.topology.stacksize, Imat=g.Imat[2], xi=g.xi[2])
    pbcs = (P.OceanBC(P.Impenetrable(P.FreeSlip()), P.Insulating()),
            P.OceanBC(P.Penetrable(P.FreeSlip()), P.Insulating()))
    m = P.HBModel(P.HomogeneousBox(prob.Lx, prob.Ly, prob.H, boundary_conditions=pbcs), cʰ=1.0, αᵀ=0.0, κʰ=0.0,
                  κᶻ=0.0, fₒ=0.0, β=0.0)
    aux = P.MPIStateArray(dgrid, 8, data=odgm.state_auxiliary[0].data)
    md = dict(vert_filter=P.CutoffFilter(dgrid, 3), exp_filter=P.ExponentialFilter(dgrid, 1, 8))
    dg = P.DGModel(m, dgrid, P.RusanovNumericalFlux(), P.CentralNumericalFluxSecondOrder(),
                   P.CentralNumericalFluxGradient(), state_auxiliary=aux, modeldata=md)
    dQ = P.MPIStateArray(dgrid, 4, data=oQ[0].data)
    sol = P.LSRK144NiegemannDiehlBusch(dg, dQ, dt=120.0, t0=0.0)
    P.solve(dQ, sol, numberofsteps=nsteps)
    Qd = dQ.data.cpu().numpy()[:g.nreal]
    auxd = dg.state_auxiliary.data.cpu().numpy()[:g.nreal]
    # the C twin of the oracle on the same arrays
    c = cref.CRefHB.from_grid(model, g, "rusanov")
    cref.use_all_cores_hb()
    cq, caux = oQ[0].data.copy(), odgm.state_auxiliary[0].data.copy()
    osol = oode.LSRK144NiegemannDiehlBusch(odgm, oQ, dt=120.0)
    c.lsrk_steps(cq, np.zeros_like(cq), caux, 120.0, osol.RKA, osol.RKB, nsteps)

    def stats(v):
        v = v.ravel()
        mean = v.mean()
        return v.min(), v.max(), mean, np.sqrt(np.sum((v - mean) ** 2) / (v.size - 1))
    x = [g.vgeo[:, ogrids._x1], g.vgeo[:, ogrids._x2], g.vgeo[:, ogrids._x3]]
    Qe = np.moveaxis(prob.init_state(x[0], x[1], x[2], 120.0 * nsteps, model), 0, 1)
    M = g.vgeo[:, ogrids._M][:, None, :]
    res = {"stats": {("Q", 0): stats(Qd[:, 0]), ("Q", 2): stats(Qd[:, 2]), ("aux", 1): stats(auxd[:, 1]),
                     ("aux", 3): stats(auxd[:, 3])},
           "state_vs_twin_rel_l2": rel_l2(Qd[:, [0, 2]], cq[:, [0, 2]]),
           "aux_vs_twin_rel_l2": rel_l2(auxd[:, [1, 3]], caux[:, [1, 3]]),
           "u2_max": float(np.abs(Qd[:, 1]).max()), "theta_max": float(np.abs(Qd[:, 3]).max()),
           "error_vs_exact": float(np.sqrt(np.sum(M * (Qd - Qe) ** 2)) / np.sqrt(np.sum(M * Qe ** 2))),
           "launches": dg.kernel_launches()}
    dg.close()
    return res
