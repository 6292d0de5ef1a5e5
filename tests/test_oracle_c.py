"""The C twin of the oracle (oracle/c/dg_ref.c, the CPU baseline that bench.py times) against
the NumPy oracle, which is itself pinned on the reference's golden numbers."""
import numpy as np
import pytest

import __graft_entry__ as ge
from oracle import cref, dgmodel as odg, atmos as oatmos, odesolvers as oode, mpistatearrays as omsa
from tests import parity


@pytest.fixture(scope="module", autouse=True)
def built():
    ge.build()


def _run(model, g, Q0, nf):
    dgm = odg.DGModel(model, [g], nf, skip_zero_viscosity=True)
    q = omsa.MPIStateArray.from_grid(g, 5)
    np.moveaxis(q.data[:g.nreal], 1, 0)[...] = Q0
    dq = q.similar()
    dgm([dq], [q], 0.0, 1, 0)
    c = cref.CRefDG(model, g, nf)
    aux = dgm.state_auxiliary[0].data.copy()
    cq, cdq = q.data.copy(), np.full_like(q.data, np.nan)
    c.tendency(cdq, cq, aux, 1.0, 0.0)
    assert parity.rel_l2(cdq[:g.nreal], dq.realdata) < 1e-13
    assert parity.rel_l2(aux[:g.nreal], dgm.state_auxiliary[0].realdata) < 1e-14
    # two LSRK54 steps
    sol = oode.LSRK54CarpenterKennedy(dgm, [q], dt=1e-5 if model.orientation == "none" else 0.5)
    oode.solve([q], sol, numberofsteps=2)
    cdq[...] = 0
    c.lsrk_steps(cq, cdq, aux, float(sol.dt), sol.RKA, sol.RKB, 2)
    assert parity.rel_l2(cq[:g.nreal], q.realdata) < 1e-13


@pytest.mark.parametrize("nf", ["rusanov", "central"])
def test_c_twin_vortex(nf):
    model, gs, setup, dt = parity.vortex_setup((3, 3, 2))
    g = gs[0]
    from oracle import grids as ogrids
    Q0 = setup(g.vgeo[:g.nreal, ogrids._x1], g.vgeo[:g.nreal, ogrids._x2],
               g.vgeo[:g.nreal, ogrids._x3], np.float64(0))
    _run(model, g, Q0, nf)


def test_c_twin_baroclinic_wave():
    model, gs = parity.gcm_setup(3, 2)
    g = gs[0]
    dgm = odg.DGModel(model, [g], "rusanov")
    aux = np.moveaxis(dgm.state_auxiliary[0].data[:g.nreal], 1, 0)
    _run(model, g, oatmos.init_baroclinic_wave(model, aux), "rusanov")


@pytest.mark.parametrize("kind", ["held_suarez", "smagorinsky_box_every", "constant_dynamic_box"])
def test_c_twin_second_order_path(kind):
    """Gradient pass + viscous fluxes (+ HeldSuarezForcing / RayleighSponge) of the C twin against the
    NumPy oracle: BASELINE.json configs[3] physics (Smagorinsky on the sphere, horizontal diffusion
    direction), the LES box (every direction, no-slip / free-slip walls) and ConstantDynamicViscosity
    with divergence.  This is what gives config (4) a CPU baseline in bench.py."""
    if kind == "held_suarez":
        model, gs = parity.gcm_setup(3, 3, turbulence=("smagorinsky", 0.21))
        model.sources = ("gravity", "coriolis", "held_suarez",
                         ("rayleigh_sponge", 30e3, 12e3, 1 / 60 / 15, (0.0, 0.0, 0.0), 2.0))
        dd, dt = "horizontal", 0.5
    else:
        turb = ("smagorinsky", 0.21) if kind.startswith("smag") else ("constant_dynamic", 50.0, True)
        model, gs = parity.box_setup((3, 2, 3), turbulence=turb)
        dd, dt = "every", 0.01
    g = gs[0]
    dgm = odg.DGModel(model, [g], "rusanov", diffusion_direction=dd)
    aux0 = np.moveaxis(dgm.state_auxiliary[0].data[:g.nreal], 1, 0)
    Q0 = oatmos.init_baroclinic_wave(model, aux0) if kind == "held_suarez" else parity.bubble_state(model, g, aux0)
    q = omsa.MPIStateArray.from_grid(g, 5)
    np.moveaxis(q.data[:g.nreal], 1, 0)[...] = Q0
    dq = q.similar()
    dgm([dq], [q], 0.0, 1, 0)
    c = cref.CRefDG(model, g, "rusanov", second_order=True, diffusion_direction=dd)
    aux = dgm.state_auxiliary[0].data.copy()
    cq, cdq = q.data.copy(), np.full_like(q.data, np.nan)
    c.tendency(cdq, cq, aux, 1.0, 0.0)
    assert parity.rel_l2(c.gradflux[:g.nreal], dgm.state_gradient_flux[0].realdata) < 1e-13
    assert parity.rel_l2(cdq[:g.nreal], dq.realdata) < 1e-13
    sol = oode.LSRK54CarpenterKennedy(dgm, [q], dt=dt)
    oode.solve([q], sol, numberofsteps=2)
    cdq[...] = 0
    c.lsrk_steps(cq, cdq, aux, float(sol.dt), sol.RKA, sol.RKB, 2)
    assert parity.rel_l2(cq[:g.nreal], q.realdata) < 1e-13


@pytest.mark.parametrize("workload", ["baroclinic_wave", "held_suarez"])
def test_c_twin_on_package_arrays(workload):
    """bench.py's CPU arm and full-size parity check drive the C twin through `ref_params_for` (package
    balance law -> ref_params) over the package's own host-built arrays: same tendency as the NumPy oracle
    on the oracle's grid (the two grid builders agree to rounding, tests/test_host_mesh.py)."""
    import torch
    import bench
    P = ge.load_package()
    from climatemachine_jl_b200 import atmos_init as ai
    ne, nvert = 3, 2
    grid, _ = bench.build_grid(P, workload, ne, nvert, 0, 1, "cpu")
    model = bench.gcm_model(P, workload)
    aux0 = ai.init_state_auxiliary(model, grid)
    Q0 = ai.baroclinic_wave(model, grid, aux0)
    skip = workload != "held_suarez"
    R = bench.ref_params_for(P, model, skip)
    npy = lambda t: np.ascontiguousarray(t.numpy())
    c = cref.CRefDG.from_arrays(R, npy(grid.vgeo), npy(grid.sgeo), npy(grid.vmapM), npy(grid.vmapP),
                                npy(grid.elemtobndy), grid.D_host, grid.nrealelem)
    Q = npy(Q0)
    dQ = np.full_like(Q, np.nan)
    c.tendency(dQ, Q, npy(aux0.data), 1.0, 0.0)
    # oracle on its own grid
    turb = ("smagorinsky", 0.21) if workload == "held_suarez" else ("constant_kinematic", 0.0, False)
    omodel, gs = parity.gcm_setup(ne, nvert, turbulence=turb)
    if workload == "held_suarez":
        omodel.sources = ("gravity", "coriolis", "held_suarez",
                          ("rayleigh_sponge", 30e3, 12e3, 1 / 60 / 15, (0.0, 0.0, 0.0), 2.0))
    g = gs[0]
    dgm = odg.DGModel(omodel, [g], "rusanov", diffusion_direction="horizontal", skip_zero_viscosity=skip)
    oaux = np.moveaxis(dgm.state_auxiliary[0].data[:g.nreal], 1, 0)
    q = omsa.MPIStateArray.from_grid(g, 5)
    np.moveaxis(q.data[:g.nreal], 1, 0)[...] = oatmos.init_baroclinic_wave(omodel, oaux)
    dq = q.similar()
    dgm([dq], [q], 0.0, 1, 0)
    assert parity.rel_l2(Q, q.realdata) < 1e-12
    assert parity.rel_l2(dQ, dq.realdata) < 1e-9


def test_extended_precision_twin_and_conditioning():
    """oracle/c/libdgref_ld.so = the same C code in x87 extended precision.  (i) It is the same
    algorithm: on a state pushed 1 % off balance its result equals the Float64 twin's to 1e-14.
    (ii) The balanced baroclinic-wave state is ill-conditioned in *relative* L2: the reference's own
    Float64 arithmetic is ~1e-12 away from the exactly rounded tendency at ne = 6 x 2 (and further on
    finer meshes), which is why the 1e-12 bar is applied to the unbalanced state and the balanced one
    is judged against this yardstick (tests/bench_checks.py)."""
    from tests import bench_checks

    def rl(a, b):
        a, b = np.asarray(a, dtype=np.longdouble), np.asarray(b, dtype=np.longdouble)
        return float(np.sqrt(np.sum((a - b) ** 2) / np.sum(b ** 2)))
    model, gs = parity.gcm_setup(6, 2)
    g = gs[0]
    dgm = odg.DGModel(model, [g], "rusanov", skip_zero_viscosity=True)
    aux = dgm.state_auxiliary[0].data
    A = np.moveaxis(aux[:g.nreal], 1, 0)
    Q0 = oatmos.init_baroclinic_wave(model, A)
    c = cref.CRefDG(model, g, "rusanov")
    res = {}
    for name in ("balanced", "unbalanced"):
        q0 = Q0.copy()
        if name == "unbalanced":
            pert, du, dw = bench_checks.unbalance(q0, A[0:3], np)
            q0 = q0 * pert
            q0[1] += du * q0[0]
            q0[3] += dw * q0[0]
        q = np.zeros((g.nelem, 5, 125))
        q[:g.nreal] = np.moveaxis(q0, 0, 1)
        d = np.full_like(q, np.nan)
        c.tendency(d, q.copy(), aux.copy())
        t = c.tendency_extended(q, aux.copy())
        res[name] = rl(d[:g.nreal], t[:g.nreal])
    assert res["unbalanced"] < 1e-14, res
    assert 1e-13 < res["balanced"] < 1e-11, res


def test_c_twin_ocean_hbmodel():
    """The ocean twin (oracle/c/hb_ref.c: vertical filters, gradient pass with the convective-adjustment switch,
    stack integrals, Rusanov + central second-order fluxes, flux-based ocean boundary conditions, LSRK144) against
    the NumPy oracle, which reproduces the reference's ocean-gyre regression values (tests/test_oracle_ocean.py).
    This is what gives BASELINE.json configs[4] a CPU baseline in bench.py."""
    model, gs, prob = parity.ocean_setup(1, (4, 3, 3))
    g = gs[0]
    dgm = odg.DGModel(model, [g], "rusanov")
    q = odg.init_ode_state(dgm, lambda x1, x2, x3, a, t: prob.init_state(x1, x2, x3), 0.0)
    sol = oode.LSRK144NiegemannDiehlBusch(dgm, q, dt=120.0)
    oode.solve(q, sol, numberofsteps=2)           # spin-up: w, pkin, wind stress, convective adjustment active
    c = cref.CRefHB.from_grid(model, g, "rusanov")
    cq, caux = q[0].data.copy(), dgm.state_auxiliary[0].data.copy()
    cdq = np.full_like(cq, np.nan)
    dq = [q[0].similar()]
    dgm(dq, q, 0.0, 1, 0)
    c.tendency(cdq, cq, caux, 1.0, 0.0)
    assert parity.rel_l2(cq, q[0].data) < 1e-14                       # filtered in place
    assert parity.rel_l2(c.gradflux, dgm.state_gradient_flux[0].data) < 1e-13
    assert parity.rel_l2(caux[:, 1:4], dgm.state_auxiliary[0].data[:, 1:4]) < 1e-13
    assert parity.rel_l2(cdq, dq[0].data) < 1e-13
    # increment form, then two LSRK144 steps from the same state
    dgm(dq, q, 0.0, 0.5, 2.0)
    c.tendency(cdq, cq, caux, 0.5, 2.0)
    assert parity.rel_l2(cdq, dq[0].data) < 1e-13
    sol2 = oode.LSRK144NiegemannDiehlBusch(dgm, q, dt=120.0)
    oode.solve(q, sol2, numberofsteps=2)
    cdq[...] = 0
    c.lsrk_steps(cq, cdq, caux, 120.0, sol2.RKA, sol2.RKB, 2)
    assert parity.rel_l2(cq, q[0].data) < 1e-13


def test_bench_ocean_twin_runs_on_package_arrays():
    """bench.py's ocean CPU baseline: the C twin over the PACKAGE's host-built ocean mesh, initial state, filter
    matrices and stack-integral operator (the arrays the GPU arm hands to libcmdg), a small box, two LSRK144 steps;
    and the same state stepped by the NumPy oracle on the oracle's own mesh of that box."""
    import bench
    P = ge.load_package()
    c, Q, aux, dt, (rka, rkb), nreal = bench.ocean_twin(P, 3, 3)
    assert nreal == 27 and Q.shape == (27, 4, 125)
    Q0 = Q.copy()
    dQ = np.zeros_like(Q)
    c.lsrk_steps(Q, dQ, aux, dt, rka, rkb, 2)
    assert np.isfinite(Q).all() and np.isfinite(aux).all()
    assert np.abs(Q - Q0).max() > 0
    # the NumPy oracle on its own mesh of the same box (bench.py: Lx = Ly = 4e6 m, H = 1000 m at N = 1)
    from oracle import ocean as oocean, topologies as otp, grids as ogrids
    Lx, Ly, H = 4e6, 4e6, 1000.0
    br = (np.linspace(0, Lx, 4), np.linspace(0, Ly, 4), np.linspace(-H, 0, 4))
    topos = otp.StackedBrickTopology(1, br, periodicity=(False, False, False), boundary=((1, 1), (1, 1), (2, 3)))
    g = ogrids.Grid(topos[0], 4)
    prob = oocean.OceanGyre(Lx, Ly, H)
    xi = g.xi[2]
    model = oocean.HBModel(prob, vert_filter=oocean.cutoff_filter_matrix(xi, 3),
                           exp_filter=oocean.exponential_filter_matrix(xi, 1, 8))
    dgm = odg.DGModel(model, [g], "rusanov")
    q = odg.init_ode_state(dgm, lambda x1, x2, x3, a, t: prob.init_state(x1, x2, x3), 0.0)
    sol = oode.LSRK144NiegemannDiehlBusch(dgm, q, dt=dt)
    oode.solve(q, sol, numberofsteps=2)
    assert parity.rel_l2(Q, q[0].data) < 1e-12


@pytest.mark.parametrize("nf,level,golden", [
    ("rusanov", 3, 2.0160333422867591e-02), ("rusanov", 4, 6.6360317881818034e-04),
    ("central", 3, 1.1680141169828175e-01), ("central", 4, 2.6414127301659534e-03)])
def test_c_twin_isentropic_vortex_golden_levels_3_4(nf, level, golden):
    """The finer rows of the reference's isentropic-vortex error table (test/Numerics/DGMethods/Euler/
    isentropicvortex.jl:103-111: 3-D Float64, levels 3 and 4 = 20 x 20 x 1 and 40 x 40 x 1 elements) reproduced by the
    C twin -- the NumPy oracle pins levels 1 and 2 (tests/test_oracle_golden.py) and is too slow beyond.  Same recipe
    as `test_run` (:306-442): dt = dx_min / c(300 K) / N^2 snapped to hit timeend = 2 L / 10 / 150, mass-weighted L2
    distance to the translated vortex; the reference's own gate is rtol = sqrt(eps)."""
    from oracle import topologies as otp, grids as ogrids, atmos as oat, mpistatearrays as msa
    FT = np.float64
    ps = oat.Params(FT)
    setup = oat.IsentropicVortexSetup(ps, FT)
    L = setup.domain_halflength
    ne = 2 ** (level - 1) * 5
    br = (np.linspace(-L, L, ne + 1), np.linspace(-L, L, ne + 1), np.linspace(-L, L, 2))
    g = ogrids.Grid(otp.BrickTopology(1, br, periodicity=(True, True, True))[0], 4, FT=FT)
    model = oat.DryAtmosModel(FT, orientation="none", ref_state=None, turbulence=("constant_dynamic", 0.0, False),
                              sources=())
    dgm = odg.DGModel(model, [g], nf, skip_zero_viscosity=True)
    timeend = FT(2 * L / 10 / setup.translation_speed)
    dt = min(float(np.min(np.diff(b))) for b in br) / oat.soundspeed_air(ps, setup.T_inf) / 4 ** 2
    nsteps = int(np.ceil(timeend / dt))
    dt = timeend / nsteps
    init = lambda x1, x2, x3, a, t: setup(x1, x2, x3, FT(t))
    q = odg.init_ode_state(dgm, init, 0)
    sol = oode.LSRK54CarpenterKennedy(dgm, q, dt=dt, t0=0)
    c = cref.CRefDG(model, g, nf)
    cref.use_all_cores()
    cq = q[0].data.copy()
    c.lsrk_steps(cq, np.zeros_like(cq), dgm.state_auxiliary[0].data.copy(), float(dt), sol.RKA, sol.RKB, nsteps)
    qe = odg.init_ode_state(dgm, init, timeend)
    q[0].data[...] = cq
    err = msa.euclidean_distance(q, qe)
    assert err == pytest.approx(golden, rel=1e-8), (err, golden)
