"""min_node_distance pinned on the analytic expectation of the reference's
test/Numerics/Mesh/min_node_distance.jl:17-83 (warped stacked brick, N = 4 and mixed orders are
tested there; the path here is N = 4), plus sanity checks of the Courant numbers."""
import numpy as np

from oracle import courant, grids, topologies as tp


def _warp(x1, x2, x3):
    x1 = np.where(x1 >= 0.5, 0.5 + 2 * (x1 - 0.5), x1)
    x2 = np.where(x2 >= 0.5, 0.5 + 2 * (x2 - 0.5), x2)
    x3 = np.where(x3 >= 1.5, 1.5 + 2 * (x3 - 1.5), x3)
    return x1, x2, x3


def test_min_node_distance_reference_expectation():
    Neh, Nev, N = 10, 4, 4
    br = (np.linspace(0, 1, Neh + 1), np.linspace(0, 1, Neh + 1), np.linspace(1, 2, Nev + 1))
    topo = tp.StackedBrickTopology(1, br, periodicity=(False, False, False))[0]
    g = grids.Grid(topo, N, meshwarp=_warp)
    dxi = g.xi[0][1] - g.xi[0][0]
    hmnd, vmnd = dxi / (2 * Neh), dxi / (2 * Nev)
    assert np.isclose(courant.min_node_distance(g, "every"), hmnd, rtol=1e-12)
    assert np.isclose(courant.min_node_distance(g, "vertical"), vmnd, rtol=1e-12)
    assert np.isclose(courant.min_node_distance(g, "horizontal"), hmnd, rtol=1e-12)


def test_courant_numbers_uniform_flow():
    """Uniform state on a uniform box: nondiffusive Courant = dt (|u| + c) / dx_min."""
    from oracle import atmos, dgmodel as odg
    from tests import parity
    model, gs, setup, dt = parity.vortex_setup((2, 2, 2))
    g = gs[0]
    ps = model.ps
    T, p, u = 300.0, 1e5, (30.0, -40.0, 0.0)
    ρ = p / (ps.R_d * T)
    e = 0.5 * (u[0] ** 2 + u[1] ** 2) + ps.cv_d * (T - ps.T_0)
    Q = np.zeros((g.nelem, 5, g.Np))
    Q[:, 0], Q[:, 1], Q[:, 2], Q[:, 4] = ρ, ρ * u[0], ρ * u[1], ρ * e
    aux = odg.DGModel(model, [g], "rusanov").state_auxiliary[0].data
    GF = np.zeros((g.nelem, 9, g.Np))
    dx = courant.min_node_distance(g)
    c = float(atmos.soundspeed_air(ps, np.float64(T)))
    assert np.isclose(courant.courant(model, g, Q, aux, GF, dt, "nondiffusive"), dt * (50.0 + c) / dx, rtol=1e-12)
    assert np.isclose(courant.courant(model, g, Q, aux, GF, dt, "advective"), dt * 50.0 / dx, rtol=1e-12)
    assert courant.courant(model, g, Q, aux, GF, dt, "diffusive") == 0.0


def test_atmos_courant_reference_known_answers():
    """test/Numerics/DGMethods/courant.jl (3-D, Float64): AtmosModel (LES config: FlatOrientation),
    NoReferenceState, ConstantDynamicViscosity(2, WithDivergence()), Gravity; u = (150 x1, 150 x1, 0),
    T = 300 K, p = 1e5 Pa on the 10 x 10 x 4 stacked brick [0,1]^2 x [1,2], dt = 1/200.  Expected:
    nondiffusive horizontal = dt (|(150,150,0)| + c_s) / dx_h and vertical = dt c_s / dx_v (rtol 1e-4),
    diffusive horizontal / vertical = dt (mu / rho) / dx^2 (`isapprox`)."""
    from oracle import atmos, dgmodel as odg
    Neh, Nev = 10, 4
    br = (np.linspace(0, 1, Neh + 1), np.linspace(0, 1, Neh + 1), np.linspace(1, 2, Nev + 1))
    topo = tp.StackedBrickTopology(1, br, periodicity=(False, False, False), connectivity="full")[0]
    g = grids.Grid(topo, 4)
    μ = 2.0
    model = atmos.DryAtmosModel(np.float64, orientation="flat", ref_state=None,
                                turbulence=("constant_dynamic", μ, True), sources=("gravity",),
                                bcs=("freeslip",))
    ps = model.ps
    dgm = odg.DGModel(model, [g], "rusanov")
    aux = dgm.state_auxiliary[0].data
    T, p = 300.0, 1e5
    ρ = float(atmos.air_density(ps, np.float64(T), np.float64(p)))
    x1 = g.vgeo[:, grids._x1]
    Q = np.zeros((g.nelem, 5, g.Np))
    u = 150.0 * x1
    Q[:, 0], Q[:, 1], Q[:, 2] = ρ, ρ * u, ρ * u
    Q[:, 4] = ρ * atmos.total_energy(ps, (u * u + u * u) / 2, np.float64(0), np.float64(T))
    GF = dgm.state_gradient_flux[0].data            # the diffusive number of a constant closure ignores it
    dt = 1 / 200
    dx_h = courant.min_node_distance(g, "horizontal")
    dx_v = courant.min_node_distance(g, "vertical")
    c_s = float(atmos.soundspeed_air(ps, np.float64(T)))
    c_h = dt * (np.linalg.norm([150.0, 150.0, 0.0]) + c_s) / dx_h
    c_v = dt * c_s / dx_v
    d_h = dt * (μ / ρ) / dx_h ** 2
    d_v = dt * (μ / ρ) / dx_v ** 2
    assert np.isclose(courant.courant(model, g, Q, aux, GF, dt, "nondiffusive", "horizontal"), c_h, rtol=1e-4)
    assert np.isclose(courant.courant(model, g, Q, aux, GF, dt, "nondiffusive", "vertical"), c_v, rtol=1e-4)
    assert np.isclose(courant.courant(model, g, Q, aux, GF, dt, "diffusive", "horizontal"), d_h, rtol=1.5e-8)
    assert np.isclose(courant.courant(model, g, Q, aux, GF, dt, "diffusive", "vertical"), d_v, rtol=1.5e-8)
