"""Pins the oracle's HBModel (fluxes, ocean boundary conditions incl. flux-based second-order
BCs, in-tendency vertical filters, column integrals, LSRK144) on the reference's own regression
values: ``test/Ocean/refvals/test_ocean_gyre_refvals.jl`` (``short``), produced by
``test/Ocean/HydrostaticBoussinesq/test_ocean_gyre_short.jl`` through
``experiments/OceanBoxGCM/simple_box.jl`` (5x5x5 elements, N = 4, dt = 120 s, 1 h)."""
import numpy as np
import pytest

from oracle import topologies as tp, grids, ocean, dgmodel, odesolvers, mpistatearrays as msa

# [field, min, max, mean, std] and the number of digits the reference test compares
REF_SHORT = {
    ("Q", 0): (-1.56752732465427965791e-02, 1.68514505893861757380e-02, -2.29247099512640793717e-03, 3.25160701902235671490e-03),
    ("Q", 1): (-3.79775189267773510826e-02, 6.43003815969073189152e-03, -1.34097283250433057383e-02, 1.20337368194613283934e-02),
    ("Q", 2): (-1.53249855976142324021e-01, 1.57769367725846931805e-01, -6.47997524778634250456e-06, 1.03985821375781786746e-01),
    ("Q", 3): (1.07842251824037485168e-05, 9.00370181779731204585e+00, 2.49971752106072075961e+00, 2.19711465681760964586e+00),
    ("aux", 0): (0.0, 1.00000000000000011642e+06, 5.0e+05, 2.92779390974978974555e+05),
    ("aux", 1): (-9.71167446044889304474e-05, 8.57892392958965760040e-05, 7.22305481630802806057e-08, 3.80815409312154189983e-05),
    ("aux", 2): (-8.99913049767501638243e-01, 0.0, -3.32109083459310172604e-01, 2.56226532893150116266e-01),
    ("aux", 3): (-8.17753668537623428815e-05, 8.25631396299614581233e-05, -1.10719884491561329431e-09, 5.14580958965247714965e-05),
}
DIGITS = {("Q", 2): (12, 12, 11, 12), ("aux", 1): (12, 12, 11, 12), ("aux", 3): (12, 11, 10, 12)}


def gyre_setup(csize=1, nelem=(5, 5, 5)):
    Lx, Ly, H = 1e6, 1e6, 1000.0
    br = (np.linspace(0, Lx, nelem[0] + 1), np.linspace(0, Ly, nelem[1] + 1), np.linspace(-H, 0, nelem[2] + 1))
    topos = tp.StackedBrickTopology(csize, br, periodicity=(False, False, False),
                                    boundary=((1, 1), (1, 1), (2, 3)))
    gs = [grids.Grid(t, 4) for t in topos]
    prob = ocean.OceanGyre(Lx, Ly, H)
    xi = gs[0].xi[2]
    model = ocean.HBModel(prob, vert_filter=ocean.cutoff_filter_matrix(xi, 3),
                          exp_filter=ocean.exponential_filter_matrix(xi, 1, 8))
    return model, gs, prob


def close_digits(a, b, digits):
    if b == 0:
        return abs(a) < 10.0 ** (-digits)
    return abs(a - b) <= abs(b) * 10.0 ** (-(digits - 1)) * 5


def test_ocean_gyre_short_refvals():
    model, gs, prob = gyre_setup()
    dg = dgmodel.DGModel(model, gs, "rusanov")

    def init(x1, x2, x3, a, t):
        return prob.init_state(x1, x2, x3)

    Q = dgmodel.init_ode_state(dg, init, 0.0)
    sol = odesolvers.LSRK144NiegemannDiehlBusch(dg, Q, dt=120.0, t0=0.0)
    odesolvers.solve(Q, sol, timeend=3600.0)
    assert sol.steps == 30
    for (name, ivar), ref in REF_SHORT.items():
        arr = Q[0] if name == "Q" else dg.state_auxiliary[0]
        got = ocean.statecheck(arr, ivar)
        digs = DIGITS.get((name, ivar), (12, 12, 12, 12))
        for g, r, d in zip(got, ref, digs):
            # we demand 2 digits fewer than the reference's own same-machine gate
            assert close_digits(g, r, d - 2), (name, ivar, got, ref)


def test_filter_matrices():
    r, _ = __import__("oracle.elements", fromlist=["x"]).lglpoints(np.float64, 4)
    F = ocean.cutoff_filter_matrix(r, 3)
    # polynomials of degree < 3 pass, the degree-4 Legendre mode is removed
    for p in range(3):
        assert np.allclose(F @ r ** p, r ** p, atol=1e-13)
    P4 = np.polynomial.legendre.legval(r, [0, 0, 0, 0, 1])
    assert np.allclose(F @ P4, 0, atol=1e-13)
    E = ocean.exponential_filter_matrix(r, 1, 8)
    assert np.allclose(E @ np.ones(5), 1) and np.allclose(E @ r, r)
    assert np.allclose(E @ P4, np.finfo(float).eps * P4, atol=1e-15)
