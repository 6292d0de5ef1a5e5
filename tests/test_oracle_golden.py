"""Pins the oracle's DG tendency + LSRK54 + dry-air thermodynamics against the
reference's own end-to-end golden numbers.

``/root/reference/test/Numerics/DGMethods/Euler/isentropicvortex.jl:60-238`` tabulates
the mass-weighted L2 error of the isentropic vortex after ``timeend = 2L/10/150`` for
each numerical flux and refinement level, and checks them with
``rtol = sqrt(eps(FT))`` (``:278``).  The same run through the oracle must land on
the same numbers (we demand 1e-9, far tighter than the reference's own gate).
Also checks that an emulated 3-rank run (Hilbert partition + halo exchange) gives
the same answer as the single-rank run.
"""
import numpy as np
import pytest

from oracle import topologies as tp, grids, atmos, dgmodel, odesolvers, mpistatearrays as msa

# expected_error[Float64, 3, flux, level]
GOLDEN_F64_3D = {
    ("rusanov", 1): 3.7918869862613858e+00, ("rusanov", 2): 6.5816485664822677e-01,
    ("central", 1): 6.5903683487905749e+00, ("central", 2): 9.2513872939749997e-01,
    ("roe", 1): 4.0766143963611068e+00, ("roe", 2): 4.3942394181655547e-01,
}
# expected_error[Float32, 3, flux, 1]  (isentropicvortex.jl:150-238)
GOLDEN_F32_3D = {("rusanov", 1): 3.7918186187744141e+00, ("central", 1): 6.5903329849243164e+00,
                 ("roe", 1): 4.0765657424926758e+00}


def vortex_error(level, nf, FT=np.float64, csize=1):
    """test_run of isentropicvortex.jl:306-442 (dims = 3)."""
    ps = atmos.Params(FT)
    setup = atmos.IsentropicVortexSetup(ps, FT)
    L = setup.domain_halflength
    ne = 2 ** (level - 1) * 5

    def rng(n):  # range(-L; length = n + 1, stop = L)
        return np.linspace(-L, L, n + 1).astype(FT)

    br = (rng(ne), rng(ne), rng(1))
    topos = tp.BrickTopology(csize, br, periodicity=(True, True, True))
    gs = [grids.Grid(t, 4, FT=FT) for t in topos]
    model = atmos.DryAtmosModel(FT, orientation="none", ref_state=None,
                                turbulence=("constant_dynamic", 0.0, False), sources=())
    dg = dgmodel.DGModel(model, gs, nf)
    timeend = FT(2 * L / 10 / setup.translation_speed)
    elementsize = min(float(np.min(np.diff(b))) for b in br)
    dt = elementsize / atmos.soundspeed_air(ps, setup.T_inf) / 4 ** 2
    nsteps = int(np.ceil(timeend / dt))
    dt = timeend / nsteps

    def init(x1, x2, x3, a, t):
        return setup(x1, x2, x3, FT(t))

    Q = dgmodel.init_ode_state(dg, init, 0)
    lsrk = odesolvers.LSRK54CarpenterKennedy(dg, Q, dt=dt, t0=0)
    odesolvers.solve(Q, lsrk, timeend=timeend)
    Qe = dgmodel.init_ode_state(dg, init, timeend)
    return msa.euclidean_distance(Q, Qe)


@pytest.mark.parametrize("nf", ["rusanov", "central", "roe"])
def test_isentropic_vortex_level1_golden(nf):
    err = vortex_error(1, nf)
    assert err == pytest.approx(GOLDEN_F64_3D[(nf, 1)], rel=1e-9)


def test_isentropic_vortex_level2_golden():
    err = vortex_error(2, "rusanov")
    assert err == pytest.approx(GOLDEN_F64_3D[("rusanov", 2)], rel=1e-9)


def test_isentropic_vortex_three_ranks_matches_golden():
    err = vortex_error(1, "rusanov", csize=3)
    assert err == pytest.approx(GOLDEN_F64_3D[("rusanov", 1)], rel=1e-9)


def test_lsrk54_order_of_accuracy():
    """ode_tests_convergence.jl:15-45: LSRK54 converges at 4th order on dq/dt = a q."""
    class S:
        def __init__(self):
            self.data = np.ones((1, 1, 1))
            self.nreal = 1

        @property
        def realdata(self):
            return self.data

        def similar(self):
            s = S()
            s.data = np.zeros_like(self.data)
            return s

    a = -1.3

    def rhs(dQ, Q, t, increment=False):
        for dq, q in zip(dQ, Q):
            dq.data[...] = a * q.data + (dq.data if increment else 0)

    errs = []
    for n in (8, 16, 32):
        q = [S()]
        sol = odesolvers.LSRK54CarpenterKennedy(rhs, q, dt=1.0 / n, t0=0.0)
        odesolvers.solve(q, sol, timeend=1.0)
        errs.append(abs(q[0].data[0, 0, 0] - np.exp(a)))
    rates = np.log2(np.array(errs[:-1]) / np.array(errs[1:]))
    assert np.all(np.abs(rates - 4) < 0.3)


@pytest.mark.parametrize("method", ["LSRK54CarpenterKennedy", "LSRK144NiegemannDiehlBusch"])
def test_lsrk_convergence_reference_problem(method):
    """ode_tests_convergence.jl:15-45 as written there: dq/dt = q cos(t) (time-dependent, so the RKC
    abscissae matter), exact q0 exp(sin t), final time 20, dt = 2^-6, 2^-7; the observed rate must be
    within 0.17 of the expected order 4 (`explicit_methods`, ode_tests_common.jl:4-5)."""
    class S:
        def __init__(self, v=1.0):
            self.data = np.full((1, 1, 1), v)
            self.nreal = 1

        @property
        def realdata(self):
            return self.data

        def similar(self):
            return S(0.0)

    def rhs(dQ, Q, t, increment=False):
        for dq, q in zip(dQ, Q):
            dq.data[...] = q.data * np.cos(t) + (dq.data if increment else 0)

    errs = []
    for k in (6, 7):
        q = [S()]
        sol = getattr(odesolvers, method)(rhs, q, dt=2.0 ** -k, t0=0.0)
        odesolvers.solve(q, sol, timeend=20.0)
        errs.append(abs(q[0].data[0, 0, 0] - np.exp(np.sin(20.0))))
    rate = np.log2(errs[0] / errs[1])
    assert abs(rate - 4) <= 0.17, (errs, rate)


@pytest.mark.parametrize("nf", ["rusanov", "central", "roe"])
def test_isentropic_vortex_float32_level1_golden(nf):
    """The Float32 rows of the same table (isentropicvortex.jl:195-206): the oracle run entirely in
    Float32 lands within the test's own gate, rtol = sqrt(eps(Float32)) = 3.5e-4."""
    err = float(vortex_error(1, nf, FT=np.float32))
    assert err == pytest.approx(GOLDEN_F32_3D[(nf, 1)], rel=float(np.sqrt(np.finfo(np.float32).eps)))
