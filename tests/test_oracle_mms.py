"""Pins the oracle's second-order (Navier-Stokes) path on the reference's own golden value:
``test/Numerics/DGMethods/compressible_Navier_Stokes/mms_bc_atmos.jl`` (3-D, level 1) -- AtmosModel,
DryModel, NoOrientation, NoReferenceState, ``ConstantDynamicViscosity(1/100, WithDivergence())``,
``InitStateBC`` on every face, a manufactured source, warped 4 x 4 x 4 brick of order 4, Rusanov /
CentralNumericalFluxSecondOrder / CentralNumericalFluxGradient, LSRK54 to t = 1 with dt = 1/800.
Expected ``euclidean_distance(Q, Q_exact(1)) = 3.3983777728925593e-02`` (``expected_result[2, 1]``,
checked there with ``isapprox``, i.e. rtol = sqrt(eps)).

The manufactured source is not copied from ``mms_solution_generated.jl``: it is derived here with
sympy from the manufactured solution and the compressible Navier-Stokes fluxes (the derivation the
reference documents in ``mms_solution.jl``).  The test's two overrides are reproduced: ``T_0 = 0`` and
``total_specific_enthalpy = 0`` (no enthalpy diffusion).

What this pins: volume/interface gradient kernels, the gradient-flux map, the viscous stress with
the divergence term, the central second-order flux and its boundary variant, the gradient boundary
state, curved-element metrics, LSRK54 with a time-dependent right-hand side.  It runs the 4000
tendency evaluations of the reference run, so it is the slowest CPU test (about three minutes).
"""
import numpy as np
import pytest
import sympy as sp

from oracle import atmos as oatmos, dgmodel as odg, grids as G, topologies as tp
from oracle import odesolvers as oode, mpistatearrays as msa

EXPECTED_3D_LEVEL1 = 3.3983777728925593e-02    # mms_bc_atmos.jl:214-217


class TimePolynomial:
    """A vector field sum_ij cos(pi t)^i sin(pi t)^j a_ij(x, y, z): the spatial coefficients are
    evaluated once per set of nodes (cached on the coordinate bytes), a time level costs a few axpys."""

    def __init__(self, exprs, t, X):
        c, s = sp.symbols("c s", real=True)
        self.terms = []
        for e in exprs:
            e = sp.expand(e.subs({sp.cos(sp.pi * t): c, sp.sin(sp.pi * t): s}))
            assert not e.has(t)
            self.terms.append([(i, j, sp.lambdify(X, co, "numpy", cse=True))
                               for (i, j), co in sp.Poly(e, c, s).terms()])
        self.cache = {}

    def __call__(self, tt, xx, yy, zz):
        key = (xx.shape, hash(xx.tobytes()), hash(yy.tobytes()), hash(zz.tobytes()))
        if key not in self.cache:
            self.cache[key] = [[(i, j, np.broadcast_to(np.asarray(f(xx, yy, zz), dtype=np.float64), xx.shape).copy())
                                for i, j, f in comp] for comp in self.terms]
        c, s = np.cos(np.pi * tt), np.sin(np.pi * tt)
        return np.stack([sum(c ** i * s ** j * a for i, j, a in comp) for comp in self.cache[key]])


def manufactured(gamma, mu):
    """(Q_exact, S) as callables of (t, x, y, z): the 3-D branch of mms_solution.jl:17-23 and
    S = dQ/dt + div(F1 + F2) for the compressible Navier-Stokes equations with constant dynamic
    viscosity and no heat conduction."""
    x, y, z, t = sp.symbols("x y z t", real=True)
    pi = sp.pi
    rho = sp.cos(pi * t) * sp.sin(pi * x) * sp.cos(pi * y) * sp.cos(pi * z) + 3
    U = sp.cos(pi * t) * rho * sp.sin(pi * x) * sp.cos(pi * y) * sp.cos(pi * z)
    V = sp.cos(pi * t) * rho * sp.sin(pi * x) * sp.cos(pi * y) * sp.cos(pi * z)
    W = sp.cos(pi * t) * rho * sp.sin(pi * x) * sp.cos(pi * y) * sp.sin(pi * z)
    E = sp.cos(pi * t) * sp.sin(pi * x) * sp.cos(pi * y) * sp.cos(pi * z) + 100
    vel = [sp.cancel(U / rho), sp.cancel(V / rho), sp.cancel(W / rho)]       # polynomials in cos(pi t)
    P = (gamma - 1) * (E - (U * vel[0] + V * vel[1] + W * vel[2]) / 2)
    X = [x, y, z]
    grad = [[sp.diff(vel[c], X[d]) for d in range(3)] for c in range(3)]     # grad[c][d] = d u_c / d x_d
    div = grad[0][0] + grad[1][1] + grad[2][2]
    tau = [[mu * (grad[c][d] + grad[d][c]) - (sp.Rational(2, 3) * mu * div if c == d else 0)
            for d in range(3)] for c in range(3)]
    mom = [U, V, W]
    Q = [rho, U, V, W, E]
    S = [sp.diff(q, t) for q in Q]
    for d in range(3):
        F = [mom[d]]
        for c in range(3):
            F.append(vel[d] * mom[c] + (P if c == d else 0) - tau[c][d])
        F.append(vel[d] * (E + P) - sum(vel[c] * tau[c][d] for c in range(3)))
        for s in range(5):
            S[s] = S[s] + sp.diff(F[s], X[d])
    return TimePolynomial(Q, t, X), TimePolynomial(S, t, X)


class MMSAtmosModel(oatmos.DryAtmosModel):
    """The AtmosModel of mms_bc_atmos.jl:121-139 with its problem-specific hooks."""

    def __init__(self, exact, src):
        ps = oatmos.Params(np.float64)
        ps.T_0 = np.float64(0)                  # CLIMAParameters.Planet.T_0(::EarthParameterSet) = 0 (:34)
        super().__init__(np.float64, orientation="none", ref_state=None,
                         turbulence=("constant_dynamic", 1 / 100, True), sources=(), bcs=("initstate",),
                         params=ps)
        self.exact, self.src, self.t = exact, src, 0.0

    def gradient_argument(self, Q, aux):
        G = super().gradient_argument(Q, aux)
        G[3] = 0                                # total_specific_enthalpy(::PhaseDry, e_tot) = 0 (:48-49)
        return G

    def source(self, Q, aux):
        return self.src(self.t, aux[0], aux[1], aux[2])

    # InitStateBC (src/Atmos/Model/bc_initstate.jl:12-45): the plus state is the exact solution at the
    # face node, for the first-order, gradient and second-order fluxes alike
    def boundary_state(self, kind, bctag, n, Qm, auxm):
        Qp = self.exact(self.t, auxm[0], auxm[1], auxm[2])
        auxp = auxm.copy()
        self.nodal_update_aux(Qp, auxp)
        return Qp, auxp

    # boundary_flux_second_order! (NumericalFluxes.jl:920-967): flux_second_order!(state+, diff+ = diff-, aux+)
    def boundary_flux_second_order(self, bctag, n, Qm, gm, auxm):
        Qp, auxp = self.boundary_state("second", bctag, n, Qm, auxm)
        return self.flux_second_order(Qp, gm, auxp)


def warp3d(x1, x2, x3):
    """mms_bc_atmos.jl:262-271."""
    return (x1 + (x1 - 1 / 2) * np.cos(2 * np.pi * x2 * x3) / 4,
            x2 + np.exp(np.sin(2 * np.pi * (x1 * x2 + x3))) / 20,
            x3 + x1 / 4 + x2 ** 2 / 2 + np.sin(x1 * x2 * x3))


def _setup(ne=4):
    ps = oatmos.Params(np.float64)
    gamma = sp.Rational(7, 5)
    assert abs(float(ps.cp_d / ps.cv_d) - 1.4) < 1e-14      # the generated source uses gamma = 1.4
    exact, src = manufactured(gamma, sp.Rational(1, 100))
    br = tuple(np.linspace(0.0, 1.0, ne + 1) for _ in range(3))
    topo = tp.BrickTopology(1, br, periodicity=(False, False, False))[0]
    g = G.Grid(topo, 4, meshwarp=warp3d)
    model = MMSAtmosModel(exact, src)
    dgm = odg.DGModel(model, [g], "rusanov")
    return g, model, dgm, exact


def _state(g, exact, t):
    q = msa.MPIStateArray.from_grid(g, 5)
    vg = g.vgeo[:g.nreal]
    np.moveaxis(q.data[:g.nreal], 1, 0)[...] = exact(t, vg[:, G._x1], vg[:, G._x2], vg[:, G._x3])
    return q


def _weighted_distance(g, a, b):
    M = g.vgeo[:g.nreal, G._M][:, None, :]
    return float(np.sqrt(np.sum(M * (a.data[:g.nreal] - b.data[:g.nreal]) ** 2)))


def test_manufactured_source_balances_the_discrete_operator():
    """Consistency of the derived source: the DG tendency of the exact solution (fluxes + source) is
    the time derivative of the exact solution up to the discretisation error, which shrinks with
    resolution."""
    errs = []
    for ne in (2, 4):
        g, model, dgm, exact = _setup(ne)
        t = 0.3
        Q = _state(g, exact, t)
        dQ = Q.similar()
        dgm([dQ], [Q], t, 1, 0)
        h = 1e-6
        dQdt = (_state(g, exact, t + h).data - _state(g, exact, t - h).data) / (2 * h)
        M = g.vgeo[:g.nreal, G._M][:, None, :]
        errs.append(float(np.sqrt(np.sum(M * (dQ.data[:g.nreal] - dQdt[:g.nreal]) ** 2))))
    assert errs[1] < errs[0] / 4, errs


@pytest.mark.timeout(900)
def test_mms_bc_atmos_3d_level1_golden_error():
    g, model, dgm, exact = _setup(4)
    Q = _state(g, exact, 0.0)
    timeend = 1.0
    dt = 5e-3 / 4
    nsteps = int(np.ceil(timeend / dt))
    dt = timeend / nsteps
    assert nsteps == 800

    sol = oode.LSRK54CarpenterKennedy(dgm, [Q], dt=dt, t0=0.0)
    oode.solve([Q], sol, timeend=timeend)
    err = _weighted_distance(g, Q, _state(g, exact, timeend))
    assert abs(err - EXPECTED_3D_LEVEL1) <= 1e-7 * EXPECTED_3D_LEVEL1, err
