"""Pins the oracle's hyperdiffusion *kernels* (gradients stored by the gradient kernels, divergence of
gradients, gradients of Laplacians, CentralNumericalFluxDivergence / CentralNumericalFluxHigherOrder,
the DGModel schedule with its three extra exchanges) on the reference's golden value of
``test/Numerics/DGMethods/advection_diffusion/periodic_3D_hyperdiffusion.jl``: 3-D,
``HorizontalDirection``, level 1 (4^3 elements of order 4 on the periodic box [0, 2 pi]^3), constant
hyperdiffusion tensor, LSRK54 to t = 1 with dt = dx^4 / 25 / sum(D):
``euclidean_distance(Q, Q_exact) = 1.9244127301149615e-02`` (``expected_result[3, 1, Float64,
HorizontalDirection]``, checked there with ``isapprox``).

The balance law of that test (one scalar, flux = eta = H grad(lap rho)) is restated below; the
AtmosModel hooks of DryBiharmonic sit on the same kernels (tests/test_oracle_hyperdiffusion.py).
``direction = HorizontalDirection()`` of the reference's DGModel also restricts the tendency kernels
to xi1 / xi2 and faces 1-4; the oracle's tendency is EveryDirection-only, so the test hands it a grid
whose vertical metric terms and top / bottom face weights are zero, which is the same operator."""
import copy

import numpy as np

from oracle import dgmodel as odg, grids as G, topologies as tp, courant
from oracle import odesolvers as oode, mpistatearrays as msa

EXPECTED = 1.9244127301149615e-02        # periodic_3D_hyperdiffusion.jl:244


class HyperDiffusionLaw:
    """AdvectionDiffusion{3}(ConstantHyperDiffusion; advection = false, diffusion = false,
    hyperdiffusion = true) (advection_diffusion_model.jl:92-330)."""
    S, A, G, GF = 1, 1, 1, 0
    ngradlap, nhyper, hyper_G = 1, 3, 0

    def __init__(self, H):
        self.H = np.asarray(H, dtype=np.float64)
        self.t = 0.0

    def viscous(self):
        return True

    def init_state_auxiliary(self, dg):
        pass

    def nodal_update_aux(self, Q, aux):
        pass

    def flux_first_order(self, Q, aux):
        return np.zeros((3, 1) + Q.shape[1:])

    def flux_second_order(self, Q, GF, aux):
        return np.zeros((3, 1) + Q.shape[1:])

    def flux_hyperdiffusive(self, Q, Hs):
        return Hs[:, None]                                   # flux.rho += eta

    def source(self, Q, aux):
        return np.zeros_like(Q)

    def gradient_argument(self, Q, aux):
        return Q[0:1].copy()                                 # transform.rho = state.rho

    def gradient_flux(self, gradG, Q, aux):
        return np.zeros((0,) + Q.shape[1:])

    def transform_post_gradient_laplacian(self, gradlap, Q, aux):
        g = gradlap[:, 0]                                    # eta = H * grad(lap rho)
        return np.stack([self.H[i, 0] * g[0] + self.H[i, 1] * g[1] + self.H[i, 2] * g[2] for i in range(3)])


def test_periodic_3d_hyperdiffusion_horizontal_level1_golden_error():
    D = np.array([[9, 3, 5], [3, 7, 4], [5, 4, 10]], dtype=np.float64) / 50 / 100
    xr = np.linspace(0.0, 2 * np.pi, 5)
    topo = tp.StackedBrickTopology(1, (xr, xr, xr), periodicity=(True, True, True), connectivity="full")[0]
    g = G.Grid(topo, 4)
    dx = courant.min_node_distance(g)
    dt = dx ** 4 / 25 / D.sum()
    dt = 1.0 / np.ceil(1.0 / dt)
    # HorizontalDirection of the whole operator: no xi3 metric terms, no top / bottom face terms
    gh = copy.copy(g)
    gh.vgeo = g.vgeo.copy()
    gh.sgeo = g.sgeo.copy()
    for c in (G._xi3x1, G._xi3x2, G._xi3x3):
        gh.vgeo[:, c] = 0
    gh.sgeo[:, 4:6, :, G._sM] = 0
    law = HyperDiffusionLaw(D)
    dgm = odg.DGModel(law, [gh], "central", diffusion_direction="horizontal")
    k = np.array([1.0, 2.0, 3.0])
    kD = np.outer(k, k) * D
    c = (k[:2] ** 2).sum() * kD[:2, :2].sum()

    def exact(t):
        q = msa.MPIStateArray.from_grid(g, 1)
        x = [g.vgeo[:, G._x1], g.vgeo[:, G._x2], g.vgeo[:, G._x3]]
        q.data[:, 0] = np.sin(k[0] * x[0] + k[1] * x[1] + k[2] * x[2]) * np.exp(-c * t)
        return q
    Q = exact(0.0)
    sol = oode.LSRK54CarpenterKennedy(dgm, [Q], dt=dt, t0=0.0)
    oode.solve([Q], sol, timeend=1.0)
    Qe = exact(1.0)
    M = g.vgeo[:g.nreal, G._M][:, None, :]
    err = float(np.sqrt(np.sum(M * (Q.data[:g.nreal] - Qe.data[:g.nreal]) ** 2)))
    assert abs(err - EXPECTED) <= 1e-7 * EXPECTED, err
