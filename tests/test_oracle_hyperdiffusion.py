"""Oracle hyperdiffusion passes (DryBiharmonic, horizontal direction).  The reference pins these
kernels only through convergence tests of a separate advection-diffusion balance law
(test/Numerics/DGMethods/advection_diffusion/hyperdiffusion_*.jl): "parity unpinned" for the
AtmosModel hooks.  Here: the same kind of analytic check -- on a periodic box the discrete
horizontal Laplacian / grad-Laplacian of a smooth field converge to the exact ones -- plus
consistency properties of the restatement."""
import numpy as np
import pytest

from oracle import atmos, dgmodel as odg, grids as G, topologies as tp, mpistatearrays as msa


def _setup(ne, csize=1, L=1000.0):
    br = (np.linspace(0, L, ne + 1), np.linspace(0, L, ne + 1), np.linspace(0, L, 3))
    topos = tp.StackedBrickTopology(csize, br, periodicity=(True, True, True), boundary=((0, 0), (0, 0), (0, 0)))
    gs = [G.Grid(t, 4) for t in topos]
    model = atmos.DryAtmosModel(np.float64, orientation="flat",
                                ref_state=dict(T_surf=300.0, T_min=220.0, H_t=8e3, subtract_off=True),
                                turbulence=("constant_kinematic", 0.0, False), sources=("gravity",),
                                hyperdiffusion=("dry_biharmonic", 3600.0))
    dgm = odg.DGModel(model, gs, "rusanov", diffusion_direction="horizontal")
    return model, gs, dgm


def _state(model, gs, dgm, L=1000.0):
    """rho = 1, u = (sin(kx) cos(ky), 0.3 cos(kx), 0.2 sin(ky)), T uniform: u_h = (u1, u2, 0)."""
    ps = model.ps
    Qs = []
    k = 2 * np.pi / L
    for g, aux in zip(gs, dgm.state_auxiliary):
        q = msa.MPIStateArray.from_grid(g, 5)
        x, y = g.vgeo[:, G._x1], g.vgeo[:, G._x2]
        u = [np.sin(k * x) * np.cos(k * y), 0.3 * np.cos(k * x), 0.2 * np.sin(k * y)]
        Φ = aux.data[:, model.a_Φ]
        e = 0.5 * (u[0] ** 2 + u[1] ** 2 + u[2] ** 2) + Φ + ps.cv_d * (290.0 - ps.T_0)
        q.data[:, 0] = 1.0
        for d in range(3):
            q.data[:, 1 + d] = u[d]
        q.data[:, 4] = e
        Qs.append(q)
    return Qs, k


def test_lengthscale_horizontal_uniform_box():
    model, gs, dgm = _setup(4)
    g = gs[0]
    Δ = dgm.state_auxiliary[0].data[:g.nreal, model.a_Δh]
    assert np.allclose(Δ, (1000.0 / 4) / 4, rtol=1e-12)      # element width / N


def _errors(ne):
    model, gs, dgm = _setup(ne)
    Q, k = _state(model, gs, dgm)
    dQ = [q.similar() for q in Q]
    dgm(dQ, Q, 0.0, 1, 0)
    g = gs[0]
    x, y = g.vgeo[:g.nreal, G._x1], g.vgeo[:g.nreal, G._x2]
    lap = dgm.Qhypervisc_div[0].data[:g.nreal]
    # horizontal Laplacian of u_h1 = sin(kx) cos(ky) is -2 k^2 u_h1; of u_h2 = 0.3 cos(kx): -k^2 u_h2
    e1 = np.max(np.abs(lap[:, 0] + 2 * k * k * np.sin(k * x) * np.cos(k * y))) / (2 * k * k)
    e2 = np.max(np.abs(lap[:, 1] + k * k * 0.3 * np.cos(k * x))) / (0.3 * k * k)
    assert np.max(np.abs(lap[:, 2])) < 1e-7 * k * k          # vertical velocity is projected out
    # hyperdiffusive state: nu4 * d/dx_d (lap u_h c) at d + 3 c
    ν4 = ((1000.0 / ne / 4) / 2) ** 4 / 2 / 3600.0
    H = dgm.Qhypervisc_grad[0].data[:g.nreal]
    exact = ν4 * (-2 * k ** 3) * np.cos(k * x) * np.cos(k * y)           # d/dx lap u_h1
    e3 = np.max(np.abs(H[:, 0] - exact)) / (ν4 * 2 * k ** 3)
    assert np.max(np.abs(H[:, 2])) < 1e-9 * ν4 * k ** 3                   # no vertical derivative
    return e1, e2, e3


def test_horizontal_laplacian_and_hyperflux_converge():
    """Central-flux DG of degree 4: Laplacian converges at order ~3, its gradient at order ~2."""
    c, f = _errors(4), _errors(8)
    assert f[0] < 4e-3 and f[1] < 8e-3 and f[2] < 5e-2
    assert c[0] / f[0] > 5 and c[1] / f[1] > 5      # ~2^3 per halving
    assert c[2] / f[2] > 2.5                        # ~2^2 per halving


def test_emulated_ranks_agree():
    """3 emulated ranks (Hilbert partition + the three extra exchanges) reproduce the single-rank
    tendency."""
    model, gs1, dgm1 = _setup(4, 1)
    Q1, _ = _state(model, gs1, dgm1)
    dQ1 = [q.similar() for q in Q1]
    dgm1(dQ1, Q1, 0.0, 1, 0)
    model3, gs3, dgm3 = _setup(4, 3)
    Q3, _ = _state(model3, gs3, dgm3)
    dQ3 = [q.similar() for q in Q3]
    dgm3(dQ3, Q3, 0.0, 1, 0)
    # compare through coordinates (element order differs between partitions)
    def key(g, arr):
        c = np.round(g.vgeo[:g.nreal, [G._x1, G._x2, G._x3]].mean(axis=2), 6)
        return {tuple(ci): a for ci, a in zip(c, arr[:g.nreal])}
    ref = key(gs1[0], dQ1[0].data)
    scale = np.abs(dQ1[0].data).max(axis=(0, 2))[:, None]
    for g, d in zip(gs3, dQ3):
        for kk, a in key(g, d.data).items():
            assert np.max(np.abs(a - ref[kk]) / scale) < 1e-12
