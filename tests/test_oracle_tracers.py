"""Oracle: passive tracers (`NTracers{N}`, src/Atmos/Model/tracers.jl, tendencies_tracers.jl) -- the
part of BASELINE.json configs[0] (tutorials/Atmos/risingbubble.jl) the north star does not name.
Parity unpinned (the reference has no golden value for tracers); property tests of the restatement:
layout sizes of SURVEY 8.0, passivity, linearity, constant mixing ratio under pure advection, and the
delta_chi scaling of the diffusive part."""
import numpy as np
import pytest

from tests import parity
from oracle import dgmodel as odg, atmos as oatmos, mpistatearrays as omsa, topologies as tp, grids as ogrids


def _setup(tracers, turbulence=("smagorinsky", 0.21), nf="rusanov"):
    br = (np.linspace(0, 1500, 4), np.linspace(0, 1000, 3), np.linspace(0, 1500, 4))
    topo = tp.StackedBrickTopology(1, br, periodicity=(True, True, False), boundary=((0, 0), (0, 0), (1, 2)))[0]
    g = ogrids.Grid(topo, 4)
    m = oatmos.DryAtmosModel(np.float64, orientation="flat",
                             ref_state=dict(T_surf=300.0, T_min=220.0, H_t=8e3, subtract_off=True),
                             turbulence=turbulence, sources=("gravity",), bcs=("freeslip", "noslip"),
                             tracers=tracers)
    return g, m, odg.DGModel(m, [g], nf)


def _chi(aux):
    x, y, z = aux[0], aux[1], aux[2]
    return [0.01 * (1 + np.sin(2 * np.pi * x / 1500) * np.cos(np.pi * z / 1500)),
            0.02 * np.exp(-((x - 700) ** 2 + (z - 600) ** 2) / 300 ** 2),
            0.5 + 0 * x,
            0.03 * np.cos(2 * np.pi * y / 1000)]


def _tendency(g, m, dgm, chi_scale=1.0):
    aux = np.moveaxis(dgm.state_auxiliary[0].data[:g.nreal], 1, 0)
    Q5 = parity.bubble_state(m, g, aux)
    Q0 = Q5 if not m.NT else np.concatenate([Q5, np.stack([Q5[0] * c * chi_scale for c in _chi(aux)[:m.NT]])])
    q = omsa.MPIStateArray.from_grid(g, m.S)
    np.moveaxis(q.data[:g.nreal], 1, 0)[...] = Q0
    omsa.ghost_exchange([q])
    dq = q.similar()
    dgm([dq], [q], 0.0, 1, 0)
    return q, dq


def test_rising_bubble_layout_sizes():
    """S = 9, A = 21, G = 9, GF = 22 for Smagorinsky + NTracers{4}."""
    g, m, dgm = _setup((1.0, 2.0, 3.0, 4.0))
    assert (m.S, m.A, m.G, m.GF) == (9, 21, 9, 22)
    a = dgm.state_auxiliary[0].data
    assert np.all(a[:g.nreal, m.a_δχ] == np.array([1.0, 2.0, 3.0, 4.0])[None, :, None])


@pytest.mark.parametrize("nf", ["rusanov", "central"])
def test_tracers_are_passive_and_linear(nf):
    g, m, dgm = _setup((1.0, 2.0, 3.0, 4.0), nf=nf)
    q, dq = _tendency(g, m, dgm)
    g0, m0, dgm0 = _setup(None, nf=nf)
    _, dq0 = _tendency(g0, m0, dgm0)
    assert np.array_equal(dq.realdata[:, :5], dq0.realdata)              # the flow does not see them
    assert np.abs(dq.realdata[:, 5:]).max() > 0
    _, dq2 = _tendency(g, m, dgm, chi_scale=2.0)
    scale = np.abs(dq.realdata[:, 5:]).max()
    assert np.abs(dq2.realdata[:, 5:] - 2 * dq.realdata[:, 5:]).max() <= 1e-13 * scale


def test_constant_mixing_ratio_follows_the_density_with_the_central_flux():
    """chi = const, no diffusion of a constant: with the central flux (no wave-speed mismatch between
    mass and tracers) the tracer tendency is chi times the mass tendency."""
    g, m, dgm = _setup((1.0, 2.0, 3.0), nf="central")
    q, dq = _tendency(g, m, dgm)
    t = dq.realdata
    assert np.abs(t[:, 7] - 0.5 * t[:, 0]).max() <= 1e-12 * np.abs(t[:, 0]).max()


def test_diffusive_part_scales_with_delta_chi():
    """Tendency(delta) - tendency(delta = 0) is linear in delta_chi and vanishes for constant chi."""
    g, m0, dgm0 = _setup((0.0, 0.0, 0.0, 0.0))
    _, d0 = _tendency(g, m0, dgm0)
    _, m1, dgm1 = _setup((1.0, 1.0, 1.0, 1.0))
    _, d1 = _tendency(g, m1, dgm1)
    _, m3, dgm3 = _setup((3.0, 3.0, 3.0, 3.0))
    _, d3 = _tendency(g, m3, dgm3)
    diff1 = d1.realdata[:, 5:] - d0.realdata[:, 5:]
    diff3 = d3.realdata[:, 5:] - d0.realdata[:, 5:]
    assert np.abs(diff1[:, [0, 1, 3]]).max() > 0
    assert np.abs(diff3 - 3 * diff1).max() <= 1e-10 * np.abs(diff1).max()
    assert np.abs(diff1[:, 2]).max() <= 1e-12 * np.abs(diff1).max()       # the constant tracer does not diffuse


def test_tracer_tendency_conditioning():
    """Why the device parity bar of a single tracer column is 1e-11 rather than 1e-12: perturbing
    aux.theta_v by one unit of round-off (what another exp/log implementation does) moves the tracer
    tendencies of the Smagorinsky run by ~1e-12 x delta_chi through N^2 -> Richardson correction -> D_t,
    while the five dynamic states move by ~1e-15."""
    g, m, dgm = _setup((1.0, 2.0, 3.0, 4.0))
    q, dq = _tendency(g, m, dgm)
    g2, m2, dgm2 = _setup((1.0, 2.0, 3.0, 4.0))
    orig = m2.nodal_update_aux
    rng = np.random.default_rng(0)

    def noisy(Q, aux):
        orig(Q, aux)
        aux[m2.a_θv] = aux[m2.a_θv] * (1 + 1.1e-16 * rng.choice([-1, 0, 1], size=aux[m2.a_θv].shape))
    m2.nodal_update_aux = noisy
    _, dq2 = _tendency(g2, m2, dgm2)
    dyn = parity.rel_l2(dq2.realdata[:, :5], dq.realdata[:, :5])
    trc = [parity.rel_l2(dq2.realdata[:, 5 + i], dq.realdata[:, 5 + i]) for i in range(4)]
    assert dyn < 1e-13
    assert 1e-14 < max(trc) < 1e-10, trc


def test_emulated_ranks_agree_with_tracers():
    """3 emulated ranks (Hilbert partition, ghost exchange of the 9-column state and of the 22-column
    gradient flux) reproduce the single-rank tendency, tracers included."""
    def run(csize):
        br = (np.linspace(0, 1500, 4), np.linspace(0, 1000, 4), np.linspace(0, 1500, 4))
        topos = tp.StackedBrickTopology(csize, br, periodicity=(True, True, False),
                                        boundary=((0, 0), (0, 0), (1, 2)))
        gs = [ogrids.Grid(t, 4) for t in topos]
        m = oatmos.DryAtmosModel(np.float64, orientation="flat",
                                 ref_state=dict(T_surf=300.0, T_min=220.0, H_t=8e3, subtract_off=True),
                                 turbulence=("smagorinsky", 0.21), sources=("gravity",),
                                 bcs=("freeslip", "noslip"), tracers=(1.0, 2.0, 3.0, 4.0))
        dgm = odg.DGModel(m, gs, "rusanov")
        Q = []
        for g, aux in zip(gs, dgm.state_auxiliary):
            a = np.moveaxis(aux.data[:g.nreal], 1, 0)
            Q5 = parity.bubble_state(m, g, a)
            q = omsa.MPIStateArray.from_grid(g, m.S)
            np.moveaxis(q.data[:g.nreal], 1, 0)[...] = np.concatenate([Q5, np.stack([Q5[0] * c for c in _chi(a)])])
            Q.append(q)
        dQ = [q.similar() for q in Q]
        dgm(dQ, Q, 0.0, 1, 0)
        out = {}
        for g, d in zip(gs, dQ):
            c = np.round(g.vgeo[:g.nreal][:, [ogrids._x1, ogrids._x2, ogrids._x3]].mean(axis=2), 6)
            out.update({tuple(ci): arr for ci, arr in zip(c, d.data[:g.nreal])})
        return out
    one, three = run(1), run(3)
    assert one.keys() == three.keys()
    scale = np.max([np.abs(v).max(axis=1) for v in one.values()], axis=0)[:, None]
    for k, v in one.items():
        assert np.max(np.abs(three[k] - v) / scale) < 1e-12
