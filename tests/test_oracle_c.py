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
