"""HBModel on the reference's 3-D hydrostatic spin-down regression (test/Ocean/HydrostaticBoussinesq/
test_3D_spindown.jl with test/Ocean/refvals/3D_hydrostatic_spindown_refvals.jl, `explicit`): SimpleBox (Fixed),
5 x 5 x 8 elements of order 4 on 1e6 x 1e6 x 400 m, periodic in x and y, free-slip bottom / penetrable free-slip
surface, c_h = 1, no buoyancy, no diffusion, no rotation, LSRK144 with dt = 120 s to t = 86400 s (720 steps =
10 080 evaluations).  The NumPy oracle is far too slow for that, so the run is made by its C twin (oracle/c/hb_ref.c),
which is first held against the NumPy oracle on this very configuration (it exercises the free-slip / penetrable /
insulating boundary branches and the periodic connectivity that the ocean-gyre regression does not)."""
import numpy as np

import __graft_entry__ as ge
from oracle import cref, dgmodel as odg, grids, ocean, odesolvers as oode, topologies as tp
from tests import parity
from tests.test_oracle_ocean import close_digits

# (min, max, mean, std) of the `explicit` block of 3D_hydrostatic_spindown_refvals.jl
REF = {
    ("Q", 0): (-9.58544066049463849843e-01, 9.58544066049465071089e-01, -6.13908923696726568442e-17,
               4.45400263687296238402e-01),
    ("Q", 2): (-8.52732886154656810618e-01, 8.52845586939211197652e-01, 2.20052243093959998331e-14,
               6.02992088522925295813e-01),
    ("aux", 1): (-4.04553460063758398447e-04, 4.04714358463272711169e-04, 4.75730566051879549438e-19,
                 1.63958655681888576441e-04),
    ("aux", 3): (-2.01164684799271339293e-04, 2.01041968159484089494e-04, -2.10942374678779754294e-20,
                 1.42228420244455277133e-04),
}
DIGITS = (12, 12, 0, 12)       # `parr` of the refvals file: the means are round-off and not compared


def spindown_setup(nelem=(5, 5, 8)):
    Lx, Ly, H = 1e6, 1e6, 400.0
    br = (np.linspace(0, Lx, nelem[0] + 1), np.linspace(0, Ly, nelem[1] + 1), np.linspace(-H, 0, nelem[2] + 1))
    topos = tp.StackedBrickTopology(1, br, periodicity=(True, True, False), boundary=((0, 0), (0, 0), (1, 2)))
    g = grids.Grid(topos[0], 4)
    prob = ocean.SimpleBox(Lx, Ly, H)
    xi = g.xi[2]
    model = ocean.HBModel(prob, ch=1.0, alphaT=0.0, kappah=0.0, kappaz=0.0, f0=0.0, beta=0.0,
                          bcs=(("freeslip", "insulating"), ("penetrable_freeslip", "insulating")),
                          vert_filter=ocean.cutoff_filter_matrix(xi, 3),
                          exp_filter=ocean.exponential_filter_matrix(xi, 1, 8))
    return model, g, prob


def test_spindown_c_twin_matches_numpy_oracle():
    ge.build()
    model, g, prob = spindown_setup((3, 2, 3))
    dgm = odg.DGModel(model, [g], "rusanov")
    q = odg.init_ode_state(dgm, lambda x1, x2, x3, a, t: prob.init_state(x1, x2, x3, 0.0, model), 0.0)
    sol = oode.LSRK144NiegemannDiehlBusch(dgm, q, dt=120.0)
    oode.solve(q, sol, numberofsteps=1)
    c = cref.CRefHB.from_grid(model, g, "rusanov")
    cq, caux = q[0].data.copy(), dgm.state_auxiliary[0].data.copy()
    cdq = np.full_like(cq, np.nan)
    dq = [q[0].similar()]
    dgm(dq, q, 0.0, 1, 0)
    c.tendency(cdq, cq, caux, 1.0, 0.0)
    assert parity.rel_l2(c.gradflux, dgm.state_gradient_flux[0].data) < 1e-13
    assert parity.rel_l2(caux[:, 1:4], dgm.state_auxiliary[0].data[:, 1:4]) < 1e-12
    assert parity.rel_l2(cdq, dq[0].data) < 1e-13
    sol2 = oode.LSRK144NiegemannDiehlBusch(dgm, q, dt=120.0)
    oode.solve(q, sol2, numberofsteps=2)
    cdq[...] = 0
    c.lsrk_steps(cq, cdq, caux, 120.0, sol2.RKA, sol2.RKB, 2)
    assert parity.rel_l2(cq, q[0].data) < 1e-13


def test_3d_hydrostatic_spindown_refvals():
    ge.build()
    model, g, prob = spindown_setup()
    dgm = odg.DGModel(model, [g], "rusanov")
    q = odg.init_ode_state(dgm, lambda x1, x2, x3, a, t: prob.init_state(x1, x2, x3, 0.0, model), 0.0)
    sol = oode.LSRK144NiegemannDiehlBusch(dgm, q, dt=120.0)
    c = cref.CRefHB.from_grid(model, g, "rusanov")
    cref.use_all_cores_hb()
    Q, aux = q[0].data.copy(), dgm.state_auxiliary[0].data.copy()
    dQ = np.zeros_like(Q)
    nsteps = int(round(86400.0 / 120.0))
    assert nsteps == 720
    c.lsrk_steps(Q, dQ, aux, 120.0, sol.RKA, sol.RKB, nsteps)
    # error against the analytic solution (the reference asserts < 0.005 and reports 0.0011289879366523504)
    x = [g.vgeo[:, grids._x1], g.vgeo[:, grids._x2], g.vgeo[:, grids._x3]]
    Qe = np.moveaxis(prob.init_state(x[0], x[1], x[2], 86400.0, model), 0, 1)
    M = g.vgeo[:, grids._M][:, None, :]
    err = np.sqrt(np.sum(M * (Q - Qe) ** 2)) / np.sqrt(np.sum(M * Qe ** 2))
    assert abs(err - 0.0011289879366523504) < 1e-11, err

    def stats(v):
        v = v.ravel()
        mean = v.mean()
        return v.min(), v.max(), mean, np.sqrt(np.sum((v - mean) ** 2) / (v.size - 1))
    for (name, ivar), ref in REF.items():
        got = stats((Q if name == "Q" else aux)[:, ivar, :])
        for gv, r, d in zip(got, ref, DIGITS):
            if d:
                # 2 digits fewer than the reference's own same-machine gate, as for the other ocean regressions
                assert close_digits(gv, r, d - 2), (name, ivar, got, ref)
    # fields the reference holds at exactly zero or at round-off
    assert np.abs(Q[:, 1]).max() < 1e-12 and np.all(Q[:, 3] == 0) and np.all(aux[:, 2] == 0)
