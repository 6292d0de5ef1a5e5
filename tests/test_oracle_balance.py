"""Pins Gravity + HydrostaticState (discrete pressure-gradient balance of the reference density) on
the reference's own property test ``test/Atmos/Model/discrete_hydrostatic_balance.jl``: initialised to
the reference state (``subtract_off = false``, ``ConstantDynamicViscosity(0)``, source = Gravity), the
LES box and the GCM shell stay balanced to ``100 eps`` relative
(``euclidean_distance(Q, Qinit) / norm(Qinit) <= 100 * eps(FT)``) for the Central and Roe fluxes and
the isothermal and decaying temperature profiles.  The reference integrates to t = 100 s; here 10
LSRK54 steps at Courant number 0.1 (the property is per step) on the same meshes: polynomial order 4,
domain height 50 km, LES box 50 km^3 with resolution height / 12 (3 x 3 x 3 elements), GCM shell with
3 horizontal x 3 vertical elements per panel."""
import numpy as np
import pytest

from oracle import atmos as oatmos, dgmodel as odg, grids as G, topologies as tp
from oracle import odesolvers as oode, mpistatearrays as msa

HEIGHT = 50e3


def _grid(config):
    ps = oatmos.Params(np.float64)
    if config == "LES":
        br = tuple(np.linspace(0.0, HEIGHT, 4) for _ in range(3))
        topo = tp.StackedBrickTopology(1, br, periodicity=(True, True, False),
                                       boundary=((0, 0), (0, 0), (1, 2)))[0]
        return G.Grid(topo, 4), "flat"
    a = float(ps.planet_radius)
    topo = tp.StackedCubedSphereTopology(1, 3, np.linspace(a, a + HEIGHT, 4), boundary=(1, 2))[0]
    return G.Grid(topo, 4, meshwarp=tp.equiangular_cubed_sphere_warp), "spherical"


@pytest.mark.parametrize("config", ["LES", "GCM"])
@pytest.mark.parametrize("nf", ["central", "roe"])
@pytest.mark.parametrize("profile", ["isothermal", "decaying"])
def test_reference_state_stays_balanced(config, nf, profile):
    ps = oatmos.Params(np.float64)
    g, orientation = _grid(config)
    T_surf = float(ps.T_surf_ref)
    # IsothermalProfile(param_set, FT) = DecayingTemperatureProfile(T_surf_ref, T_surf_ref);
    # DecayingTemperatureProfile{FT}(param_set): 290 K -> 220 K over H_t = R_d T_surf / grav
    T_min = T_surf if profile == "isothermal" else float(ps.T_min_ref)
    H_t = float(ps.R_d) * T_surf / float(ps.grav)
    model = oatmos.DryAtmosModel(np.float64, orientation=orientation,
                                 ref_state=dict(T_surf=T_surf, T_min=T_min, H_t=H_t, subtract_off=False),
                                 turbulence=("constant_dynamic", 0.0, False), sources=("gravity",),
                                 bcs=("freeslip", "freeslip"))
    dgm = odg.DGModel(model, [g], nf, diffusion_direction="horizontal")
    aux = dgm.state_auxiliary[0].data
    Q = msa.MPIStateArray.from_grid(g, 5)
    Q.data[:, 0] = aux[:, model.a_ref["ρ"]]
    Q.data[:, 4] = aux[:, model.a_ref["ρe"]]
    Q0 = Q.data.copy()
    # Courant number 0.1 on the smallest node distance with the surface sound speed
    vg = g.vgeo[:g.nreal]
    x = np.stack([vg[:, G._x1], vg[:, G._x2], vg[:, G._x3]], axis=-1).reshape(g.nreal, 5, 5, 5, 3)
    dmin = min(np.linalg.norm(np.diff(x, axis=ax), axis=-1).min() for ax in (1, 2, 3))
    dt = 0.1 * dmin / float(oatmos.soundspeed_air(ps, np.float64(T_surf)))
    sol = oode.LSRK54CarpenterKennedy(dgm, [Q], dt=dt, t0=0.0)
    oode.solve([Q], sol, numberofsteps=10)
    M = vg[:, G._M][:, None, :]
    err = np.sqrt(np.sum(M * (Q.data[:g.nreal] - Q0[:g.nreal]) ** 2)) / np.sqrt(np.sum(M * Q0[:g.nreal] ** 2))
    assert err <= 100 * np.finfo(np.float64).eps, err
