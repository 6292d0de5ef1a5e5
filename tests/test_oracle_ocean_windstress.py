"""A second HBModel regression pin: ``test/Ocean/refvals/test_windstress_refvals.jl`` (``explicit_cpu``),
produced by ``test/Ocean/HydrostaticBoussinesq/test_windstress_short.jl`` through
``experiments/OceanBoxGCM/simple_box.jl``: HomogeneousBox (theta = 20, at rest, jet-like wind stress),
1000 km x 1000 km x 400 m, 5^3 elements of order 4, boundary conditions (NoSlip, Insulating) on the sides,
(FreeSlip, Insulating) at the bottom, (KinematicStress, Insulating) at the surface, LSRK144 with
dt = 180 s for one hour.  Exercises the free-slip and insulating ocean boundary conditions, which the
ocean-gyre pin does not."""
import numpy as np

from oracle import topologies as tp, grids, ocean, dgmodel, odesolvers
from tests.test_oracle_ocean import close_digits

# [min, max, mean, std] of Q: u[1], u[2], eta, theta and s_aux: y, w, pkin, wz0 (refVals.explicit_cpu)
REF = {
    ("Q", 0): (-3.74270752639261211625e-02, 3.72763215301363109999e-02, -1.40392694287316997219e-06, 5.29521849931491837837e-03),
    ("Q", 1): (-7.13776376792300999707e-03, 6.54949226335132476257e-03, -5.45004642311143043000e-06, 9.84283148803805074678e-04),
    ("Q", 2): (-5.75146380759523553894e-03, 5.06819905742867966164e-03, -1.61399463692112299070e-05, 1.58594255118538803723e-03),
    ("aux", 0): (0.0, 1.00000000000000011642e+06, 5.0e+05, 2.92779390974978974555e+05),
    ("aux", 1): (-1.91903846873268650749e-05, 1.88201368003043060118e-05, 6.88331165208082022114e-08, 1.81657738735220659856e-06),
    ("aux", 2): (-1.60000000000003339551e+00, 0.0, -8.00000000000017919000e-01, 4.68447025559975749331e-01),
    ("aux", 3): (-1.79359066322475932405e-06, 1.42978124248321576704e-06, -3.01455602004428728418e-09, 6.97246485535118604352e-07),
}
# digits the reference compares (parr): eta mean 11, wz0 mean 10, everything else 12
DIGITS = {("Q", 2): (12, 12, 11, 12), ("aux", 3): (12, 12, 10, 12)}


class HomogeneousBox:
    """src/Ocean/OceanProblems/homogeneous_box.jl:15-66."""

    def __init__(self, Lx, Ly, H, tau0=1e-1):
        self.Lx, self.Ly, self.H, self.tau0 = float(Lx), float(Ly), float(H), tau0

    def init_state(self, x, y, z):
        Q = np.zeros((4,) + y.shape)
        Q[3] = 20.0
        return Q

    def kinematic_stress(self, y, rho):
        return [(self.tau0 / rho) * np.cos(y * np.pi / self.Ly), 0 * y]


def test_windstress_short_explicit_refvals():
    Lx, Ly, H = 1e6, 1e6, 400.0
    br = (np.linspace(0, Lx, 6), np.linspace(0, Ly, 6), np.linspace(-H, 0, 6))
    topo = tp.StackedBrickTopology(1, br, periodicity=(False, False, False), boundary=((1, 1), (1, 1), (2, 3)))
    g = grids.Grid(topo[0], 4)
    prob = HomogeneousBox(Lx, Ly, H)
    xi = g.xi[2]
    model = ocean.HBModel(prob, bcs=(("noslip", "insulating"), ("freeslip", "insulating"),
                                     ("kinematic_stress", "insulating")),
                          vert_filter=ocean.cutoff_filter_matrix(xi, 3),
                          exp_filter=ocean.exponential_filter_matrix(xi, 1, 8))
    dg = dgmodel.DGModel(model, [g], "rusanov")
    Q = dgmodel.init_ode_state(dg, lambda x1, x2, x3, a, t: prob.init_state(x1, x2, x3), 0.0)
    sol = odesolvers.LSRK144NiegemannDiehlBusch(dg, Q, dt=180.0, t0=0.0)
    odesolvers.solve(Q, sol, timeend=3600.0)
    assert sol.steps == 20
    for (name, ivar), ref in REF.items():
        arr = Q[0] if name == "Q" else dg.state_auxiliary[0]
        got = ocean.statecheck(arr, ivar)
        digs = DIGITS.get((name, ivar), (12, 12, 12, 12))
        for gv, r, d in zip(got, ref, digs):
            assert close_digits(gv, r, d - 2), (name, ivar, got, ref)
    # theta stays 20 to round-off (the reference lists its std as 2.6e-13 with 0 digits compared)
    th = ocean.statecheck(Q[0], 3)
    assert abs(th[0] - 20) < 1e-10 and abs(th[1] - 20) < 1e-10 and th[3] < 1e-11
