"""Pins the oracle's cubed-sphere metric terms and the element-local strong derivative (what builds
grad Phi for Gravity / the hydrostatic reference state, `auxiliary_field_gradient!`) on the golden values
of the reference's ``test/Numerics/DGMethods/grad_test_sphere.jl``: a = r^3 on the equiangular shell
0.5 <= r <= 1, N = 4, base mesh 4 horizontal x 2 vertical elements per panel;
``euclidean_distance(exact_aux, aux)`` = 4.3759489495202896e-04 (level 1), 2.9065372851175251e-05
(level 2) for EveryDirection (checked there with ``isapprox``)."""
import numpy as np
import pytest

from oracle import dgmodel as odg, grids as G, topologies as tp

EXPECTED = {1: 4.3759489495202896e-04, 2: 2.9065372851175251e-05}     # grad_test_sphere.jl:77-80


class _NoModel:
    """Just enough of a balance law for DGModel.local_gradient."""
    S = G_ = GF = A = 0


@pytest.mark.parametrize("level", [1, 2])
def test_sphere_gradient_golden_error(level):
    ne_h, ne_v = 2 ** (level - 1) * 4, 2 ** (level - 1) * 2
    topo = tp.StackedCubedSphereTopology(1, ne_h, np.linspace(0.5, 1.0, ne_v + 1))[0]
    g = G.Grid(topo, 4, meshwarp=tp.equiangular_cubed_sphere_warp)
    vg = g.vgeo[:g.nreal]
    x = [vg[:, G._x1], vg[:, G._x2], vg[:, G._x3]]
    r = np.sqrt(x[0] ** 2 + x[1] ** 2 + x[2] ** 2)        # hypot(x, y, z)
    a = r ** 3
    exact = [3 * r ** 2 * xi / r for xi in x]
    got = odg.DGModel.local_gradient(None, g, a)
    M = vg[:, G._M]
    err = np.sqrt(sum(np.sum(M * (gd - ex) ** 2) for gd, ex in zip(got, exact)))
    assert err == pytest.approx(EXPECTED[level], rel=1e-7), err


def test_box_gradient_is_exact_for_the_reference_polynomial():
    """test/Numerics/DGMethods/grad_test.jl (3-D, polynomial order (4, 4), EveryDirection):
    a = x^2 + y^3 + z^2 y^2 - x y z on the 5^3 stacked brick [0, 3]^3 is differentiated exactly
    (`Array(aux.grad_a) ≈ Array(exact.grad_a)`)."""
    br = tuple(np.linspace(0.0, 3.0, 6) for _ in range(3))
    topo = tp.StackedBrickTopology(1, br, periodicity=(False, False, False), connectivity="full")[0]
    g = G.Grid(topo, 4)
    vg = g.vgeo[:g.nreal]
    x, y, z = vg[:, G._x1], vg[:, G._x2], vg[:, G._x3]
    a = x ** 2 + y ** 3 + z ** 2 * y ** 2 - x * y * z
    exact = [2 * x - y * z, 3 * y ** 2 + 2 * z ** 2 * y - x * z, 2 * z * y ** 2 - x * y]
    got = odg.DGModel.local_gradient(None, g, a)
    num = np.sqrt(sum(np.sum((gd - ex) ** 2) for gd, ex in zip(got, exact)))
    den = np.sqrt(sum(np.sum(ex ** 2) for ex in exact))
    assert num <= np.sqrt(np.finfo(np.float64).eps) * den      # isapprox's default rtol
    assert num <= 1e-12 * den
