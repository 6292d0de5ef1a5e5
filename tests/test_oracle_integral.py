"""The oracle's stack integral (the HBModel's column operator) on the reference's known-answer test
``test/Numerics/DGMethods/integral_test.jl`` (3-D): integrands a = x + y and
b = 2x + sin(x) y - (z - 1)^2 y^2 on the 5^3 stacked brick [0, 3]^3 of order 4 integrate upwards to
a_int = x z + y z and b_int = 2 x z + sin(x) y z - (1 + (z - 1)^3) y^2 / 3 (`≈`), and the reverse
integral is the top value minus the upward one."""
import numpy as np

from oracle import grids as G, topologies as tp, ocean


def test_indefinite_stack_integral_reference_known_answers():
    br = tuple(np.linspace(0.0, 3.0, 6) for _ in range(3))
    topo = tp.StackedBrickTopology(1, br, periodicity=(True, True, True), connectivity="full")[0]
    g = G.Grid(topo, 4)
    vg = g.vgeo[:g.nreal]
    x, y, z = vg[:, G._x1], vg[:, G._x2], vg[:, G._x3]
    kern = np.stack([x + y, 2 * x + np.sin(x) * y - (z - 1) ** 2 * y ** 2])
    out = ocean.indefinite_stack_integral(g, kern)
    nv = g.topology.stacksize
    got = out.reshape(2, g.nreal, g.Np)
    exact = np.stack([x * z + y * z, 2 * x * z + np.sin(x) * y * z - (1 + (z - 1) ** 3) * y ** 2 / 3])
    assert np.allclose(got, exact, rtol=1.5e-8, atol=1e-12)
    assert np.max(np.abs(got - exact)) <= 1e-12 * np.max(np.abs(exact))
    # reverse integral (kernel_reverse_indefinite_stack_integral!, :1992-2046): top value minus the integral
    top = out[:, :, nv - 1, g.Nq[2] - 1, :]
    rev = (top[:, :, None, None, :] - out).reshape(2, g.nreal, g.Np)
    zt = 3.0
    exact_top = np.stack([x * zt + y * zt, 2 * x * zt + np.sin(x) * y * zt - (1 + (zt - 1) ** 3) * y ** 2 / 3])
    assert np.max(np.abs(rev - (exact_top - exact))) <= 1e-12 * np.max(np.abs(exact))
