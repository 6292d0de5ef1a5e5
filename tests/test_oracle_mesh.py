"""Pins the oracle's mesh / grid restatement against the reference's own tests.

Golden connectivity tables below are the expectations written out in
``/root/reference/test/Numerics/Mesh/mpi_connect.jl:36-124``,
``mpi_connect_stacked.jl:36-116``, ``mpi_connect_stacked_3d.jl:35-219`` and
``mpi_connectfull.jl`` (bit-exact integer gates); the analytic checks restate
``test/Numerics/Mesh/Elements.jl`` / ``Grids.jl`` / ``mpi_connect_sphere.jl:63-83``.
"""
import numpy as np
import pytest

from oracle import elements, brickmesh as bm, topologies as tp, grids


def M(s):
    return np.array([[int(t) for t in row.split()] for row in s.strip().splitlines()], dtype=np.int64)


GLOBAL_2D_COORD = {
    1: [[0, 1, 0, 1], [5, 5, 6, 6]], 2: [[1, 2, 1, 2], [5, 5, 6, 6]],
    3: [[1, 2, 1, 2], [6, 6, 7, 7]], 4: [[0, 1, 0, 1], [6, 6, 7, 7]],
    5: [[0, 1, 0, 1], [7, 7, 8, 8]], 6: [[0, 1, 0, 1], [8, 8, 9, 9]],
    7: [[1, 2, 1, 2], [8, 8, 9, 9]], 8: [[1, 2, 1, 2], [7, 7, 8, 8]],
    9: [[2, 3, 2, 3], [7, 7, 8, 8]], 10: [[2, 3, 2, 3], [8, 8, 9, 9]],
    11: [[3, 4, 3, 4], [8, 8, 9, 9]], 12: [[3, 4, 3, 4], [7, 7, 8, 8]],
    13: [[3, 4, 3, 4], [6, 6, 7, 7]], 14: [[2, 3, 2, 3], [6, 6, 7, 7]],
    15: [[2, 3, 2, 3], [5, 5, 6, 6]], 16: [[3, 4, 3, 4], [5, 5, 6, 6]],
}
GLOBAL_2D_FACE = M("""
1 2 2 1 1 1 2 2 2 2 2 2 2 2 2 2
1 1 1 1 1 1 1 1 1 1 2 2 2 1 1 2
4 4 4 4 4 4 4 4 4 4 4 4 4 4 4 4
3 3 3 3 3 3 3 3 3 3 3 3 3 3 3 3""")
GLOBAL_2D_BNDY = M("""
1 0 0 1 1 1 0 0 0 0 0 0 0 0 0 0
0 0 0 0 0 0 0 0 0 0 2 2 2 0 0 2
0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0
0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0""")


def _check_common(t, nreal, globalelems, gcoord, gface, gbndy, e2e, n2r, n2recv, n2send):
    assert t.nelem == len(globalelems)
    assert t.nreal == nreal
    exp_coord = np.stack([np.array(gcoord[g]) for g in globalelems], axis=2)
    assert np.array_equal(t.elemtocoord, exp_coord)
    ge = np.array(globalelems) - 1
    assert np.array_equal(t.elemtoface[:, :nreal], gface[:, ge[:nreal]])
    assert np.array_equal(t.elemtoelem, e2e)
    assert np.array_equal(t.elemtobndy, gbndy[:, ge])
    assert np.all(t.elemtoordr == 1)
    assert t.nabrtorank == n2r
    assert t.nabrtorecv == n2recv
    assert t.nabrtosend == n2send
    assert sorted(set(t.exteriorelems) | set(t.interiorelems)) == list(range(1, nreal + 1))
    assert sorted(set(t.sendelems.tolist())) == t.exteriorelems.tolist()
    assert not (set(t.exteriorelems) & set(t.interiorelems))


def test_mpi_connect_golden():
    """mpi_connect.jl: 2-D brick, 3 ranks, :face connectivity."""
    topos = tp.BrickTopology(3, (np.arange(0, 5), np.arange(5, 10)),
                             boundary=((1, 2), (3, 4)), periodicity=(False, True),
                             connectivity="face")
    exp = [
        (5, [1, 2, 3, 4, 5, 6, 7, 8, 14, 15], M("""
            1 1 4 2 3 4 7 8 9 10
            2 10 9 3 8 6 7 8 9 10
            6 7 2 1 4 6 7 8 9 10
            4 3 8 5 6 6 7 8 9 10"""), [1, 2], [(1, 3), (4, 5)], [(1, 4), (5, 6)]),
        (5, [6, 7, 8, 9, 10, 1, 2, 3, 5, 11, 12, 14, 15], M("""
            1 1 9 3 2 2 7 8 3 10 11 12 13
            2 5 4 11 10 6 7 8 9 1 2 12 13
            9 3 8 12 4 6 7 8 9 10 11 12 13
            6 7 2 5 13 6 7 8 9 10 11 12 13"""), [0, 2], [(1, 4), (5, 8)], [(1, 3), (4, 5)]),
        (6, [11, 12, 13, 14, 15, 16, 2, 3, 9, 10], M("""
            10 9 4 8 7 5 7 8 9 10
            1 2 3 3 6 4 7 8 9 10
            2 3 6 5 10 1 7 8 9 10
            6 1 2 9 4 3 7 8 9 10"""), [0, 1], [(1, 2), (3, 4)], [(1, 2), (3, 6)]),
    ]
    for t, (nreal, ge, e2e, n2r, rr, ss) in zip(topos, exp):
        _check_common(t, nreal, ge, GLOBAL_2D_COORD, GLOBAL_2D_FACE, GLOBAL_2D_BNDY, e2e, n2r, rr, ss)


def test_mpi_connectfull_golden():
    """mpi_connectfull.jl: same mesh, :full (vertex) connectivity."""
    topos = tp.BrickTopology(3, (np.arange(0, 5), np.arange(5, 10)),
                             boundary=((1, 2), (3, 4)), periodicity=(False, True),
                             connectivity="full")
    exp = [
        (5, [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 14, 15], M("""
            1 1 4 2 3 4 6 5 8 7 3 2
            2 12 11 3 8 7 10 9 9 10 11 12
            6 7 2 1 4 5 8 3 11 9 12 10
            4 3 8 5 6 1 2 7 10 12 9 11"""), M("""
            1 2 2 1 1 1 2 2 2 2 2 2
            1 1 1 1 1 1 1 1 2 2 2 2
            4 4 4 4 4 4 4 4 4 4 4 4
            3 3 3 3 3 3 3 3 3 3 3 3"""), [1, 2], [(1, 5), (6, 7)], [(1, 5), (6, 7)]),
        (5, [6, 7, 8, 9, 10, 1, 2, 3, 4, 5, 11, 12, 13, 14, 15, 16], M("""
            1 1 10 3 2 2 6 9 3 4 5 4 14 8 7 15
            2 5 4 12 11 7 15 14 8 3 1 2 3 13 16 4
            10 3 8 14 4 1 2 7 6 9 12 13 16 15 5 11
            6 7 2 5 15 9 8 3 10 1 16 11 12 4 14 13"""), M("""
            1 2 2 2 2 1 2 2 1 1 2 2 2 2 2 2
            1 1 1 1 1 1 1 1 1 1 2 2 2 1 1 2
            4 4 4 4 4 4 4 4 4 4 4 4 4 4 4 4
            3 3 3 3 3 3 3 3 3 3 3 3 3 3 3 3"""), [0, 2], [(1, 5), (6, 11)], [(1, 5), (6, 9)]),
        (6, [11, 12, 13, 14, 15, 16, 2, 3, 7, 8, 9, 10], M("""
            12 11 4 8 7 5 7 8 9 10 10 9
            1 2 3 3 6 4 5 4 12 11 2 1
            2 3 6 5 12 1 9 7 10 8 4 11
            6 1 2 11 4 3 8 10 7 9 12 5"""), M("""
            2 2 2 2 2 2 1 1 1 1 2 2
            2 2 2 1 1 2 1 1 1 1 1 1
            4 4 4 4 4 4 4 4 4 4 4 4
            3 3 3 3 3 3 3 3 3 3 3 3"""), [0, 1], [(1, 2), (3, 6)], [(1, 2), (3, 8)]),
    ]
    for t, (nreal, ge, e2e, e2f, n2r, rr, ss) in zip(topos, exp):
        _check_common(t, nreal, ge, GLOBAL_2D_COORD, GLOBAL_2D_FACE, GLOBAL_2D_BNDY, e2e, n2r, rr, ss)
        assert np.array_equal(t.elemtoface, e2f)


def test_mpi_connect_stacked_golden():
    """mpi_connect_stacked.jl: 2-D stacked brick on 3 ranks."""
    topos = tp.StackedBrickTopology(3, (np.arange(2, 6), np.arange(4, 7)),
                                    periodicity=(False, True), boundary=((1, 2), (3, 4)),
                                    connectivity="face")
    gcoord = {1: [[2, 3, 2, 3], [4, 4, 5, 5]], 2: [[2, 3, 2, 3], [5, 5, 6, 6]],
              3: [[3, 4, 3, 4], [4, 4, 5, 5]], 4: [[3, 4, 3, 4], [5, 5, 6, 6]],
              5: [[4, 5, 4, 5], [4, 4, 5, 5]], 6: [[4, 5, 4, 5], [5, 5, 6, 6]]}
    gface = M("""
        1 1 2 2 2 2
        1 1 1 1 2 2
        4 4 4 4 4 4
        3 3 3 3 3 3""")
    gbndy = M("""
        1 1 0 0 0 0
        0 0 0 0 2 2
        0 0 0 0 0 0
        0 0 0 0 0 0""")
    exp = [
        (2, [1, 2, 3, 4], M("""
            1 2 3 4
            3 4 3 4
            2 1 3 4
            2 1 3 4"""), [1], [(1, 2)], [(1, 2)]),
        (2, [3, 4, 1, 2, 5, 6], M("""
            3 4 1 2 5 6
            5 6 3 4 1 2
            2 1 3 4 5 6
            2 1 3 4 5 6"""), [0, 2], [(1, 2), (3, 4)], [(1, 2), (3, 4)]),
        (2, [5, 6, 3, 4], M("""
            3 4 3 4
            1 2 3 4
            2 1 3 4
            2 1 3 4"""), [1], [(1, 2)], [(1, 2)]),
    ]
    for t, (nreal, ge, e2e, n2r, rr, ss) in zip(topos, exp):
        _check_common(t, nreal, ge, gcoord, gface, gbndy, e2e, n2r, rr, ss)


def test_mpi_connect_stacked_3d_golden():
    """mpi_connect_stacked_3d.jl: 3x3x3 stacked brick on 2 ranks."""
    topos = tp.StackedBrickTopology(
        2, (np.arange(1, 5), np.arange(5, 9), np.arange(9, 13)),
        periodicity=(False, True, False), boundary=((1, 2), (3, 4), (5, 6)),
        connectivity="face")
    # global element numbering of the test: horizontal column c = 1..9, 3 levels each
    colxy = [(1, 5), (2, 5), (2, 6), (1, 6), (1, 7), (2, 7), (3, 7), (3, 6), (3, 5)]
    gcoord = {}
    for c, (x, y) in enumerate(colxy):
        for l in range(3):
            z = 9 + l
            gcoord[3 * c + l + 1] = [[x, x + 1] * 4, [y, y, y + 1, y + 1] * 2, [z] * 4 + [z + 1] * 4]
    gface = M("""
        1 1 1 2 2 2 2 2 2 1 1 1 1 1 1 2 2 2 2 2 2 2 2 2 2 2 2
        1 1 1 1 1 1 1 1 1 1 1 1 1 1 1 1 1 1 2 2 2 2 2 2 2 2 2
        4 4 4 4 4 4 4 4 4 4 4 4 4 4 4 4 4 4 4 4 4 4 4 4 4 4 4
        3 3 3 3 3 3 3 3 3 3 3 3 3 3 3 3 3 3 3 3 3 3 3 3 3 3 3
        5 6 6 5 6 6 5 6 6 5 6 6 5 6 6 5 6 6 5 6 6 5 6 6 5 6 6
        5 5 6 5 5 6 5 5 6 5 5 6 5 5 6 5 5 6 5 5 6 5 5 6 5 5 6""")
    gbndy = M("""
        1 1 1 0 0 0 0 0 0 1 1 1 1 1 1 0 0 0 0 0 0 0 0 0 0 0 0
        0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 2 2 2 2 2 2 2 2 2
        0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0
        0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0
        5 0 0 5 0 0 5 0 0 5 0 0 5 0 0 5 0 0 5 0 0 5 0 0 5 0 0
        0 0 6 0 0 6 0 0 6 0 0 6 0 0 6 0 0 6 0 0 6 0 0 6 0 0 6""")
    exp = [
        (12, list(range(1, 19)) + list(range(22, 28)), M("""
            1 2 3 1 2 3 10 11 12 4 5 6 7 8 9 16 17 18 19 20 21 22 23 24
            4 5 6 22 23 24 19 20 21 7 8 9 13 14 15 16 17 18 1 2 3 4 5 6
            13 14 15 16 17 18 4 5 6 1 2 3 13 14 15 16 17 18 19 20 21 22 23 24
            10 11 12 7 8 9 16 17 18 13 14 15 13 14 15 16 17 18 19 20 21 22 23 24
            1 1 2 2 4 5 3 7 8 4 10 11 5 14 15 6 17 18 7 20 21 8 23 24
            2 3 1 5 6 2 8 9 3 11 12 4 13 14 5 16 17 6 19 20 7 22 23 8"""),
         [1], [(1, 12)], [(1, 12)]),
        (15, list(range(13, 28)) + list(range(1, 13)), M("""
            1 2 3 1 2 3 4 5 6 22 23 24 19 20 21 4 5 6 19 20 21 22 23 24 7 8 9
            4 5 6 7 8 9 1 2 3 4 5 6 7 8 9 16 17 18 19 20 21 22 23 24 25 26 27
            25 26 27 22 23 24 10 11 12 13 14 15 7 8 9 16 17 18 19 20 21 22 23 24 25 26 27
            16 17 18 19 20 21 13 14 15 7 8 9 10 11 12 16 17 18 19 20 21 22 23 24 25 26 27
            1 1 2 2 4 5 3 7 8 4 10 11 5 13 14 6 17 18 7 20 21 8 23 24 9 26 27
            2 3 1 5 6 2 8 9 3 11 12 4 14 15 5 16 17 6 19 20 7 22 23 8 25 26 9"""),
         [0], [(1, 12)], [(1, 12)]),
    ]
    for t, (nreal, ge, e2e, n2r, rr, ss) in zip(topos, exp):
        _check_common(t, nreal, ge, gcoord, gface, gbndy, e2e, n2r, rr, ss)


def test_hilbertcode_known():
    # BrickMesh.jl docstring examples / test/Numerics/Mesh/BrickMesh.jl hilbert tests
    assert bm.hilbertcode([0, 0], bits=1) == [0, 0]
    assert bm.hilbertcode([0, 1], bits=1) == [0, 1]
    assert bm.hilbertcode([1, 1], bits=1) == [1, 0]
    assert bm.hilbertcode([1, 0], bits=1) == [1, 1]
    assert bm.hilbertcode([0, 0], bits=2) == [0, 0]
    assert bm.hilbertcode([1, 0], bits=2) == [0, 1]
    assert bm.hilbertcode([1, 1], bits=2) == [0, 2]
    assert bm.hilbertcode([0, 1], bits=2) == [0, 3]
    assert bm.hilbertcode([0, 2], bits=2) == [1, 0]
    assert bm.hilbertcode([0, 3], bits=2) == [1, 1]
    assert bm.hilbertcode([1, 3], bits=2) == [1, 2]
    assert bm.hilbertcode([1, 2], bits=2) == [1, 3]
    assert bm.hilbertcode([2, 2], bits=2) == [2, 0]
    assert bm.hilbertcode([2, 3], bits=2) == [2, 1]
    assert bm.hilbertcode([3, 3], bits=2) == [2, 2]
    assert bm.hilbertcode([3, 2], bits=2) == [2, 3]
    assert bm.hilbertcode([3, 1], bits=2) == [3, 0]
    assert bm.hilbertcode([2, 1], bits=2) == [3, 1]
    assert bm.hilbertcode([2, 0], bits=2) == [3, 2]
    assert bm.hilbertcode([3, 0], bits=2) == [3, 3]


def test_lgl_and_derivative_exactness():
    """test/Numerics/Mesh/Elements.jl 'Operators': P6' = D P6 on LGL(6) points etc."""
    for N in range(1, 9):
        r, w = elements.lglpoints(np.float64, N)
        assert abs(w.sum() - 2) < 1e-14
        assert r[0] == -1 and r[-1] == 1
        # exact for polynomials of degree <= 2N-1
        for p in range(0, 2 * N):
            exact = (1 - (-1) ** (p + 1)) / (p + 1)
            assert abs(np.dot(w, r ** p) - exact) < 1e-13
        D = elements.spectralderivative(r)
        for p in range(1, N + 1):
            assert np.allclose(D @ r ** p, p * r ** (p - 1), atol=1e-12)
    r, w = elements.lglpoints(np.float64, 4)
    assert np.allclose(r, [-1, -np.sqrt(3 / 7), 0, np.sqrt(3 / 7), 1], atol=1e-16)
    assert np.allclose(w, [0.1, 49 / 90, 32 / 45, 49 / 90, 0.1], atol=1e-16)
    r6, _ = elements.lglpoints(np.float64, 6)
    P6 = (-5 + 105 * r6 ** 2 - 315 * r6 ** 4 + 231 * r6 ** 6) / 16
    DP6 = (210 * r6 - 1260 * r6 ** 3 + 1386 * r6 ** 5) / 16
    assert np.allclose(elements.spectralderivative(r6) @ P6, DP6, atol=1e-12)


def test_indefinite_integral_matrix():
    r, w = elements.lglpoints(np.float64, 4)
    I = elements.indefinite_integral_interpolation_matrix(r, w)
    for p in range(0, 4):
        exact = (r ** (p + 1) - (-1.0) ** (p + 1)) / (p + 1)
        assert np.allclose(I @ r ** p, exact, atol=1e-13)


def test_grid_mass_and_metrics_box():
    """test/Numerics/Mesh/Grids.jl: mass matrix sums to the volume, surface mass to the area."""
    t = tp.BrickTopology(1, (np.linspace(0, 2, 3), np.linspace(-1, 1, 4), np.linspace(0, 3, 3)),
                         periodicity=(True, False, True))[0]
    g = grids.Grid(t, 3)
    assert abs(g.vgeo[:, grids._M, :].sum() - 2 * 2 * 3) < 1e-12
    assert np.allclose(g.vgeo[:, grids._M, :] * g.vgeo[:, grids._MI, :], 1)
    # constant metrics of an affine brick: dxi1/dx1 = 2/h1
    assert np.allclose(g.vgeo[:, grids._xi1x1, :], 2 / 1.0)
    assert np.allclose(g.vgeo[:, grids._xi2x2, :], 2 / (2 / 3))
    assert np.allclose(g.vgeo[:, grids._xi3x3, :], 2 / 1.5)
    assert np.allclose(g.vgeo[:, grids._xi1x2, :], 0, atol=1e-13)
    # faces: normals +-e_d, surface mass sums to face areas
    assert np.allclose(g.sgeo[:, 0, :, grids._n1], -1)
    assert np.allclose(g.sgeo[:, 5, :, grids._n3], 1)
    nel = t.nelem
    assert abs(g.sgeo[:, 0, :, grids._sM].sum() - nel * (2 / 3) * 1.5) < 1e-12
    # interior faces see the same coordinates from both sides (mod periodic shift)
    x1, x2, x3 = g.coords()
    flat2 = x2.reshape(-1)
    interior = g.elemtobndy == 0
    vm = g.vmapM[interior] - 1
    vp = g.vmapP[interior] - 1
    assert np.allclose(flat2[vm], flat2[vp])


@pytest.mark.parametrize("csize", [1, 3])
def test_cubed_sphere_vmap_coordinates_match(csize):
    """mpi_connect_sphere.jl:63-83: x(vmap-) == x(vmap+) on interior faces, before and
    (for ghosts) after an exchange of the coordinates -- validates orientation flips."""
    from oracle import mpistatearrays as msa
    Nhorz, Nstack, N = 3, 2, 3
    Rrange = np.cumsum(np.arange(1, Nstack + 2)).astype(np.float64)
    topos = tp.StackedCubedSphereTopology(csize, Nhorz, Rrange, boundary=(1, 2),
                                          connectivity="face" if csize == 1 else "full")
    gs = [grids.Grid(t, N, meshwarp=tp.equiangular_cubed_sphere_warp) for t in topos]
    arrs = []
    for g in gs:
        a = msa.MPIStateArray.from_grid(g, 3)
        a.data[:g.nreal] = g.vgeo[:g.nreal, [grids._x1, grids._x2, grids._x3], :]
        arrs.append(a)
    msa.ghost_exchange(arrs)
    for g, a in zip(gs, arrs):
        interior = g.elemtobndy == 0
        interior[g.nreal:] = False
        vm = g.vmapM[interior] - 1
        vp = g.vmapP[interior] - 1
        for c in range(3):
            flat = np.moveaxis(a.data, 1, 0)[c].reshape(-1)
            assert np.allclose(flat[vm], flat[vp], rtol=1e-12, atol=1e-12)
        # radius of warped nodes equals the stack radius range
        r = np.sqrt((a.data[:g.nreal] ** 2).sum(axis=1))
        assert r.min() > Rrange[0] - 1e-12 and r.max() < Rrange[-1] + 1e-12


def test_cubed_sphere_surface_area():
    Rrange = np.array([1.0, 1.5])
    t = tp.StackedCubedSphereTopology(1, 4, Rrange)[0]
    g = grids.Grid(t, 4, meshwarp=tp.equiangular_cubed_sphere_warp)
    bot = g.elemtobndy[:, 4] == 1
    top = g.elemtobndy[:, 5] == 1
    assert abs(g.sgeo[bot, 4, :, grids._sM].sum() / (4 * np.pi) - 1) < 1e-6
    assert abs(g.sgeo[top, 5, :, grids._sM].sum() / (4 * np.pi * 1.5 ** 2) - 1) < 1e-6
    vol = 4 / 3 * np.pi * (1.5 ** 3 - 1)
    assert abs(g.vgeo[:, grids._M, :].sum() / vol - 1) < 1e-6
