"""Mesh topologies (test infrastructure -- see oracle/__init__.py).

Restates ``src/Numerics/Mesh/Topologies.jl`` for a serial emulation of an
``csize``-rank run (each constructor returns a list with one topology per rank):

* ``BoxElementTopology`` fields, interior/exterior split <- ``Topologies.jl:20-260``
* ``BrickTopology``              <- ``Topologies.jl:470-533``
* ``StackedBrickTopology``       <- ``Topologies.jl:632-800`` (DSS tables omitted:
  they feed ``Mesh/DSS.jl`` only, which is off the hot path)
* ``cubedshellmesh``             <- ``Topologies.jl:1170-1224``
* ``CubedShellTopology``         <- ``Topologies.jl:1067-1128``
* ``cubed_sphere_warp`` (equiangular) <- ``Topologies.jl:1254-1299``
* ``StackedCubedSphereTopology`` <- ``Topologies.jl:1620-1790``
* ``equiangular_cubed_sphere_warp`` <- ``Topologies.jl:1301-1310``
"""
import math

import numpy as np

from . import brickmesh as bm


class Topology:
    """Per-rank topology record with the reference's field names."""

    def __init__(self, dim, conn, nb, stacksize=0, periodicstack=False):
        self.dim = dim
        self.nelem = conn.nelem
        self.nreal = conn.nreal
        self.nghost = conn.nghost
        self.elems = (1, conn.nelem)
        self.realelems = (1, conn.nreal)
        self.ghostelems = (conn.nreal + 1, conn.nelem)
        self.ghostfaces = conn.ghostfaces
        self.sendelems = conn.sendelems
        self.sendfaces = conn.sendfaces
        self.elemtocoord = conn.elemtocoord
        self.elemtoelem = conn.elemtoelem
        self.elemtoface = conn.elemtoface
        self.elemtoordr = conn.elemtoordr
        self.elemtobndy = conn.elemtobndy
        self.nabrtorank = conn.nabrtorank
        self.nabrtorecv = conn.nabrtorecv
        self.nabrtosend = conn.nabrtosend
        self.nb = nb
        self.stacksize = stacksize
        self.periodicstack = periodicstack
        # Topologies.jl:251-252
        self.exteriorelems = np.array(sorted(set(int(e) for e in conn.sendelems)), dtype=np.int64)
        ext = set(self.exteriorelems.tolist())
        self.interiorelems = np.array(
            [e for e in range(1, conn.nreal + 1) if e not in ext], dtype=np.int64)


def BrickTopology(csize, elemrange, boundary=None, periodicity=None, connectivity="face"):
    d = len(elemrange)
    if boundary is None:
        boundary = tuple((1, 1) for _ in range(d))
    if periodicity is None:
        periodicity = tuple(False for _ in range(d))
    meshes = [bm.brickmesh(elemrange, periodicity, part=r + 1, numparts=csize, boundary=boundary)
              for r in range(csize)]
    parts = bm.partition(csize, meshes)
    m4 = [p[:4] for p in parts]
    if connectivity == "face":
        conns = bm.connectmesh(csize, m4)
    else:
        conns = bm.connectmeshfull(csize, m4)
    topos = []
    for c in conns:
        b2e, _ = bm.enumerateboundaryfaces(c.elemtoelem, c.elemtobndy, periodicity, boundary)
        topos.append(Topology(d, c, len(b2e)))
    return topos


def _stack(basetopos, dim, stack_coords_fn, stacksize, periodic_vert, boundary_vert,
           ordr_map, periodicity_for_enum, boundary_for_enum):
    """Shared extrusion logic of Stacked{Brick,CubedSphere}Topology."""
    nvert = 2 ** dim
    nface = 2 * dim
    nhf = 2 * (dim - 1)
    out = []
    for bt in basetopos:
        nreal = bt.nreal * stacksize
        nghost = bt.nghost * stacksize
        ntot = nreal + nghost
        sendelems = np.zeros(len(bt.sendelems) * stacksize, dtype=np.int64)
        for i in range(len(bt.sendelems)):
            for j in range(stacksize):
                sendelems[stacksize * i + j] = stacksize * (bt.sendelems[i] - 1) + j + 1
        ghostfaces = np.zeros((nface, nghost), dtype=bool)
        for i in range(bt.nghost):
            for j in range(stacksize):
                ghostfaces[:nhf, stacksize * i + j] = bt.ghostfaces[:nhf, i]
        sendfaces = np.zeros((nface, len(sendelems)), dtype=bool)
        for i in range(len(bt.sendelems)):
            for j in range(stacksize):
                sendfaces[:nhf, stacksize * i + j] = bt.sendfaces[:nhf, i]
        elemtocoord = stack_coords_fn(bt)
        elemtoelem = np.tile(np.arange(1, ntot + 1, dtype=np.int64), (nface, 1))
        elemtoface = np.tile(np.arange(1, nface + 1, dtype=np.int64)[:, None], (1, ntot))
        elemtoordr = np.ones((nface, ntot), dtype=np.int64)
        elemtobndy = np.zeros((nface, ntot), dtype=np.int64)
        for i in range(1, bt.nreal + 1):
            for j in range(1, stacksize + 1):
                e1 = stacksize * (i - 1) + j
                for f in range(nhf):
                    e2 = stacksize * (bt.elemtoelem[f, i - 1] - 1) + j
                    elemtoelem[f, e1 - 1] = e2
                    elemtoface[f, e1 - 1] = bt.elemtoface[f, i - 1]
                    elemtoordr[f, e1 - 1] = ordr_map(int(bt.elemtoordr[f, i - 1]))
                et = e1 + 1
                eb = e1 - 1
                f_of_top_nbr = nhf + 1      # neighbour above shows its bottom face
                f_of_bot_nbr = nhf + 2
                if j == stacksize:
                    et = stacksize * (i - 1) + 1 if periodic_vert else e1
                    f_of_top_nbr = f_of_top_nbr if periodic_vert else nhf + 2
                if j == 1:
                    eb = stacksize * (i - 1) + stacksize if periodic_vert else e1
                    f_of_bot_nbr = f_of_bot_nbr if periodic_vert else nhf + 1
                elemtoelem[nhf, e1 - 1] = eb
                elemtoelem[nhf + 1, e1 - 1] = et
                elemtoface[nhf, e1 - 1] = f_of_bot_nbr
                elemtoface[nhf + 1, e1 - 1] = f_of_top_nbr
        for i in range(1, bt.nelem + 1):
            for j in range(1, stacksize + 1):
                e1 = stacksize * (i - 1) + j
                elemtobndy[:nhf, e1 - 1] = bt.elemtobndy[:nhf, i - 1]
                bb = bt_ = 0
                if j == stacksize and not periodic_vert:
                    bt_ = boundary_vert[1]
                if j == 1 and not periodic_vert:
                    bb = boundary_vert[0]
                elemtobndy[nhf, e1 - 1] = bb
                elemtobndy[nhf + 1, e1 - 1] = bt_
        nabrtorecv = [(stacksize * (a - 1) + 1, stacksize * b) for a, b in bt.nabrtorecv]
        nabrtosend = [(stacksize * (a - 1) + 1, stacksize * b) for a, b in bt.nabrtosend]
        b2e, _ = bm.enumerateboundaryfaces(elemtoelem, elemtobndy,
                                           periodicity_for_enum, boundary_for_enum)
        conn = bm.Connected(
            nelem=ntot, nreal=nreal, nghost=nghost, ghostfaces=ghostfaces,
            sendelems=sendelems, sendfaces=sendfaces, elemtocoord=elemtocoord,
            elemtovert=None, elemtoelem=elemtoelem, elemtoface=elemtoface,
            elemtoordr=elemtoordr, elemtobndy=elemtobndy, nabrtorank=bt.nabrtorank,
            nabrtorecv=nabrtorecv, nabrtosend=nabrtosend)
        out.append(Topology(dim, conn, len(b2e), stacksize=stacksize, periodicstack=periodic_vert))
    return out


def StackedBrickTopology(csize, elemrange, boundary=None, periodicity=None, connectivity="full"):
    dim = len(elemrange)
    assert dim >= 2
    if boundary is None:
        boundary = tuple((1, 1) for _ in range(dim))
    if periodicity is None:
        periodicity = tuple(False for _ in range(dim))
    if dim == 2 and connectivity == "full":
        # 1-D base meshes have no vertex-only neighbours; the reference's
        # connectmeshfull asserts dim == 2, so 2-D stacked bricks use :face.
        connectivity = "face"
    base = BrickTopology(csize, elemrange[:dim - 1], boundary=boundary[:dim - 1],
                         periodicity=periodicity[:dim - 1], connectivity=connectivity)
    stack = np.asarray(elemrange[dim - 1])
    stacksize = len(stack) - 1
    nvert = 2 ** dim
    nbv = 2 ** (dim - 1)

    def coords(bt):
        T = np.result_type(bt.elemtocoord.dtype, stack.dtype)
        ec = np.zeros((dim, nvert, bt.nelem * stacksize), dtype=T)
        for i in range(bt.nelem):
            for j in range(stacksize):
                e = stacksize * i + j
                for v in range(nbv):
                    ec[:dim - 1, v, e] = bt.elemtocoord[:dim - 1, v, i]
                    ec[:dim - 1, nbv + v, e] = bt.elemtocoord[:dim - 1, v, i]
                    ec[dim - 1, v, e] = stack[j]
                    ec[dim - 1, nbv + v, e] = stack[j + 1]
        return ec

    def ordr(o):
        assert o == 1
        return o

    return _stack(base, dim, coords, stacksize, periodicity[dim - 1], boundary[dim - 1],
                  ordr, periodicity, boundary)


def cubedshellmesh(Ne, part=1, numparts=1):
    """``Topologies.jl:1170-1224``: flattened-cross cubed shell, 2-D elements in 3-D."""
    nglob = 6 * Ne * Ne
    first, last = bm.linearpartition(nglob, part, numparts)
    nloc = last - first + 1
    elemtovert = np.zeros((4, nloc), dtype=np.int64)
    elemtocoord = np.zeros((2, 4, nloc), dtype=np.int64)
    bx = [0, Ne, 2 * Ne, Ne, Ne, Ne]
    by = [0, 0, 0, Ne, 2 * Ne, 3 * Ne]

    def vertmap(a, b, c):  # 1-based args -> 1-based linear index
        return (a - 1) + (Ne + 1) * ((b - 1) + (Ne + 1) * (c - 1)) + 1

    for le in range(nloc):
        e = first - 1 + le
        i = e % Ne + 1
        j = (e // Ne) % Ne + 1
        blck = e // (Ne * Ne) + 1
        elemtocoord[0, :, le] = bx[blck - 1] + np.array([i - 1, i, i - 1, i])
        elemtocoord[1, :, le] = by[blck - 1] + np.array([j - 1, j - 1, j, j])
        for n in range(1, 5):
            ix = i + (n - 1) % 2
            jx = j + (n - 1) // 2
            if blck == 1:
                v = vertmap(1, Ne + 2 - ix, jx)
            elif blck == 2:
                v = vertmap(ix, 1, jx)
            elif blck == 3:
                v = vertmap(Ne + 1, ix, jx)
            elif blck == 4:
                v = vertmap(ix, jx, Ne + 1)
            elif blck == 5:
                v = vertmap(ix, Ne + 1, Ne + 2 - jx)
            else:
                v = vertmap(ix, Ne + 2 - jx, 1)
            elemtovert[n - 1, le] = v
    elemtobndy = np.zeros((4, nloc), dtype=np.int64)
    return elemtovert, elemtocoord, elemtobndy, [], np.arange(first, last + 1)


def CubedShellTopology(csize, Neside, T=np.float64, connectivity="full"):
    meshes = [cubedshellmesh(Neside, part=r + 1, numparts=csize) for r in range(csize)]
    parts = bm.partition(csize, [m[:4] for m in meshes], globords=[m[4] for m in meshes])
    m4 = []
    for p in parts:
        elemtovert = p[0]
        nelem = elemtovert.shape[1]
        ec = np.zeros((3, 4, nelem), dtype=T)
        for e in range(nelem):
            for n in range(4):
                v = int(elemtovert[n, e]) - 1
                i = v % (Neside + 1)
                j = (v // (Neside + 1)) % (Neside + 1)
                k = v // ((Neside + 1) ** 2)
                ec[:, n, e] = (2 * np.array([i, j, k]) - Neside) / Neside
        m4.append((elemtovert, ec, p[2], p[3]))
    if connectivity == "face":
        conns = bm.connectmesh(csize, m4, dim=2)
    else:
        conns = bm.connectmeshfull(csize, m4, dim=2)
    return [Topology(2, c, 0) for c in conns]


def cubed_sphere_warp(a, b, c, R=None):
    """Equiangular gnomonic warp (``Topologies.jl:1254-1299``), scalar version."""
    if R is None:
        R = max(abs(a), abs(b), abs(c))

    def f(sR, xi, eta):
        X, Y = math.tan(math.pi * xi / 4), math.tan(math.pi * eta / 4)
        z1 = sR / math.sqrt(X * X + Y * Y + 1)
        return z1, X * z1, Y * z1

    absv = (abs(a), abs(b), abs(c))
    fdim = int(np.argmax(absv)) + 1
    if fdim == 1 and a < 0:
        x1, x2, x3 = f(-R, b / a, c / a)
    elif fdim == 2 and b < 0:
        x2, x1, x3 = f(-R, a / b, c / b)
    elif fdim == 1 and a > 0:
        x1, x2, x3 = f(R, b / a, c / a)
    elif fdim == 2 and b > 0:
        x2, x1, x3 = f(R, a / b, c / b)
    elif fdim == 3 and c > 0:
        x3, x2, x1 = f(R, b / c, a / c)
    elif fdim == 3 and c < 0:
        x3, x2, x1 = f(-R, b / c, a / c)
    else:
        raise ValueError("invalid case for cubed_sphere_warp")
    return x1, x2, x3


def cubed_sphere_warp_vec(a, b, c):
    """Vectorised ``cubed_sphere_warp`` with ``R = max(|a|,|b|,|c|)``; the same
    branch order (and therefore the same tie-breaking: ``argmax`` takes the
    first maximum) as the scalar version."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64)
    absv = np.stack([np.abs(a), np.abs(b), np.abs(c)])
    R = absv.max(axis=0)
    fdim = np.argmax(absv, axis=0)
    x1 = np.empty_like(a)
    x2 = np.empty_like(a)
    x3 = np.empty_like(a)

    def f(sR, xi, eta):
        X, Y = np.tan(np.pi * xi / 4), np.tan(np.pi * eta / 4)
        z1 = sR / np.sqrt(X * X + Y * Y + 1)
        return z1, X * z1, Y * z1

    with np.errstate(divide="ignore", invalid="ignore"):
        for d, sgn in ((0, -1), (1, -1), (0, 1), (1, 1), (2, 1), (2, -1)):
            comp = (a, b, c)[d]
            m = (fdim == d) & ((comp < 0) if sgn < 0 else (comp > 0))
            if not m.any():
                continue
            am, bm_, cm, Rm = a[m], b[m], c[m], R[m]
            if d == 0:
                z1, z2, z3 = f(sgn * Rm, bm_ / am, cm / am)
                x1[m], x2[m], x3[m] = z1, z2, z3
            elif d == 1:
                z1, z2, z3 = f(sgn * Rm, am / bm_, cm / bm_)
                x2[m], x1[m], x3[m] = z1, z2, z3
            else:
                z1, z2, z3 = f(sgn * Rm, bm_ / cm, am / cm)
                x3[m], x2[m], x1[m] = z1, z2, z3
    return x1, x2, x3


equiangular_cubed_sphere_warp = cubed_sphere_warp_vec


def StackedCubedSphereTopology(csize, Nhorz, Rrange, boundary=(1, 1), connectivity="full"):
    Rrange = np.asarray(Rrange, dtype=np.float64)
    base = CubedShellTopology(csize, Nhorz, Rrange.dtype, connectivity=connectivity)
    dim = 3
    stacksize = len(Rrange) - 1

    def coords(bt):
        ec = np.zeros((3, 8, bt.nelem * stacksize), dtype=Rrange.dtype)
        for i in range(bt.nelem):
            for j in range(stacksize):
                e = stacksize * i + j
                ec[:, :4, e] = bt.elemtocoord[:, :, i] * Rrange[j]
                ec[:, 4:, e] = bt.elemtocoord[:, :, i] * Rrange[j + 1]
        return ec

    def ordr(o):
        assert o in (1, 2)
        return 1 if o == 1 else 3

    return _stack(base, dim, coords, stacksize, False, boundary, ordr, (False,), (boundary,))
