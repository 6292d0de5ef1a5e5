"""DiscontinuousSpectralElementGrid (test infrastructure -- see oracle/__init__.py).

Restates, for 3-D hexahedral elements:

* ``vgeo``/``sgeo`` column ids     <- ``Grids.jl:76-146``, ``GeometricFactors.jl:18-67``
* ``mappings`` (vmap-/vmap+)       <- ``Grids.jl:559-637``
* ``commmapping`` (vmapsend/recv)  <- ``Grids.jl:761-811``
* ``creategrid!`` (trilinear)      <- ``Metrics.jl:85-114``
* ``compute_reference_to_physical_coord_jacobian!`` <- ``Metrics.jl:215-264``
* ``computemetric!`` (curl-invariant, Kopriva 2006) <- ``Metrics.jl:431-722``
* ``computegeometry`` weights, vMI, sM, MH <- ``Grids.jl:1028-1154``
* grid constructor                 <- ``Grids.jl:267-413``

Byte layout (see oracle/__init__.py): ``vgeo`` has NumPy shape
``(nelem, 25, Np)`` == Julia ``Np x 25 x nelem``; ``sgeo`` has shape
``(nelem, 6, Nfp, 5)`` == Julia ``5 x Nfp x 6 x nelem``; ``vmapM``/``vmapP``
have shape ``(nelem, 6, Nfp)`` == Julia ``Nfp x 6 x nelem`` and hold 1-based
Int64 linear node ids ``Np*(e-1)+n``; ``elemtobndy`` has shape ``(nelem, 6)``.
"""
import numpy as np

from . import elements

# 0-based column ids into vgeo (reference ids are these + 1)
(_xi1x1, _xi2x1, _xi3x1, _xi1x2, _xi2x2, _xi3x2, _xi1x3, _xi2x3, _xi3x3,
 _M, _MI, _MH, _x1, _x2, _x3, _JcV) = range(16)
_nvgeo = 25  # VolumeGeometry.array carries 9 extra dx/dxi columns (GeometricFactors.jl:60-67)
_n1, _n2, _n3, _sM, _vMI = range(5)
_nsgeo = 5


def mappings(N, elemtoelem, elemtoface, elemtoordr):
    """vmap-/vmap+ for 3-D elements; returns arrays of shape (nelem, 6, Nfp)."""
    nfaces, nelem = elemtoelem.shape
    assert nfaces == 6
    Nq = [n + 1 for n in N]
    Np = Nq[0] * Nq[1] * Nq[2]
    p = np.arange(1, Np + 1).reshape(Nq, order="F")
    fmask = [
        p[0, :, :].ravel(order="F"), p[Nq[0] - 1, :, :].ravel(order="F"),
        p[:, 0, :].ravel(order="F"), p[:, Nq[1] - 1, :].ravel(order="F"),
        p[:, :, 0].ravel(order="F"), p[:, :, Nq[2] - 1].ravel(order="F"),
    ]
    Nfp = [Np // q for q in Nq]
    # inds[f][a, b] -> 0-based linear face dof number
    inds = []
    for f in range(6):
        d = f // 2
        shp = [Nq[j] for j in range(3) if j != d]
        inds.append(np.arange(shp[0] * shp[1]).reshape(shp, order="F"))
    maxNfp = max(Nfp)
    vmapM = np.zeros((nelem, 6, maxNfp), dtype=np.int64)
    vmapP = np.zeros((nelem, 6, maxNfp), dtype=np.int64)
    for e1 in range(nelem):
        for f1 in range(6):
            e2 = int(elemtoelem[f1, e1])
            f2 = int(elemtoface[f1, e1]) - 1
            o2 = int(elemtoordr[f1, e1])
            d1, d2 = f1 // 2, f2 // 2
            assert Nfp[d1] == Nfp[d2]
            n = Nfp[d1]
            vmapM[e1, f1, :n] = Np * e1 + fmask[f1][:n]
            if o2 == 1:
                vmapP[e1, f1, :n] = Np * (e2 - 1) + fmask[f2][:n]
            elif o2 == 3:
                flip = inds[f2][::-1, :].ravel(order="F")
                vmapP[e1, f1, :n] = Np * (e2 - 1) + fmask[f2][flip]
            else:
                raise NotImplementedError(f"Orientation '{o2}' with dim 3 not supported yet")
    return vmapM, vmapP


def commmapping(N, commelems, commfaces, nabrtocomm):
    """Linear node ids to communicate (``Grids.jl:761-811``).

    ``commelems``: 1-based element ids, ``commfaces``: bool ``(nface, ncomm)``,
    ``nabrtocomm``: list of inclusive 1-based ranges into ``commelems``.
    Returns ``(vmapC, nabrtovmapC)`` (1-based ids; inclusive 1-based ranges).
    """
    nface = commfaces.shape[0]
    d = nface // 2
    Nq = [n + 1 for n in N]
    Np = int(np.prod(Nq))
    vmapC = []
    nabrtovmapC = []
    e = 0
    # node multi-indices in linear order (first index fastest)
    ci = np.stack(np.unravel_index(np.arange(Np), Nq, order="F"), axis=1)
    for (a, b) in nabrtocomm:
        rbegin = len(vmapC) + 1
        for ne in range(a, b + 1):
            ce = int(commelems[ne - 1])
            add = np.zeros(Np, dtype=bool)
            for j in range(d):
                if commfaces[2 * j, e]:
                    add |= ci[:, j] == 0
                if commfaces[2 * j + 1, e]:
                    add |= ci[:, j] == Nq[j] - 1
            vmapC.extend(((ce - 1) * Np + 1 + np.nonzero(add)[0]).tolist())
            e += 1
        nabrtovmapC.append((rbegin, len(vmapC)))
    return np.array(vmapC, dtype=np.int64), nabrtovmapC


def _contract(D, x, axis):
    """sum_n D[a, n] x[..n..] along reference axis ``axis`` (0 = xi1 fastest).

    ``x`` has shape (nelem, Nq3, Nq2, Nq1); reference axis j is numpy axis 3-j.
    Accumulates in the same n = 1..Nq order as the reference loops.
    """
    ax = 3 - axis
    xm = np.moveaxis(x, ax, -1)  # (..., n)
    out = np.zeros(xm.shape, dtype=x.dtype)
    for n in range(D.shape[1]):
        out += D[:, n] * xm[..., n:n + 1]
    return np.moveaxis(out, -1, ax)


def computegeometry(elemtocoord, D, xi, w, meshwarp=None, FT=np.float64):
    """vgeo (nelem, 25, Np) and sgeo (nelem, 6, Nfp, 5) as the reference builds them."""
    FT = np.dtype(FT).type
    d, nvert, nelem = elemtocoord.shape
    assert d == 3
    Nq = [len(x) for x in xi]
    Np = Nq[0] * Nq[1] * Nq[2]
    Nfp = [Np // q for q in Nq]
    assert Nfp[0] == Nfp[1] == Nfp[2], "mixed polynomial order not restated"
    e2c = elemtocoord.astype(FT)
    x1r, x2r, x3r = [np.asarray(x, dtype=FT) for x in xi]
    # a) trilinear blend, arrays shaped (nelem, k, j, i)
    r = x1r[None, None, None, :]
    s = x2r[None, None, :, None]
    t = x3r[None, :, None, None]
    X = []
    for n in range(3):
        c = [e2c[n, v, :][:, None, None, None] for v in range(8)]
        X.append(((1 - r) * (1 - s) * (1 - t) * c[0] + (1 + r) * (1 - s) * (1 - t) * c[1]
                  + (1 - r) * (1 + s) * (1 - t) * c[2] + (1 + r) * (1 + s) * (1 - t) * c[3]
                  + (1 - r) * (1 - s) * (1 + t) * c[4] + (1 + r) * (1 - s) * (1 + t) * c[5]
                  + (1 - r) * (1 + s) * (1 + t) * c[6] + (1 + r) * (1 + s) * (1 + t) * c[7]) / 8)
    x1, x2, x3 = X
    # b) warp
    if meshwarp is not None:
        shp = x1.shape
        a, b, c = meshwarp(x1.ravel(), x2.ravel(), x3.ravel())
        x1 = np.asarray(a, dtype=FT).reshape(shp)
        x2 = np.asarray(b, dtype=FT).reshape(shp)
        x3 = np.asarray(c, dtype=FT).reshape(shp)
    D1, D2, D3 = [np.asarray(Dj, dtype=FT) for Dj in D]
    # c) dx/dxi
    x1r_, x2r_, x3r_ = _contract(D1, x1, 0), _contract(D1, x2, 0), _contract(D1, x3, 0)
    x1s_, x2s_, x3s_ = _contract(D2, x1, 1), _contract(D2, x2, 1), _contract(D2, x3, 1)
    x1t_, x2t_, x3t_ = _contract(D3, x1, 2), _contract(D3, x2, 2), _contract(D3, x3, 2)
    # d) metric terms
    JcV = np.sqrt(x1t_ ** 2 + x2t_ ** 2 + x3t_ ** 2)  # hypot
    J = (x1r_ * (x2s_ * x3t_ - x3s_ * x2t_) + x2r_ * (x3s_ * x1t_ - x1s_ * x3t_)
         + x3r_ * (x1s_ * x2t_ - x2s_ * x1t_))
    JI2 = 1 / (2 * J)
    yzr = x2 * x3r_ - x3 * x2r_
    yzs = x2 * x3s_ - x3 * x2s_
    yzt = x2 * x3t_ - x3 * x2t_
    zxr = x3 * x1r_ - x1 * x3r_
    zxs = x3 * x1s_ - x1 * x3s_
    zxt = x3 * x1t_ - x1 * x3t_
    xyr = x1 * x2r_ - x2 * x1r_
    xys = x1 * x2s_ - x2 * x1s_
    xyt = x1 * x2t_ - x2 * x1t_
    # accumulation order of Metrics.jl:548-573: D1 terms, then D2, then D3
    xi2x1 = -_contract(D1, yzt, 0)
    xi3x1 = _contract(D1, yzs, 0)
    xi2x2 = -_contract(D1, zxt, 0)
    xi3x2 = _contract(D1, zxs, 0)
    xi2x3 = -_contract(D1, xyt, 0)
    xi3x3 = _contract(D1, xys, 0)
    xi1x1 = _contract(D2, yzt, 1)
    xi3x1 = xi3x1 - _contract(D2, yzr, 1)
    xi1x2 = _contract(D2, zxt, 1)
    xi3x2 = xi3x2 - _contract(D2, zxr, 1)
    xi1x3 = _contract(D2, xyt, 1)
    xi3x3 = xi3x3 - _contract(D2, xyr, 1)
    xi1x1 = xi1x1 - _contract(D3, yzs, 2)
    xi2x1 = xi2x1 + _contract(D3, yzr, 2)
    xi1x2 = xi1x2 - _contract(D3, zxs, 2)
    xi2x2 = xi2x2 + _contract(D3, zxr, 2)
    xi1x3 = xi1x3 - _contract(D3, xys, 2)
    xi2x3 = xi2x3 + _contract(D3, xyr, 2)
    xi1x1, xi2x1, xi3x1 = xi1x1 * JI2, xi2x1 * JI2, xi3x1 * JI2
    xi1x2, xi2x2, xi3x2 = xi1x2 * JI2, xi2x2 * JI2, xi3x2 * JI2
    xi1x3, xi2x3, xi3x3 = xi1x3 * JI2, xi2x3 * JI2, xi3x3 * JI2
    # inverse of dxi/dx -> dx/dxi columns (vgeo 17..25)
    a11 = xi2x2 * xi3x3 - xi2x3 * xi3x2
    a12 = xi1x3 * xi3x2 - xi1x2 * xi3x3
    a13 = xi1x2 * xi2x3 - xi1x3 * xi2x2
    a21 = xi2x3 * xi3x1 - xi2x1 * xi3x3
    a22 = xi1x1 * xi3x3 - xi1x3 * xi3x1
    a23 = xi1x3 * xi2x1 - xi1x1 * xi2x3
    a31 = xi2x1 * xi3x2 - xi2x2 * xi3x1
    a32 = xi1x2 * xi3x1 - xi1x1 * xi3x2
    a33 = xi1x1 * xi2x2 - xi1x2 * xi2x1
    det = xi1x1 * a11 + xi2x1 * a12 + xi3x1 * a13
    idet = 1.0 / det
    x1xi1 = idet * (a11 * a11 + a12 * a12 + a13 * a13)
    x1xi2 = idet * (a11 * a21 + a12 * a22 + a13 * a23)
    x1xi3 = idet * (a11 * a31 + a12 * a32 + a13 * a33)
    x2xi1 = idet * (a21 * a11 + a22 * a12 + a23 * a13)
    x2xi2 = idet * (a21 * a21 + a22 * a22 + a23 * a23)
    x2xi3 = idet * (a21 * a31 + a22 * a32 + a23 * a33)
    x3xi1 = idet * (a31 * a11 + a32 * a12 + a33 * a13)
    x3xi2 = idet * (a31 * a21 + a32 * a22 + a33 * a23)
    x3xi3 = idet * (a31 * a31 + a32 * a32 + a33 * a33)

    sgeo = np.zeros((nelem, 6, Nfp[0], _nsgeo), dtype=FT)

    def face(f, sl, sign, m1, m2, m3):
        nn1 = (sign * J[sl] * m1[sl]).reshape(nelem, -1)
        nn2 = (sign * J[sl] * m2[sl]).reshape(nelem, -1)
        nn3 = (sign * J[sl] * m3[sl]).reshape(nelem, -1)
        sJ = np.sqrt(nn1 ** 2 + nn2 ** 2 + nn3 ** 2)
        sgeo[:, f, :, _n1] = nn1 / sJ
        sgeo[:, f, :, _n2] = nn2 / sJ
        sgeo[:, f, :, _n3] = nn3 / sJ
        sgeo[:, f, :, _sM] = sJ

    S = slice(None)
    face(0, (S, S, S, 0), -1, xi1x1, xi1x2, xi1x3)
    face(1, (S, S, S, Nq[0] - 1), 1, xi1x1, xi1x2, xi1x3)
    face(2, (S, S, 0, S), -1, xi2x1, xi2x2, xi2x3)
    face(3, (S, S, Nq[1] - 1, S), 1, xi2x1, xi2x2, xi2x3)
    face(4, (S, 0, S, S), -1, xi3x1, xi3x2, xi3x3)
    face(5, (S, Nq[2] - 1, S, S), 1, xi3x1, xi3x2, xi3x3)

    w1, w2, w3 = [np.asarray(x, dtype=FT) for x in w]
    Mw = w3[:, None, None] * w2[None, :, None] * w1[None, None, :]
    M = J * Mw[None]
    MI = 1 / M
    MIr = MI  # (nelem,k,j,i)
    sgeo[:, 0, :, _vMI] = MIr[:, :, :, 0].reshape(nelem, -1)
    sgeo[:, 1, :, _vMI] = MIr[:, :, :, Nq[0] - 1].reshape(nelem, -1)
    sgeo[:, 2, :, _vMI] = MIr[:, :, 0, :].reshape(nelem, -1)
    sgeo[:, 3, :, _vMI] = MIr[:, :, Nq[1] - 1, :].reshape(nelem, -1)
    sgeo[:, 4, :, _vMI] = MIr[:, 0, :, :].reshape(nelem, -1)
    sgeo[:, 5, :, _vMI] = MIr[:, Nq[2] - 1, :, :].reshape(nelem, -1)
    # surface quadrature weights (Grids.jl:1103-1114)
    sw = [
        (w3[:, None] * w2[None, :]).ravel(), (w3[:, None] * w2[None, :]).ravel(),
        (w3[:, None] * w1[None, :]).ravel(), (w3[:, None] * w1[None, :]).ravel(),
        (w2[:, None] * w1[None, :]).ravel(), (w2[:, None] * w1[None, :]).ravel(),
    ]
    for f in range(6):
        sgeo[:, f, :, _sM] *= sw[f][None, :]
    # horizontal metrics (Grids.jl:1133-1154)
    MHw = np.broadcast_to((w2[None, :, None] * w1[None, None, :]), (Nq[2], Nq[1], Nq[0]))
    Jb = M / Mw[None]
    MH = MHw[None] * np.sqrt((Jb * xi3x1) ** 2 + (Jb * xi3x2) ** 2 + (Jb * xi3x3) ** 2)

    vgeo = np.zeros((nelem, _nvgeo, Np), dtype=FT)
    cols = [xi1x1, xi2x1, xi3x1, xi1x2, xi2x2, xi3x2, xi1x3, xi2x3, xi3x3,
            M, MI, MH, x1, x2, x3, JcV,
            x1xi1, x2xi1, x3xi1, x1xi2, x2xi2, x3xi2, x1xi3, x2xi3, x3xi3]
    for c, arr in enumerate(cols):
        vgeo[:, c, :] = arr.reshape(nelem, Np)
    return vgeo, sgeo


class Grid:
    """One rank's DiscontinuousSpectralElementGrid (3-D)."""

    def __init__(self, topology, polynomialorder, FT=np.float64, meshwarp=None):
        assert topology.dim == 3
        if isinstance(polynomialorder, int):
            N = (polynomialorder,) * 3
        elif len(polynomialorder) == 2:
            N = (polynomialorder[0], polynomialorder[0], polynomialorder[1])
        else:
            N = tuple(polynomialorder)
        self.topology = topology
        self.N = N
        self.FT = np.dtype(FT).type
        self.Nq = tuple(n + 1 for n in N)
        self.Np = int(np.prod(self.Nq))
        self.Nfp = self.Np // self.Nq[0]
        self.nface = 6
        t = topology
        self.nelem, self.nreal = t.nelem, t.nreal
        self.vmapM, self.vmapP = mappings(N, t.elemtoelem, t.elemtoface, t.elemtoordr)
        self.vmaprecv, self.nabrtovmaprecv = commmapping(
            N, np.arange(t.nreal + 1, t.nelem + 1), t.ghostfaces, t.nabrtorecv)
        self.vmapsend, self.nabrtovmapsend = commmapping(
            N, t.sendelems, t.sendfaces, t.nabrtosend)
        xw = [elements.lglpoints(FT, n) for n in N]
        self.xi = [x for x, _ in xw]
        self.w = [w for _, w in xw]
        self.D = [elements.spectralderivative(x) for x in self.xi]
        self.Imat = [elements.indefinite_integral_interpolation_matrix(x, w)
                     for x, w in zip(self.xi, self.w)]
        self.vgeo, self.sgeo = computegeometry(t.elemtocoord, self.D, self.xi, self.w,
                                               meshwarp, FT)
        self.elemtobndy = np.ascontiguousarray(t.elemtobndy.T)  # (nelem, 6)
        self.interiorelems = t.interiorelems
        self.exteriorelems = t.exteriorelems
        self.nabrtorank = t.nabrtorank

    # convenience views -------------------------------------------------
    def coords(self):
        from .grids import _x1, _x2, _x3
        return self.vgeo[:, _x1, :], self.vgeo[:, _x2, :], self.vgeo[:, _x3, :]

    def min_node_distance_box(self):
        """Minimum element edge length * min LGL gap / 2 (affine bricks only)."""
        raise NotImplementedError
