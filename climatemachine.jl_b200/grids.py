"""Harness-side ``DiscontinuousSpectralElementGrid`` construction on the device (torch; setup
code, not the hot path).  In production Julia builds these arrays
(src/Numerics/Mesh/Grids.jl:267-413, Metrics.jl:85-114, 215-264, 431-722) and hands their
device pointers to ``cmdg_bind_grid``; the harness needs them at benchmark sizes without a
Julia runtime.  ``tests/test_host_mesh.py`` compares every array with the oracle's.
"""
import numpy as np
import torch

from .dgmodel import DiscontinuousSpectralElementGrid


def lglpoints(N):
    """Legendre-Gauss-Lobatto nodes/weights (float64): roots of P_N' polished by Newton."""
    if N == 1:
        return np.array([-1.0, 1.0]), np.array([1.0, 1.0])
    from numpy.polynomial import legendre as L
    c = np.zeros(N + 1)
    c[N] = 1
    x = np.sort(L.legroots(L.legder(c)))
    for _ in range(3):
        d1 = L.legval(x, L.legder(c))
        d2 = L.legval(x, L.legder(c, 2))
        x = x - d1 / d2
    x = np.concatenate(([-1.0], x, [1.0]))
    x = (x - x[::-1]) / 2
    w = 2 / (N * (N + 1) * L.legval(x, c) ** 2)
    return x, w


def spectralderivative(r):
    """Barycentric differentiation matrix D[j, k] = l_k'(r_j) (Elements.jl:60-82)."""
    n = len(r)
    diff = r[:, None] - r[None, :]
    np.fill_diagonal(diff, 1.0)
    wb = 1.0 / np.prod(diff, axis=1)
    D = (wb[None, :] / wb[:, None]) / diff
    np.fill_diagonal(D, 0.0)
    inv = 1.0 / (r[:, None] - r[None, :] + np.eye(n))
    np.fill_diagonal(inv, 0.0)
    D[np.diag_indices(n)] = inv.sum(axis=1)
    return D


def indefinite_integral_interpolation_matrix(r, w):
    """Grids.jl:1184-1206: row n integrates the interpolant from r[0] to r[n]."""
    Nq = len(r)
    diff = r[:, None] - r[None, :]
    np.fill_diagonal(diff, 1.0)
    wb = 1.0 / np.prod(diff, axis=1)
    out = np.zeros((Nq, Nq))
    for n in range(1, Nq):
        rdst = (1 - r) / 2 * r[0] + (1 + r) / 2 * r[n]
        I = np.zeros((Nq, Nq))
        for k in range(Nq):
            d = rdst[k] - r
            hit = np.nonzero(d == 0)[0]
            if hit.size:
                I[k, hit[0]] = 1.0
            else:
                t = wb / d
                I[k] = t / t.sum()
        out[n] = (r[n] - r[0]) / 2 * (w @ I)
    return out


def _face_tables(Nq):
    Np = Nq ** 3
    p = np.arange(Np).reshape((Nq, Nq, Nq), order="F")
    fmask = np.stack([p[0].ravel(order="F"), p[Nq - 1].ravel(order="F"),
                      p[:, 0].ravel(order="F"), p[:, Nq - 1].ravel(order="F"),
                      p[:, :, 0].ravel(order="F"), p[:, :, Nq - 1].ravel(order="F")])
    flip = np.arange(Nq * Nq).reshape((Nq, Nq), order="F")[::-1, :].ravel(order="F")
    return fmask, flip


def mappings(N, elemtoelem, elemtoface, elemtoordr):
    Nq = N + 1
    Np = Nq ** 3
    fmask, flip = _face_tables(Nq)
    nelem = elemtoelem.shape[1]
    e = np.arange(nelem)
    vmapM = (Np * e)[:, None, None] + fmask[None, :, :] + 1
    pat = np.stack([fmask, fmask[:, flip]], axis=1)            # [f2, flipped?, n]
    f2 = elemtoface.T - 1
    o2 = elemtoordr.T
    if not np.all((o2 == 1) | (o2 == 3)):
        raise NotImplementedError("face orientation other than 1 or 3")
    vmapP = Np * (elemtoelem.T - 1)[:, :, None] + pat[f2, (o2 == 3).astype(np.int64)] + 1
    return vmapM.astype(np.int64), vmapP.astype(np.int64)


def commmapping(N, commelems, commfaces, nabrtocomm):
    Nq = N + 1
    Np = Nq ** 3
    idx = np.stack(np.unravel_index(np.arange(Np), (Nq, Nq, Nq), order="F"))
    onface = np.stack([idx[0] == 0, idx[0] == Nq - 1, idx[1] == 0, idx[1] == Nq - 1,
                       idx[2] == 0, idx[2] == Nq - 1])           # (6, Np)
    add = (commfaces.T[:, :, None] & onface[None, :, :]).any(axis=1)   # (ncomm, Np)
    ids = (np.asarray(commelems, dtype=np.int64)[:, None] - 1) * Np + np.arange(Np)[None, :] + 1
    counts = add.sum(axis=1)
    vmapC = ids[add]
    csum = np.concatenate(([0], np.cumsum(counts)))
    ranges = [(int(csum[a - 1]) + 1, int(csum[b])) for a, b in nabrtocomm]
    return vmapC.astype(np.int64), ranges


def _deriv(D, x, axis):
    """sum_n D[a, n] x[.., n, ..] along reference axis (0 = fastest = last torch dim)."""
    ax = x.dim() - 1 - axis
    xm = x.movedim(ax, -1)
    return torch.matmul(xm, D.T).movedim(-1, ax)


def computegeometry(elemtocoord, N, FT, meshwarp, device, chunk=32768):
    """vgeo (nelem, 25, Np), sgeo (nelem, 6, Nfp, 5) on ``device`` in the reference layout."""
    Nq = N + 1
    Np, Nfp = Nq ** 3, Nq ** 2
    xi_np, w_np = lglpoints(N)
    D_np = spectralderivative(xi_np)
    xi = torch.as_tensor(xi_np, dtype=FT, device=device)
    w = torch.as_tensor(w_np, dtype=FT, device=device)
    D = torch.as_tensor(D_np, dtype=FT, device=device)
    nelem = elemtocoord.shape[2]
    vgeo = torch.zeros((nelem, 25, Np), dtype=FT, device=device)
    sgeo = torch.zeros((nelem, 6, Nfp, 5), dtype=FT, device=device)
    r = xi.view(1, 1, 1, Nq)
    s = xi.view(1, 1, Nq, 1)
    t = xi.view(1, Nq, 1, 1)
    Mw = w.view(Nq, 1, 1) * w.view(1, Nq, 1) * w.view(1, 1, Nq)
    sw = [(w.view(Nq, 1) * w.view(1, Nq)).reshape(-1)] * 6
    for c0 in range(0, nelem, chunk):
        c1 = min(nelem, c0 + chunk)
        ne = c1 - c0
        e2c = torch.as_tensor(np.ascontiguousarray(elemtocoord[:, :, c0:c1]), device=device).to(FT)
        X = []
        for n in range(3):
            c = [e2c[n, v].view(ne, 1, 1, 1) for v in range(8)]
            X.append(((1 - r) * (1 - s) * (1 - t) * c[0] + (1 + r) * (1 - s) * (1 - t) * c[1]
                      + (1 - r) * (1 + s) * (1 - t) * c[2] + (1 + r) * (1 + s) * (1 - t) * c[3]
                      + (1 - r) * (1 - s) * (1 + t) * c[4] + (1 + r) * (1 - s) * (1 + t) * c[5]
                      + (1 - r) * (1 + s) * (1 + t) * c[6] + (1 + r) * (1 + s) * (1 + t) * c[7]) / 8)
        x1, x2, x3 = X
        if meshwarp is not None:
            x1, x2, x3 = meshwarp(x1, x2, x3)
        xr = [_deriv(D, x, 0) for x in (x1, x2, x3)]
        xs = [_deriv(D, x, 1) for x in (x1, x2, x3)]
        xt = [_deriv(D, x, 2) for x in (x1, x2, x3)]
        JcV = torch.sqrt(xt[0] ** 2 + xt[1] ** 2 + xt[2] ** 2)
        J = (xr[0] * (xs[1] * xt[2] - xs[2] * xt[1]) + xr[1] * (xs[2] * xt[0] - xs[0] * xt[2])
             + xr[2] * (xs[0] * xt[1] - xs[1] * xt[0]))
        JI2 = 1 / (2 * J)
        # curl-invariant metric terms (Kopriva 2006; Metrics.jl:431-722)
        yzr, yzs, yzt = (x2 * xr[2] - x3 * xr[1], x2 * xs[2] - x3 * xs[1], x2 * xt[2] - x3 * xt[1])
        zxr, zxs, zxt = (x3 * xr[0] - x1 * xr[2], x3 * xs[0] - x1 * xs[2], x3 * xt[0] - x1 * xt[2])
        xyr, xys, xyt = (x1 * xr[1] - x2 * xr[0], x1 * xs[1] - x2 * xs[0], x1 * xt[1] - x2 * xt[0])
        d0, d1, d2 = (lambda a: _deriv(D, a, 0)), (lambda a: _deriv(D, a, 1)), (lambda a: _deriv(D, a, 2))
        xi1x1 = (d1(yzt) - d2(yzs)) * JI2
        xi2x1 = (-d0(yzt) + d2(yzr)) * JI2
        xi3x1 = (d0(yzs) - d1(yzr)) * JI2
        xi1x2 = (d1(zxt) - d2(zxs)) * JI2
        xi2x2 = (-d0(zxt) + d2(zxr)) * JI2
        xi3x2 = (d0(zxs) - d1(zxr)) * JI2
        xi1x3 = (d1(xyt) - d2(xys)) * JI2
        xi2x3 = (-d0(xyt) + d2(xyr)) * JI2
        xi3x3 = (d0(xys) - d1(xyr)) * JI2
        a11 = xi2x2 * xi3x3 - xi2x3 * xi3x2
        a12 = xi1x3 * xi3x2 - xi1x2 * xi3x3
        a13 = xi1x2 * xi2x3 - xi1x3 * xi2x2
        a21 = xi2x3 * xi3x1 - xi2x1 * xi3x3
        a22 = xi1x1 * xi3x3 - xi1x3 * xi3x1
        a23 = xi1x3 * xi2x1 - xi1x1 * xi2x3
        a31 = xi2x1 * xi3x2 - xi2x2 * xi3x1
        a32 = xi1x2 * xi3x1 - xi1x1 * xi3x2
        a33 = xi1x1 * xi2x2 - xi1x2 * xi2x1
        idet = 1.0 / (xi1x1 * a11 + xi2x1 * a12 + xi3x1 * a13)
        inv = [idet * (a11 * a11 + a12 * a12 + a13 * a13), idet * (a21 * a11 + a22 * a12 + a23 * a13),
               idet * (a31 * a11 + a32 * a12 + a33 * a13), idet * (a11 * a21 + a12 * a22 + a13 * a23),
               idet * (a21 * a21 + a22 * a22 + a23 * a23), idet * (a31 * a21 + a32 * a22 + a33 * a23),
               idet * (a11 * a31 + a12 * a32 + a13 * a33), idet * (a21 * a31 + a22 * a32 + a23 * a33),
               idet * (a31 * a31 + a32 * a32 + a33 * a33)]
        M = J * Mw
        MI = 1 / M
        MH = (w.view(1, Nq, 1) * w.view(1, 1, Nq)) * torch.sqrt(
            (J * xi3x1) ** 2 + (J * xi3x2) ** 2 + (J * xi3x3) ** 2)
        cols = [xi1x1, xi2x1, xi3x1, xi1x2, xi2x2, xi3x2, xi1x3, xi2x3, xi3x3, M, MI,
                MH.expand_as(M), x1, x2, x3, JcV] + inv
        vg = vgeo[c0:c1]
        for c, arr in enumerate(cols):
            vg[:, c, :] = arr.reshape(ne, Np)
        sg = sgeo[c0:c1]
        faces = [((Ellipsis, 0), -1, (xi1x1, xi1x2, xi1x3)), ((Ellipsis, Nq - 1), 1, (xi1x1, xi1x2, xi1x3)),
                 ((slice(None), slice(None), 0), -1, (xi2x1, xi2x2, xi2x3)),
                 ((slice(None), slice(None), Nq - 1), 1, (xi2x1, xi2x2, xi2x3)),
                 ((slice(None), 0), -1, (xi3x1, xi3x2, xi3x3)), ((slice(None), Nq - 1), 1, (xi3x1, xi3x2, xi3x3))]
        for f, (sl, sign, m) in enumerate(faces):
            nn = [(sign * J[sl] * mm[sl]).reshape(ne, Nfp) for mm in m]
            sJ = torch.sqrt(nn[0] ** 2 + nn[1] ** 2 + nn[2] ** 2)
            for c in range(3):
                sg[:, f, :, c] = nn[c] / sJ
            sg[:, f, :, 3] = sJ * sw[f]
            sg[:, f, :, 4] = MI[sl].reshape(ne, Nfp)
    return vgeo, sgeo, D_np, xi_np, w_np


def build_grid(topology, N, FT=torch.float64, meshwarp=None, device="cuda"):
    """Topology (climatemachine.jl_b200/topologies.py) -> device grid for ``DGModel``."""
    t = topology
    vmapM, vmapP = mappings(N, t.elemtoelem, t.elemtoface, t.elemtoordr)
    vmaprecv, nabrtovmaprecv = commmapping(N, np.arange(t.nreal + 1, t.nelem + 1), t.ghostfaces,
                                           t.nabrtorecv)
    vmapsend, nabrtovmapsend = commmapping(N, t.sendelems, t.sendfaces, t.nabrtosend)
    vgeo, sgeo, D, xi, w = computegeometry(t.elemtocoord, N, FT, meshwarp, device)
    g = DiscontinuousSpectralElementGrid.__new__(DiscontinuousSpectralElementGrid)
    dev = torch.device(device)
    g.N, g.Nq, g.Np, g.Nfp = N, N + 1, (N + 1) ** 3, (N + 1) ** 2
    g.vgeo, g.sgeo, g.FT = vgeo, sgeo, FT
    g.nelem, g.nrealelem = t.nelem, t.nreal
    tt = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.int64).to(dev)
    g.vmapM, g.vmapP = tt(vmapM), tt(vmapP)
    g.elemtobndy = tt(t.elemtobndy.T)
    g.D = torch.as_tensor(np.ascontiguousarray(D.T), dtype=FT).to(dev)   # Julia memory order
    g.D_host, g.xi, g.w = D, xi, w
    g.Imat = torch.as_tensor(np.ascontiguousarray(indefinite_integral_interpolation_matrix(xi, w).T),
                             dtype=FT).to(dev)   # Julia memory order
    g.interiorelems, g.exteriorelems = tt(t.interiorelems), tt(t.exteriorelems)
    g.vmapsend, g.vmaprecv = tt(vmapsend), tt(vmaprecv)
    g.nabrtorank = [int(r) for r in t.nabrtorank]
    g.nabrtovmapsend, g.nabrtovmaprecv = nabrtovmapsend, nabrtovmaprecv
    g.nvertelem = t.stacksize
    g.device = dev
    g.topology = t
    return g
