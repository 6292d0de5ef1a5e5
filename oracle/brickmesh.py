"""Brick meshes, Hilbert partition and face connectivity (test infrastructure).

Restates ``src/Numerics/Mesh/BrickMesh.jl`` for a *serial emulation* of an
``csize``-rank MPI run: every function that is collective in the reference
takes the data of all ranks at once and returns one result per rank.

* ``linearpartition``   <- ``BrickMesh.jl:17-18``
* ``hilbertcode``       <- ``BrickMesh.jl:40-93`` (Skilling 2004)
* ``centroidtocode``    <- ``BrickMesh.jl:112-155``
* ``brickmesh``         <- ``BrickMesh.jl:272-348``
* ``partition``         <- ``BrickMesh.jl:449-522`` (getpartition) and ``:531-652``
* ``vertsortandorder``  <- ``BrickMesh.jl:685-791``
* ``connectmesh``       <- ``BrickMesh.jl:827-1084``
* ``enumerateboundaryfaces`` <- ``BrickMesh.jl:1092-1128``
* ``connectmeshfull``   <- ``BrickMesh.jl:1170-1480``

All element / face / vertex numbers stored in returned arrays are **1-based**
exactly as in the reference.  Arrays keep Julia's index order in their *shape*
here (e.g. ``elemtoelem[f-1, e-1]`` has shape ``(nface, nelem)``); the grid
layer transposes to Julia memory order where bytes matter.
"""
from fractions import Fraction

import numpy as np

U64 = (1 << 64) - 1


def linearpartition(n, p, np_):
    """1-based inclusive range (first, last) of piece ``p`` (1-based) of ``1:n``."""
    return ((p - 1) * n) // np_ + 1, (p * n) // np_


def hilbertcode(Y, bits=64):
    """Hilbert integer of integer axes ``Y`` (python ints, ``bits`` bits each)."""
    X = [int(y) for y in Y]
    n = len(X)
    mask = (1 << bits) - 1
    M = 1 << (bits - 1)
    Q = M
    for _ in range(bits - 1):
        P = Q - 1
        for i in range(n):
            if X[i] & Q:
                X[0] ^= P
            else:
                t = (X[0] ^ X[i]) & P
                X[0] ^= t
                X[i] ^= t
        Q >>= 1
    for i in range(1, n):
        X[i] ^= X[i - 1]
    t = 0
    Q = M
    for _ in range(bits - 1):
        if X[n - 1] & Q:
            t ^= Q - 1
        Q >>= 1
    for i in range(n):
        X[i] ^= t
    H = [0] * n
    for i in range(n):
        for j in range(bits):
            k = i * bits + j
            bit = (X[n - 1 - (k % n)] >> (k // n)) & 1
            H[n - 1 - i] |= bit << j
    return [h & mask for h in H]


def centroidtocode(elemtocoord_all, bits=64):
    """Hilbert codes of element centroids over the union of all ranks.

    ``elemtocoord_all``: array ``(d, nvert, nelem_total)``.  Returns a list of
    ``d``-tuples of python ints (most significant word first).  The reference
    converts ``c in [0,1]`` with ``floor(typemax(UInt64) * BigFloat(c))``
    (``BrickMesh.jl:146-150``); we do the same with exact rationals.
    """
    d, nvert, nelem = elemtocoord_all.shape
    centroids = elemtocoord_all.sum(axis=1) / nvert  # (d, nelem) float64
    cmin = centroids.min(axis=1) if nelem > 0 else np.zeros(d)
    cmax = centroids.max(axis=1) if nelem > 0 else np.zeros(d)
    csize = cmax - cmin
    if not np.any(csize):
        csize = np.ones(d)
    else:
        for i in range(d):
            if csize[i] == 0:
                csize[i] = csize.max()
    tmax = (1 << bits) - 1
    codes = []
    for e in range(nelem):
        c = (centroids[:, e] - cmin) / csize  # float64 arithmetic as in Julia
        X = [int(Fraction(float(ci)) * tmax) for ci in c]  # floor of exact product
        codes.append(tuple(hilbertcode(X, bits)))
    return codes


def _fmask(d):
    """Face masks over the 2^d corners, 0-based, shape (nfacevert, nface)."""
    nvert = 2 ** d
    p = np.arange(nvert).reshape((2,) * d, order="F")
    cols = []
    for f in range(2 * d):
        idx = [slice(None)] * d
        idx[f // 2] = f % 2
        cols.append(p[tuple(idx)].ravel(order="F"))
    return np.stack(cols, axis=1)


def brickmesh(x, periodic, part=1, numparts=1, boundary=None):
    """Cartesian brick, piece ``part`` of ``numparts`` (``BrickMesh.jl:272-348``)."""
    d = len(x)
    if boundary is None:
        boundary = tuple((1, 1) for _ in range(d))
    x = [np.asarray(xi) for xi in x]
    nvert = 2 ** d
    nface = 2 * d
    nelemdim = [len(xi) - 1 for xi in x]
    nvertdim = [len(xi) for xi in x]
    first, last = linearpartition(int(np.prod(nelemdim)), part, numparts)
    nloc = last - first + 1
    T = np.result_type(*[xi.dtype for xi in x])
    elemtovert = np.zeros((nvert, nloc), dtype=np.int64)
    elemtocoord = np.zeros((d, nvert, nloc), dtype=T)
    elemtobndy = np.zeros((nface, nloc), dtype=np.int64)
    faceconnections = []
    fmask = _fmask(d)

    def vertnum(vc):  # 1-based linear index, first index fastest
        return int(np.ravel_multi_index(vc, nvertdim, order="F")) + 1

    for e in range(nloc):
        ec = np.unravel_index(first - 1 + e, nelemdim, order="F")  # 0-based
        corners = []
        for v in range(nvert):
            off = np.unravel_index(v, (2,) * d, order="F")
            vc = tuple(ec[j] + off[j] for j in range(d))
            corners.append(vc)
            elemtovert[v, e] = vertnum(vc)
            for j in range(d):
                elemtocoord[j, v, e] = x[j][vc[j]]
        for i in range(d):
            if not periodic[i] and ec[i] == 0:
                elemtobndy[2 * i, e] = boundary[i][0]
            if not periodic[i] and ec[i] == nelemdim[i] - 1:
                elemtobndy[2 * i + 1, e] = boundary[i][1]
        for i in range(d):
            if periodic[i] and ec[i] == nelemdim[i] - 1:
                # corners of the (virtual) neighbour whose i-index range is 1:2
                ncorn = []
                for v in range(nvert):
                    off = np.unravel_index(v, (2,) * d, order="F")
                    vc = tuple(off[j] if j == i else ec[j] + off[j] for j in range(d))
                    ncorn.append(vertnum(vc))
                verts = [ncorn[k] for k in fmask[:, 2 * i]]
                faceconnections.append([e + 1, 2 * i + 2] + verts)
    return elemtovert, elemtocoord, elemtobndy, faceconnections


def partition(csize, meshes, globords=None):
    """Hilbert-curve partition of per-rank meshes (``BrickMesh.jl:449-652``).

    ``meshes``: list (len csize) of ``(elemtovert, elemtocoord, elemtobndy,
    faceconnections)``.  Returns the list of re-partitioned meshes, each
    ``(elemtovert, elemtocoord, elemtobndy, faceconnections, globord)``.

    Net effect of getpartition + Alltoallv + local re-sort: the globally
    Hilbert-sorted element list is cut into ``linearpartition`` chunks, and
    each rank holds its chunk in code order.
    """
    d, nvert, _ = meshes[0][1].shape
    nface = 2 * d
    nfacevert = 2 ** (d - 1)
    ev = np.concatenate([m[0] for m in meshes], axis=1)
    ec = np.concatenate([m[1] for m in meshes], axis=2)
    eb = np.concatenate([m[2] for m in meshes], axis=1)
    ntot = ev.shape[1]
    efc = np.zeros((nfacevert, nface, ntot), dtype=np.int64)
    off = 0
    for m in meshes:
        for fc in m[3]:
            efc[:, fc[1] - 1, off + fc[0] - 1] = fc[2:]
        off += m[0].shape[1]
    if globords is not None:
        go = np.concatenate(globords)
    codes = centroidtocode(ec)
    order = sorted(range(ntot), key=lambda e: (codes[e], e))
    out = []
    for r in range(csize):
        first, last = linearpartition(ntot, r + 1, csize)
        idx = order[first - 1:last]
        nev, nec, neb, nefc = ev[:, idx], ec[:, :, idx], eb[:, idx], efc[:, :, idx]
        nfc = []
        for e in range(len(idx)):
            for f in range(nface):
                if nefc[0, f, e] > 0:
                    nfc.append([e + 1, f + 1] + [int(v) for v in nefc[:, f, e]])
        ngo = go[idx] if globords is not None else None
        out.append((nev, nec, neb, nfc, ngo))
    return out


def vertsortandorder(*v):
    """Sorted vertex tuple and ordering code (``BrickMesh.jl:685-791``)."""
    def mmf(x, y):
        return (y, x, True) if y < x else (x, y, False)

    if len(v) == 1:
        return (v[0],), 1
    if len(v) == 2:
        a, b, s1 = mmf(v[0], v[1])
        return (a, b), (2 if s1 else 1)
    if len(v) == 4:
        a, b, c, d = v
        a, b, s1 = mmf(a, b)
        c, d, s2 = mmf(c, d)
        a, c, s3 = mmf(a, c)
        b, d, s4 = mmf(b, d)
        b, c, s5 = mmf(b, c)
        table = {
            (False, False, False, False, False): 1,
            (False, False, False, False, True): 2,
            (True, False, False, False, False): 3,
            (False, False, True, True, True): 4,
            (True, True, False, False, True): 5,
            (False, False, True, True, False): 6,
            (True, True, True, True, True): 7,
            (True, True, True, True, False): 8,
        }
        key = (s1, s2, s3, s4, s5)
        if key not in table:
            raise ValueError(f"Problem finding vertex ordering {v} with flips {key}")
        return (a, b, c, d), table[key]
    raise ValueError("unsupported number of face vertices")


class Connected:
    """Result record of ``connectmesh`` for one rank (fields as in the reference)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def connectmesh(csize, meshes, dim=None):
    """Face-neighbour connectivity for every rank (``BrickMesh.jl:827-1084``).

    ``meshes[r] = (elemtovert, elemtocoord, elemtobndy, faceconnections)``.
    """
    d = dim if dim is not None else meshes[0][1].shape[0]
    nface, nfacevert = 2 * d, 2 ** (d - 1)
    fmask = _fmask(d)
    # global face table: key -> list of (rank, e, f, o)
    table = {}
    for r, m in enumerate(meshes):
        elemtovert = m[0]
        nelem = elemtovert.shape[1]
        keys = {}
        for e in range(nelem):
            for f in range(nface):
                fv, o = vertsortandorder(*[int(elemtovert[k, e]) for k in fmask[:, f]])
                keys[(e + 1, f + 1)] = (fv, o)
        for fc in m[3]:
            fv, o = vertsortandorder(*fc[2:])
            keys[(fc[0], fc[1])] = (tuple(int(t) for t in fv), o)
        for (e, f), (fv, o) in keys.items():
            table.setdefault(fv, []).append((r, e, f, o))
    # match pairs (sorted columns: equal keys adjacent, pairs swap info)
    nbr = {}  # (r,e,f) -> (nr, ne, nf, no, mo)
    for fv, lst in table.items():
        lst.sort()
        if len(lst) == 2:
            a, b = lst
            nbr[a[:3]] = (b[0], b[1], b[2], b[3], a[3])
            nbr[b[:3]] = (a[0], a[1], a[2], a[3], b[3])
        elif len(lst) == 1:
            a = lst[0]
            nbr[a[:3]] = (a[0], a[1], a[2], a[3], a[3])
        else:
            raise ValueError("non-manifold face")
    out = []
    for crank, m in enumerate(meshes):
        elemtovert, elemtocoord, elemtobndy = m[0], m[1], m[2]
        nelem = elemtovert.shape[1]
        cols = []  # (ME, MF, MO, NR, NE, NF, NO)
        for e in range(1, nelem + 1):
            for f in range(1, nface + 1):
                nr, ne, nf, no, mo = nbr[(crank, e, f)]
                cols.append([e, f, mo, nr, ne, nf, no])
        # send lists: order by (NR, ME)
        sendelems, sendfaces = [], []
        counts = [0] * (csize + 1)
        counts[0] = 1 if cols else 0
        sr, se = -1, 0
        for c in sorted(cols, key=lambda c: (c[3], c[0], c[1])):
            r, e, f = c[3], c[0], c[1]
            if r != crank:
                if not (sr == r and se == e):
                    counts[r + 1] += 1
                    sendelems.append(e)
                    sendfaces.append([False] * nface)
                    sr, se = r, e
                sendfaces[-1][f - 1] = True
        sendstarts = np.cumsum(counts)
        nabrtosendrank = [r for r in range(csize) if sendstarts[r + 1] - sendstarts[r] > 0]
        nabrtosend = [(int(sendstarts[r]), int(sendstarts[r + 1] - 1))
                      for r in range(csize) if sendstarts[r + 1] - sendstarts[r] > 0]
        # ghost lists: order by (NR, NE)
        counts = [0] * (csize + 1)
        counts[0] = 1 if cols else 0
        sr, se = -1, 0
        nghost = 0
        ghostfaces = []
        ghost_src = []  # (rank, remote local elem)
        for c in sorted(cols, key=lambda c: (c[3], c[4], c[5])):
            r, e, f = c[3], c[4], c[5]
            if r != crank:
                if not (sr == r and se == e):
                    nghost += 1
                    counts[r + 1] += 1
                    sr, se = r, e
                    ghostfaces.append([False] * nface)
                    ghost_src.append((r, e))
                c[4] = nelem + nghost
                c[3] = crank
                ghostfaces[-1][f - 1] = True
        recvstarts = np.cumsum(counts)
        nabrtorecvrank = [r for r in range(csize) if recvstarts[r + 1] - recvstarts[r] > 0]
        nabrtorecv = [(int(recvstarts[r]), int(recvstarts[r + 1] - 1))
                      for r in range(csize) if recvstarts[r + 1] - recvstarts[r] > 0]
        assert nabrtorecvrank == nabrtosendrank
        ntot = nelem + nghost
        elemtoelem = np.tile(np.arange(1, ntot + 1, dtype=np.int64), (nface, 1))
        elemtoface = np.tile(np.arange(1, nface + 1, dtype=np.int64)[:, None], (1, ntot))
        elemtoordr = np.ones((nface, ntot), dtype=np.int64)
        for me, mf, mo, nr, ne, nf, no in cols:
            elemtoelem[mf - 1, me - 1] = ne
            elemtoface[mf - 1, me - 1] = nf
            if d == 2:
                elemtoordr[mf - 1, me - 1] = 1 if no == mo else 2
            else:
                if no != 1 or mo != 1:
                    raise NotImplementedError("TODO add support for other orientations")
                elemtoordr[mf - 1, me - 1] = 1
        newcoord = np.zeros(elemtocoord.shape[:2] + (ntot,), dtype=elemtocoord.dtype)
        newbndy = np.zeros((nface, ntot), dtype=np.int64)
        newcoord[:, :, :nelem] = elemtocoord
        newbndy[:, :nelem] = elemtobndy
        for g, (r, e) in enumerate(ghost_src):
            newcoord[:, :, nelem + g] = meshes[r][1][:, :, e - 1]
            newbndy[:, nelem + g] = meshes[r][2][:, e - 1]
        out.append(Connected(
            nelem=ntot, nreal=nelem, nghost=nghost,
            ghostfaces=np.array(ghostfaces, dtype=bool).reshape(nghost, nface).T,
            sendelems=np.array(sendelems, dtype=np.int64),
            sendfaces=np.array(sendfaces, dtype=bool).reshape(len(sendelems), nface).T,
            elemtocoord=newcoord, elemtovert=None,
            elemtoelem=elemtoelem, elemtoface=elemtoface, elemtoordr=elemtoordr,
            elemtobndy=newbndy, nabrtorank=nabrtorecvrank,
            nabrtorecv=nabrtorecv, nabrtosend=nabrtosend))
    return out


def enumerateboundaryfaces(elemtoelem, elemtobndy, periodicity, boundary):
    """In-place ``elemtoelem`` update at boundary faces (``BrickMesh.jl:1092-1128``)."""
    nb = 0
    for i, per in enumerate(periodicity):
        if not per:
            nb = max(nb, *boundary[i])
    assert nb <= 6
    bndytoelem = [[] for _ in range(nb)]
    bndytoface = [[] for _ in range(nb)]
    nface, nelem = elemtoelem.shape
    N = [0] * nb
    for e in range(nelem):
        for f in range(nface):
            dd = int(elemtobndy[f, e])
            assert 0 <= dd <= nb
            if dd != 0:
                N[dd - 1] += 1
                elemtoelem[f, e] = N[dd - 1]
                bndytoelem[dd - 1].append(e + 1)
                bndytoface[dd - 1].append(f + 1)
    return bndytoelem, bndytoface


def connectmeshfull(csize, meshes, dim=2):
    """Vertex-neighbour ("full") connectivity, 2-D base meshes only
    (``BrickMesh.jl:1170-1480``)."""
    assert dim == 2
    nvert = 4
    nfaces = 4
    fmask = _fmask(dim)
    nfvert = 2
    nelemv = [m[0].shape[1] for m in meshes]
    offset = np.concatenate(([0], np.cumsum(nelemv)))  # 0-based offsets
    evg = np.concatenate([m[0] for m in meshes], axis=1)
    rankg = np.concatenate([np.full(n, r) for r, n in enumerate(nelemv)])
    lclg = np.concatenate([np.arange(1, n + 1) for n in nelemv])
    ecg = np.concatenate([m[1] for m in meshes], axis=2)
    ebg = np.concatenate([m[2] for m in meshes], axis=1)
    nelemg = evg.shape[1]
    nvertg = int(evg.max())
    # periodic vertex identification
    vconng = []
    for m in meshes:
        seen = []
        for fc in m[3]:
            e, f, v = fc[0], fc[1], fc[2:]
            fv = [int(m[0][k, e - 1]) for k in fmask[:, f - 1]]
            fv, _ = vertsortandorder(*fv)
            v, _ = vertsortandorder(*v)
            for i in range(nfvert):
                pair = (fv[i], v[i])
                if pair not in seen:
                    seen.append(pair)
        vconng.extend(seen)
    gldofv = -np.ones(nvertg + 1, dtype=np.int64)  # 1-based
    pmarker = [-1] * len(vconng)
    for i, (v1, v2) in enumerate(vconng):
        if gldofv[v1] == -1 and gldofv[v2] == -1:
            gldofv[v1] = gldofv[v2] = min(v1, v2)
            pmarker[i] = 1
    for i, (v1, v2) in enumerate(vconng):
        if pmarker[i] == -1:
            idv = min(gldofv[v1], gldofv[v2])
            gldofv[v1] = gldofv[v2] = idv
    for i in range(1, nvertg + 1):
        if gldofv[i] == -1:
            gldofv[i] = i
    evg_orig = evg.copy()
    evg = gldofv[evg]
    vertgtoprocs = [[] for _ in range(nvertg + 1)]
    vertgtolelem = [[] for _ in range(nvertg + 1)]
    for icls in range(nelemg):
        for ivt in range(nvert):
            gvt = evg[ivt, icls]
            vertgtoprocs[gvt].append(int(rankg[icls]))
            vertgtolelem[gvt].append(int(lclg[icls]))
    out = []
    for crank, m in enumerate(meshes):
        elemtovert, elemtocoord, elemtobndy = m[0], m[1], m[2]
        nelem = elemtovert.shape[1]
        sendel = [[] for _ in range(csize)]
        recvel = [[] for _ in range(csize)]
        nsend = nghost = 0
        for icls in range(1, nelem + 1):
            for ivt in range(nvert):
                gvt = gldofv[elemtovert[ivt, icls - 1]]
                for ip, proc in enumerate(vertgtoprocs[gvt]):
                    if proc != crank:
                        lcell = vertgtolelem[gvt][ip]
                        if lcell not in recvel[proc]:
                            recvel[proc].append(lcell)
                            nghost += 1
                        if icls not in sendel[proc]:
                            sendel[proc].append(icls)
                            nsend += 1
        nabrtorank, nabrtosend, nabrtorecv = [], [], []
        newsendelems = []
        st_s = st_r = 1
        for ipr in range(csize):
            if sendel[ipr]:
                sendel[ipr].sort()
                newsendelems.extend(sendel[ipr])
                nabrtosend.append((st_s, st_s + len(sendel[ipr]) - 1))
                st_s += len(sendel[ipr])
                nabrtorank.append(ipr)
            if recvel[ipr]:
                recvel[ipr].sort()
                nabrtorecv.append((st_r, st_r + len(recvel[ipr]) - 1))
                st_r += len(recvel[ipr])
        sendfaces = np.zeros((nfaces, nsend), dtype=bool)
        ghostfaces = np.zeros((nfaces, nghost), dtype=bool)
        ntot = nelem + nghost
        newev = np.zeros((nvert, ntot), dtype=np.int64)
        newec = np.zeros(elemtocoord.shape[:2] + (ntot,), dtype=elemtocoord.dtype)
        neweb = np.zeros((nfaces, ntot), dtype=np.int64)
        newev[:, :nelem] = elemtovert
        newec[:, :, :nelem] = elemtocoord
        neweb[:, :nelem] = elemtobndy
        ctrg = ctrs = 0
        for ipr in range(csize):
            for icls in recvel[ipr]:
                g = offset[ipr] + icls - 1
                newev[:, nelem + ctrg] = evg_orig[:, g]
                newec[:, :, nelem + ctrg] = ecg[:, :, g]
                neweb[:, nelem + ctrg] = ebg[:, g]
                vmarker = [crank in vertgtoprocs[evg[ivt, g]] for ivt in range(nvert)]
                for fc in range(nfaces):
                    if any(vmarker[k] for k in fmask[:, fc]):
                        ghostfaces[fc, ctrg] = True
                ctrg += 1
            for icls in sendel[ipr]:
                vmarker = [ipr in vertgtoprocs[gldofv[elemtovert[ivt, icls - 1]]]
                           for ivt in range(nvert)]
                for fc in range(nfaces):
                    if any(vmarker[k] for k in fmask[:, fc]):
                        sendfaces[fc, ctrs] = True
                ctrs += 1
        cols = []
        for e in range(ntot):
            for f in range(nfaces):
                fv, o = vertsortandorder(*[int(newev[k, e]) for k in fmask[:, f]])
                key = tuple(int(gldofv[t]) for t in fv)
                cols.append((key, o, e + 1, f + 1))
        # sortslices(A, by = x -> x[1:nfvert]) is a stable sort on the key
        cols.sort(key=lambda c: c[0])
        elemtoelem = np.zeros((nfaces, ntot), dtype=np.int64)
        elemtoface = np.zeros((nfaces, ntot), dtype=np.int64)
        elemtoordr = np.zeros((nfaces, ntot), dtype=np.int64)
        j = 0
        n = len(cols)
        while j < n:
            key, o, lel, lfc = cols[j]
            if j + 1 < n and cols[j + 1][0] == key:
                _, o2, nel, nfc = cols[j + 1]
                elemtoelem[lfc - 1, lel - 1] = nel
                elemtoface[lfc - 1, lel - 1] = nfc
                elemtoelem[nfc - 1, nel - 1] = lel
                elemtoface[nfc - 1, nel - 1] = lfc
                od = 1 if o == o2 else 2
                elemtoordr[lfc - 1, lel - 1] = od
                elemtoordr[nfc - 1, nel - 1] = od
                j += 2
            else:
                elemtoelem[lfc - 1, lel - 1] = lel
                elemtoface[lfc - 1, lel - 1] = lfc
                elemtoordr[lfc - 1, lel - 1] = 1
                j += 1
        uv = gldofv[newev]
        uniq = np.unique(uv)
        remap = {int(v): i + 1 for i, v in enumerate(uniq)}
        newuv = np.vectorize(lambda v: remap[int(v)])(uv).astype(np.int64)
        out.append(Connected(
            nelem=ntot, nreal=nelem, nghost=nghost,
            ghostfaces=ghostfaces, sendelems=np.array(newsendelems, dtype=np.int64),
            sendfaces=sendfaces, elemtocoord=newec, elemtovert=newuv,
            elemtoelem=elemtoelem, elemtoface=elemtoface, elemtoordr=elemtoordr,
            elemtobndy=neweb, nabrtorank=nabrtorank,
            nabrtorecv=nabrtorecv, nabrtosend=nabrtosend))
    return out
