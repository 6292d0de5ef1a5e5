"""Host-side mesh topologies for the harness (vectorised NumPy; setup code, not the hot path).

In production the topology comes from Julia (``ClimateMachine.Mesh.Topologies``) and only the
resulting arrays cross the C ABI.  The Python harness needs the same arrays to drive libcmdg
from ``bench.py`` at full problem sizes, so this module re-derives them -- independently of the
test oracle -- with array operations instead of per-element loops:

* Hilbert-curve ordering and equal contiguous partition of elements
  (src/Numerics/Mesh/BrickMesh.jl:40-155, 449-652)
* face matching, ghost / send lists (``connectmesh`` / ``connectmeshfull``, :827-1084, :1170-1480)
* ``BrickTopology``, ``StackedBrickTopology``, ``StackedCubedSphereTopology``
  (src/Numerics/Mesh/Topologies.jl:470-533, 632-800, 1067-1224, 1620-1790)

All returned index arrays are 1-based, shaped like the reference's (``elemtoelem[f, e]`` has
shape ``(nface, nelem)``).  ``tests/test_host_mesh.py`` checks them bit-for-bit against the
oracle (which is itself pinned on the reference's golden connectivity tables).
"""
from fractions import Fraction

import numpy as np


def linearpartition(n, p, np_):
    return ((p - 1) * n) // np_ + 1, (p * n) // np_


def hilbert_codes(X, bits=64):
    """Skilling's transform, vectorised over rows.  ``X``: (n, d) uint64.  Returns (n, d)
    uint64 code words, most significant first."""
    X = np.array(X, dtype=np.uint64, copy=True)
    n, d = X.shape
    one = np.uint64(1)
    M = one << np.uint64(bits - 1)
    Q = M
    for _ in range(bits - 1):
        P = Q - one
        for i in range(d):
            hit = (X[:, i] & Q) != 0
            X[hit, 0] ^= P
            t = (X[:, 0] ^ X[:, i]) & P
            t[hit] = 0
            X[:, 0] ^= t
            X[:, i] ^= t
        Q >>= one
    for i in range(1, d):
        X[:, i] ^= X[:, i - 1]
    t = np.zeros(n, dtype=np.uint64)
    Q = M
    for _ in range(bits - 1):
        hit = (X[:, d - 1] & Q) != 0
        t[hit] ^= Q - one
        Q >>= one
    X ^= t[:, None]
    H = np.zeros_like(X)
    for i in range(d):
        for j in range(bits):
            k = i * bits + j
            bit = (X[:, d - 1 - (k % d)] >> np.uint64(k // d)) & one
            H[:, d - 1 - i] |= bit << np.uint64(j)
    return H


def centroid_codes(elemtocoord):
    """Hilbert codes of centroids (``centroidtocode``): ``floor(typemax(UInt64) * c)`` with
    ``c`` the Float64 centroid normalised to the bounding box, computed exactly."""
    d, nvert, nelem = elemtocoord.shape
    cent = elemtocoord.sum(axis=1) / nvert
    cmin, cmax = cent.min(axis=1), cent.max(axis=1)
    size = cmax - cmin
    if not np.any(size):
        size = np.ones(d)
    else:
        size = np.where(size == 0, size.max(), size)
    c = (cent - cmin[:, None]) / size[:, None]  # (d, nelem) in [0, 1]
    # exact floor((2^64-1) * c): c = m * 2^e with 53-bit integer m
    m, e = np.frexp(c)
    mi = (m * (1 << 53)).astype(np.int64).astype(object)
    ex = (e - 53).astype(np.int64)
    tmax = (1 << 64) - 1
    X = np.zeros((nelem, d), dtype=np.uint64)
    flat_m, flat_e = mi.ravel(), ex.ravel()
    vals = np.empty(flat_m.shape, dtype=object)
    for idx in range(flat_m.size):  # python big ints: ~1 us each
        mm, ee = int(flat_m[idx]), int(flat_e[idx])
        prod = mm * tmax
        vals[idx] = (prod << ee) if ee >= 0 else (prod >> (-ee))
    X[:, :] = np.array(vals.reshape(d, nelem).T.tolist(), dtype=np.uint64)
    return hilbert_codes(X)


class Topology:
    """Per-rank result with the reference's BoxElementTopology field names."""
    pass


def _fmask(d):
    nvert = 2 ** d
    p = np.arange(nvert).reshape((2,) * d, order="F")
    cols = []
    for f in range(2 * d):
        idx = [slice(None)] * d
        idx[f // 2] = f % 2
        cols.append(p[tuple(idx)].ravel(order="F"))
    return np.stack(cols, axis=1)  # (nfacevert, nface)


def _pair_faces(keys, orient):
    """keys: (F, nfv) sorted vertex ids per face (row = elem*nface + face).  Returns for each
    face row the index of the matching row (itself if unmatched)."""
    F = keys.shape[0]
    order = np.lexsort(tuple(keys[:, c] for c in range(keys.shape[1] - 1, -1, -1)))
    ks = keys[order]
    same = np.all(ks[1:] == ks[:-1], axis=1)
    mate = np.arange(F)
    first = np.nonzero(same)[0]
    # non-manifold guard: a key may appear at most twice
    if first.size and np.any(first[1:] == first[:-1] + 1):
        raise ValueError("non-manifold face")
    a, b = order[first], order[first + 1]
    mate[a] = b
    mate[b] = a
    return mate


def _build_rank_topologies(dim, nranks, rank, elemtovert, vert_ident, elemtocoord, elemtobndy,
                           order, connectivity, face_keys=None):
    """Generic conforming quad/hex mesh -> Topology of ``rank``.

    ``elemtovert``: (nvert, nelem) global vertex ids (1-based) in the *original* element order;
    ``vert_ident``: maps vertex id -> identified id (periodicity); ``order``: permutation giving
    the global Hilbert order of elements.
    """
    nvert, nelem = elemtovert.shape
    nface = 2 * dim
    fm = _fmask(dim)
    nfv = fm.shape[0]
    # global numbering in Hilbert order
    ev = elemtovert[:, order]
    ec = elemtocoord[:, :, order]
    eb = elemtobndy[:, order]
    owner = np.empty(nelem, dtype=np.int64)
    localid = np.empty(nelem, dtype=np.int64)
    starts = []
    for r in range(nranks):
        a, b = linearpartition(nelem, r + 1, nranks)
        owner[a - 1:b] = r
        localid[a - 1:b] = np.arange(1, b - a + 2)
        starts.append(a - 1)
    # face keys: vertex ids of each face (row = e*nface + f); ``face_keys`` (original element
    # order) overrides them where the reference re-keys periodic faces (faceconnections)
    if face_keys is not None:
        fv_id = face_keys.reshape(nelem, nface, nfv)[order].reshape(nelem * nface, nfv)
    else:
        fv_raw = ev[fm.T.reshape(-1), :].reshape(nface, nfv, nelem)      # [f, v, e]
        fv_raw = np.moveaxis(fv_raw, 2, 0).reshape(nelem * nface, nfv)
        fv_id = vert_ident[fv_raw - 1]
    keys = np.sort(fv_id, axis=1)
    if nfv == 2:
        orient = np.where(fv_id[:, 1] < fv_id[:, 0], 2, 1)
    else:
        orient = np.ones(nelem * nface, dtype=np.int64)
    mate = _pair_faces(keys, orient)
    me = np.repeat(np.arange(nelem), nface)
    mf = np.tile(np.arange(nface), nelem)
    ne, nf = me[mate], mf[mate]
    nordr = np.where(orient[mate] == orient, 1, 2) if nfv == 2 else np.ones_like(me)
    unmatched = mate == np.arange(nelem * nface)
    nordr[unmatched] = 1

    mine = np.nonzero(owner == rank)[0]
    nreal = mine.size
    lo = starts[rank]
    T = Topology()
    T.dim = dim
    if connectivity == "face":
        rows = (mine[:, None] * nface + np.arange(nface)[None, :]).ravel()
        nbr_e = ne[rows]
        remote = owner[nbr_e] != rank
        # ghosts: unique remote neighbours ordered by (rank, remote local id) == global order
        ghosts = np.unique(nbr_e[remote])
        gfaces = np.zeros((nface, ghosts.size), dtype=bool)
        gpos = np.searchsorted(ghosts, nbr_e[remote])
        gfaces[nf[rows][remote], gpos] = True
        # send elements: ordered by (neighbour rank, local elem), unique per (rank, elem)
        se_rank = owner[nbr_e[remote]]
        se_elem = me[rows][remote]
        pairs = np.unique(np.stack([se_rank, se_elem], axis=1), axis=0)
        sfaces = np.zeros((nface, pairs.shape[0]), dtype=bool)
        pkey = pairs[:, 0] * nelem + pairs[:, 1]
        spos = np.searchsorted(pkey, se_rank * nelem + se_elem)
        sfaces[mf[rows][remote], spos] = True
        sendelems = pairs[:, 1] - lo + 1
        send_rank = pairs[:, 0]
    else:
        if dim != 2:
            raise NotImplementedError("vertex connectivity is for 2-D base meshes")
        # vertex -> set of ranks
        vid = vert_ident[ev - 1]                         # (nvert, nelem)
        vr = np.unique(np.stack([vid.ravel(), np.tile(owner, nvert)], axis=1), axis=0)
        nv = int(vert_ident.max()) + 1
        # incidence matrix vertex x rank (nranks is small)
        inc = np.zeros((nv, nranks), dtype=bool)
        inc[vr[:, 0], vr[:, 1]] = True
        touches = inc[vid]                               # (nvert, nelem, nranks)
        # ghosts: remote elements having a vertex touched by my rank
        gmask = touches[:, :, rank].any(axis=0) & (owner != rank)
        ghosts = np.nonzero(gmask)[0]
        vmark = touches[:, ghosts, rank]                 # (nvert, nghost)
        gfaces = np.stack([vmark[fm[:, f]].any(axis=0) for f in range(nface)])
        # send: my elements having a vertex touched by another rank r, per r
        t_mine = touches[:, mine, :]                     # (nvert, nreal, nranks)
        pr, pe = [], []
        sf = []
        for r in range(nranks):
            if r == rank:
                continue
            m = t_mine[:, :, r].any(axis=0)
            idx = np.nonzero(m)[0]
            pr.append(np.full(idx.size, r))
            pe.append(idx)
            vm = t_mine[:, idx, r]
            sf.append(np.stack([vm[fm[:, f]].any(axis=0) for f in range(nface)]))
        send_rank = np.concatenate(pr) if pr else np.zeros(0, dtype=np.int64)
        sendelems = (np.concatenate(pe) + 1) if pe else np.zeros(0, dtype=np.int64)
        sfaces = np.concatenate(sf, axis=1) if sf else np.zeros((nface, 0), dtype=bool)
    nghost = ghosts.size
    ntot = nreal + nghost
    # local numbering: real elements then ghosts
    g2l = np.full(nelem, -1, dtype=np.int64)
    g2l[mine] = np.arange(1, nreal + 1)
    g2l[ghosts] = nreal + np.arange(1, nghost + 1)
    loc = np.concatenate([mine, ghosts])
    elemtoelem = np.tile(np.arange(1, ntot + 1, dtype=np.int64), (nface, 1))
    elemtoface = np.tile(np.arange(1, nface + 1, dtype=np.int64)[:, None], (1, ntot))
    elemtoordr = np.ones((nface, ntot), dtype=np.int64)
    sel = loc if connectivity == "full" else mine
    rows = (sel[:, None] * nface + np.arange(nface)[None, :])
    nb_l = g2l[ne[rows]]                                  # (nsel, nface), -1 if not local
    ok = (nb_l > 0) & ~unmatched[rows]
    ee = np.broadcast_to(np.arange(sel.size)[:, None], rows.shape)
    ff = np.broadcast_to(np.arange(nface)[None, :], rows.shape)
    elemtoelem[ff[ok], ee[ok]] = nb_l[ok]
    elemtoface[ff[ok], ee[ok]] = nf[rows][ok] + 1
    elemtoordr[ff[ok], ee[ok]] = nordr[rows][ok]
    T.nelem, T.nreal, T.nghost = ntot, nreal, nghost
    T.elemtoelem, T.elemtoface, T.elemtoordr = elemtoelem, elemtoface, elemtoordr
    T.elemtocoord = ec[:, :, loc]
    T.elemtobndy = eb[:, loc].copy()
    T.elemtovert_global = ev[:, loc]
    T.ghostfaces = gfaces
    T.sendelems = sendelems.astype(np.int64)
    T.sendfaces = sfaces
    granks = owner[ghosts]
    T.nabrtorank = sorted(set(granks.tolist()) | set(send_rank.tolist()))
    T.nabrtorecv, T.nabrtosend = [], []
    for r in T.nabrtorank:
        idx = np.nonzero(granks == r)[0]
        T.nabrtorecv.append((int(idx[0]) + 1, int(idx[-1]) + 1))
        idx = np.nonzero(send_rank == r)[0]
        T.nabrtosend.append((int(idx[0]) + 1, int(idx[-1]) + 1))
    T.stacksize = 0
    _finish(T)
    return T


def _finish(T):
    ext = np.unique(T.sendelems)
    T.exteriorelems = ext.astype(np.int64)
    mask = np.ones(T.nreal, dtype=bool)
    mask[ext - 1] = False
    T.interiorelems = (np.nonzero(mask)[0] + 1).astype(np.int64)


def _enumerate_boundary_faces(elemtoelem, elemtobndy):
    """``enumerateboundaryfaces!``: boundary faces get a per-tag running number in elemtoelem
    (column-major traversal: e outer, f inner)."""
    nface, nelem = elemtoelem.shape
    flat_b = elemtobndy.T.ravel()           # e-major, f inner
    flat_e = elemtoelem.T.ravel().copy()
    for tag in np.unique(flat_b[flat_b != 0]):
        m = flat_b == tag
        flat_e[m] = np.arange(1, m.sum() + 1)
    elemtoelem[:, :] = flat_e.reshape(nelem, nface).T


def brick_topology(elemrange, periodicity, boundary=None, rank=0, nranks=1, connectivity="face"):
    """``BrickTopology`` (Topologies.jl:470-533) for ``rank`` of ``nranks``."""
    d = len(elemrange)
    if boundary is None:
        boundary = tuple((1, 1) for _ in range(d))
    x = [np.asarray(r) for r in elemrange]
    ne = [len(r) - 1 for r in x]
    nv = [len(r) for r in x]
    nelem = int(np.prod(ne))
    nvert = 2 ** d
    eidx = np.stack(np.unravel_index(np.arange(nelem), ne, order="F"))      # (d, nelem)
    off = np.stack(np.unravel_index(np.arange(nvert), (2,) * d, order="F"))  # (d, nvert)
    vc = eidx[:, None, :] + off[:, :, None]                                   # (d, nvert, nelem)
    elemtovert = np.ravel_multi_index(tuple(vc), nv, order="F") + 1           # (nvert, nelem)
    T_ = np.result_type(*[r.dtype for r in x])
    elemtocoord = np.stack([x[j][vc[j]] for j in range(d)]).astype(T_)
    elemtobndy = np.zeros((2 * d, nelem), dtype=np.int64)
    for i in range(d):
        if not periodicity[i]:
            elemtobndy[2 * i, eidx[i] == 0] = boundary[i][0]
            elemtobndy[2 * i + 1, eidx[i] == ne[i] - 1] = boundary[i][1]
    # periodic identification of vertices
    nvtot = int(np.prod(nv))
    vidx = np.stack(np.unravel_index(np.arange(nvtot), nv, order="F"))
    for i in range(d):
        if periodicity[i]:
            vidx[i] = np.where(vidx[i] == nv[i] - 1, 0, vidx[i])
    vert_ident = np.ravel_multi_index(tuple(vidx), nv, order="F")
    codes = centroid_codes(elemtocoord.astype(np.float64))
    order = np.lexsort(tuple(codes[:, c] for c in range(d - 1, -1, -1)))
    face_keys = None
    if connectivity == "face":
        # the reference keys a periodic high face by the vertices of the wrapped neighbour's
        # low face (``faceconnections``); every other face keeps its own vertices
        fm = _fmask(d)
        nface, nfv = 2 * d, fm.shape[0]
        fvc = vc[:, fm.T.reshape(-1), :].reshape(d, nface, nfv, nelem).copy()   # [dim, f, v, e]
        for i in range(d):
            if periodicity[i]:
                top = eidx[i] == ne[i] - 1
                fvc[i, 2 * i + 1, :, top] = 0
        fk = np.ravel_multi_index(tuple(fvc), nv, order="F")                      # [f, v, e]
        face_keys = np.moveaxis(fk, 2, 0).reshape(nelem * nface, nfv)
    T = _build_rank_topologies(d, nranks, rank, elemtovert, vert_ident, elemtocoord,
                               elemtobndy, order, connectivity, face_keys=face_keys)
    _enumerate_boundary_faces(T.elemtoelem, T.elemtobndy)
    return T


def _stack(base, dim, stack_coords, stacksize, periodic_vert, boundary_vert, ordr_map):
    nface = 2 * dim
    nhf = 2 * (dim - 1)
    nreal, nghost = base.nreal * stacksize, base.nghost * stacksize
    ntot = nreal + nghost
    T = Topology()
    T.dim = dim
    j = np.arange(stacksize)
    T.sendelems = (stacksize * (base.sendelems[:, None] - 1) + j[None, :] + 1).ravel()
    T.ghostfaces = np.zeros((nface, nghost), dtype=bool)
    T.ghostfaces[:nhf] = np.repeat(base.ghostfaces[:nhf], stacksize, axis=1)
    T.sendfaces = np.zeros((nface, T.sendelems.size), dtype=bool)
    T.sendfaces[:nhf] = np.repeat(base.sendfaces[:nhf], stacksize, axis=1)
    T.elemtocoord = stack_coords(base)
    e2e = np.tile(np.arange(1, ntot + 1, dtype=np.int64), (nface, 1))
    e2f = np.tile(np.arange(1, nface + 1, dtype=np.int64)[:, None], (1, ntot))
    e2o = np.ones((nface, ntot), dtype=np.int64)
    e2b = np.zeros((nface, ntot), dtype=np.int64)
    i = np.arange(base.nreal)
    e1 = (stacksize * i[:, None] + j[None, :])                      # 0-based (nreal_b, stack)
    for f in range(nhf):
        e2 = stacksize * (base.elemtoelem[f, :base.nreal] - 1)[:, None] + j[None, :] + 1
        e2e[f, e1.ravel()] = e2.ravel()
        e2f[f, e1.ravel()] = np.repeat(base.elemtoface[f, :base.nreal], stacksize)
        e2o[f, e1.ravel()] = np.repeat(ordr_map(base.elemtoordr[f, :base.nreal]), stacksize)
    eb = e1 - 1 + 1      # element below, 1-based id = e1(0-based) ; above = e1 + 2
    below = e1.copy()    # 1-based id of element j-1  == 0-based id of element j
    above = e1 + 2
    fb = np.full(e1.shape, nhf + 2)
    ft = np.full(e1.shape, nhf + 1)
    if periodic_vert:
        below[:, 0] = e1[:, -1] + 1
        above[:, -1] = e1[:, 0] + 1
    else:
        below[:, 0] = e1[:, 0] + 1
        above[:, -1] = e1[:, -1] + 1
        fb[:, 0] = nhf + 1
        ft[:, -1] = nhf + 2
    e2e[nhf, e1.ravel()] = below.ravel()
    e2e[nhf + 1, e1.ravel()] = above.ravel()
    e2f[nhf, e1.ravel()] = fb.ravel()
    e2f[nhf + 1, e1.ravel()] = ft.ravel()
    e2b[:nhf] = np.repeat(base.elemtobndy[:nhf], stacksize, axis=1)
    if not periodic_vert:
        e2b[nhf, 0::stacksize] = boundary_vert[0]
        e2b[nhf + 1, stacksize - 1::stacksize] = boundary_vert[1]
    T.nelem, T.nreal, T.nghost = ntot, nreal, nghost
    T.elemtoelem, T.elemtoface, T.elemtoordr, T.elemtobndy = e2e, e2f, e2o, e2b
    T.nabrtorank = list(base.nabrtorank)
    T.nabrtorecv = [(stacksize * (a - 1) + 1, stacksize * b) for a, b in base.nabrtorecv]
    T.nabrtosend = [(stacksize * (a - 1) + 1, stacksize * b) for a, b in base.nabrtosend]
    T.stacksize = stacksize
    _enumerate_boundary_faces(T.elemtoelem, T.elemtobndy)
    _finish(T)
    return T


def stacked_brick_topology(elemrange, periodicity, boundary=None, rank=0, nranks=1,
                           connectivity="full"):
    """``StackedBrickTopology`` (Topologies.jl:632-800)."""
    dim = len(elemrange)
    if boundary is None:
        boundary = tuple((1, 1) for _ in range(dim))
    if dim == 2:
        connectivity = "face"
    base = brick_topology(elemrange[:dim - 1], periodicity[:dim - 1], boundary[:dim - 1],
                          rank, nranks, connectivity)
    stack = np.asarray(elemrange[dim - 1])
    stacksize = len(stack) - 1
    nbv = 2 ** (dim - 1)

    def coords(b):
        T_ = np.result_type(b.elemtocoord.dtype, stack.dtype)
        ec = np.zeros((dim, 2 * nbv, b.nelem, stacksize), dtype=T_)
        ec[:dim - 1, :nbv] = b.elemtocoord[:dim - 1, :, :, None]
        ec[:dim - 1, nbv:] = b.elemtocoord[:dim - 1, :, :, None]
        ec[dim - 1, :nbv] = stack[None, None, :-1]
        ec[dim - 1, nbv:] = stack[None, None, 1:]
        return ec.reshape(dim, 2 * nbv, b.nelem * stacksize)

    def ordr(o):
        assert np.all(o == 1)
        return o

    return _stack(base, dim, coords, stacksize, periodicity[dim - 1], boundary[dim - 1], ordr)


def cubed_shell_topology(Ne, rank=0, nranks=1, connectivity="full"):
    """``CubedShellTopology`` (Topologies.jl:1067-1224): 6 Ne^2 quads embedded in 3-D."""
    nelem = 6 * Ne * Ne
    e = np.arange(nelem)
    i, j, blck = e % Ne + 1, (e // Ne) % Ne + 1, e // (Ne * Ne) + 1
    bx = np.array([0, Ne, 2 * Ne, Ne, Ne, Ne])[blck - 1]
    by = np.array([0, 0, 0, Ne, 2 * Ne, 3 * Ne])[blck - 1]
    flat = np.zeros((2, 4, nelem), dtype=np.int64)
    flat[0] = bx[None, :] + np.stack([i - 1, i, i - 1, i])
    flat[1] = by[None, :] + np.stack([j - 1, j - 1, j, j])

    def vertmap(a, b, c):
        return (a - 1) + (Ne + 1) * ((b - 1) + (Ne + 1) * (c - 1)) + 1

    elemtovert = np.zeros((4, nelem), dtype=np.int64)
    for n in range(1, 5):
        ix, jx = i + (n - 1) % 2, j + (n - 1) // 2
        one = np.ones_like(ix)
        v = np.select(
            [blck == 1, blck == 2, blck == 3, blck == 4, blck == 5, blck == 6],
            [vertmap(one, Ne + 2 - ix, jx), vertmap(ix, one, jx), vertmap((Ne + 1) * one, ix, jx),
             vertmap(ix, jx, (Ne + 1) * one), vertmap(ix, (Ne + 1) * one, Ne + 2 - jx),
             vertmap(ix, Ne + 2 - jx, one)])
        elemtovert[n - 1] = v
    codes = centroid_codes(flat.astype(np.float64))
    order = np.lexsort((codes[:, 1], codes[:, 0]))
    v0 = elemtovert - 1
    vi, vj, vk = v0 % (Ne + 1), (v0 // (Ne + 1)) % (Ne + 1), v0 // ((Ne + 1) ** 2)
    elemtocoord = (2 * np.stack([vi, vj, vk]).astype(np.float64) - Ne) / Ne
    nvtot = (Ne + 1) ** 3
    vert_ident = np.arange(nvtot)
    T = _build_rank_topologies(2, nranks, rank, elemtovert, vert_ident, elemtocoord,
                               np.zeros((4, nelem), dtype=np.int64), order, connectivity)
    return T


def stacked_cubed_sphere_topology(Nhorz, Rrange, boundary=(1, 1), rank=0, nranks=1,
                                  connectivity="full"):
    """``StackedCubedSphereTopology`` (Topologies.jl:1620-1790)."""
    Rrange = np.asarray(Rrange, dtype=np.float64)
    base = cubed_shell_topology(Nhorz, rank, nranks, connectivity)
    stacksize = len(Rrange) - 1

    def coords(b):
        ec = np.zeros((3, 8, b.nelem, stacksize))
        ec[:, :4] = b.elemtocoord[:, :, :, None] * Rrange[None, None, None, :-1]
        ec[:, 4:] = b.elemtocoord[:, :, :, None] * Rrange[None, None, None, 1:]
        return ec.reshape(3, 8, b.nelem * stacksize)

    def ordr(o):
        return np.where(o == 1, 1, 3)

    return _stack(base, 3, coords, stacksize, False, boundary, ordr)


def cubed_sphere_warp(a, b, c):
    """Equiangular gnomonic warp (Topologies.jl:1254-1299) on torch or numpy arrays with
    ``R = max(|a|,|b|,|c|)``."""
    import torch
    if isinstance(a, torch.Tensor):
        xp, where, stack = torch, torch.where, torch.stack
        absv = stack([a.abs(), b.abs(), c.abs()])
        R = absv.max(dim=0).values
        fdim = absv.argmax(dim=0)     # first maximum, like Julia's argmax
        tan, sqrt, pi = torch.tan, torch.sqrt, np.pi
    else:
        absv = np.stack([np.abs(a), np.abs(b), np.abs(c)])
        R = absv.max(axis=0)
        fdim = absv.argmax(axis=0)
        where, tan, sqrt, pi = np.where, np.tan, np.sqrt, np.pi

    def f(sR, xi, eta):
        X, Y = tan(pi * xi / 4), tan(pi * eta / 4)
        z1 = sR / sqrt(X * X + Y * Y + 1)
        return z1, X * z1, Y * z1

    one = R * 0 + 1
    sa = where(a < 0, -one, one)
    sb = where(b < 0, -one, one)
    sc = where(c < 0, -one, one)
    safe = lambda v: where(v == 0, one, v)
    # fdim == 0: x1,x2,x3 = f(+-R, b/a, c/a)
    p1, p2, p3 = f(sa * R, b / safe(a), c / safe(a))
    # fdim == 1: x2,x1,x3 = f(+-R, a/b, c/b)
    q2, q1, q3 = f(sb * R, a / safe(b), c / safe(b))
    # fdim == 2: x3,x2,x1 = f(+-R, b/c, a/c)
    r3, r2, r1 = f(sc * R, b / safe(c), a / safe(c))
    x1 = where(fdim == 0, p1, where(fdim == 1, q1, r1))
    x2 = where(fdim == 0, p2, where(fdim == 1, q2, r2))
    x3 = where(fdim == 0, p3, where(fdim == 1, q3, r3))
    return x1, x2, x3
