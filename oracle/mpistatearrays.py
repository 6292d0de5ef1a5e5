"""MPIStateArray and its ghost exchange (test infrastructure -- see oracle/__init__.py).

Restates ``src/Arrays/MPIStateArrays.jl``:

* layout ``data[n, s, e]`` (``:46-172``) -> NumPy ``data[e, s, n]`` (same bytes)
* ``begin_ghost_exchange!`` / ``end_ghost_exchange!`` (``:411-483``) with
  ``kernel_fillsendbuf!`` / ``kernel_transferrecvbuf!`` (``:837-871``) and the
  per-neighbour message slices of ``__Irecv!`` / ``__Isend!`` (``:485-514``) --
  emulated serially over the list of all ranks' arrays
* ``norm`` (``:583-604``), ``euclidean_distance`` (``:628-644``), weights =
  ``vgeo[:, _M, :]`` (``DGMethods/create_states.jl:16-17``)
"""
import numpy as np


class MPIStateArray:
    def __init__(self, FT, Np, nstate, nelem, nreal, vmaprecv, vmapsend,
                 nabrtorank, nabrtovmaprecv, nabrtovmapsend, weights=None):
        self.Np, self.nstate, self.nelem, self.nreal = Np, nstate, nelem, nreal
        self.data = np.zeros((nelem, nstate, Np), dtype=FT)
        self.vmaprecv = np.asarray(vmaprecv, dtype=np.int64)
        self.vmapsend = np.asarray(vmapsend, dtype=np.int64)
        self.nabrtorank = list(nabrtorank)
        self.nabrtovmaprecv = list(nabrtovmaprecv)
        self.nabrtovmapsend = list(nabrtovmapsend)
        self.weights = weights  # (nelem, Np) or None

    @classmethod
    def from_grid(cls, grid, nstate, FT=None):
        from .grids import _M
        FT = FT or grid.FT
        return cls(FT, grid.Np, nstate, grid.nelem, grid.nreal, grid.vmaprecv,
                   grid.vmapsend, grid.nabrtorank, grid.nabrtovmaprecv,
                   grid.nabrtovmapsend, weights=grid.vgeo[:, _M, :])

    @property
    def realdata(self):
        return self.data[:self.nreal]

    def similar(self, nstate=None):
        return MPIStateArray(self.data.dtype, self.Np, nstate or self.nstate, self.nelem,
                             self.nreal, self.vmaprecv, self.vmapsend, self.nabrtorank,
                             self.nabrtovmaprecv, self.nabrtovmapsend, self.weights)

    # kernel_fillsendbuf!: sendbuf[s, i] = buf[n, s, e], (e, n) = fldmod1(vmapsend[i], Np)
    def fillsendbuf(self):
        e, n = np.divmod(self.vmapsend - 1, self.Np)
        return self.data[e, :, n]  # (nsend, nstate) == Julia nstate x nsend bytes

    # kernel_transferrecvbuf!
    def transferrecvbuf(self, recvbuf):
        e, n = np.divmod(self.vmaprecv - 1, self.Np)
        self.data[e, :, n] = recvbuf


def ghost_exchange(arrays):
    """begin_ + end_ghost_exchange! for all ranks at once (rank = list index)."""
    sendbufs = [a.fillsendbuf() for a in arrays]
    for r, a in enumerate(arrays):
        recv = np.zeros((len(a.vmaprecv), a.nstate), dtype=a.data.dtype)
        for n, nbr in enumerate(a.nabrtorank):
            b = arrays[nbr]
            m = b.nabrtorank.index(r)
            s0, s1 = b.nabrtovmapsend[m]
            r0, r1 = a.nabrtovmaprecv[n]
            assert s1 - s0 == r1 - r0, "send/recv message size mismatch"
            recv[r0 - 1:r1] = sendbufs[nbr][s0 - 1:s1]
        a.transferrecvbuf(recv)


def norm(arrays, weighted=True):
    """2-norm over real elements of all ranks (``MPIStateArrays.jl:583-604``)."""
    if not isinstance(arrays, (list, tuple)):
        arrays = [arrays]
    tot = 0.0
    for a in arrays:
        d = a.realdata.astype(np.float64)
        if weighted and a.weights is not None:
            tot += float(np.sum(d * d * a.weights[:a.nreal, None, :]))
        else:
            tot += float(np.sum(d * d))
    return np.sqrt(tot)


def euclidean_distance(A, B):
    """sqrt(sum M (A-B)^2) over real elements of all ranks (``:628-644``)."""
    if not isinstance(A, (list, tuple)):
        A, B = [A], [B]
    tot = 0.0
    for a, b in zip(A, B):
        d = (a.realdata - b.realdata).astype(np.float64)
        E = d * d
        if a.weights is not None:
            E = E * a.weights[:a.nreal, None, :]
        tot += float(np.sum(E))
    return np.sqrt(tot)
