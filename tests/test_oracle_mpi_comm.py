"""The oracle's ghost exchange (pack -> per-neighbour messages -> unpack) on the reference's own
known-answer test ``test/Arrays/mpi_comm.jl:23-157``: three ranks with hand-written ``vmapsend`` /
``vmaprecv`` / neighbour ranges, two states, nine nodes per element; the ghost nodes must hold the
neighbours' values afterwards (exact integers)."""
import numpy as np

from oracle import mpistatearrays as msa

RANKS = [
    dict(numreal=4, numghost=3, nabrtorank=[1, 2],
         vmaprecv=[37, 38, 39, 40, 42, 43, 44, 45, 46, 49, 52, 53, 54, 57, 60, 61, 62, 63],
         vmapsend=[3, 6, 9, 10, 11, 12, 19, 22, 25, 34, 35, 36, 1, 2, 3, 28, 31, 34],
         nabrtovmaprecv=[(1, 13), (14, 18)], nabrtovmapsend=[(1, 12), (13, 18)],
         expected=[1001, 1002, 1003, 1004, 1006, 1007, 1008, 1009, 1010, 1013, 1016, 1017, 1018,
                   2003, 2006, 2007, 2008, 2009]),
    dict(numreal=2, numghost=4, nabrtorank=[0],
         vmaprecv=[21, 24, 27, 28, 29, 30, 37, 40, 43, 52, 53, 54],
         vmapsend=[1, 2, 3, 4, 6, 7, 8, 9, 10, 13, 16, 17, 18],
         nabrtovmaprecv=[(1, 12)], nabrtovmapsend=[(1, 13)],
         expected=[3, 6, 9, 10, 11, 12, 19, 22, 25, 34, 35, 36]),
    dict(numreal=1, numghost=2, nabrtorank=[0],
         vmaprecv=[10, 11, 12, 19, 22, 25], vmapsend=[3, 6, 7, 8, 9],
         nabrtovmaprecv=[(1, 6)], nabrtovmapsend=[(1, 5)],
         expected=[1, 2, 3, 28, 31, 34]),
]


def test_mpi_comm_known_answers():
    Np, shift = 9, 100
    arrays = []
    for crank, r in enumerate(RANKS):
        nelem = r["numreal"] + r["numghost"]
        a = msa.MPIStateArray(np.int64, Np, 2, nelem, r["numreal"], r["vmaprecv"], r["vmapsend"],
                              r["nabrtorank"], r["nabrtovmaprecv"], r["nabrtovmapsend"])
        a.data[...] = -1
        vals = crank * 1000 + np.arange(1, Np * r["numreal"] + 1).reshape(r["numreal"], Np)
        a.data[:r["numreal"], 0] = vals
        a.data[:r["numreal"], 1] = vals + shift
        arrays.append(a)
    msa.ghost_exchange(arrays)
    for a, r in zip(arrays, RANKS):
        e, n = np.divmod(np.asarray(r["vmaprecv"]) - 1, Np)
        assert np.array_equal(a.data[e, 0, n], r["expected"])
        assert np.array_equal(a.data[e, 1, n], shift + np.asarray(r["expected"]))
        # only the listed ghost nodes were written
        untouched = np.ones((a.nelem, Np), dtype=bool)
        untouched[:r["numreal"]] = False
        untouched[e, n] = False
        assert np.all(a.data[:, 0][untouched] == -1)
