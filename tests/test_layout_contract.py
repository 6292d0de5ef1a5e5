"""The layout contract of the drop-in boundary: the reference's Diagnostics,
VTK / NetCDF writers and `Checkpoint.jl:33-64` read `Q.data`, `dg.state_auxiliary.data` as Julia
`Np x nstate x nelem` column-major arrays; the host mirror's tensors and what libcmdg is handed must have
exactly those bytes, and a checkpoint-style dump / restore must round-trip them."""
import io

import numpy as np
import torch

from tests import parity
from oracle import grids as ogrids, topologies as tp


def _grid():
    br = (np.linspace(0, 1, 3), np.linspace(0, 1, 2), np.linspace(0, 2, 3))
    topo = tp.StackedBrickTopology(1, br, periodicity=(True, True, False), boundary=((0, 0), (0, 0), (1, 2)))[0]
    return ogrids.Grid(topo, 4)


def test_mpistatearray_bytes_are_julia_column_major():
    P = parity.pkg()
    g = _grid()
    dgrid = parity.device_grid(g, device="cpu")
    Np, S, ne = dgrid.Np, 5, dgrid.nelem
    Q = P.MPIStateArray(dgrid, S)
    assert Q.data.shape == (ne, S, Np) and Q.data.is_contiguous()
    # Julia: Q.data[n, s, e] (1-based) lives at linear offset (n-1) + Np*((s-1) + S*(e-1))
    flat = Q.data.view(-1)
    for (n, s, e) in ((1, 1, 1), (7, 3, 2), (Np, S, ne)):
        off = (n - 1) + Np * ((s - 1) + S * (e - 1))
        flat[off] = 1000 * e + 10 * s + n / 1000
        assert Q.data[e - 1, s - 1, n - 1] == flat[off]
    # every (state, element) column is one contiguous Np-run: what the VTK writer slices
    col = Q.data[1, 2]
    assert col.is_contiguous() and col.data_ptr() == Q.data.data_ptr() + 8 * Np * (2 + S * 1)
    # realview(Q): the first nrealelem elements, a prefix of the buffer (MPIStateArrays.jl:174-186)
    assert Q.realdata.data_ptr() == Q.data.data_ptr() and Q.realdata.shape[0] == dgrid.nrealelem


def test_grid_arrays_keep_reference_ids_and_index_base():
    g = _grid()
    dgrid = parity.device_grid(g, device="cpu")
    # vgeo column ids (Grids.jl:76-92, 1-based there): xi{m}x{d} at 3(d-1)+m, M 10, MI 11, MH 12, x1..x3 13..15, JcV 16
    assert (ogrids._xi1x1, ogrids._xi2x1, ogrids._xi3x1, ogrids._xi1x2, ogrids._M, ogrids._MI, ogrids._MH,
            ogrids._x1, ogrids._x3, ogrids._JcV) == (0, 1, 2, 3, 9, 10, 11, 12, 14, 15)
    assert (ogrids._n1, ogrids._n2, ogrids._n3, ogrids._sM, ogrids._vMI) == (0, 1, 2, 3, 4)   # Grids.jl:129-146
    assert tuple(dgrid.vgeo.shape) == (dgrid.nelem, 25, 125) and tuple(dgrid.sgeo.shape) == (dgrid.nelem, 6, 25, 5)
    # index arrays stay 1-based Int64 as in Julia: vmap- of element e, face f enumerates that face's own nodes
    assert dgrid.vmapM.dtype == torch.int64 and int(dgrid.vmapM.min()) == 1
    e = 2
    own = dgrid.vmapM[e - 1].reshape(-1) - 1
    assert int(own.min()) // 125 == e - 1 and int(own.max()) // 125 == e - 1
    assert int(dgrid.elemtobndy.max()) == 2 and int(dgrid.elemtobndy.min()) == 0


def test_checkpoint_style_round_trip():
    """Checkpoint.jl:53-61 stores `Array(Q.data)` and `Array(dg.state_auxiliary.data)`; restoring copies
    them back into `Q.data` (`:110-124`).  Dump in Julia memory order, restore, compare bytes."""
    P = parity.pkg()
    g = _grid()
    dgrid = parity.device_grid(g, device="cpu")
    rng = np.random.default_rng(0)
    Q = P.MPIStateArray(dgrid, 5, data=rng.standard_normal((dgrid.nelem, 5, dgrid.Np)))
    buf = io.BytesIO()
    # a Julia reader sees an Np x 5 x nelem Float64 array: Fortran-order dump of the transposed view
    julia_view = Q.data.numpy().transpose(2, 1, 0)
    assert julia_view.flags.f_contiguous
    buf.write(julia_view.tobytes(order="F"))
    raw = np.frombuffer(buf.getvalue(), dtype=np.float64)
    assert np.array_equal(raw, Q.data.numpy().reshape(-1))
    restored = P.MPIStateArray(dgrid, 5, data=raw.reshape(dgrid.nelem, 5, dgrid.Np))
    assert torch.equal(restored.data, Q.data)
