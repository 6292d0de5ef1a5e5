"""Multi-GPU parity (run under torchrun, one rank per GPU): the oracle emulates all ranks
serially (Hilbert partition, ghost lists, halo exchange); every GPU rank runs libcmdg on its
own partition with the NCCL halo exchange and compares with the oracle's arrays of that rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_parity.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tests import parity  # noqa: E402
from oracle import dgmodel as odg, atmos as oatmos, odesolvers as oode, mpistatearrays as omsa  # noqa: E402
from oracle import grids as ogrids  # noqa: E402


def run_case(name, model, gs, Q0s, nf, dt, nsteps, rank, world, skip_zero_viscosity, diffusion_direction="every"):
    res = parity.multi_rank_case(model, gs, Q0s, nf, dt, nsteps, rank, world, skip_zero_viscosity,
                                 diffusion_direction)
    if rank == 0:
        print(f"MULTI_GPU_PARITY {name} world={world} halo_exact={res['halo_exact']} "
              f"tendency_rel_l2={res['tendency_rel_l2']:.3e} state_rel_l2={res['state_rel_l2']:.3e}", flush=True)
    assert res["halo_exact"], name
    assert res["tendency_rel_l2"] <= 1e-12 and res["state_rel_l2"] <= 1e-12, (name, res)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{torch.cuda.current_device()}"))
    # periodic-box vortex
    model, gs, setup, dt = parity.vortex_setup((4, 4, 3), csize=world)
    Q0s = [setup(g.vgeo[:g.nreal, ogrids._x1], g.vgeo[:g.nreal, ogrids._x2],
                 g.vgeo[:g.nreal, ogrids._x3], np.float64(0)) for g in gs]
    run_case("vortex", model, gs, Q0s, "rusanov", dt, 3, rank, world, True)
    # baroclinic wave on the cubed sphere (vertex-connectivity ghost layer, panel flips)
    model, gs = parity.gcm_setup(3, 2, csize=world)
    tmp = odg.DGModel(model, gs, "rusanov")
    Q0s = [oatmos.init_baroclinic_wave(model, np.moveaxis(a.data[:g.nreal], 1, 0))
           for g, a in zip(gs, tmp.state_auxiliary)]
    run_case("baroclinic_wave", model, gs, Q0s, "rusanov", 0.5, 3, rank, world, True)
    # second-order path with both exchanges (Q and the gradient flux)
    model, gs = parity.gcm_setup(3, 2, csize=world, turbulence=("smagorinsky", 0.21))
    tmp = odg.DGModel(model, gs, "rusanov")
    Q0s = [oatmos.init_baroclinic_wave(model, np.moveaxis(a.data[:g.nreal], 1, 0))
           for g, a in zip(gs, tmp.state_auxiliary)]
    run_case("held_suarez_like", model, gs, Q0s, "rusanov", 0.5, 2, rank, world, False, "horizontal")
    # DryBiharmonic hyperdiffusion: four exchanges per evaluation (Q, grad, Laplacian, total F2)
    model, gs = parity.gcm_setup(3, 2, csize=world, hyperdiffusion=("dry_biharmonic", 8 * 3600.0))
    tmp = odg.DGModel(model, gs, "rusanov", diffusion_direction="horizontal")
    Q0s = [oatmos.init_baroclinic_wave(model, np.moveaxis(a.data[:g.nreal], 1, 0))
           for g, a in zip(gs, tmp.state_auxiliary)]
    run_case("baroclinic_wave_hyperdiffusion", model, gs, Q0s, "rusanov", 0.5, 2, rank, world, False, "horizontal")
    run_ocean(rank, world)
    crash_propagation(rank, world)
    if world == 3:
        mpi_comm_known_answer(rank)
    dist.destroy_process_group()


def crash_propagation(rank, world):
    """check_for_crashes across ranks (MPIStateArrays.jl:910-935): a NaN on the LAST rank only is seen by every
    rank (ncclAllReduce of the flag), the failing rank reports itself, the others 'ErrorOnRemoteNode'."""
    P = parity.pkg()
    model, gs, setup, dt = parity.vortex_setup((4, 4, 3), csize=world)
    g = gs[rank]
    odgm = odg.DGModel(model, gs, "rusanov", skip_zero_viscosity=True)
    dgrid = parity.device_grid(g, device=f"cuda:{torch.cuda.current_device()}")
    aux = P.MPIStateArray(dgrid, model.A, data=odgm.state_auxiliary[rank].data)
    dg = P.DGModel(parity.device_model(model), dgrid, P.RusanovNumericalFlux(), P.CentralNumericalFluxSecondOrder(),
                   P.CentralNumericalFluxGradient(), state_auxiliary=aux, skip_zero_viscosity=True)
    uid = [P.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    dg.comm_init(uid[0], rank, world)
    Q = P.MPIStateArray(dgrid, 5)
    Q.data.fill_(1.0)
    ok0 = dg.check_for_crashes(Q, raise_on_failure=False) == (False, False)
    if rank == world - 1:
        Q.data[0, 4, 5] = float("nan")
    local, anyb = dg.check_for_crashes(Q, raise_on_failure=False)
    ok = ok0 and anyb and (local == (rank == world - 1))
    raised = None
    try:
        dg.check_for_crashes(Q)
    except FloatingPointError:
        raised = "local"
    except P.ErrorOnRemoteNode:
        raised = "remote"
    ok = ok and raised == ("local" if rank == world - 1 else "remote")
    flag = torch.tensor([0.0 if ok else 1.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"MULTI_GPU_PARITY crash_propagation world={world} ok={float(flag) == 0.0}", flush=True)
    assert float(flag) == 0.0
    dg.close()


def mpi_comm_known_answer(rank):
    """Device twin of the reference's own halo known-answer test test/Arrays/mpi_comm.jl:23-157 (three
    ranks, hand-written vmapsend / vmaprecv / neighbour ranges, two states): the same integers go
    through cmdg_exchange_begin / cmdg_exchange_end (pack kernel -> ncclSend/Recv -> unpack kernel).
    The reference test uses 9 nodes per element; libcmdg is compiled for Np = 125, so node n of element
    e of the reference's arrays is embedded as node n of element e here (linear id (e-1)*125 + n)."""
    from tests.test_oracle_mpi_comm import RANKS
    P = parity.pkg()
    r = RANKS[rank]
    Np9, Np, shift = 9, 125, 100
    nreal, nelem = r["numreal"], r["numreal"] + r["numghost"]

    def embed(ids):
        e, n = np.divmod(np.asarray(ids, dtype=np.int64) - 1, Np9)
        return e * Np + n + 1
    # a grid whose real elements are isolated (every face a wall): only the halo maps matter here
    ii = np.arange(Np).reshape((5, 5, 5), order="F")
    fmask = np.stack([ii[0].ravel(order="F"), ii[4].ravel(order="F"), ii[:, 0].ravel(order="F"),
                      ii[:, 4].ravel(order="F"), ii[:, :, 0].ravel(order="F"), ii[:, :, 4].ravel(order="F")])
    vmapM = (Np * np.arange(nelem))[:, None, None] + fmask[None] + 1
    grid = P.DiscontinuousSpectralElementGrid(
        4, np.ones((nelem, 25, Np)), np.ones((nelem, 6, 25, 5)), vmapM, vmapM, np.ones((nelem, 6), dtype=np.int64),
        np.eye(5), nreal, interiorelems=np.zeros(0, dtype=np.int64), exteriorelems=np.arange(1, nreal + 1),
        vmapsend=embed(r["vmapsend"]), vmaprecv=embed(r["vmaprecv"]), nabrtorank=r["nabrtorank"],
        nabrtovmapsend=r["nabrtovmapsend"], nabrtovmaprecv=r["nabrtovmaprecv"],
        device=f"cuda:{torch.cuda.current_device()}")
    model = P.AtmosModel(boundaryconditions=(P.AtmosBC(),))
    dg = P.DGModel(model, grid, P.RusanovNumericalFlux(), P.CentralNumericalFluxSecondOrder(),
                   P.CentralNumericalFluxGradient(), skip_zero_viscosity=True)
    uid = [P.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    dg.comm_init(uid[0], rank, 3)
    data = np.full((nelem, 2, Np), -1.0)
    vals = rank * 1000 + np.arange(1, Np9 * nreal + 1).reshape(nreal, Np9)
    data[:nreal, 0, :Np9] = vals
    data[:nreal, 1, :Np9] = vals + shift
    A = P.MPIStateArray(grid, 2, data=data)
    dg.ghost_exchange(A)
    got = A.data.cpu().numpy()
    e, n = np.divmod(np.asarray(r["vmaprecv"]) - 1, Np9)
    ok = np.array_equal(got[e, 0, n], np.asarray(r["expected"], dtype=np.float64)) and \
        np.array_equal(got[e, 1, n], shift + np.asarray(r["expected"], dtype=np.float64))
    # only the listed ghost nodes were written
    untouched = np.ones((nelem, Np), dtype=bool)
    untouched[:nreal, :Np9] = False
    untouched[e, n] = False
    ok = ok and bool(np.all(got[:, 0][untouched] == -1.0))
    flag = torch.tensor([0.0 if ok else 1.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"MULTI_GPU_PARITY mpi_comm_known_answer world=3 exact={float(flag) == 0.0}", flush=True)
    assert float(flag) == 0.0, "mpi_comm.jl known answers not reproduced by the device halo exchange"
    dg.close()


def run_ocean(rank, world):
    """HBModel on a partitioned box: Q and gradient-flux exchanges, ghost column integrals."""
    P = parity.pkg()
    model, gs, prob = parity.ocean_setup(world, (4, 4, 3))
    odgm = odg.DGModel(model, gs, "rusanov")
    oQ = odg.init_ode_state(odgm, lambda x1, x2, x3, a, t: prob.init_state(x1, x2, x3), 0.0)
    osol = oode.LSRK144NiegemannDiehlBusch(odgm, oQ, dt=120.0)
    oode.solve(oQ, osol, numberofsteps=2)       # spin-up so that every term is active
    g = gs[rank]
    dg, dgrid = parity.device_ocean_dg(model, g, odgm, rank=rank)
    uid = [P.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    dg.comm_init(uid[0], rank, world)
    data = oQ[rank].data.copy()
    data[g.nreal:] = np.nan
    dQ = P.MPIStateArray(dgrid, 4, data=data)
    odQ = [q.similar() for q in oQ]
    odgm(odQ, oQ, 0.0, 1, 0)
    dT = P.MPIStateArray(dgrid, 4)
    dg(dT, dQ, None, 0.0, 1.0, 0.0)
    r1 = parity.rel_l2(dT.realdata.cpu().numpy(), odQ[rank].realdata)
    osol2 = oode.LSRK144NiegemannDiehlBusch(odgm, oQ, dt=120.0)
    oode.solve(oQ, osol2, numberofsteps=2)
    dsol = P.LSRK144NiegemannDiehlBusch(dg, dQ, dt=120.0)
    P.solve(dQ, dsol, numberofsteps=2)
    r2 = parity.rel_l2(dQ.realdata.cpu().numpy(), oQ[rank].realdata)
    res = torch.tensor([r1, r2], dtype=torch.float64, device="cuda")
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"MULTI_GPU_PARITY ocean_hb world={world} tendency_rel_l2={float(res[0]):.3e} "
              f"state_rel_l2={float(res[1]):.3e}", flush=True)
    assert float(res[0]) <= 1e-12 and float(res[1]) <= 1e-12, res
    dg.close()


if __name__ == "__main__":
    main()
