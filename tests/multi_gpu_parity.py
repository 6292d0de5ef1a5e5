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
    P = parity.pkg()
    odgm = odg.DGModel(model, gs, nf, skip_zero_viscosity=skip_zero_viscosity,
                       diffusion_direction=diffusion_direction)
    oQ = []
    for g, q0 in zip(gs, Q0s):
        q = omsa.MPIStateArray.from_grid(g, 5)
        np.moveaxis(q.data[:g.nreal], 1, 0)[...] = q0
        oQ.append(q)
    g = gs[rank]
    dgrid = parity.device_grid(g, device=f"cuda:{torch.cuda.current_device()}")
    m = parity.device_model(model)
    aux = P.MPIStateArray(dgrid, model.A, data=odgm.state_auxiliary[rank].data)
    dd = P.HorizontalDirection() if diffusion_direction == "horizontal" else P.EveryDirection()
    dg = P.DGModel(m, dgrid, getattr(P, parity.NF[nf])(), P.CentralNumericalFluxSecondOrder(),
                   P.CentralNumericalFluxGradient(), state_auxiliary=aux, diffusion_direction=dd,
                   skip_zero_viscosity=skip_zero_viscosity)
    uid = [P.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    dg.comm_init(uid[0], rank, world)
    # ghost elements of the device state start as NaN: only the exchange may fill them
    data = oQ[rank].data.copy()
    data[g.nreal:] = np.nan
    dQ = P.MPIStateArray(dgrid, 5, data=data)
    # halo exchange known answer: ghost face nodes must equal the oracle's after exchange
    omsa.ghost_exchange(oQ)
    dg.ghost_exchange(dQ)
    e, n = np.divmod(g.vmaprecv - 1, g.Np)
    got = dQ.data.cpu().numpy()[e, :, n]
    assert np.array_equal(got, oQ[rank].data[e, :, n]), "halo exchange mismatch"
    # tendency
    odQ = [q.similar() for q in oQ]
    odgm(odQ, oQ, 0.0, 1, 0)
    dT = P.MPIStateArray(dgrid, 5)
    dT.data.fill_(float("nan"))
    dg(dT, dQ, None, 0.0, 1.0, 0.0)
    r1 = parity.rel_l2(dT.realdata.cpu().numpy(), odQ[rank].realdata)
    # fused steps
    osol = oode.LSRK54CarpenterKennedy(odgm, oQ, dt=dt)
    oode.solve(oQ, osol, numberofsteps=nsteps)
    dsol = P.LSRK54CarpenterKennedy(dg, dQ, dt=dt)
    P.solve(dQ, dsol, numberofsteps=nsteps)
    r2 = parity.rel_l2(dQ.realdata.cpu().numpy(), oQ[rank].realdata)
    res = torch.tensor([r1, r2], dtype=torch.float64, device="cuda")
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"MULTI_GPU_PARITY {name} world={world} tendency_rel_l2={float(res[0]):.3e} "
              f"state_rel_l2={float(res[1]):.3e}", flush=True)
    assert float(res[0]) <= 1e-12 and float(res[1]) <= 1e-12, (name, res)
    dg.close()


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
    dist.destroy_process_group()


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
