"""N > 1 host-side logic on CPU with the gloo backend (world_size 2): every rank builds its own
partition with the harness's vectorised topology/grid code, the ranks exchange the coordinates
of their send nodes with torch.distributed point-to-point calls using exactly the per-neighbour
ranges libcmdg's NCCL exchange uses, and check that what arrives matches the ghost nodes
(the `mpi_connect_sphere.jl` property, across processes)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, kind, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import __graft_entry__ as ge
        ge.load_package()
        from climatemachine_jl_b200 import topologies as tp, grids as gr
        if kind == "sphere":
            R = np.array([1.0, 1.5, 2.0])
            topo = tp.stacked_cubed_sphere_topology(4, R, (1, 2), rank, world)
            grid = gr.build_grid(topo, 4, torch.float64, tp.cubed_sphere_warp, "cpu")
        else:
            br = (np.linspace(0, 1, 5), np.linspace(0, 1, 4), np.linspace(0, 1, 3))
            topo = tp.brick_topology(br, (True, True, True), None, rank, world)
            grid = gr.build_grid(topo, 4, torch.float64, None, "cpu")
        Np = grid.Np
        # element counts must add up to the global mesh
        n = torch.tensor([grid.nrealelem])
        dist.all_reduce(n)
        assert int(n) == (6 * 16 * 2 if kind == "sphere" else 4 * 3 * 2)
        # state = node coordinates on real elements, NaN on ghosts
        X = torch.full((grid.nelem, 3, Np), float("nan"), dtype=torch.float64)
        X[:grid.nrealelem] = grid.vgeo[:grid.nrealelem, 12:15]
        # pack (kernel_fillsendbuf!), send/recv per neighbour range, unpack (kernel_transferrecvbuf!)
        vs, vr = grid.vmapsend - 1, grid.vmaprecv - 1
        send = X[vs // Np, :, vs % Np].contiguous()
        recv = torch.empty((vr.numel(), 3), dtype=torch.float64)
        reqs = []
        for nb, (s0, s1), (r0, r1) in zip(grid.nabrtorank, grid.nabrtovmapsend, grid.nabrtovmaprecv):
            reqs.append(dist.isend(send[s0 - 1:s1].contiguous(), nb))
            reqs.append(dist.irecv(recv[r0 - 1:r1], nb))
        for r in reqs:
            r.wait()
        X[vr // Np, :, vr % Np] = recv
        # every interior face of a real element now sees matching coordinates on both sides
        interior = grid.elemtobndy[:grid.nrealelem] == 0
        vm = (grid.vmapM[:grid.nrealelem][interior] - 1).reshape(-1)
        vp = (grid.vmapP[:grid.nrealelem][interior] - 1).reshape(-1)
        for c in range(3):
            flat = X[:, c, :].reshape(-1)
            a, b = flat[vm], flat[vp]
            assert torch.isfinite(b).all(), "a face looks at a ghost node that was never received"
            if kind == "sphere":
                assert torch.allclose(a, b, rtol=1e-12, atol=1e-12)
            else:   # periodic box: coordinates agree modulo the period (1.0)
                d = (a - b).abs()
                assert ((d < 1e-12) | ((d - 1.0).abs() < 1e-12)).all()
        # interior/exterior split: exterior == elements that send
        ext = set(grid.exteriorelems.tolist())
        assert ext == set(topo.sendelems.tolist())
        assert ext.isdisjoint(set(grid.interiorelems.tolist()))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))


@pytest.mark.parametrize("kind", ["sphere", "box"])
def test_two_rank_partition_and_exchange_gloo(kind):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kind, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"
