"""Dev tool: time the fused LSRK54 stage on the headline workload for several values of an
environment knob read by cmdg_create (e.g. CMDG_PF, the L2 prefetch distance), building the grid
once.  Usage:  python tools/sweep.py [--workload W] [--ne N] [--steps K] KNOB v1 v2 ...
Prints one line per value: us per stage, GDOF/s, algorithmic GB/s."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import __graft_entry__ as ge

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="baroclinic_wave")
ap.add_argument("--ne", type=int, default=32)
ap.add_argument("--nvert", type=int, default=10)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--hyperdiffusion", action="store_true")
ap.add_argument("knob")
ap.add_argument("values", nargs="+")
args = ap.parse_args()

P = ge.load_package()
torch.cuda.set_device(0)
case = bench.build_case(P, args.workload, args.ne, args.nvert, 0, 1, "cuda:0", hyper=args.hyperdiffusion)
case["dg"].close()
grid, model, ai = case["grid"], case["model"], case["ai"]
aux0 = ai.init_state_auxiliary(model, grid, exchange=None)
case["aux"].data.copy_(aux0.data)
if args.workload in ("baroclinic_wave", "held_suarez"):
    Q0 = ai.baroclinic_wave(model, grid, case["aux"]).clone()
else:
    Q0 = ai.isentropic_vortex(model, grid, 0.0).clone()
_, bpn = bench.algorithmic_bytes_per_node(args.workload)
nodes = grid.nrealelem * 125
print(f"lib={P._lib.LIB_PATH} workload={args.workload} ne={args.ne} nelem={grid.nrealelem}", flush=True)
for v in args.values:
    os.environ[args.knob] = v
    second = args.hyperdiffusion or args.workload == "held_suarez"
    dg = P.DGModel(model, grid, P.RusanovNumericalFlux(), P.CentralNumericalFluxSecondOrder(),
                   P.CentralNumericalFluxGradient(), state_auxiliary=case["aux"],
                   diffusion_direction=P.HorizontalDirection() if second else None,
                   skip_zero_viscosity=not second, write_aux_diagnostics=True)
    Q = P.MPIStateArray(grid, 5)
    Q.data[:grid.nrealelem] = Q0
    sol = P.LSRK54CarpenterKennedy(dg, Q, dt=case["dt"], t0=0.0)
    sol.dostep(Q, 0.0, nsteps=5)
    torch.cuda.synchronize()
    best = 1e30
    dg.set_timing(True)
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sol.dostep(Q, 0.0, nsteps=args.steps)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    st = best / (args.steps * 5)
    print(f"{args.knob}={v:>6}: {st*1e3:8.1f} us/stage  {nodes*5/st/1e6:7.2f} GDOF/s  "
          f"{nodes*bpn/st/1e6:7.1f} GB/s algorithmic  norm={P.norm(Q):.12e}  per-stage us: "
          + " ".join(f"{k}={ms / (args.steps * 5) * 1e3:.1f}" for k, (ms, n) in dg.kernel_class_ms().items() if n),
          flush=True)
    del sol, dg, Q
