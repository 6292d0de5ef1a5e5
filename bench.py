#!/usr/bin/env python
"""Benchmark of the DG tendency + LSRK54 hot path (contract: see the task's bench.py section).

    python bench.py --gpus N --steps K --warmup W            # libcmdg on N B200s
    python bench.py --impl reference --gpus N --steps K ...  # restated reference CPU path

Workload (BASELINE.json configs[2], the one the metric is quoted on): dry baroclinic wave on
the cubed sphere, N = 4, 6 x 32^2 horizontal x 10 vertical elements, Rusanov flux, LSRK54,
Float64, synthetic analytic initial state; weak scaling keeps ~61 440 elements per GPU
(ne = 32, 45, 64, 90 for 1, 2, 4, 8 GPUs).  A "step" is one full LSRK54 step = 5 fused
tendency+stage-update kernels (+ halo exchange when N > 1).

One JSON line is printed by rank 0.  `value` = GDOF/s = (real elements x Np x 5 states) x
(tendency evaluations) / time, whole job; inputs resident in HBM.  `e2e` = the same metric
through cmdg_lsrk_steps_host with pinned HOST buffers: every step copies the state H2D, runs
one step, copies it back.

Besides the contract's keys the line carries (all measured in the same run):
  parity                   untimed check before the timed region, at every --gpus N, on the same ranks /
                           NCCL transport: partitioned cubed sphere (ne = 6 x 2) through cmdg_exchange_*,
                           cmdg_tendency and cmdg_lsrk_steps against the oracle's emulated-N-rank run
                           (tests/parity.py::multi_rank_case); the run FAILS if a bar is missed
  parity_fullsize_rel_l2   N = 1: one tendency of the headline state at the headline size (ne = 32 x 10)
                           against the C twin of the oracle on the same arrays
  reference_schedule       the headline with skip_zero_viscosity = false (the reference always runs the
                           nu = 0 gradient pass and the gradient-flux exchange)
  sustained_100            100 further steps when --steps < 100
  secondary                BASELINE.json configs[3] (Held-Suarez + Smagorinsky) and configs[4] (ocean
                           HBModel) at the same N: GDOF/s and device ms per kernel class
  cpu_baseline             N = 1: the C/OpenMP restatement of the reference schedule on the SAME mesh
                           (ne = 32 x 10), all host cores; also for Held-Suarez inside `secondary`
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NP, NSTATE = 125, 5
WEAK_NE = {1: 32, 2: 45, 4: 64, 8: 90}
METRIC = "DG tendency GDOF/s (fused LSRK54 stage; steps/s in lsrk54_steps_per_s)"
TRAFFIC_SOURCE = ("static: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this "
                  "kernel (profiles/traffic.json names the file) per node x this run's nodes; not re-measured here")


def algorithmic_bytes_per_node(workload):
    """compulsory bytes per node of one fused tendency+stage launch
    (Float64).  Stage 1 of every step has beta = RKA[1] = 0 and does not read dQ."""
    w = 8
    if workload == "ocean_gyre":
        # HBModel (S = 4, A_vol = y,w,pkin,wz0, A_face = w,pkin, 9 gradient-flux columns read by
        # the tendency kernel on both sides), LSRK144: 14 stages, stage 1 skips the dQ read.
        # Whole evaluation incl. gradient pass (262.4), column integrals (40), filters (48): 872.4.
        S, A_vol, A_face, GFu = 4, 4, 2, 9
        b_eval = w * (3 * S + 11 + A_vol) + 1.2 * (w * (5 + S + A_face) + 8) + w * GFu + 1.2 * w * GFu
        b_stage = b_eval + w * S
        return b_eval, (13 * b_stage + (b_stage - w * S)) / 14
    if workload == "rising_bubble":
        # dynamics as row (4) without the Held-Suarez coordinates (A_vol = Phi, grad Phi, rho_ref, p_ref, Delta = 7;
        # A_face = Phi, p_ref, Delta, grad Phi = 6); the tracer columns have their own kernels (TRACER_BYTES_PER_NODE)
        S, GF = 5, 10
        b_tend = w * (3 * S + 11 + 7) + 1.2 * (w * (5 + S + 6) + 8) + w * GF + 1.2 * w * GF
        b_grad = w * (S + 2 + 9 + GF) + 1.2 * (w * (5 + S + 2) + 8)
        b_stage = b_tend + w * S
        return b_tend + b_grad, (13 * b_stage + (b_stage - w * S)) / 14
    if workload == "held_suarez":
        # Euler part + gradient pass + gradient-flux reads in the tendency pass.
        # Tendency launch only (the roofline kernel): A_vol = Phi, grad Phi, rho_ref, p_ref, Delta,
        # coord (HS latitude) = 10; A_face = Phi, p_ref, Delta, grad Phi = 6 ... the formula of 8(d).
        S, GF = 5, 10
        b_tend = w * (3 * S + 11 + 10) + 1.2 * (w * (5 + S + 6) + 8) + w * GF + 1.2 * w * GF
        b_grad = w * (S + 2 + 9 + GF) + 1.2 * (w * (5 + S + 2) + 8)
        b_stage = b_tend + w * S
        # returned pair: (whole evaluation incl. gradient pass, tendency launch averaged over stages)
        return b_tend + b_grad, (4 * b_stage + (b_stage - w * S)) / 5
    if workload == "baroclinic_wave":
        S, A_vol, A_face = 5, 6, 2
    else:  # isentropic vortex, Euler-minimal
        S, A_vol, A_face = 5, 0, 0
    b_eval = w * (S + S + S + 11 + A_vol) + 1.2 * (w * (5 + S + A_face) + 8)
    b_stage = b_eval + w * S
    b_stage_first = b_stage - w * S
    return b_eval, (4 * b_stage + b_stage_first) / 5


# passive tracers (NTracers{4}), same counting rules.  tracer_tendency_kernel: rho, rho u (4) + rho chi (4) + dQ chi
# read (4) and written (4) + new state (4) + own diffusive flux (12) + geometry (11), per face node face geometry (5),
# the neighbour's rho, rho u, rho chi (8), its diffusive flux (12) and one index; tracer_gradient_kernel: rho, rho chi
# (5) + nu (3) + delta_chi (4) + 9 metrics, writes the diffusive flux (12) (+ grad chi (12) on the last stage), per face
# node face geometry (5) + the neighbour's rho, rho chi (5) + one index
TRACER_BYTES_PER_NODE = {"tracer_tendency": 8 * (4 + 4 + 4 + 4 + 4 + 12 + 11) + 1.2 * (8 * (5 + 8 + 12) + 8),
                         "tracer_gradient": 8 * (5 + 3 + 4 + 9 + 12) + 1.2 * (8 * (5 + 5) + 8)}
GRADIENT_BYTES_PER_NODE = 8 * (5 + 2 + 9 + 10) + 1.2 * (8 * (5 + 5 + 2) + 8)   # SURVEY 8(d) B_2nd gradient pass


# ----------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.idx), "-lms", os.environ.get("BENCH_SMI_MS", "20")],
                stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        inside = [s for (t, s) in self.samples if t0 is not None and t0 <= t <= t1]
        where = "timed region"
        if len(inside) < 3:   # short region: use every sample since the warm-up started (GPU busy)
            inside, where = [s for (_, s) in self.samples], "warm-up + timed region"
        for s in inside:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": where}


# ----------------------------------------------------------------------------------------
# libcmdg arm
# ----------------------------------------------------------------------------------------
def gcm_model(P, workload, hyper=False):
    """BASELINE.json configs[2] / configs[3] balance laws (host mirror objects)."""
    common = dict(orientation=P.SphericalOrientation(),
                  ref_state=P.HydrostaticState(P.DecayingTemperatureProfile(290.0, 220.0, 8e3)),
                  boundaryconditions=(P.AtmosBC(), P.AtmosBC()),
                  hyperdiffusion=P.DryBiharmonic(8 * 3600.0) if hyper else None)
    if workload == "held_suarez":
        # configs[3] as tutorials/Atmos/heldsuarez.jl:160-201 sets it, explicit LSRK54, hyperdiffusion off
        # by default: Smagorinsky(0.21), horizontal diffusion direction, Gravity + Coriolis +
        # HeldSuarezForcing + RayleighSponge(30 km, 12 km, 1/900 s)
        return P.AtmosModel(turbulence=P.SmagorinskyLilly(0.21),
                            source=(P.Gravity(), P.Coriolis(), P.HeldSuarezForcing(),
                                    P.RayleighSponge(30e3, 12e3, 1 / 60 / 15, (0.0, 0.0, 0.0), 2.0)), **common)
    return P.AtmosModel(turbulence=P.ConstantKinematicViscosity(0.0), source=(P.Gravity(), P.Coriolis()), **common)


def build_grid(P, workload, ne, nvert, rank, nranks, device):
    import numpy as np
    import torch
    from climatemachine_jl_b200 import topologies as tp, grids as gr
    if workload in ("baroclinic_wave", "held_suarez"):
        ps = P.EarthParameterSet()
        R = np.linspace(ps.planet_radius, ps.planet_radius + 30e3, nvert + 1)
        topo = tp.stacked_cubed_sphere_topology(ne, R, (1, 2), rank, nranks)
        return gr.build_grid(topo, 4, torch.float64, tp.cubed_sphere_warp, device), None
    if workload == "ocean_gyre":
        # BASELINE.json configs[4]: OceanBoxGCM HBModel, 20 x 20 x 50 elements per GPU
        # (experiments/OceanBoxGCM/homogeneous_box.jl:11-21), box replicated fx x fy times horizontally (weak_box)
        fx, fy = weak_box(nranks)
        prob = P.OceanGyre(4e6 * fx, 4e6 * fy, 1000.0)
        br = (np.linspace(0, prob.Lˣ, ne * fx + 1), np.linspace(0, prob.Lʸ, ne * fy + 1),
              np.linspace(-prob.H, 0, nvert + 1))
        topo = tp.stacked_brick_topology(br, (False, False, False), ((1, 1), (1, 1), (2, 3)), rank, nranks)
        return gr.build_grid(topo, 4, torch.float64, None, device), prob
    if workload == "rising_bubble":
        # BASELINE.json configs[0] (tutorials/Atmos/risingbubble.jl) at benchmark size: 500 m elements, 10 km deep
        # (20 levels as the tutorial), ne x ne/4 elements horizontally per GPU (the tutorial: 20 x 1; weak_box copies), periodic x / y
        fx, fy = weak_box(nranks)
        nx, ny = ne * fx, max(ne // 4, 1) * fy
        br = (np.linspace(0, 500.0 * nx, nx + 1), np.linspace(0, 500.0 * ny, ny + 1), np.linspace(0, 10000.0, nvert + 1))
        topo = tp.stacked_brick_topology(br, (True, True, False), ((0, 0), (0, 0), (1, 2)), rank, nranks)
        return gr.build_grid(topo, 4, torch.float64, None, device), None
    L = 0.05
    br = tuple(np.linspace(-L, L, ne + 1) for _ in range(3))
    topo = tp.brick_topology(br, (True, True, True), None, rank, nranks)
    return gr.build_grid(topo, 4, torch.float64, None, device), None


def build_case(P, workload, ne, nvert, rank, nranks, device, hyper=False, skip_zero_viscosity=True,
               grid=None, prob=None):
    import numpy as np
    from climatemachine_jl_b200 import atmos_init as ai
    if grid is None:
        grid, prob = build_grid(P, workload, ne, nvert, rank, nranks, device)
    nf = (P.RusanovNumericalFlux(), P.CentralNumericalFluxSecondOrder(), P.CentralNumericalFluxGradient())
    if workload == "ocean_gyre":
        model = P.HBModel(prob, cʰ=float(np.sqrt(9.81 * prob.H)))
        Q0, aux = ai.ocean_gyre_state(prob, grid)
        md = dict(vert_filter=P.CutoffFilter(grid, 3), exp_filter=P.ExponentialFilter(grid, 1, 8))
        dg = P.DGModel(model, grid, *nf, state_auxiliary=aux, modeldata=md)
        return dict(grid=grid, model=model, dg=dg, aux=aux, dt=55.0, ai=ai, Q0=Q0, prob=prob, skip=False)
    if workload in ("baroclinic_wave", "held_suarez"):
        model = gcm_model(P, workload, hyper)
        dt = 0.4     # s; vertical acoustic CFL ~0.3
        second = workload == "held_suarez" or hyper or not skip_zero_viscosity
        aux = P.MPIStateArray(grid, model.number_states("Auxiliary"))
        # experiments/TestCase/baroclinic_wave.jl:258, tutorials/Atmos/heldsuarez.jl:252: diffdir = HorizontalDirection()
        dg = P.DGModel(model, grid, *nf, state_auxiliary=aux, diffusion_direction=P.HorizontalDirection(),
                       skip_zero_viscosity=not second, write_aux_diagnostics=True)
        return dict(grid=grid, model=model, dg=dg, aux=aux, dt=dt, ai=ai, skip=not second)
    if workload == "rising_bubble":
        # as the tutorial ships it: SmagorinskyLilly(0.21), DryAdiabaticProfile(300 K, 0 K), Gravity, NTracers{4}
        # with delta_chi = (1, 2, 3, 4), Rusanov, LSRK144NiegemannDiehlBusch
        model = P.AtmosModel(orientation=P.FlatOrientation(),
                             ref_state=P.HydrostaticState(P.DryAdiabaticProfile(300.0, 0.0)),
                             turbulence=P.SmagorinskyLilly(0.21), source=(P.Gravity(),),
                             boundaryconditions=(P.AtmosBC(), P.AtmosBC()), tracers=P.NTracers((1.0, 2.0, 3.0, 4.0)))
        aux = P.MPIStateArray(grid, model.number_states("Auxiliary"))
        dg = P.DGModel(model, grid, *nf, state_auxiliary=aux, write_aux_diagnostics=True)
        return dict(grid=grid, model=model, dg=dg, aux=aux, dt=0.4, ai=ai, skip=False)
    model = P.AtmosModel()
    aux = P.MPIStateArray(grid, model.number_states("Auxiliary"))
    dg = P.DGModel(model, grid, *nf, state_auxiliary=aux, skip_zero_viscosity=skip_zero_viscosity,
                   write_aux_diagnostics=True)
    return dict(grid=grid, model=model, dg=dg, aux=aux, dt=(2 * 0.05 / ne) / 347.2 / 16, ai=ai,
                skip=skip_zero_viscosity)


def init_case(P, case, workload, rank, world, dist):
    """NCCL communicator of this handle, auxiliary state, initial condition, solver."""
    dg, grid, model, ai = case["dg"], case["grid"], case["model"], case["ai"]
    if world > 1:
        uid = [P.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        dg.comm_init(uid[0], rank, world)
    ocean = workload == "ocean_gyre"
    Q = P.MPIStateArray(grid, 4 if ocean else model.number_states("Prognostic"))
    if ocean:
        Q.data[:grid.nrealelem] = case["Q0"]
        if world > 1:
            dg.ghost_exchange(case["aux"])
        sol = P.LSRK144NiegemannDiehlBusch(dg, Q, dt=case["dt"], t0=0.0)
    else:
        if "aux0" not in case:
            ex = (lambda arr: dg.ghost_exchange(arr)) if world > 1 else None
            case["aux0"] = ai.init_state_auxiliary(model, grid, exchange=ex)
        case["aux"].data.copy_(case["aux0"].data)
        if workload in ("baroclinic_wave", "held_suarez"):
            # (Held-Suarez starts from rest + noise in the tutorial; the baroclinic-wave state gives
            # the friction, relaxation and sponge terms something to act on -- synthetic either way)
            Q.data[:grid.nrealelem] = ai.baroclinic_wave(model, grid, case["aux"])
        elif workload == "rising_bubble":
            Q.data[:grid.nrealelem] = ai.rising_bubble(model, grid, case["aux"])
        else:
            Q.data[:grid.nrealelem] = ai.isentropic_vortex(model, grid, 0.0)
        if workload == "rising_bubble":
            sol = P.LSRK144NiegemannDiehlBusch(dg, Q, dt=case["dt"], t0=0.0)
        else:
            sol = P.LSRK54CarpenterKennedy(dg, Q, dt=case["dt"], t0=0.0)
    return Q, sol


def time_steps(P, case, Q, sol, steps, warmup, world, dev, dist, clocks=None, kernel_events=True):
    """Device-resident timing of `steps` steps: barrier + synchronize on both sides, CUDA events on the
    launching stream.  kernel_events: the library additionally brackets every launch with CUDA events (per
    kernel-class durations for the roofline); those records sit between back-to-back kernels and cost a few
    microseconds per stage, so `value` is taken from a pass without them and the kernel durations from a
    second pass of the same steps right after it."""
    import numpy as np
    import torch
    dg = case["dg"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sol.dostep(Q, 0.0, nsteps=max(warmup, 3))
    barrier()
    norm0 = P.norm(Q)
    if clocks is not None:
        # keep the GPU busy until the sampler has produced its first lines (nvidia-smi start-up); the
        # decision is taken collectively: every rank must run the same number of steps, or the halo
        # exchanges would no longer pair up
        t_w = time.perf_counter()
        while True:
            more = torch.tensor([1.0 if (len(clocks.samples) < 2 and time.perf_counter() - t_w < 3.0) else 0.0],
                                device=dev)
            if world > 1:
                dist.all_reduce(more, op=dist.ReduceOp.MIN)
            if float(more) == 0.0:
                break
            sol.dostep(Q, 0.0, nsteps=5)
            torch.cuda.synchronize()
    dg.set_timing(bool(kernel_events))
    l0 = dg.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if world > 1:
        # the ranks leave the host barrier up to milliseconds apart, and a rank that starts early then waits for its
        # neighbours inside the first halo -- with K = 20 that skew was 5 % of the timed region at 8 GPUs.  A
        # stream-ordered all-reduce right before the start event aligns the DEVICE timelines of the ranks (the host
        # barrier + synchronize on both sides of the timed region stay as they are).
        dist.all_reduce(torch.zeros(1, device=dev))
    tc0 = time.perf_counter()
    e0.record()
    sol.dostep(Q, 0.0, nsteps=steps)
    e1.record()
    barrier()
    tc1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    launches = dg.kernel_launches() - l0
    kern_ms, kern_n = dg.last_kernel_ms() if kernel_events else (0.0, 0)
    classes = dg.kernel_class_ms() if kernel_events else {}
    dg.set_timing(False)
    norm1 = P.norm(Q)
    assert np.isfinite(norm1), "state blew up"
    return dict(ms=ms, launches=launches, kern_ms=kern_ms, kern_n=kern_n, classes=classes,
                norm_ratio=norm1 / norm0, tc=(tc0, tc1))


def reduce_max_sum(vals, world, dev, dist):
    import torch
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    if world == 1:
        return list(vals), list(vals)
    tmax, tsum = t.clone(), t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    return [float(x) for x in tmax], [float(x) for x in tsum]


def weak_box(world):
    """Horizontal replication (fx, fy) of the per-GPU box for the weak-scaling series of the box workloads (ocean,
    rising bubble): as square as a power of two allows -- 1 x 1, 2 x 1, 2 x 2, 4 x 2 for 1, 2, 4, 8 GPUs.  The
    reference's partition is a Hilbert curve over the horizontal columns (kept, BrickMesh.jl:449-522): on a box that
    only grows in x (round 2 until its last day: 160 x 20 columns at 8 GPUs) it cuts elongated, fragmented parts --
    8 150 of 20 000 elements exterior and six neighbours on the worst rank, against 1 950 and three on the 2 x 2 box."""
    fx = 1
    while fx * fx < world:
        fx *= 2
    if world % fx:
        return world, 1
    return fx, world // fx


def workload_name(workload, ne, nvert, world, hyper=False):
    if workload == "baroclinic_wave":
        return f"dry baroclinic wave, cubed sphere ne={ne} x {nvert} vertical, N=4, Rusanov, LSRK54, dt=0.4 s"
    if workload == "held_suarez":
        return ("Held-Suarez dry GCM + SmagorinskyLilly(0.21) (gradient pass + viscous fluxes, horizontal diffusion "
                "direction), sources Gravity/Coriolis/HeldSuarezForcing/RayleighSponge, cubed sphere "
                f"ne={ne} x {nvert}, N=4, Rusanov, LSRK54, dt=0.4 s")
    if workload == "ocean_gyre":
        fx, fy = weak_box(world)
        return (f"OceanBoxGCM HBModel ocean gyre, {ne * fx}x{ne * fy}x{nvert} elements, N=4, Rusanov, LSRK144 "
                "(a step = 14 stages), dt=55 s")
    if workload == "rising_bubble":
        fx, fy = weak_box(world)
        return (f"rising thermal bubble LES (tutorials/Atmos/risingbubble.jl as shipped: SmagorinskyLilly, NTracers{{4}}, "
                f"DryAdiabaticProfile), {ne * fx}x{max(ne // 4, 1) * fy}x{nvert} elements of 500 m, N=4, Rusanov, LSRK144 "
                "(a step = 14 stages), dt=0.4 s")
    return f"isentropic vortex, periodic box {ne}^3, N=4, Rusanov, LSRK54"


def default_mesh(workload, world, args):
    if workload == "rising_bubble":
        return (args.ne or 64), (20 if args.nvert == 10 else args.nvert)
    if workload == "ocean_gyre":
        return (args.ne or 20), (50 if args.nvert == 10 else args.nvert)
    if workload in ("baroclinic_wave", "held_suarez"):
        return (args.ne or WEAK_NE.get(world, int(round(32 * world ** 0.5)))), args.nvert
    return (args.ne or int(round(64 * world ** (1 / 3)))), args.nvert


def secondary_run(P, args, workload, rank, world, dev, dist, steps, grid=None, skip=True, hyper=False,
                  aux0=None):
    """One more workload at the same N (device-resident timing only): returns rank 0's summary dict."""
    import torch
    ne, nvert = default_mesh(workload, world, args)
    case = build_case(P, workload, ne, nvert, rank, world, dev, hyper=hyper, skip_zero_viscosity=skip, grid=grid)
    if aux0 is not None and aux0.data.shape == case["aux"].data.shape:
        case["aux0"] = aux0
    Q, sol = init_case(P, case, workload, rank, world, dist)
    ocean = workload == "ocean_gyre"
    nstate, nstage = (4, 14) if ocean else ((9, 14) if workload == "rising_bubble" else (5, 5))
    r = time_steps(P, case, Q, sol, steps, 3, world, dev, dist, kernel_events=False)
    rk = time_steps(P, case, Q, sol, steps, 0, world, dev, dist, kernel_events=True)
    r["kern_ms"], r["kern_n"], r["classes"] = rk["kern_ms"], rk["kern_n"], rk["classes"]
    nodes_local = case["grid"].nrealelem * NP
    cls = r["classes"]
    other = [k for k in cls if k not in ("tendency", "gradient", "hyper_divergence", "hyper_flux")]
    red, _ = reduce_max_sum(
        [r["ms"], r["kern_ms"], max(cls["gradient"][0], 0.0), max(cls["hyper_divergence"][0], 0.0),
         max(cls["hyper_flux"][0], 0.0)] + [max(cls[k][0], 0.0) for k in other], world, dev, dist)
    ms, kern_ms, g_ms, hd_ms, hf_ms = red[:5]
    (_,), (nodes,) = reduce_max_sum([float(nodes_local)], world, dev, dist)
    evals = nstage * steps
    b_eval, b_launch = algorithmic_bytes_per_node(workload)
    peak = hbm_peak()[0]
    out = {"workload": workload_name(workload, ne, nvert, world), "value": nodes * nstate * evals / (ms * 1e-3) / 1e9,
           "unit": "GDOF/s", "steps": steps, "ms_per_step": ms / steps, "steps_per_s": steps / (ms * 1e-3),
           "nelem_total": int(nodes / NP), "skip_zero_viscosity": bool(case["skip"]),
           "kernel_ms_per_stage": {"tendency": kern_ms / evals},
           "roofline_tendency_frac": nodes_local * b_launch / (kern_ms / evals * 1e-3) / 1e9 / peak,
           "gpu_launches": int(r["launches"]), "norm_ratio": r["norm_ratio"]}
    if g_ms > 0:
        out["kernel_ms_per_stage"]["gradient"] = g_ms / evals
        out["roofline_gradient_frac"] = nodes_local * GRADIENT_BYTES_PER_NODE / (g_ms / evals * 1e-3) / 1e9 / peak
    if hd_ms > 0:
        out["kernel_ms_per_stage"]["hyper_divergence"] = hd_ms / evals
        out["kernel_ms_per_stage"]["hyper_flux"] = hf_ms / evals
    for k, v in zip(other, red[5:]):
        if v > 0:
            out["kernel_ms_per_stage"][k] = v / evals
            if k in TRACER_BYTES_PER_NODE:
                out[f"roofline_{k}_frac"] = nodes_local * TRACER_BYTES_PER_NODE[k] / (v / evals * 1e-3) / 1e9 / peak
    host = None
    if args.keep_host_copy:
        host = (case, Q)
    else:
        case["dg"].close()
        del case, Q, sol
        torch.cuda.empty_cache()
    return out, host


def bind_to_gpu_numa(local):
    """One process per GPU: run on the cores that are local to this rank's GPU and prefer that NUMA node for this
    process's memory, so that the pinned host buffers of the end-to-end leg and the copy engines' DMA stay on the
    GPU's side of the socket interconnect.  Round 1 left every rank on the default policy and the 8-rank end-to-end
    step piled up on one socket's memory.  When the container's cpuset has no core on the GPU's node (an 8-GPU box
    that hands out the cores of one socket only) the affinity cannot follow the GPU, but the memory policy still can:
    set_mempolicy(MPOL_PREFERRED, node) -- it falls back silently when the node is not allowed."""
    info = {"numa_cpus": None}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:          # 00000000:1B:00.0 -> 0000:1b:00.0
            bus = bus[4:]
        info["pci"] = bus
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            if a:
                cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        info["numa_cpus"] = len(cpus)
        if cpus:
            os.sched_setaffinity(0, cpus)
        node = -1
        try:
            with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
                node = int(f.read().strip())
        except Exception:
            pass
        if node >= 0:
            import ctypes
            libc = ctypes.CDLL(None, use_errno=True)
            mask = (ctypes.c_ulong * 16)()
            mask[node // 64] = 1 << (node % 64)
            MPOL_PREFERRED, SYS_set_mempolicy = 1, 238      # x86_64
            rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), 16 * 64 + 1)
            info["mem_node"] = node if rc == 0 else "errno %d" % ctypes.get_errno()
    except Exception as e:       # no NUMA information: keep the default placement
        info["why"] = str(e)[:80]
    return info


def hbm_peak():
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_file):
        return json.load(open(peaks_file))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    P = ge.load_package()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun)"
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    numa = bind_to_gpu_numa(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    args.keep_host_copy = False
    extras = args.workload == "baroclinic_wave" and not args.hyperdiffusion and not args.ne and not args.headline_only

    # ---- untimed parity block (tests/bench_checks.py: the oracle is the checker, never the thing timed) ----
    parity = None
    if not args.no_parity:
        from tests import bench_checks
        parity = bench_checks.parity_block(rank, world, dev)

    ocean = args.workload == "ocean_gyre"
    nstate, nstage = (4, 14) if ocean else ((9, 14) if args.workload == "rising_bubble" else (5, 5))
    ne, nvert = default_mesh(args.workload, world, args)
    args.nvert = nvert
    if args.emulate_rank:
        er, ew = [int(x) for x in args.emulate_rank.split("/")]
        ne = args.ne or WEAK_NE.get(ew, ne)
        case = build_case(P, args.workload, ne, nvert, er, ew, dev, hyper=args.hyperdiffusion)
    else:
        case = build_case(P, args.workload, ne, nvert, rank, world, dev, hyper=args.hyperdiffusion)
    dg, grid = case["dg"], case["grid"]
    Q, sol = init_case(P, case, args.workload, rank, world, dist)
    if args.emulate_rank:
        ng = grid.nelem - grid.nrealelem
        Q.data[grid.nrealelem:] = Q.data[:ng]
        case["aux"].data[grid.nrealelem:] = case["aux"].data[:ng]
    nreal = grid.nrealelem
    nodes_local = nreal * NP

    # ---- device-resident timing -------------------------------------------------------
    clocks = ClockSampler(local)
    clocks.start()
    r = time_steps(P, case, Q, sol, args.steps, args.warmup, world, dev, dist, clocks=clocks, kernel_events=False)
    rk = time_steps(P, case, Q, sol, args.steps, 0, world, dev, dist, kernel_events=True)
    clk = clocks.stop(r["tc"][0], rk["tc"][1])
    r["kern_ms"], r["kern_n"], r["classes"], r["ms_events"] = rk["kern_ms"], rk["kern_n"], rk["classes"], rk["ms"]
    sustained = None
    if args.steps < 100 and not args.headline_only:
        r100 = time_steps(P, case, Q, sol, 100, 0, world, dev, dist, kernel_events=False)
        sustained = r100

    # ---- end to end through host buffers ------------------------------------------------
    e2e_steps = max(2, min(args.steps, 10))
    Qh = torch.empty((nreal, nstate, NP), dtype=torch.float64).pin_memory()
    Qh.copy_(Q.realdata)
    sol.dostep_host(Qh, 0.0, nsteps=1)   # warm up (allocates the library's device state)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sol.dostep_host(Qh, 0.0, nsteps=1)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_local_s = e2e_s

    vals = [r["ms"], e2e_s, float(r["kern_ms"]), sustained["ms"] if sustained else 0.0,
            float(clk["sm_mhz"] or 0.0)]
    tmax, _ = reduce_max_sum(vals, world, dev, dist)
    _, (nodes,) = reduce_max_sum([float(nodes_local)], world, dev, dist)
    ms, e2e_s, kern_ms, sus_ms = tmax[0], tmax[1], tmax[2], tmax[3]
    clocks_per_rank = None
    if world > 1:
        allc = [None] * world
        dist.all_gather_object(allc, {"rank": rank, "sm_mhz": clk["sm_mhz"], "power_w_max": clk.get("power_w_max"),
                                      "reasons": clk["reasons"], "kernel_ms_per_stage": r["kern_ms"] / (nstage * args.steps),
                                      "host_placement": numa, "e2e_ms_per_step": e2e_local_s / e2e_steps * 1e3})
        clocks_per_rank = allc

    # ---- full-size parity + same-mesh CPU baseline (N = 1) ----------------------------------------
    fullsize, cpu_b = None, None
    if world == 1 and not ocean and args.workload != "vortex" and not args.hyperdiffusion:
        from tests import bench_checks
        host = bench_checks.host_arrays(case, Q)
        if not args.no_parity:
            fullsize = bench_checks.fullsize_parity(P, case, Q, host)
        if not args.no_cpu_baseline:
            cpu_b = cpu_baseline_from_host(P, case, host, args.workload, ne, nvert, args.cpu_budget)
        del host
    elif world == 1 and ocean and not args.no_cpu_baseline:
        try:
            cpu_b = ocean_cpu_baseline(P, ne, nvert, args.cpu_budget)
        except Exception as e:
            cpu_b = {"error": str(e)[:200]}

    # ---- the other schedules / configs at the same N ---------------------------------------------
    ref_sched, secondary = None, None
    if extras:
        keep_grid = grid
        aux0 = case.get("aux0")
        dg.close()
        del case, dg, sol, Q, Qh
        torch.cuda.empty_cache()
        ssteps = max(5, min(args.steps, 20))
        ref_sched, _ = secondary_run(P, args, "baroclinic_wave", rank, world, dev, dist, ssteps, grid=keep_grid,
                                     skip=False, aux0=aux0)
        secondary = {}
        if world == 1 and not args.no_cpu_baseline:
            args.keep_host_copy = True
        hs, keep = secondary_run(P, args, "held_suarez", rank, world, dev, dist, ssteps, grid=keep_grid)
        args.keep_host_copy = False
        if keep is not None:
            from tests import bench_checks
            hcase, hQ = keep
            host = bench_checks.host_arrays(hcase, hQ)
            if not args.no_parity:
                hs["parity_fullsize"] = bench_checks.fullsize_parity(P, hcase, hQ, host)
                hs["parity_fullsize_rel_l2"] = hs["parity_fullsize"]["tendency_rel_l2"]
            hs["cpu_baseline"] = cpu_baseline_from_host(P, hcase, host, "held_suarez", ne, nvert,
                                                        min(args.cpu_budget, 10.0))
            hcase["dg"].close()
            del host, hcase, hQ, keep
            torch.cuda.empty_cache()
        secondary["held_suarez"] = hs
        del keep_grid
        torch.cuda.empty_cache()
        secondary["ocean_gyre"], _ = secondary_run(P, args, "ocean_gyre", rank, world, dev, dist, max(2, ssteps // 4))
        if world == 1 and not args.no_cpu_baseline:
            try:
                secondary["ocean_gyre"]["cpu_baseline"] = ocean_cpu_baseline(
                    P, *default_mesh("ocean_gyre", 1, args), min(args.cpu_budget, 8.0))
            except Exception as e:      # the baseline is reported beside the GPU number, never in its way
                secondary["ocean_gyre"]["cpu_baseline"] = {"error": str(e)[:200]}
        secondary["rising_bubble_tracers"], _ = secondary_run(P, args, "rising_bubble", rank, world, dev, dist,
                                                              max(2, ssteps // 4))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    evals = nstage * args.steps
    dof = nodes * nstate
    value = dof * evals / (ms * 1e-3) / 1e9
    b_eval, b_launch_node = algorithmic_bytes_per_node(args.workload)
    peak, peak_src = hbm_peak()
    launches_per_stage = r["kern_n"] / evals if evals else 1
    # with N > 1 a stage is split in an exterior and an interior launch that run CONCURRENTLY on two streams:
    # the sum of their durations is not a wall time, so the roofline fraction is only stated for N = 1
    stage_ms = kern_ms / evals
    achieved = nodes_local * b_launch_node / (stage_ms * 1e-3) / 1e9
    traffic = None
    tf = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tf):
        try:
            traffic = json.load(open(tf)).get(args.workload, {}).get("dram_bytes_per_node")
            traffic = traffic * nodes_local if traffic else None
        except Exception:
            traffic = None
    second = ocean or args.workload == "held_suarez" or args.hyperdiffusion
    out = {
        "metric": METRIC,
        "value": value, "unit": "GDOF/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "lsrk54_steps_per_s": args.steps / (ms * 1e-3),
        "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, ne, nvert, world),
                   "hyperdiffusion": ("DryBiharmonic(8 h), horizontal (3 extra kernels + 2 extra exchanges per evaluation)"
                                      if args.hyperdiffusion else "off"),
                   "nelem_total": int(nodes / NP), "dof_total": int(dof),
                   "cache": "inputs larger than L2 (Q+dQ+Qout+aux+geometry = %.0f MB per GPU vs 126 MB L2)"
                            % (nodes_local * 8 * (3 * nstate + 16 + 10 + 4.8 + (10 if ocean else 0)) / 1e6),
                   "skip_zero_viscosity": not second, "parallelism": f"element partition x{world}"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": TRAFFIC_SOURCE,
                     "peak_source": peak_src,
                     "kernel": "hb_tendency_kernel<double,5,RUSANOV>" if ocean else "dg_tendency_kernel<double,5,RUSANOV,...>",
                     "algorithmic_bytes_per_node_per_launch": b_launch_node,
                     "kernel_ms_per_stage": stage_ms, "launches_per_stage": launches_per_stage,
                     "timing": ("per-launch CUDA events in a second pass of the same %d steps right after the pass `value` "
                                "is taken from (the event records between back-to-back kernels cost a few us per stage): "
                                "%.4f ms/step with them" % (args.steps, r["ms_events"] / args.steps))},
        "e2e": {"value": dof * nstage * e2e_steps / e2e_s / 1e9, "unit": "GDOF/s",
                "h2d_bytes_per_step": int(nodes_local * nstate * 8),
                "d2h_bytes_per_step": int(nodes_local * nstate * 8),
                "api": "cmdg_lsrk_steps_host (pinned host state in/out every step)",
                "ms_per_step": e2e_s / e2e_steps * 1e3, "host_placement": numa},
        "gpu_launches": int(r["launches"]),
        "clocks": clk,
        "norm_ratio": r["norm_ratio"],
    }
    if world > 1:
        out["roofline"]["note"] = ("N > 1: kernel_ms_per_stage sums the exterior and interior launches, which overlap "
                                   "on two streams; frac is a lower bound, see the N = 1 line for the kernel's roofline")
        out["clocks_per_rank"] = clocks_per_rank
    if sustained:
        out["sustained_100"] = {"value": dof * nstage * 100 / (sus_ms * 1e-3) / 1e9, "unit": "GDOF/s",
                                "ms_per_step": sus_ms / 100, "lsrk54_steps_per_s": 100 / (sus_ms * 1e-3)}
    if parity is not None:
        out["parity"] = parity
    if fullsize is not None:
        out["parity_fullsize_rel_l2"] = fullsize["tendency_rel_l2"]
        out["parity_fullsize"] = fullsize
    if ref_sched is not None:
        out["reference_schedule"] = ref_sched
    if secondary is not None:
        out["secondary"] = secondary
    if cpu_b is not None:
        out["cpu_baseline"] = cpu_b
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------
# restated reference CPU path (oracle/c/dg_ref.c) -- the only place bench.py touches oracle/
# ----------------------------------------------------------------------------------------
def ref_params_for(P, model, dg_skip):
    """cref.ref_params of a package AtmosModel (the C twin's view of the same balance law)."""
    from oracle import cref
    from climatemachine_jl_b200 import atmos_init as ai
    lay = ai.aux_layout(model)
    p = model.param_set
    R = cref.ref_params()
    R.R_d, R.cp_d, R.cv_d, R.T_0, R.MSLP, R.grav, R.Omega = p.R_d, p.cp_d, p.cv_d, p.T_0, p.MSLP, p.grav, p.Omega
    R.naux = lay["A"]
    R.a_Phi, R.a_gradPhi = lay.get("Φ", -1), lay.get("∇Φ", -1)
    R.a_ref_rho, R.a_ref_p = lay.get("ref_ρ", -1), lay.get("ref_p", -1)
    R.a_theta_v, R.a_T = lay["θ_v"], lay["T"]
    R.subtract_off = int(isinstance(model.ref_state, P.HydrostaticState) and model.ref_state.subtract_off)
    R.gravity = int(any(isinstance(s, P.Gravity) for s in model.source))
    R.coriolis = int(any(isinstance(s, P.Coriolis) for s in model.source))
    R.nf_first = 0
    for i, bc in enumerate(model.boundaryconditions):
        R.bc_kind[i] = 1 if isinstance(bc.momentum.drag, P.FreeSlip) else 2
    R.second_order = int(not dg_skip)
    t = model.turbulence
    if isinstance(t, P.SmagorinskyLilly):
        R.turbulence, R.turb_param, R.ngradflux = 2, t.C_smag, 10
    elif isinstance(t, P.ConstantKinematicViscosity):
        R.turbulence, R.turb_param, R.with_divergence, R.ngradflux = 0, t.ν, int(t.with_divergence), 9
    else:
        R.turbulence, R.turb_param, R.with_divergence, R.ngradflux = 1, t.ρν, int(t.with_divergence), 9
    R.horizontal_diffusion = 1      # diffdir = HorizontalDirection() in the GCM drivers
    R.a_Delta = lay.get("Δ", -1)
    R.inv_Pr_turb, R.day = p.inv_Pr_turb, p.day
    for s in model.source:
        if isinstance(s, P.HeldSuarezForcing):
            R.held_suarez = 1
        if isinstance(s, P.RayleighSponge):
            R.sponge = 1
            R.sponge_z_max, R.sponge_z_sponge, R.sponge_alpha_max, R.sponge_gamma = s.z_max, s.z_sponge, s.α_max, s.γ
            for i in range(3):
                R.sponge_u[i] = s.u_relaxation[i]
    return R


def time_cpu(c, cref, Q, dQ, aux, dt, budget_s=None, steps=None, warmup=1):
    import numpy as np
    from oracle import odesolvers as oode
    rka = [float(x) for x in oode._conv(np.float64, oode.LSRK54_RKA)]
    rkb = [float(x) for x in oode._conv(np.float64, oode.LSRK54_RKB)]
    cref.use_all_cores()     # torchrun exports OMP_NUM_THREADS=1
    t0 = time.perf_counter()
    c.lsrk_steps(Q, dQ, aux, dt, rka, rkb, max(warmup, 1))
    one = (time.perf_counter() - t0) / max(warmup, 1)
    n = steps if steps is not None else max(1, min(2000, int(budget_s / max(one, 1e-6))))
    t0 = time.perf_counter()
    c.lsrk_steps(Q, dQ, aux, dt, rka, rkb, n)
    el = time.perf_counter() - t0
    assert np.isfinite(Q).all()
    return n, el


def cpu_baseline_from_host(P, case, host, workload, ne, nvert, budget_s):
    """Times the C restatement of the reference's schedule (all host threads) on the SAME mesh and
    state the GPU arm ran (host copies of its arrays), for a bounded number of steps."""
    from oracle import cref
    R = ref_params_for(P, case["model"], case["skip"])
    c = cref.CRefDG.from_arrays(R, host["vgeo"], host["sgeo"], host["vmapM"], host["vmapP"], host["elemtobndy"],
                                host["D"], host["nreal"])
    Q, aux = host["Q"].copy(), host["aux"].copy()
    n, el = time_cpu(c, cref, Q, Q * 0, aux, case["dt"], budget_s=budget_s)
    dof = host["nreal"] * NP * NSTATE
    return {"value": dof * 5 * n / el / 1e9, "unit": "GDOF/s", "cores": cref.lib().ref_num_threads(),
            "kind": "port", "same_config": True,
            "sample": (f"{workload}: ne={ne} x {nvert} vertical ({host['nreal']} elements, {dof} DOF) -- the GPU arm's "
                       f"mesh and state --, {n} LSRK54 steps in {el:.1f} s; C/OpenMP restatement of the reference's "
                       "kernel schedule (oracle/c/dg_ref.c), "
                       + ("gradient pass + viscous fluxes included" if R.second_order
                          else "nu=0 gradient pass skipped as in the GPU arm")),
            "ms_per_step": el / n * 1e3}


def ocean_twin(P, ne, nvert):
    """The C/OpenMP restatement of the reference's HBModel schedule (oracle/c/hb_ref.c) on the GPU arm's N = 1
    ocean mesh and initial state (built on the host by the package's builders): (twin, Q, aux, dt, tableau)."""
    import numpy as np
    from oracle import cref
    from climatemachine_jl_b200 import atmos_init as ai, balance_laws as bl
    from climatemachine_jl_b200.dgmodel import _LSRK144
    grid, prob = build_grid(P, "ocean_gyre", ne, nvert, 0, 1, "cpu")
    m = P.HBModel(prob, cʰ=float(np.sqrt(9.81 * prob.H)))
    Q0, aux = ai.ocean_gyre_state(prob, grid)
    vel = {1: "noslip", 2: "freeslip", 3: "penetrable_freeslip", 4: "kinematic_stress"}
    temp = {1: "insulating", 2: "temperature_flux"}
    bcs = [(vel[v], temp[t]) for v, t in (bl.ocean_bc_codes(bc) for bc in prob.boundary_conditions)]
    Pm = cref.hb_params_from(m.param_set.grav, m.ρₒ, m.cʰ, m.cᶻ, m.αᵀ, m.νʰ, m.νᶻ, m.κʰ, m.κᶻ, m.κᶜ, m.fₒ, m.β,
                             prob.Lʸ, prob.τₒ, prob.λʳ, prob.θᴱ, bcs, grid.nvertelem)
    npy = lambda t: np.ascontiguousarray(t.cpu().numpy())
    c = cref.CRefHB(Pm, npy(grid.vgeo), npy(grid.sgeo), npy(grid.vmapM), npy(grid.vmapP), npy(grid.elemtobndy),
                    grid.D_host, npy(grid.Imat).T, P.CutoffFilter(grid, 3).filter_matrix,
                    P.ExponentialFilter(grid, 1, 8).filter_matrix, grid.nrealelem)
    rka = np.array([float(x) for x in _LSRK144[0]])
    rkb = np.array([float(x) for x in _LSRK144[1]])
    return c, npy(Q0), npy(aux.data), 55.0, (rka, rkb), grid.nrealelem


def ocean_cpu_baseline(P, ne, nvert, budget_s):
    """`cpu_baseline` of the ocean workload: LSRK144 steps of the twin, all host threads, until the budget is used."""
    from oracle import cref
    c, Q, aux, dt, (rka, rkb), nreal = ocean_twin(P, ne, nvert)
    cores = cref.use_all_cores_hb()
    dQ = Q * 0
    c.lsrk_steps(Q, dQ, aux, dt, rka, rkb, 1)          # warm-up (page faults, thread pool)
    n, t0 = 0, time.perf_counter()
    while True:
        c.lsrk_steps(Q, dQ, aux, dt, rka, rkb, 1)
        n += 1
        el = time.perf_counter() - t0
        if el >= budget_s or n >= 50:
            break
    dof = nreal * NP * 4
    return {"value": dof * 14 * n / el / 1e9, "unit": "GDOF/s", "cores": cores, "kind": "port", "same_config": True,
            "sample": (f"ocean_gyre: {ne} x {ne} x {nvert} elements ({nreal} elements, {dof} DOF) -- the GPU arm's N = 1 "
                       f"mesh and initial state --, {n} LSRK144 steps (14 evaluations each) in {el:.1f} s; C/OpenMP "
                       "restatement of the reference's HBModel schedule (oracle/c/hb_ref.c: vertical filters, gradient "
                       "pass, stack integrals, tendency)"),
            "ms_per_step": el / n * 1e3}


def run_reference_ocean(args):
    """`--impl reference --workload ocean_gyre`: the ocean twin, one LSRK144 step per bench step."""
    import __graft_entry__ as ge
    from oracle import cref
    P = ge.load_package()
    ne, nvert = default_mesh("ocean_gyre", 1, args)
    c, Q, aux, dt, (rka, rkb), nreal = ocean_twin(P, ne, nvert)
    cores = cref.use_all_cores_hb()
    dQ = Q * 0
    c.lsrk_steps(Q, dQ, aux, dt, rka, rkb, max(args.warmup, 1))
    t0 = time.perf_counter()
    c.lsrk_steps(Q, dQ, aux, dt, rka, rkb, args.steps)
    el = time.perf_counter() - t0
    dof = nreal * NP * 4
    value = dof * 14 * args.steps / el / 1e9
    sample = (f"one LSRK144 step (14 evaluations) per bench step on the GPU arm's N=1 ocean mesh; C/OpenMP restatement "
              f"of the reference's HBModel schedule (Julia is not available; oracle/c/hb_ref.c), {cores} host threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC,
        "value": value, "unit": "GDOF/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": el / args.steps * 1e3,
        "lsrk54_steps_per_s": args.steps / el, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name("ocean_gyre", ne, nvert, 1), "hyperdiffusion": "off",
                   "nelem_total": int(nreal), "dof_total": int(dof),
                   "skip_zero_viscosity": False, "parallelism": f"OpenMP x{cores}"},
        "cpu_baseline": {"value": value, "unit": "GDOF/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


def run_reference(args):
    """`--impl reference`: the restated reference CPU path on the GPU arm's own mesh (ne = 32 x 10 by
    default; the mesh is built on the host by the package's vectorised builder), all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "ocean_gyre":
        return run_reference_ocean(args)
    import numpy as np
    import __graft_entry__ as ge
    from oracle import cref
    P = ge.load_package()
    workload = args.workload if args.workload in ("baroclinic_wave", "held_suarez", "vortex") else "baroclinic_wave"
    ne, nvert = default_mesh(workload, 1, args)
    from climatemachine_jl_b200 import atmos_init as ai
    grid, _ = build_grid(P, workload, ne, nvert, 0, 1, "cpu")
    model = gcm_model(P, workload) if workload != "vortex" else P.AtmosModel()
    skip = workload != "held_suarez"
    if workload == "vortex":
        aux0 = P.MPIStateArray(grid, model.number_states("Auxiliary"))
        aux0.data[:, 0:3] = grid.vgeo[:, 12:15]
        Q0 = ai.isentropic_vortex(model, grid, 0.0)
        dt = (2 * 0.05 / ne) / 347.2 / 16
    else:
        aux0 = ai.init_state_auxiliary(model, grid)
        Q0 = ai.baroclinic_wave(model, grid, aux0)
        dt = 0.4
    R = ref_params_for(P, model, skip)
    if workload == "vortex":
        R.horizontal_diffusion = 0
    npy = lambda t: np.ascontiguousarray(t.numpy())
    c = cref.CRefDG.from_arrays(R, npy(grid.vgeo), npy(grid.sgeo), npy(grid.vmapM), npy(grid.vmapP),
                                npy(grid.elemtobndy), grid.D_host, grid.nrealelem)
    Q = np.zeros((grid.nelem, 5, NP))
    Q[:grid.nrealelem] = npy(Q0)
    n, el = time_cpu(c, cref, Q, np.zeros_like(Q), npy(aux0.data), dt, steps=args.steps, warmup=max(args.warmup, 1))
    nreal = grid.nrealelem
    dof = nreal * NP * NSTATE
    value = dof * 5 * args.steps / el / 1e9
    cores = cref.lib().ref_num_threads()
    sample = (f"one LSRK54 step per bench step on the GPU arm's N=1 mesh; C/OpenMP restatement of the reference's "
              f"kernel schedule (Julia is not available; oracle/c/dg_ref.c), {cores} host threads, "
              + ("gradient pass + viscous fluxes included" if R.second_order else "nu=0 gradient pass skipped as in the GPU arm"))
    print(json.dumps({
        "impl": "reference", "metric": METRIC,
        "value": value, "unit": "GDOF/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": el / args.steps * 1e3,
        "lsrk54_steps_per_s": args.steps / el, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(workload, ne, nvert, 1), "hyperdiffusion": "off",
                   "nelem_total": int(nreal), "dof_total": int(dof),
                   "skip_zero_viscosity": bool(skip), "parallelism": f"OpenMP x{cores}",
                   "sample_of": (None if args.gpus == 1 else
                                 f"one GPU's share of the {args.gpus}-GPU weak-scaling mesh (ne = {WEAK_NE.get(args.gpus, '?')}): "
                                 f"the N = 1 mesh has the same elements per GPU")},
        "cpu_baseline": {"value": value, "unit": "GDOF/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="baroclinic_wave",
                    choices=["baroclinic_wave", "vortex", "ocean_gyre", "held_suarez", "rising_bubble"])
    ap.add_argument("--ne", type=int, default=0, help="horizontal elements per cube edge / box edge")
    ap.add_argument("--nvert", type=int, default=10)
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling (SURVEY 10-3): the named mesh ne = 32 x 10 at every --gpus N instead of the "
                         "weak-scaling series ne = 32 / 45 / 64 / 90")
    ap.add_argument("--hyperdiffusion", action="store_true",
                    help="baroclinic_wave / held_suarez as the reference's drivers ship them: DryBiharmonic(8 h)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed parity blocks")
    ap.add_argument("--headline-only", action="store_true",
                    help="skip reference_schedule / secondary / sustained_100 (profiling runs)")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--emulate-rank", default="", help="diagnostic: 'R/W' builds rank R of a W-rank partition on ONE GPU "
                    "without a communicator (ghost elements hold copies of real ones)")
    args = ap.parse_args()
    if args.strong and not args.ne:
        args.ne = 32
    # stdout carries exactly one JSON line: libraries that print to fd 1 (c10d's "NCCL version ..."
    # banner on the first communicator) are sent to stderr for the duration of the run
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(saved, "w")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
