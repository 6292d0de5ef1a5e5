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
        sm, mx, reasons = [], [], set()
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
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": where}


# ----------------------------------------------------------------------------------------
# libcmdg arm
# ----------------------------------------------------------------------------------------
def build_case(P, workload, ne, nvert, rank, nranks, device, hyper=False):
    import numpy as np
    import torch
    from climatemachine_jl_b200 import topologies as tp, grids as gr, atmos_init as ai
    if workload in ("baroclinic_wave", "held_suarez"):
        ps = P.EarthParameterSet()
        R = np.linspace(ps.planet_radius, ps.planet_radius + 30e3, nvert + 1)
        topo = tp.stacked_cubed_sphere_topology(ne, R, (1, 2), rank, nranks)
        grid = gr.build_grid(topo, 4, torch.float64, tp.cubed_sphere_warp, device)
        if workload == "held_suarez":
            # BASELINE.json configs[3] as tutorials/Atmos/heldsuarez.jl:160-201 sets it, explicit
            # LSRK54, hyperdiffusion off: Smagorinsky(0.21), horizontal diffusion direction,
            # Gravity + Coriolis + HeldSuarezForcing + RayleighSponge(30 km, 12 km, 1/900 s)
            model = P.AtmosModel(orientation=P.SphericalOrientation(),
                                 ref_state=P.HydrostaticState(P.DecayingTemperatureProfile(290.0, 220.0, 8e3)),
                                 turbulence=P.SmagorinskyLilly(0.21),
                                 source=(P.Gravity(), P.Coriolis(), P.HeldSuarezForcing(),
                                         P.RayleighSponge(30e3, 12e3, 1 / 60 / 15, (0.0, 0.0, 0.0), 2.0)),
                                 boundaryconditions=(P.AtmosBC(), P.AtmosBC()),
                                 hyperdiffusion=P.DryBiharmonic(8 * 3600.0) if hyper else None)
            aux = P.MPIStateArray(grid, model.number_states("Auxiliary"))
            dg = P.DGModel(model, grid, P.RusanovNumericalFlux(), P.CentralNumericalFluxSecondOrder(),
                           P.CentralNumericalFluxGradient(), state_auxiliary=aux,
                           diffusion_direction=P.HorizontalDirection(), write_aux_diagnostics=True)
            return dict(topo=topo, grid=grid, model=model, dg=dg, aux=aux, dt=0.4, ai=ai)
        model = P.AtmosModel(orientation=P.SphericalOrientation(),
                             ref_state=P.HydrostaticState(P.DecayingTemperatureProfile(290.0, 220.0, 8e3)),
                             turbulence=P.ConstantKinematicViscosity(0.0),
                             source=(P.Gravity(), P.Coriolis()),
                             boundaryconditions=(P.AtmosBC(), P.AtmosBC()),
                             hyperdiffusion=P.DryBiharmonic(8 * 3600.0) if hyper else None)
        dt = 0.4     # s; vertical acoustic CFL ~0.3
        if hyper:
            # experiments/TestCase/baroclinic_wave.jl:179,258 as shipped: DryBiharmonic(8 h) with the
            # horizontal diffusion direction; the nu = 0 gradient pass cannot be skipped then
            aux = P.MPIStateArray(grid, model.number_states("Auxiliary"))
            dg = P.DGModel(model, grid, P.RusanovNumericalFlux(), P.CentralNumericalFluxSecondOrder(),
                           P.CentralNumericalFluxGradient(), state_auxiliary=aux,
                           diffusion_direction=P.HorizontalDirection(), write_aux_diagnostics=True)
            return dict(topo=topo, grid=grid, model=model, dg=dg, aux=aux, dt=dt, ai=ai)
    elif workload == "ocean_gyre":
        # BASELINE.json configs[4]: OceanBoxGCM HBModel, 20 x 20 x 50 elements per GPU
        # (experiments/OceanBoxGCM/homogeneous_box.jl:11-21), box widened in x with the GPU count
        nx = ne * nranks
        prob = P.OceanGyre(4e6 * nranks, 4e6, 1000.0)
        br = (np.linspace(0, prob.Lˣ, nx + 1), np.linspace(0, prob.Lʸ, ne + 1), np.linspace(-prob.H, 0, nvert + 1))
        topo = tp.stacked_brick_topology(br, (False, False, False), ((1, 1), (1, 1), (2, 3)), rank, nranks)
        grid = gr.build_grid(topo, 4, torch.float64, None, device)
        model = P.HBModel(prob, cʰ=float(np.sqrt(9.81 * prob.H)))
        Q0, aux = ai.ocean_gyre_state(prob, grid)
        md = dict(vert_filter=P.CutoffFilter(grid, 3), exp_filter=P.ExponentialFilter(grid, 1, 8))
        dg = P.DGModel(model, grid, P.RusanovNumericalFlux(), P.CentralNumericalFluxSecondOrder(),
                       P.CentralNumericalFluxGradient(), state_auxiliary=aux, modeldata=md)
        return dict(topo=topo, grid=grid, model=model, dg=dg, aux=aux, dt=55.0, ai=ai, Q0=Q0)
    else:
        L = 0.05
        br = tuple(np.linspace(-L, L, ne + 1) for _ in range(3))
        topo = tp.brick_topology(br, (True, True, True), None, rank, nranks)
        grid = gr.build_grid(topo, 4, torch.float64, None, device)
        model = P.AtmosModel()
        dt = (2 * L / ne) / 347.2 / 16
    aux = P.MPIStateArray(grid, model.number_states("Auxiliary"))
    dg = P.DGModel(model, grid, P.RusanovNumericalFlux(), P.CentralNumericalFluxSecondOrder(),
                   P.CentralNumericalFluxGradient(), state_auxiliary=aux,
                   skip_zero_viscosity=True, write_aux_diagnostics=True)
    return dict(topo=topo, grid=grid, model=model, dg=dg, aux=aux, dt=dt, ai=ai)


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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    ocean = args.workload == "ocean_gyre"
    NSTATE, NSTAGE = (4, 14) if ocean else (5, 5)
    if ocean:
        ne = args.ne or 20
        if args.nvert == 10:
            args.nvert = 50
    else:
        ne = args.ne or (WEAK_NE.get(world, int(round(32 * world ** 0.5)))
                         if args.workload in ("baroclinic_wave", "held_suarez")
                         else int(round(64 * world ** (1 / 3))))
    case = build_case(P, args.workload, ne, args.nvert, rank, world, dev, hyper=args.hyperdiffusion)
    dg, grid, model, ai = case["dg"], case["grid"], case["model"], case["ai"]
    if world > 1:
        uid = [P.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        dg.comm_init(uid[0], rank, world)
    # auxiliary state and initial condition (setup; harness-side, torch on the device)
    Q = P.MPIStateArray(grid, NSTATE)
    if ocean:
        Q.data[:grid.nrealelem] = case["Q0"]
        if world > 1:
            dg.ghost_exchange(case["aux"])
        sol = P.LSRK144NiegemannDiehlBusch(dg, Q, dt=case["dt"], t0=0.0)
    else:
        ex = (lambda arr: dg.ghost_exchange(arr)) if world > 1 else None
        aux0 = ai.init_state_auxiliary(model, grid, exchange=ex)
        case["aux"].data.copy_(aux0.data)
        if args.workload in ("baroclinic_wave", "held_suarez"):
            # (Held-Suarez starts from rest + noise in the tutorial; the baroclinic-wave state gives
            # the friction, relaxation and sponge terms something to act on -- synthetic either way)
            Q.data[:grid.nrealelem] = ai.baroclinic_wave(model, grid, case["aux"])
        else:
            Q.data[:grid.nrealelem] = ai.isentropic_vortex(model, grid, 0.0)
        sol = P.LSRK54CarpenterKennedy(dg, Q, dt=case["dt"], t0=0.0)
    nreal = grid.nrealelem
    nodes_local = nreal * NP

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------
    clocks = ClockSampler(local)
    clocks.start()
    sol.dostep(Q, 0.0, nsteps=max(args.warmup, 3))
    barrier()
    norm0 = P.norm(Q)
    # keep the GPU busy until the sampler has produced its first lines (nvidia-smi start-up)
    # (the decision is taken collectively: every rank must run the same number of steps, or the
    # halo exchanges would no longer pair up)
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
    dg.set_timing(not os.environ.get("BENCH_NO_KERNEL_TIMING"))
    l0 = dg.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    tc0 = time.perf_counter()
    e0.record()
    sol.dostep(Q, 0.0, nsteps=args.steps)
    e1.record()
    barrier()
    tc1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    launches = dg.kernel_launches() - l0
    kern_ms, kern_n = dg.last_kernel_ms()
    dg.set_timing(False)
    clk = clocks.stop(tc0, tc1)
    norm1 = P.norm(Q)
    assert np.isfinite(norm1), "state blew up"

    # ---- end to end through host buffers ------------------------------------------------
    e2e_steps = max(2, min(args.steps, 10))
    Qh = torch.empty((nreal, NSTATE, NP), dtype=torch.float64).pin_memory()
    Qh.copy_(Q.realdata)
    sol.dostep_host(Qh, 0.0, nsteps=1)   # warm up (allocates the library's device state)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sol.dostep_host(Qh, 0.0, nsteps=1)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    t = torch.tensor([ms, e2e_s, float(nodes_local), float(kern_ms), float(kern_n)],
                     dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, e2e_s, kern_ms = float(tmax[0]), float(tmax[1]), float(tmax[3])
        nodes = float(tsum[2])
    else:
        nodes = float(nodes_local)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    evals = NSTAGE * args.steps
    dof = nodes * NSTATE
    value = dof * evals / (ms * 1e-3) / 1e9
    b_eval, b_launch_node = algorithmic_bytes_per_node(args.workload)
    # roofline of the dominant kernel (dg_tendency_kernel), rank-0 launch durations
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_file):
        peak, peak_src = json.load(open(peaks_file))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    per_launch_ms = kern_ms / max(kern_n, 1)
    launches_per_stage = kern_n / evals if evals else 1
    # with N > 1 a stage is split in an exterior and an interior launch: use time per stage
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
    out = {
        "metric": "DG tendency GDOF/s (fused LSRK54 stage; steps/s in lsrk54_steps_per_s)",
        "value": value, "unit": "GDOF/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "lsrk54_steps_per_s": args.steps / (ms * 1e-3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": (f"dry baroclinic wave, cubed sphere ne={ne} x {args.nvert} vertical, N=4, "
                                "Rusanov, LSRK54, dt=0.4 s" if args.workload == "baroclinic_wave"
                                else f"Held-Suarez dry GCM + SmagorinskyLilly(0.21) (gradient pass + viscous fluxes, horizontal "
                                f"diffusion direction), sources Gravity/Coriolis/HeldSuarezForcing/RayleighSponge, cubed sphere "
                                f"ne={ne} x {args.nvert}, N=4, Rusanov, LSRK54, dt=0.4 s"
                                if args.workload == "held_suarez"
                                else f"OceanBoxGCM HBModel ocean gyre, {ne * world}x{ne}x{args.nvert} elements, N=4, "
                                "Rusanov, LSRK144 (a step = 14 stages), dt=55 s" if ocean
                                else f"isentropic vortex, periodic box {ne}^3, N=4, Rusanov, LSRK54"),
                   "hyperdiffusion": ("DryBiharmonic(8 h), horizontal (3 extra kernels + 2 extra exchanges per evaluation)"
                                      if args.hyperdiffusion else "off"),
                   "nelem_total": int(nodes / NP), "dof_total": int(dof),
                   "cache": "inputs larger than L2 (Q+dQ+Qout+aux+geometry = %.0f MB per GPU vs 126 MB L2)"
                            % (nodes_local * 8 * (3 * NSTATE + case["aux"].nstate + 10 + 4.8 + (10 if ocean else 0)) / 1e6),
                   "skip_zero_viscosity": not ocean and args.workload != "held_suarez" and not args.hyperdiffusion, "parallelism": f"element partition x{world}"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "kernel": "hb_tendency_kernel<double,5,RUSANOV>" if ocean else "dg_tendency_kernel<double,5,RUSANOV,...>",
                     "algorithmic_bytes_per_node_per_launch": b_launch_node,
                     "kernel_ms_per_stage": stage_ms, "launches_per_stage": launches_per_stage},
        "e2e": {"value": dof * NSTAGE * e2e_steps / e2e_s / 1e9, "unit": "GDOF/s",
                "h2d_bytes_per_step": int(nodes_local * NSTATE * 8),
                "d2h_bytes_per_step": int(nodes_local * NSTATE * 8),
                "api": "cmdg_lsrk_steps_host (pinned host state in/out every step)",
                "ms_per_step": e2e_s / e2e_steps * 1e3},
        "gpu_launches": int(launches),
        "clocks": clk,
        "norm_ratio": norm1 / norm0,
    }
    if world == 1 and not args.no_cpu_baseline and not ocean and args.workload != "held_suarez" \
            and not args.hyperdiffusion:
        out["cpu_baseline"] = cpu_baseline(args.workload, budget_s=args.cpu_budget)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------
# restated reference CPU path (oracle/c/dg_ref.c) -- the only place bench.py touches oracle/
# ----------------------------------------------------------------------------------------
def cpu_case(workload, ne, nvert):
    import numpy as np
    from oracle import topologies as otp, grids as ogrids, atmos as oatmos, dgmodel as odg, cref
    from oracle import odesolvers as oode
    if workload == "baroclinic_wave":
        ps = oatmos.Params()
        a = float(ps.planet_radius)
        topo = otp.StackedCubedSphereTopology(1, ne, np.linspace(a, a + 30e3, nvert + 1), boundary=(1, 2))[0]
        g = ogrids.Grid(topo, 4, meshwarp=otp.equiangular_cubed_sphere_warp)
        model = oatmos.DryAtmosModel(np.float64, orientation="spherical",
                                     ref_state=dict(T_surf=290.0, T_min=220.0, H_t=8e3, subtract_off=True),
                                     turbulence=("constant_kinematic", 0.0, False),
                                     sources=("gravity", "coriolis"), bcs=("freeslip", "freeslip"))
        dgm = odg.DGModel(model, [g], "rusanov", skip_zero_viscosity=True)
        aux = dgm.state_auxiliary[0].data
        Q0 = oatmos.init_baroclinic_wave(model, np.moveaxis(aux[:g.nreal], 1, 0))
        dt = 0.4
    else:
        L = 0.05
        br = tuple(np.linspace(-L, L, ne + 1) for _ in range(3))
        topo = otp.BrickTopology(1, br, periodicity=(True, True, True))[0]
        g = ogrids.Grid(topo, 4)
        model = oatmos.DryAtmosModel(np.float64)
        dgm = odg.DGModel(model, [g], "rusanov", skip_zero_viscosity=True)
        aux = dgm.state_auxiliary[0].data
        setup = oatmos.IsentropicVortexSetup(oatmos.Params())
        Q0 = setup(g.vgeo[:g.nreal, 12], g.vgeo[:g.nreal, 13], g.vgeo[:g.nreal, 14], np.float64(0))
        dt = (2 * L / ne) / 347.2 / 16
    Q = np.zeros((g.nelem, 5, 125))
    Q[:g.nreal] = np.moveaxis(Q0, 0, 1)
    rka = [float(x) for x in oode._conv(np.float64, oode.LSRK54_RKA)]
    rkb = [float(x) for x in oode._conv(np.float64, oode.LSRK54_RKB)]
    return cref.CRefDG(model, g, "rusanov"), Q, np.zeros_like(Q), aux.copy(), dt, rka, rkb, g.nreal, cref


def cpu_baseline(workload, budget_s=15.0, ne=None, nvert=10):
    """Times the C restatement of the reference's schedule (all host threads) on a bounded
    sample of the same workload: a coarser horizontal mesh with the same vertical stack."""
    ne = ne or (8 if workload == "baroclinic_wave" else 12)
    c, Q, dQ, aux, dt, rka, rkb, nreal, cref = cpu_case(workload, ne, nvert)
    cref.use_all_cores()
    c.lsrk_steps(Q, dQ, aux, dt, rka, rkb, 1)       # warm up
    t0 = time.perf_counter()
    c.lsrk_steps(Q, dQ, aux, dt, rka, rkb, 1)
    one = time.perf_counter() - t0
    n = max(1, min(2000, int(budget_s / max(one, 1e-6))))
    t0 = time.perf_counter()
    c.lsrk_steps(Q, dQ, aux, dt, rka, rkb, n)
    el = time.perf_counter() - t0
    dof = nreal * NP * NSTATE
    return {"value": dof * 5 * n / el / 1e9, "unit": "GDOF/s", "cores": cref.lib().ref_num_threads(),
            "kind": "port",
            "sample": (f"{workload}: ne={ne} x {nvert} vertical ({nreal} elements, {dof} DOF), "
                       f"{n} LSRK54 steps in {el:.1f} s; C/OpenMP restatement of the reference's "
                       "kernel schedule (oracle/c/dg_ref.c), nu=0 gradient pass skipped as in the GPU arm"),
            "ms_per_step": el / n * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    ne = args.ne or (8 if args.workload == "baroclinic_wave" else 12)
    c, Q, dQ, aux, dt, rka, rkb, nreal, cref = cpu_case(args.workload, ne, args.nvert)
    cref.use_all_cores()     # torchrun exports OMP_NUM_THREADS=1
    c.lsrk_steps(Q, dQ, aux, dt, rka, rkb, max(args.warmup, 1))
    t0 = time.perf_counter()
    c.lsrk_steps(Q, dQ, aux, dt, rka, rkb, args.steps)
    el = time.perf_counter() - t0
    assert np.isfinite(Q).all()
    dof = nreal * NP * NSTATE
    value = dof * 5 * args.steps / el / 1e9
    cores = cref.lib().ref_num_threads()
    sample = (f"{args.workload}: ne={ne} x {args.nvert} vertical ({nreal} elements), one LSRK54 step per "
              "bench step; C/OpenMP restatement of the reference's kernel schedule (Julia is not "
              "available; oracle/c/dg_ref.c)")
    print(json.dumps({
        "impl": "reference", "metric": "DG tendency GDOF/s (fused LSRK54 stage; steps/s in lsrk54_steps_per_s)",
        "value": value, "unit": "GDOF/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": el / args.steps * 1e3,
        "lsrk54_steps_per_s": args.steps / el, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": sample, "nelem_total": int(nreal), "dof_total": int(dof)},
        "cpu_baseline": {"value": value, "unit": "GDOF/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="baroclinic_wave", choices=["baroclinic_wave", "vortex", "ocean_gyre", "held_suarez"])
    ap.add_argument("--ne", type=int, default=0, help="horizontal elements per cube edge / box edge")
    ap.add_argument("--nvert", type=int, default=10)
    ap.add_argument("--hyperdiffusion", action="store_true",
                    help="baroclinic_wave / held_suarez as the reference's drivers ship them: DryBiharmonic(8 h)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()
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
