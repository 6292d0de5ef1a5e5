#!/bin/bash
# A/B: shipped build vs libcmdg$1.so (interleaved bench runs, ncu cycles), parity subset with the candidate,
# and the ncu launch list of the timed loop (only libcmdg's kernels)
BASE=${1:-_nosplit}
python -m pytest tests -m gpu -q -x -k "vortex_tendency or baroclinic_wave_cubed or viscous_box_second or held_suarez_forcing or dry_biharmonic or tracers_as_shipped or float32 or rising_bubble or vortex_100 or plain_c_abi" 2>&1 | tail -4
B="python bench.py --headline-only --no-parity --no-cpu-baseline --steps 100 --warmup 3"
for rep in 1 2 3; do
for v in "" $BASE; do
  CMDG_LIB=$PWD/climatemachine.jl_b200/libcmdg$v.so $B > gpurun_out/ab4_v${v}_$rep.json 2>/dev/null
done; done
for v in "" $BASE; do
CMDG_LIB=$PWD/climatemachine.jl_b200/libcmdg$v.so python bench.py --workload held_suarez --headline-only --no-parity --no-cpu-baseline --steps 40 --warmup 3 > gpurun_out/ab4_hs$v.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ab4_*.json")):
    try:
        d=json.load(open(f)); print(f, "%.2f GDOF/s %.4f ms/step kern %.4f ms/stage clk %s %s e2e %.2f"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_stage"],d["clocks"]["sm_mhz"],d["clocks"]["reasons"],d["e2e"]["value"]))
    except Exception as e: print(f,"ERR",e)
PY
M=sm__cycles_elapsed.max,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,smsp__inst_executed.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,smsp__average_warp_latency_issue_stalled_barrier.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
for v in "" $BASE; do
  CMDG_LIB=$PWD/climatemachine.jl_b200/libcmdg$v.so ncu --metrics $M --clock-control none -k regex:dg_tendency_kernel --launch-skip 12 --launch-count 3 --csv --log-file gpurun_out/ab4_ncu$v.csv python bench.py --headline-only --no-parity --no-cpu-baseline --steps 4 --warmup 3 > /dev/null 2>&1
done
python - <<PY
import csv
for v in ("","$BASE"):
    rows=[r for r in csv.reader(open(f"gpurun_out/ab4_ncu{v}.csv")) if len(r)>10]
    hdr=rows[0]; ix={n:i for i,n in enumerate(hdr)}
    agg={}
    for r in rows[1:]:
        agg.setdefault(r[ix["Metric Name"]],[]).append(r[ix["Metric Value"]])
    print("variant",v or "candidate")
    for k,vals in agg.items(): print("   ",k,vals)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dg_|lsrk|pack|scale_kernel" --launch-skip 8 -c 40 --csv --log-file gpurun_out/final_launches.csv python bench.py --headline-only --no-parity --no-cpu-baseline --steps 4 --warmup 3 > /dev/null 2>&1; echo "launch list rc=$?"
