#!/bin/bash
# candidate (packed face aux + pipelined host path) vs -DCMDG_FACE_AUX=0 build, and host pipe on/off
python -m pytest tests -m gpu -q -x -k "vortex_tendency or baroclinic_wave_cubed or viscous_box_second or held_suarez_forcing or dry_biharmonic or tracers_as_shipped or hydrostatic_balance" 2>&1 | tail -6
B="python bench.py --headline-only --no-parity --no-cpu-baseline --steps 100 --warmup 3"
for rep in 1 2; do
for v in "" _nofaux; do
  CMDG_LIB=$PWD/climatemachine.jl_b200/libcmdg$v.so $B > gpurun_out/ab3_v${v}_$rep.json 2>/dev/null
done; done
CMDG_HOST_PIPE=0 $B > gpurun_out/ab3_v_nopipe_1.json 2>/dev/null
CMDG_HOST_CHUNKS=16 $B > gpurun_out/ab3_v_chunks16_1.json 2>/dev/null
CMDG_HOST_CHUNKS=4 $B > gpurun_out/ab3_v_chunks4_1.json 2>/dev/null
python bench.py --workload held_suarez --headline-only --no-parity --no-cpu-baseline --steps 40 --warmup 3 > gpurun_out/ab3_hs.json 2>/dev/null
CMDG_LIB=$PWD/climatemachine.jl_b200/libcmdg_nofaux.so python bench.py --workload held_suarez --headline-only --no-parity --no-cpu-baseline --steps 40 --warmup 3 > gpurun_out/ab3_hs_nofaux.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ab3_*.json")):
    try:
        d=json.load(open(f)); print(f, "%.2f GDOF/s %.4f ms/step kern %.4f ms/stage clk %s %s e2e %.2f (%.3f ms)"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_stage"],d["clocks"]["sm_mhz"],d["clocks"]["reasons"],d["e2e"]["value"],d["e2e"]["ms_per_step"]))
    except Exception as e: print(f, "ERR", e)
PY
M=sm__cycles_elapsed.max,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed
for v in "" _nofaux; do
  CMDG_LIB=$PWD/climatemachine.jl_b200/libcmdg$v.so ncu --metrics $M --clock-control none -k regex:dg_tendency_kernel --launch-skip 12 --launch-count 3 --csv --log-file gpurun_out/ab3_ncu$v.csv python bench.py --headline-only --no-parity --no-cpu-baseline --steps 4 --warmup 3 > /dev/null 2>&1
done
python - <<'PY'
import csv
for v in ("","_nofaux"):
    rows=[r for r in csv.reader(open(f"gpurun_out/ab3_ncu{v}.csv")) if len(r)>10]
    hdr=rows[0]; ix={n:i for i,n in enumerate(hdr)}
    agg={}
    for r in rows[1:]:
        agg.setdefault(r[ix["Metric Name"]],[]).append(r[ix["Metric Value"]])
    print("variant",v or "candidate")
    for k,vals in agg.items(): print("   ",k,vals)
PY
