#!/bin/bash
python -m pytest tests -m gpu -q -x -k "viscous or held_suarez or hyper or bubble or tracer or courant" 2>&1 | tail -4
B="python bench.py --workload held_suarez --headline-only --no-parity --no-cpu-baseline --steps 40 --warmup 3"
for rep in 1 2; do for v in "" _gradold; do
  CMDG_LIB=$PWD/climatemachine.jl_b200/libcmdg$v.so $B > gpurun_out/grad_v${v}_$rep.json 2>/dev/null
done; done
CMDG_LIB=$PWD/climatemachine.jl_b200/libcmdg.so $B --hyperdiffusion > gpurun_out/grad_hyper_new.json 2>/dev/null
CMDG_LIB=$PWD/climatemachine.jl_b200/libcmdg_gradold.so $B --hyperdiffusion > gpurun_out/grad_hyper_old.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/grad_*.json")):
    d=json.load(open(f)); print(f, "%.2f GDOF/s %.4f ms/step tend %.4f ms/stage clk %s"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_stage"],d["clocks"]["sm_mhz"]))
PY
M=sm__cycles_elapsed.max,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,smsp__inst_executed.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum
for v in "" _gradold; do
  CMDG_LIB=$PWD/climatemachine.jl_b200/libcmdg$v.so ncu --metrics $M --clock-control none -k regex:dg_gradient_kernel --launch-skip 12 --launch-count 3 --csv --log-file gpurun_out/grad_ncu$v.csv python bench.py --workload held_suarez --headline-only --no-parity --no-cpu-baseline --steps 4 --warmup 3 > /dev/null 2>&1
done
python - <<'PY'
import csv
for v in ("","_gradold"):
    rows=[r for r in csv.reader(open(f"gpurun_out/grad_ncu{v}.csv")) if len(r)>10]
    hdr=rows[0]; ix={n:i for i,n in enumerate(hdr)}
    agg={}
    for r in rows[1:]: agg.setdefault(r[ix["Metric Name"]],[]).append(r[ix["Metric Value"]])
    print("variant",v or "new")
    for k,vals in agg.items(): print("   ",k,vals)
PY
