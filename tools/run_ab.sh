#!/bin/bash
# A/B of kernel variants on ONE box: interleaved bench runs + ncu cycle counts (clock-independent)
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
B="python bench.py --headline-only --no-parity --no-cpu-baseline --steps 100 --warmup 3"
for rep in 1 2; do
for v in "" _mb6; do
  CMDG_LIB=$PWD/climatemachine.jl_b200/libcmdg$v.so $B > gpurun_out/ab_v${v}_$rep.json 2>/dev/null
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ab_v*.json")):
    d=json.load(open(f)); print(f, "%.2f GDOF/s %.4f ms/step kern %.4f ms/stage clk %s %s"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_stage"],d["clocks"]["sm_mhz"],d["clocks"]["reasons"]))
PY
M=sm__cycles_elapsed.max,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed
for v in "" _mb6; do
  CMDG_LIB=$PWD/climatemachine.jl_b200/libcmdg$v.so ncu --metrics $M --clock-control none -k regex:dg_tendency_kernel --launch-skip 12 --launch-count 3 --csv --log-file gpurun_out/ab_ncu$v.csv python bench.py --headline-only --no-parity --no-cpu-baseline --steps 4 --warmup 3 > /dev/null 2>&1
done
python - <<'PY'
import csv
for v in ("","_mb6"):
    rows=[r for r in csv.reader(open(f"gpurun_out/ab_ncu{v}.csv")) if len(r)>10]
    hdr=rows[0]; ix={n:i for i,n in enumerate(hdr)}
    agg={}
    for r in rows[1:]:
        agg.setdefault(r[ix["Metric Name"]],[]).append(r[ix["Metric Value"]])
    print("variant",v or "base")
    for k,vals in agg.items(): print("   ",k,vals)
PY
