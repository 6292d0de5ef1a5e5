#!/bin/bash
# A/B of two builds of libcmdg on ONE box: parity subset with the candidate, interleaved bench runs,
# ncu cycle counts (clock-independent).  usage: tools/run_ab2.sh <suffix of the baseline lib> [pytest -k expr]
BASE=${1:-_base}
KEXPR=${2:-"vortex_tendency or baroclinic_wave_cubed or viscous_box_second or held_suarez_forcing or dry_biharmonic"}
python -m pytest tests -m gpu -q -x -k "$KEXPR" 2>&1 | tail -4
B="python bench.py --headline-only --no-parity --no-cpu-baseline --steps 100 --warmup 3"
for rep in 1 2; do
for v in "" $BASE; do
  CMDG_LIB=$PWD/climatemachine.jl_b200/libcmdg$v.so $B > gpurun_out/ab2_v${v}_$rep.json 2>/dev/null
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ab2_v*.json")):
    d=json.load(open(f)); print(f, "%.2f GDOF/s %.4f ms/step kern %.4f ms/stage clk %s %s e2e %.2f"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_stage"],d["clocks"]["sm_mhz"],d["clocks"]["reasons"],d["e2e"]["value"]))
PY
M=sm__cycles_elapsed.max,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed
for v in "" $BASE; do
  CMDG_LIB=$PWD/climatemachine.jl_b200/libcmdg$v.so ncu --metrics $M --clock-control none -k regex:dg_tendency_kernel --launch-skip 12 --launch-count 3 --csv --log-file gpurun_out/ab2_ncu$v.csv python bench.py --headline-only --no-parity --no-cpu-baseline --steps 4 --warmup 3 > /dev/null 2>&1
done
python - <<PY
import csv
for v in ("","$BASE"):
    rows=[r for r in csv.reader(open(f"gpurun_out/ab2_ncu{v}.csv")) if len(r)>10]
    hdr=rows[0]; ix={n:i for i,n in enumerate(hdr)}
    agg={}
    for r in rows[1:]:
        agg.setdefault(r[ix["Metric Name"]],[]).append(r[ix["Metric Value"]])
    print("variant",v or "candidate")
    for k,vals in agg.items(): print("   ",k,vals)
PY
