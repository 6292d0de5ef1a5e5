B="python bench.py --headline-only --no-parity --no-cpu-baseline --steps 50 --warmup 3"
$B > gpurun_out/em_n1.json 2>/dev/null
$B --emulate-rank 0/4 > gpurun_out/em_r0of4_identity.json 2>/dev/null
CMDG_FORCE_LIST=1 $B --emulate-rank 0/4 > gpurun_out/em_r0of4_list.json 2>/dev/null
CMDG_FORCE_LIST=1 $B --emulate-rank 1/4 > gpurun_out/em_r1of4_list.json 2>/dev/null
$B > gpurun_out/em_n1_b.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/em_*.json")):
    d=json.load(open(f)); print(f, "%.2f GDOF/s %.4f ms/step kern %.4f nelem %d clk %s"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_stage"],d["config"]["nelem_total"],d["clocks"]["sm_mhz"]))
PY
M=sm__cycles_elapsed.max,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed
for v in n1 r0_identity r0_list; do
  case $v in n1) E=""; X="";; r0_identity) E=""; X="--emulate-rank 0/4";; r0_list) E="CMDG_FORCE_LIST=1"; X="--emulate-rank 0/4";; esac
  env $E X=1 ncu --metrics $M --clock-control none -k regex:dg_tendency_kernel --launch-skip 12 --launch-count 2 --csv --log-file gpurun_out/em_ncu_$v.csv python bench.py --headline-only --no-parity --no-cpu-baseline --steps 4 --warmup 3 $X > /dev/null 2>&1
done
python - <<'PY'
import csv
for v in ("n1","r0_identity","r0_list"):
    rows=[r for r in csv.reader(open(f"gpurun_out/em_ncu_{v}.csv")) if len(r)>10]
    hdr=rows[0]; ix={n:i for i,n in enumerate(hdr)}
    agg={}
    for r in rows[1:]: agg.setdefault(r[ix["Metric Name"]],[]).append(r[ix["Metric Value"]])
    print("variant",v)
    for k,vals in agg.items(): print("   ",k,vals)
PY
