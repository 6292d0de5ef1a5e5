#!/bin/bash
# final single-GPU check of the round: whole GPU suite, smoke, the default bench line as the driver runs it,
# ncu launch list of the timed loop and one --set full capture of a whole step of the headline kernel
python -m pytest tests -m gpu -q 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/final_n1.json"))
print("%.2f GDOF/s %.4f ms/step roofline %.3f e2e %.2f (%.3f ms) clk %s %s parity %s fullsize %.2e cpu %.3f x%d"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["e2e"]["value"],d["e2e"]["ms_per_step"],d["clocks"]["sm_mhz"],d["clocks"]["reasons"],d["parity"]["green"],d["parity_fullsize_rel_l2"],d["cpu_baseline"]["value"],d["cpu_baseline"]["cores"]))
print("sustained_100 %.2f  reference_schedule %.2f"%(d["sustained_100"]["value"], d["reference_schedule"]["value"]))
for k,v in d["secondary"].items(): print("   ",k,"%.2f GDOF/s %.3f ms/step"%(v["value"],v["ms_per_step"]), {a:round(b,3) for a,b in v["kernel_ms_per_stage"].items()})
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/final_launches.csv python bench.py --headline-only --no-parity --no-cpu-baseline --steps 2 --warmup 1 > /dev/null 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:dg_tendency_kernel --launch-skip 15 --launch-count 5 -o gpurun_out/final_tend -f python bench.py --headline-only --no-parity --no-cpu-baseline --steps 3 --warmup 3 > /dev/null 2> gpurun_out/final_ncu.log; echo "ncu full rc=$?"
ls -la gpurun_out/final_tend.ncu-rep
