#!/bin/bash
# final single-GPU check of the round: racecheck over one case per kernel family, whole GPU suite, smoke, the default
# bench line as the driver runs it
bash tools/sanitize.sh racecheck; tail -4 gpurun_out/sanitizer_racecheck.log
python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/final_n1.json"))
print("%.2f GDOF/s %.4f ms/step roofline %.3f e2e %.2f (%.3f ms) clk %s %s parity %s fullsize %.2e cpu %.3f x%d"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["e2e"]["value"],d["e2e"]["ms_per_step"],d["clocks"]["sm_mhz"],d["clocks"]["reasons"],d["parity"]["green"],d["parity_fullsize_rel_l2"],d["cpu_baseline"]["value"],d["cpu_baseline"]["cores"]))
print("sustained_100 %.2f  reference_schedule %.2f"%(d["sustained_100"]["value"], d["reference_schedule"]["value"]))
for k,v in d["secondary"].items(): print("   ",k,"%.2f GDOF/s %.3f ms/step"%(v["value"],v["ms_per_step"]), {a:round(b,3) for a,b in v["kernel_ms_per_stage"].items()})
PY
