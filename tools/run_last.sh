#!/bin/bash
python bench.py --steps 20 --warmup 3 > gpurun_out/last_n1.json 2> gpurun_out/last_n1.err; echo "bench rc=$?"
tail -n 3 gpurun_out/last_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/last_n1.json"))
print("%.2f GDOF/s %.4f ms/step roofline %.3f traffic %.3e e2e %.2f clk %s parity %s cpu %.3f"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["roofline"]["traffic"],d["e2e"]["value"],d["clocks"]["sm_mhz"],d["parity"]["green"],d["cpu_baseline"]["value"]))
for k,v in d["secondary"].items(): print("   ",k,"%.2f GDOF/s"%v["value"], v.get("cpu_baseline",{}).get("value"), str(v.get("cpu_baseline",{}).get("error",""))[:100])
PY
