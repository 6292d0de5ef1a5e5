#!/bin/bash
# single-GPU check: GPU tests, A/B of the tail prefetch, ncu capture of the headline kernel
python -m pytest tests -m gpu -q 2>&1 | tail -15
B="python bench.py --headline-only --no-parity --no-cpu-baseline --steps 100 --warmup 3"
$B > gpurun_out/r2_n1_a.json 2>/dev/null
CMDG_TAILPF=0 $B > gpurun_out/r2_n1_tailpf0.json 2>/dev/null
$B > gpurun_out/r2_n1_b.json 2>/dev/null
python - <<'PY'
import json
for f in ("r2_n1_a","r2_n1_tailpf0","r2_n1_b"):
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, "%.2f GDOF/s %.4f ms/step kern %.4f ms/stage clk %s %s"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_stage"],d["clocks"]["sm_mhz"],d["clocks"]["reasons"]))
PY
ncu --set full --clock-control none --import-source on -k regex:dg_tendency_kernel --launch-skip 12 --launch-count 2 -o gpurun_out/r2_tend_v4 -f python bench.py --headline-only --no-parity --no-cpu-baseline --steps 3 --warmup 3 > /dev/null 2> gpurun_out/ncu_r2_tend.log; echo ncu rc=$?
