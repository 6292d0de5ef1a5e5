#!/bin/bash
python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1_full.json 2> gpurun_out/r2_bench_n1_full.err; echo bench rc=$?; tail -3 gpurun_out/r2_bench_n1_full.err
B="python bench.py --workload held_suarez --headline-only --no-parity --no-cpu-baseline --steps 40 --warmup 3"
for v in "" _g4 _g6 _v5 ""; do
  CMDG_LIB=$PWD/climatemachine.jl_b200/libcmdg$v.so $B > gpurun_out/mb_v${v}_$RANDOM.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/mb_v*.json")):
    d=json.load(open(f)); print(f, "%.2f GDOF/s %.4f ms/step tend %.4f ms/stage clk %s"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_stage"],d["clocks"]["sm_mhz"]))
PY
