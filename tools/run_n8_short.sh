#!/bin/bash
# 8 GPUs, short: headline (20 steps, with the parity block) and the ocean weak-scaling step with the two-chain schedule
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29831 bench.py --gpus 8 --headline-only --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/s_n8_headline.json 2> gpurun_out/s_n8_headline.err; echo "n8 headline rc=$?"
timeout 400 $TR --nproc-per-node 8 --master-port 29832 bench.py --gpus 8 --workload ocean_gyre --headline-only --no-parity --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/s_n8_ocean.json 2> gpurun_out/s_n8_ocean.err; echo "n8 ocean rc=$?"
tail -n 2 gpurun_out/s_n8_ocean.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s_n8_*.json")):
    try:
        d=json.load(open(f))
        print(f, "%.2f GDOF/s %.4f ms/step e2e %.2f (%.2f ms)"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["ms_per_step"]), d["clocks"]["sm_mhz"], d.get("parity",{}).get("green"), d["gpu_launches"])
        for c in d.get("clocks_per_rank",[]): print("    rank", c["rank"], c["sm_mhz"], round(c["kernel_ms_per_stage"],4), c.get("host_placement"), round(c.get("e2e_ms_per_step",0),2))
    except Exception as e: print(f,"ERR",e)
PY
