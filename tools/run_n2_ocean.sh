#!/bin/bash
# 2 GPUs: multi-rank parity (all paths), ocean weak-scaling step: serial vs two-chain schedule (+ timeline)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "(multi_gpu and 2)" 2>&1 | tail -5
O="--workload ocean_gyre --headline-only --no-parity --no-cpu-baseline --steps 10 --warmup 3"
CMDG_OVERLAP=0 timeout 300 $TR --nproc-per-node 2 --master-port 29811 bench.py --gpus 2 $O > gpurun_out/oc_n2_serial.json 2>gpurun_out/oc_n2_serial.err; echo "oc n2 serial rc=$?"
timeout 300 $TR --nproc-per-node 2 --master-port 29812 bench.py --gpus 2 $O > gpurun_out/oc_n2_chains.json 2>gpurun_out/oc_n2_chains.err; echo "oc n2 chains rc=$?"
CMDG_TIMELINE=gpurun_out/tl_oc_n2 timeout 300 $TR --nproc-per-node 2 --master-port 29813 bench.py --gpus 2 $O > gpurun_out/oc_n2_chains_tl.json 2>/dev/null; echo "oc n2 tl rc=$?"
tail -n 3 gpurun_out/oc_n2_chains.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/oc_n2*.json")):
    try:
        d=json.load(open(f))
        print(f, "%.2f GDOF/s %.4f ms/step e2e %.2f"%(d["value"],d["ms_per_step"],d["e2e"]["value"]), d["clocks"]["sm_mhz"], d["gpu_launches"], d["norm_ratio"])
    except Exception as e: print(f,"ERR",e)
PY
