#!/bin/bash
# 4 GPUs: multi-rank parity (world 4), the full default bench line, one strong-scaling line (ne = 32 on 4 GPUs)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "(multi_gpu and 4)" 2>&1 | tail -5
timeout 900 $TR --nproc-per-node 4 --master-port 29821 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/f_n4_full.json 2> gpurun_out/f_n4_full.err; echo "n4 full rc=$?"
tail -n 3 gpurun_out/f_n4_full.err
timeout 300 $TR --nproc-per-node 4 --master-port 29822 bench.py --gpus 4 --strong --headline-only --no-parity --no-cpu-baseline --steps 50 --warmup 3 > gpurun_out/f_n4_strong.json 2> gpurun_out/f_n4_strong.err; echo "n4 strong rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/f_n4_*.json")):
    try:
        d=json.load(open(f))
        print(f, "%.2f GDOF/s %.4f ms/step e2e %.2f (%.2f ms)"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["ms_per_step"]), d["clocks"]["sm_mhz"], d.get("parity",{}).get("green"), d["scaling"], d["config"]["nelem_total"])
        for c in d.get("clocks_per_rank",[]): print("    rank", c["rank"], c["sm_mhz"], round(c["kernel_ms_per_stage"],4), c.get("host_placement"), c.get("e2e_ms_per_step"))
        if "sustained_100" in d: print("    sustained_100 %.2f"%d["sustained_100"]["value"])
        for k,v in d.get("secondary",{}).items(): print("    ",k,"%.2f GDOF/s %.3f ms/step"%(v["value"],v["ms_per_step"]))
    except Exception as e: print(f,"ERR",e)
PY
