TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_gpu and 8" 2>&1 | tail -4
H="--headline-only --no-cpu-baseline --steps 50 --warmup 3"
python bench.py $H --no-parity > gpurun_out/n8_n1.json 2>/dev/null; echo n1 rc=$?
timeout 600 $TR --nproc-per-node 8 --master-port 29801 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/n8_full.json 2> gpurun_out/n8_full.err; echo "n8 full rc=$?"; tail -2 gpurun_out/n8_full.err
CMDG_OVERLAP=2 timeout 300 $TR --nproc-per-node 8 --master-port 29802 bench.py --gpus 8 $H --no-parity > gpurun_out/n8_mode2.json 2>/dev/null; echo "n8 mode2 rc=$?"
CMDG_TIMELINE=gpurun_out/tl_n8 timeout 300 $TR --nproc-per-node 8 --master-port 29803 bench.py --gpus 8 $H --no-parity > gpurun_out/n8_mode1.json 2>/dev/null; echo "n8 mode1 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/n8_*.json")):
    try:
        d=json.load(open(f))
        print(f, "%.2f GDOF/s %.4f ms/step e2e %.2f"%(d["value"],d["ms_per_step"],d["e2e"]["value"]), d["clocks"]["sm_mhz"], [ (c["sm_mhz"], round(c["kernel_ms_per_stage"],4)) for c in d.get("clocks_per_rank",[])], d.get("parity",{}).get("green"))
        if "secondary" in d:
            for k,v in d["secondary"].items(): print("    ",k,"%.2f GDOF/s"%v["value"])
    except Exception as e: print(f,"ERR",e)
PY
