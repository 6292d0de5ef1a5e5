TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
H="--headline-only --no-cpu-baseline --steps 50 --warmup 3"
python bench.py $H --no-parity > gpurun_out/m2_n1.json 2>/dev/null; echo n1 rc=$?
CMDG_TIMELINE=gpurun_out/tl2_n4 $TR --nproc-per-node 4 --master-port 29601 bench.py --gpus 4 $H > gpurun_out/m2_n4.json 2> gpurun_out/m2_n4.err; echo n4 rc=$?
CMDG_TAILPF=0 $TR --nproc-per-node 4 --master-port 29602 bench.py --gpus 4 $H --no-parity > gpurun_out/m2_n4_tailpf0.json 2>/dev/null; echo n4b rc=$?
$TR --nproc-per-node 4 --master-port 29603 bench.py --gpus 4 $H --no-parity > gpurun_out/m2_n4_b.json 2>/dev/null; echo n4c rc=$?
$TR --nproc-per-node 2 --master-port 29604 bench.py --gpus 2 $H --no-parity > gpurun_out/m2_n2.json 2>/dev/null; echo n2 rc=$?
python bench.py $H --no-parity > gpurun_out/m2_n1_b.json 2>/dev/null
python - <<'PY'
import json
for f in ("m2_n1","m2_n1_b","m2_n2","m2_n4","m2_n4_tailpf0","m2_n4_b"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, "%.2f GDOF/s %.4f ms/step"%(d["value"],d["ms_per_step"]), d["clocks"]["sm_mhz"], [ (c["sm_mhz"], round(c["kernel_ms_per_stage"],4)) for c in d.get("clocks_per_rank",[])], d.get("parity",{}).get("green"))
    except Exception as e: print(f, "ERR", e)
PY
