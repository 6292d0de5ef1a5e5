TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_gpu" 2>&1 | tail -4
H="--workload held_suarez --headline-only --no-cpu-baseline --steps 30 --warmup 3"
python bench.py $H --no-parity > gpurun_out/m4_hs_n1.json 2>/dev/null; echo n1 rc=$?
$TR --nproc-per-node 2 --master-port 29701 bench.py --gpus 2 $H > gpurun_out/m4_hs_n2.json 2> gpurun_out/m4_hs_n2.err; echo n2 rc=$?; tail -2 gpurun_out/m4_hs_n2.err
CMDG_OVERLAP=0 $TR --nproc-per-node 2 --master-port 29702 bench.py --gpus 2 $H --no-parity > gpurun_out/m4_hs_n2_serial.json 2>/dev/null; echo n2s rc=$?
$TR --nproc-per-node 2 --master-port 29703 bench.py --gpus 2 $H --no-parity > gpurun_out/m4_hs_n2_b.json 2>/dev/null; echo n2b rc=$?
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/m4_*.json")):
    try:
        d=json.load(open(f))
        print(f, "%.2f GDOF/s %.4f ms/step"%(d["value"],d["ms_per_step"]), d["clocks"]["sm_mhz"], d.get("parity",{}).get("green"), d.get("parity",{}).get("second_order"))
    except Exception as e: print(f,"ERR",e)
PY
