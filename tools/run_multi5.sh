TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_gpu and (2 or 4)" 2>&1 | tail -4
H="--headline-only --no-cpu-baseline --steps 50 --warmup 3"
python bench.py $H --no-parity > gpurun_out/m5_n1.json 2>/dev/null; echo n1 rc=$?
i=0
for mode in 2 1 2 1; do
  i=$((i+1))
  CMDG_OVERLAP=$mode timeout 300 $TR --nproc-per-node 4 --master-port $((29710+i)) bench.py --gpus 4 $H $( [ $i -gt 1 ] && echo --no-parity ) > gpurun_out/m5_n4_mode${mode}_$i.json 2> gpurun_out/m5_n4_$i.err; echo "n4 mode $mode rc=$?"
done
CMDG_OVERLAP=2 timeout 300 $TR --nproc-per-node 2 --master-port 29720 bench.py --gpus 2 $H --no-parity > gpurun_out/m5_n2_mode2.json 2>/dev/null; echo n2 rc=$?
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/m5_*.json")):
    try:
        d=json.load(open(f))
        print(f, "%.2f GDOF/s %.4f ms/step"%(d["value"],d["ms_per_step"]), d["clocks"]["sm_mhz"], [ (c["sm_mhz"], round(c["kernel_ms_per_stage"],4)) for c in d.get("clocks_per_rank",[])], d.get("parity",{}).get("green"))
    except Exception as e: print(f,"ERR",e)
PY
tail -3 gpurun_out/m5_n4_1.err
