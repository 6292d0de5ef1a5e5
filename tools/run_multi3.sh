TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
H="--headline-only --no-cpu-baseline --no-parity --steps 50 --warmup 3"
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_gpu or crashes" 2>&1 | tail -4
python bench.py $H > gpurun_out/m3_n1.json 2>/dev/null; echo n1 rc=$?
i=0
for envs in "X=1" "NCCL_MAX_P2P_NCHANNELS=1" "NCCL_MAX_P2P_NCHANNELS=2" "NCCL_MAX_P2P_NCHANNELS=1 NCCL_NTHREADS=256" "X=1"; do
  i=$((i+1))
  env $envs $TR --nproc-per-node 4 --master-port $((29610+i)) bench.py --gpus 4 $H > gpurun_out/m3_n4_$i.json 2>/dev/null; echo "n4 [$envs] rc=$?"
done
$TR --nproc-per-node 4 --master-port 29650 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/m3_n4_full.json 2> gpurun_out/m3_n4_full.err; echo n4full rc=$?; tail -2 gpurun_out/m3_n4_full.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/m3_*.json")):
    try:
        d=json.load(open(f))
        print(f, "%.2f GDOF/s %.4f ms/step e2e %.2f"%(d["value"],d["ms_per_step"],d["e2e"]["value"]), d["clocks"]["sm_mhz"], [ (c["sm_mhz"], round(c["kernel_ms_per_stage"],4)) for c in d.get("clocks_per_rank",[])])
    except Exception as e: print(f,"ERR",e)
PY
