python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_gpu" 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
CMDG_TIMELINE=gpurun_out/tl_n4 $TR --nproc-per-node 4 --master-port 29601 bench.py --gpus 4 --steps 20 --warmup 3 --headline-only --no-parity > gpurun_out/r2_bench_n4_tl.json 2> gpurun_out/r2_bench_n4_tl.err; echo tl rc=$?
$TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo n2 rc=$?; tail -3 gpurun_out/r2_bench_n2.err
$TR --nproc-per-node 4 --master-port 29603 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo n4 rc=$?; tail -3 gpurun_out/r2_bench_n4.err
python bench.py --steps 20 --warmup 3 --headline-only --no-parity --no-cpu-baseline > gpurun_out/r2_bench_n1_samebox.json 2>/dev/null; echo n1 rc=$?
