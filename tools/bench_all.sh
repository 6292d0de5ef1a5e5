#!/bin/bash
# Runs the round's single-GPU bench lines and the ncu launch lists (separate runs) into gpurun_out/.
# usage: tools/bench_all.sh [bench] [lists] [full]
mkdir -p gpurun_out
KREGEX='regex:(dg_|hyper_|hb_|lsrk_|filter_|courant_|pack_kernel|unpack_kernel|scale_kernel)'
what=${*:-bench lists}
if [[ $what == *bench* ]]; then
python bench.py --steps 100 --warmup 3 > gpurun_out/bench_bw.json 2> gpurun_out/bench_bw.err
python bench.py --workload held_suarez --steps 40 --warmup 3 > gpurun_out/bench_hs.json 2> gpurun_out/bench_hs.err
python bench.py --hyperdiffusion --steps 40 --warmup 3 > gpurun_out/bench_bw_hyper.json 2> gpurun_out/bench_bw_hyper.err
python bench.py --workload held_suarez --hyperdiffusion --steps 40 --warmup 3 > gpurun_out/bench_hs_hyper.json 2> gpurun_out/bench_hs_hyper.err
tail -c 600 gpurun_out/bench_*.err
cat gpurun_out/bench_bw.json gpurun_out/bench_hs.json gpurun_out/bench_bw_hyper.json gpurun_out/bench_hs_hyper.json | cut -c1-400
fi
if [[ $what == *lists* ]]; then
for w in "bw_hyper:--hyperdiffusion" "hs:--workload held_suarez" "hs_hyper:--workload held_suarez --hyperdiffusion"; do
  name=${w%%:*}; flags=${w#*:}
  BENCH_NO_KERNEL_TIMING=1 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 120 --csv \
    --log-file gpurun_out/launches_$name.csv python bench.py $flags --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$name.log 2>&1
done
fi
if [[ $what == *full* ]]; then
  # one full capture per second-order kernel (launch-skip past the warm-up)
  BENCH_NO_KERNEL_TIMING=1 ncu --set full --clock-control none --import-source on -k regex:dg_gradient_kernel -s 20 -c 1 \
    -o gpurun_out/prof_grad_hs -f python bench.py --workload held_suarez --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_grad.log 2>&1
  BENCH_NO_KERNEL_TIMING=1 ncu --set full --clock-control none --import-source on -k 'regex:(dg_gradient_kernel|hyper_)' -s 60 -c 3 \
    -o gpurun_out/prof_hyper_bw -f python bench.py --hyperdiffusion --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_hyper.log 2>&1
fi
