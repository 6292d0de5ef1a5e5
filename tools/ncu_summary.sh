#!/bin/bash
# Dev tool: key metrics of an .ncu-rep + per-source-line table.
# usage: tools/ncu_summary.sh rep.ncu-rep [kernel-mangled-substring] [sort column] [rows]
here=$(cd "$(dirname "$0")" && pwd)
rep=$1; pat=${2:-dg_tendency_kernelIdLi5ELi0ELb1ELb0ELb0}; key=${3:-"# Samples"}
ncu -i $rep --page raw --csv 2>/dev/null > /tmp/raw.csv
python - <<PY
import csv
rows=list(csv.reader(open('/tmp/raw.csv')))
h=rows[0]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum.pct_of_peak_sustained_elapsed',
'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts.sum',
'sm__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread',
'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','launch__grid_size']
for r in rows[2:]:
    print(r[h.index('Kernel Name')][:80] if 'Kernel Name' in h else '')
    for w in want:
        if w in h: print('  %-95s %s %s'%(w,r[h.index(w)],rows[1][h.index(w)]))
PY
ncu -i $rep --page source --csv 2>/dev/null > /tmp/src.csv
so=$here/../climatemachine.jl_b200/libcmdg.so
mkdir -p /tmp/cub && (cd /tmp/cub && rm -f *.cubin && cuobjdump -xelf all $so >/dev/null && nvdisasm -g -c *.cubin 2>/dev/null > /tmp/all.sass)
python - "$pat" <<'PY'
import sys
pat=sys.argv[1]
out=[];on=False
for ln in open('/tmp/all.sass'):
    if ln.startswith('//---') and '.text.' in ln:
        on = pat in ln
    if on: out.append(ln)
open('/tmp/kern1.sass','w').writelines(out)
PY
python $here/ncu_by_line.py /tmp/src.csv /tmp/kern1.sass $here/../climatemachine.jl_b200/csrc/cmdg_kernels.cuh "$key" | cut -c1-230 | head -${4:-40}
