#!/bin/bash
# Regenerates the tracked profile extracts under profiles/ from the .ncu-rep / .csv files a GPU run left in gpurun_out/.
#   usage: tools/make_profiles.sh <tag of gpu_ncu_full.sh> <round prefix, e.g. r01>
TAG=$1; R=${2:-r01}
cd "$(dirname "$0")/.."
for r in ${REPS:-hf_closest hf_shadow hf_shade cbox_closest cbox_shade sphere_box ring}; do
  rep=gpurun_out/prof_${TAG}_$r.ncu-rep
  [ -f $rep ] || { echo "missing $rep"; continue; }
  python tools/ncu_regions.py $rep 40 > profiles/${R}_ncu_$r.txt 2>&1
  ncu -i $rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; u=rows[1]; r=rows[2]
want=['smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__warps_eligible.avg.per_cycle_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_local_op_ld_lookup_miss.sum','lts__t_sectors_srcunit_tex_op_read.sum','launch__grid_size','launch__block_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__cycles_elapsed.avg.per_second']
print('--- stall / pipe / memory detail (ncu --set full, raw page)')
for k in want:
    if k in h: print(f'{k:85s} {r[h.index(k)]} {u[h.index(k)]}')
" >> profiles/${R}_ncu_$r.txt
  echo "== $r: $(grep -E 'gpu__time_duration|dram__bytes_read|dram__bytes_write|smsp__inst_executed.sum|issue_active|thread_inst_executed_per' profiles/${R}_ncu_$r.txt | awk '{print $(NF-1)}' | tr '\n' ' ')"
done
