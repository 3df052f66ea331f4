#!/bin/bash
# usage: tools/ncu_report.sh <file.ncu-rep> <lib.so> [topN]  -> key metrics + per-source-line aggregation
REP=$1; LIB=$2; TOP=${3:-40}
TMP=$(mktemp -d)
ncu -i $REP --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]; vals=rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__shared_mem_per_block_dynamic','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','smsp__inst_executed.sum','launch__grid_size','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_selected_per_issue_active.ratio','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
for k in want:
    for i,h in enumerate(hdr):
        if h==k: print(f'{k:90s} {vals[i]:>16s} {rows[1][i]}')
"
ncu -i $REP --page source --csv 2>/dev/null > $TMP/src.csv
(cd $TMP && cuobjdump -xelf all $LIB >/dev/null && nvdisasm -g -c *.cubin > dis.txt 2>/dev/null)
python $(dirname $0)/ncu_by_line.py $TMP/src.csv $TMP/dis.txt _Z16dmb_solve_kernel10KernelArgs $TOP
rm -rf $TMP
