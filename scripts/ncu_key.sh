#!/bin/bash
# key metrics of one .ncu-rep: bash scripts/ncu_key.sh file.ncu-rep
ncu -i "$1" --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]; u=rows[1]; v=rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed.sum','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__average_warp_latency_issue_stalled_barrier','l1tex__t_bytes.sum','smsp__inst_executed.sum','lts__t_sector_hit_rate.pct','launch__waves_per_multiprocessor','smsp__warp_issue_stalled_barrier_per_warp_active.pct','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_membar_per_warp_active.pct','smsp__warp_issue_stalled_wait_per_warp_active.pct','smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct','smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct','smsp__warp_issue_stalled_not_selected_per_warp_active.pct','smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct','smsp__warp_issue_stalled_sleeping_per_warp_active.pct','smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct','smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct','smsp__warp_issue_stalled_no_instruction_per_warp_active.pct','smsp__warp_issue_stalled_drain_per_warp_active.pct','smsp__warp_issue_stalled_imc_miss_per_warp_active.pct','smsp__warp_issue_stalled_selected_per_warp_active.pct']
for w in want:
    for i,n in enumerate(h):
        if n==w: print('%-80s %-12s %s'%(n,u[i],v[i]))
"
