#!/bin/bash
# round 2h: what the FMA-pipe counter and the stall reasons say when the ONLY limit is register operand delivery
# (tools/microbench/opnd.cu: pure FFMA2 / FFMA loops, 8 warps per SMSP, nothing else in the kernel)
mkdir -p gpurun_out/r2h
O=gpurun_out/r2h
M=sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_selected_per_issue_active.ratio,gpu__time_duration.sum
timeout 600 ncu --clock-control none --metrics $M --csv --log-file $O/opnd_ncu.csv tools/microbench/opnd > $O/opnd_under_ncu.txt 2>&1
echo "rc=$?"
tail -3 $O/opnd_ncu.csv
