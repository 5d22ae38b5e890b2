#!/bin/bash
# metric capture of the fp32 narrow-channel kernels inside an eager bf16 step (excitation pyramid, stems)
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__shared_mem_per_block_dynamic,smsp__inst_executed.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
LIST="python bench.py --steps 1 --warmup 2 --profile --no-graph --no-inference --no-cpu-baseline"
timeout 300 ncu --metrics $M --clock-control none -k "regex:stem_wgrad_k|narrow_wgrad_k|conv_fwd_k" -s 130 -c 75 --csv --log-file gpurun_out/r2_ncu_narrow.csv $LIST > gpurun_out/ncu_narrow.log 2>&1
echo "exit $?"; wc -l gpurun_out/r2_ncu_narrow.csv
