#!/bin/bash
# final-build ncu evidence (one GPU): (1) launch list of ONE eager bf16 step (the timed step of
# `bench.py --steps 1 --warmup 2 --profile --no-graph`, bracketed by cudaProfilerStart/Stop), (2) duration + DRAM bytes +
# tensor-pipe activity of every launch of the time-dominant conv kernels inside that step.  The command runs plainly
# first (exit 0), then under ncu with --clock-control none.
mkdir -p gpurun_out
t0=$(date +%s)
LIST="python bench.py --steps 1 --warmup 2 --profile --no-graph --no-inference --no-cpu-baseline"
$LIST > gpurun_out/ncu_final_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_final_plain.log; exit 1; }
echo "plain exit 0 at $(( $(date +%s) - t0 )) s"
timeout ${LIST_TIMEOUT:-230} ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv \
  --log-file gpurun_out/r2_final_launches.csv $LIST > gpurun_out/ncu_final_list.log 2>&1
echo "launch list exit $? at $(( $(date +%s) - t0 )) s"; wc -l gpurun_out/r2_final_launches.csv
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
timeout ${SWEEP_TIMEOUT:-170} ncu --metrics $M --clock-control none --profile-from-start off \
  -k "regex:conv_tc_ws_k|conv_tc_fwdh_k|conv_tc_wgrad2s_k|conv_tc_wt_k" -c 300 --csv \
  --log-file gpurun_out/r2_final_conv_metrics.csv $LIST > gpurun_out/ncu_final_sweep.log 2>&1
echo "sweep exit $? at $(( $(date +%s) - t0 )) s"; wc -l gpurun_out/r2_final_conv_metrics.csv
