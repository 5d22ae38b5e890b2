#!/bin/bash
# ncu session of round 2 (one GPU): (1) launch list of an eager bf16 step, (2) a metric sweep (duration, DRAM bytes and
# throughput, tensor-pipe activity, occupancy) over the kernel families at their benchmark shapes
# (profiles/tools/ncu_targets.py; the profiled range is the second call of each target), (3) --set full of the
# time-dominant kernel (conv_tc_wgrad2_k) and the FLOP-dominant one (conv_tc_wt_k).  Every command is run plainly first
# (exit 0) and then under ncu.  A first attempt with --set full over ~500 launches did not finish in 40 minutes: the sweep
# uses a metric list (3 passes per launch) instead.
mkdir -p gpurun_out
t0=$(date +%s)
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__shared_mem_per_block_dynamic
LIST="python bench.py --steps 1 --warmup 1 --profile --no-graph --no-inference --no-cpu-baseline"
$LIST > gpurun_out/ncu_list_plain.log 2>&1 &&
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2_launches.csv $LIST > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $? at $(( $(date +%s) - t0 )) s"
T="python profiles/tools/ncu_targets.py mini"
$T > gpurun_out/ncu_mini_plain.log 2>&1 &&
timeout 300 ncu --metrics $M --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r2_ncu_sweep.csv $T > gpurun_out/ncu_mini.log 2>&1
echo "sweep exit $? at $(( $(date +%s) - t0 )) s"
T="python profiles/tools/ncu_targets.py mrf"
timeout 240 ncu --set full --clock-control none --profile-from-start off -k regex:conv_tc_wt_k -c 2 -o /tmp/r2_ncu_wt -f $T > gpurun_out/ncu_wt.log 2>&1
echo "wt exit $? at $(( $(date +%s) - t0 )) s"
ncu -i /tmp/r2_ncu_wt.ncu-rep --page raw --csv > gpurun_out/r2_ncu_wt.raw.csv 2>/dev/null
T="python profiles/tools/ncu_targets.py mini"
timeout 240 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc_wgrad2_k -c 4 -o /tmp/r2_ncu_wgrad2 -f $T > gpurun_out/ncu_wgrad2.log 2>&1
echo "wgrad2 exit $? at $(( $(date +%s) - t0 )) s"
ncu -i /tmp/r2_ncu_wgrad2.ncu-rep --page raw --csv > gpurun_out/r2_ncu_wgrad2.raw.csv 2>/dev/null
[ $(stat -c %s /tmp/r2_ncu_wgrad2.ncu-rep) -lt 30000000 ] && cp /tmp/r2_ncu_wgrad2.ncu-rep gpurun_out/
du -sh gpurun_out; ls -la gpurun_out | tail -12
