#!/bin/bash
# ncu session of round 2 (one GPU): (1) launch list of an eager bf16 step, (2) --set full of every kernel family at its
# benchmark shape (profiles/tools/ncu_targets.py).  Every command is run plainly first (exit 0) and then under ncu.
mkdir -p gpurun_out
LIST="python bench.py --steps 1 --warmup 1 --profile --no-graph --no-inference --no-cpu-baseline"
$LIST > gpurun_out/ncu_list_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches.csv $LIST > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
for part in mrf disc hbm; do
  T="python profiles/tools/ncu_targets.py $part"
  $T > gpurun_out/ncu_${part}_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:tdvc -c 400 -o gpurun_out/r2_ncu_$part -f $T > gpurun_out/ncu_$part.log 2>&1
  echo "$part exit $?"
done
ls -la gpurun_out | tail -20
