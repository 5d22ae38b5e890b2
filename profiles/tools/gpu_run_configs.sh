mkdir -p gpurun_out
for cfg in stage2_1 stage2_1_latcls stage2_2; do
  timeout 60 python bench.py --steps 5 --warmup 3 --config $cfg --no-inference --no-cpu-baseline > gpurun_out/final3_$cfg.json 2> gpurun_out/final3_$cfg.err
  python -c "import json; d=json.load(open('gpurun_out/final3_$cfg.json')); print('$cfg', d['ms_per_step'], d['value'], d['gpu_launches'])"
done
