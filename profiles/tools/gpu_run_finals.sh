#!/bin/bash
# final measurements of the round on one GPU: the default bench line (stage1, with inference sweep and CPU baseline), the other
# shipped configs, and the reference arm once (bounded)
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/final_stage1.json 2> gpurun_out/final_stage1.err; echo "stage1 exit $?"
for cfg in stage2_1 stage2_1_latcls stage2_2; do
  timeout 300 python bench.py --steps 5 --warmup 3 --config $cfg --no-inference --no-cpu-baseline > gpurun_out/final_$cfg.json 2> gpurun_out/final_$cfg.err
  echo "$cfg exit $?"
done
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/final_reference_arm.json 2> gpurun_out/final_reference_arm.err; echo "reference arm exit $?"
python - <<'PY'
import json
for n in ("stage1", "stage2_1", "stage2_1_latcls", "stage2_2", "reference_arm"):
    try:
        d = json.load(open(f"gpurun_out/final_{n}.json"))
        print(n, d.get("ms_per_step"), d.get("value"), d.get("gpu_launches"), (d.get("inference") or {}).get("best"))
    except Exception as e:
        print(n, "ERR", e)
PY
