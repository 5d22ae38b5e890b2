#!/bin/bash
# last GPU call of the round: full -m gpu suite, A/B of the fused output-layer op, then the final bench lines with the faster
# setting (stage1 with inference sweep + CPU baseline first, the other shipped configs while time remains)
mkdir -p gpurun_out
(timeout 300 python -m pytest tests -m gpu -q --tb=short --maxfail=30 -p no:cacheprovider > gpurun_out/last_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/last_tests.log)
tail -3 gpurun_out/last_tests.log
ab() {
  env TDVC_FUSED_SELECT=$1 timeout 200 python bench.py --steps 10 --warmup 3 --no-inference --no-cpu-baseline > gpurun_out/last_select$1.json 2> gpurun_out/last_select$1.err
  python -c "import json; print(json.load(open('gpurun_out/last_select$1.json'))['ms_per_step'])" 2>/dev/null || echo 999
}
m1=$(ab 1); m0=$(ab 0)
echo "fused select: $m1 ms   unfused: $m0 ms"
best=$(python -c "print(1 if float('$m1') < float('$m0') - 0.03 else 0)")
echo "TDVC_FUSED_SELECT=$best for the final lines"
export TDVC_FUSED_SELECT=$best
echo $best > gpurun_out/last_fused_select_choice.txt
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/final2_stage1.json 2> gpurun_out/final2_stage1.err; echo "stage1 exit $?"
for cfg in stage2_1 stage2_1_latcls stage2_2; do
  timeout 120 python bench.py --steps 5 --warmup 3 --config $cfg --no-inference --no-cpu-baseline > gpurun_out/final2_$cfg.json 2> gpurun_out/final2_$cfg.err
  echo "$cfg exit $?"
done
python - <<'PY'
import json
for n in ("stage1", "stage2_1", "stage2_1_latcls", "stage2_2"):
    try:
        d = json.load(open(f"gpurun_out/final2_{n}.json"))
        print(n, d.get("ms_per_step"), d.get("value"), d.get("gpu_launches"), (d.get("inference") or {}).get("best"))
    except Exception as e:
        print(n, "ERR", e)
PY
