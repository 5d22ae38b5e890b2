#!/bin/bash
# GPU session D of round 2: suite with the taps-on-M weight-gradient form and the parallel fold, A/B benches.
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q --tb=short --maxfail=40 -p no:cacheprovider > gpurun_out/r2d_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2d_tests.log)
tail -4 gpurun_out/r2d_tests.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-inference > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench exit $?"
TDVC_WGRAD2_TAPSM=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-inference --no-cpu-baseline > gpurun_out/r2d_bench_notapsm.json 2> gpurun_out/r2d_bench_notapsm.err; echo "no tapsm exit $?"
python - <<'PY'
import json
for n in ("r2d_bench", "r2d_bench_notapsm"):
    try:
        d = json.load(open("gpurun_out/" + n + ".json"))
        print(n, d["ms_per_step"], d["value"], d["gpu_launches"])
    except Exception as e:
        print(n, "ERR", e)
PY
