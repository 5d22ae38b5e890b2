#!/bin/bash
# GPU session C of round 2: full -m gpu suite (incl. the bf16-resident MRF stage), then A/B benches.
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q --tb=short --maxfail=40 -p no:cacheprovider > gpurun_out/r2c_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_tests.log)
tail -4 gpurun_out/r2c_tests.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-inference > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench exit $?"
TDVC_MRF_CHAIN=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-inference --no-cpu-baseline > gpurun_out/r2c_bench_nochain.json 2> gpurun_out/r2c_bench_nochain.err; echo "no chain exit $?"
TDVC_WGRAD2_KINNER=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-inference --no-cpu-baseline > gpurun_out/r2c_bench_kinner.json 2> gpurun_out/r2c_bench_kinner.err; echo "k inner exit $?"
python - <<'PY'
import json
for n in ("r2c_bench", "r2c_bench_nochain", "r2c_bench_kinner"):
    try:
        d = json.load(open("gpurun_out/" + n + ".json"))
        print(n, d["ms_per_step"], d["value"], d["gpu_launches"])
    except Exception as e:
        print(n, "ERR", e)
PY
