#!/bin/bash
# full -m gpu suite + the default bench line; TAG names the output files
TAG=${1:-x}
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q --tb=short --maxfail=40 -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_tests.log)
tail -4 gpurun_out/${TAG}_tests.log
timeout 400 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS:---no-inference} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench.json"))
    print("${TAG}", d["ms_per_step"], d["value"], d["gpu_launches"], d["e2e"]["value"])
except Exception as e:
    print("ERR", e)
PY
