#!/bin/bash
# GPU session B of round 2: full -m gpu suite, then A/B benches of the second-generation wgrad kernel and the grouped
# frame convolutions.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=40 -p no:cacheprovider > gpurun_out/r2b_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_tests.log)
tail -4 gpurun_out/r2b_tests.log
if ! grep -q "pytest exit 0" gpurun_out/r2b_tests.log; then
  (TDVC_WGRAD2_HALOED=0 timeout 600 python -m pytest tests/test_gpu_frames.py tests/test_gpu_tc.py -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/r2b_tests_nohalo.log 2>&1; tail -3 gpurun_out/r2b_tests_nohalo.log)
fi
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench exit $?"
TDVC_WGRAD2=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-inference --no-cpu-baseline > gpurun_out/r2b_bench_oldwgrad.json 2> gpurun_out/r2b_bench_oldwgrad.err; echo "old wgrad exit $?"
TDVC_GROUPED_FRAMES=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-inference --no-cpu-baseline > gpurun_out/r2b_bench_nogframes.json 2> gpurun_out/r2b_bench_nogframes.err; echo "no gframes exit $?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-inference --no-cpu-baseline --config stage2_1_latcls > gpurun_out/r2b_bench_latcls.json 2> gpurun_out/r2b_bench_latcls.err; echo "latcls exit $?"
python - <<'PY'
import json
for n in ("r2b_bench", "r2b_bench_oldwgrad", "r2b_bench_nogframes", "r2b_bench_latcls"):
    try:
        d = json.load(open("gpurun_out/" + n + ".json"))
        print(n, d["ms_per_step"], d["value"], d["gpu_launches"])
    except Exception as e:
        print(n, "ERR", e)
PY
