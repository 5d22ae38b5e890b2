#!/bin/bash
# guarded A/B run: targeted tests of the changed kernels first, the full -m gpu suite, then short benches of the default
# build and of the variants named on the command line as  tag:ENV=VALUE[,ENV=VALUE...]
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_chain.py tests/test_gpu_frames.py -x -q --tb=short -p no:cacheprovider > gpurun_out/ab_quick.log 2>&1
rc=$?; tail -4 gpurun_out/ab_quick.log
if [ $rc -ne 0 ]; then echo "quick tests failed rc=$rc"; fi
(timeout 600 python -m pytest tests -m gpu -q --tb=short --maxfail=30 -p no:cacheprovider > gpurun_out/ab_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/ab_tests.log)
tail -3 gpurun_out/ab_tests.log
run() { # tag, env assignments
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --no-inference --no-cpu-baseline > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_$tag.json"))
    print("$tag", d["ms_per_step"], d["value"], d["gpu_launches"])
except Exception as e:
    print("$tag ERR", e)
PY
}
run default TDVC_AB=default
for spec in "$@"; do
  tag=${spec%%:*}; envs=${spec#*:}
  run $tag ${envs//,/ }
done
