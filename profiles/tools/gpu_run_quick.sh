#!/bin/bash
# guarded run for a risky kernel change: a short targeted test first (a hung kernel costs 150 s, not the suite's 20 min),
# then the full suite + bench
TAG=${1:-x}; shift
mkdir -p gpurun_out
timeout 150 python -m pytest "$@" -x -q --tb=short -p no:cacheprovider > gpurun_out/${TAG}_quick.log 2>&1
rc=$?
tail -5 gpurun_out/${TAG}_quick.log
if [ $rc -ne 0 ]; then echo "quick tests failed rc=$rc: stopping"; exit 1; fi
bash profiles/tools/gpu_run_tests_bench.sh $TAG
